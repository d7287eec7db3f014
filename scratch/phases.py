import sys, ctypes as C, numpy as np, torch
sys.path.insert(0,'/root/repo')
import pct_b200
items=[(i,j,k) for i in range(1,6) for j in range(1,6) for k in range(1,6)]
n=int(sys.argv[1]) if len(sys.argv)>1 else 4096
setting=int(sys.argv[2]) if len(sys.argv)>2 else 1
b=pct_b200.PctBatch(n, setting, item_set=items, seed=1234)
buf=torch.zeros((n,8),dtype=torch.int64,device='cuda')
b.L.pct_debug_set_timer_buffer.argtypes=[C.c_void_p,C.c_void_p]
b.L.pct_debug_set_timer_buffer(b.h, C.c_void_p(buf.data_ptr()))
b.reset()
acc=[]
for t in range(150):
    b.step(leaf_idx=b.random_policy(4321,t))
    if t>=100: acc.append(buf.cpu().numpy().copy())
a=np.stack(acc).astype(np.float64)  # T,n,8
d=np.diff(a,axis=2)
names=['load+decode','real drop','genems+draw','build cand','feasibility','persist','obs+store']
print('n',n,'setting',setting,'total mean %.0f max %.0f cycles'%(a[:,:,7].mean(), a[:,:,7].max()))
for i,nm in enumerate(names): print('%-14s mean %9.0f  p99 %9.0f  max %9.0f'%(nm, d[:,:,i].mean(), np.percentile(d[:,:,i],99), d[:,:,i].max()))
