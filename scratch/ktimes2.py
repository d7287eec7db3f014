import sys, ctypes as C, numpy as np, torch
sys.path.insert(0,'/root/repo')
import pct_b200
items=[(i,j,k) for i in range(1,6) for j in range(1,6) for k in range(1,6)]
n=4096
b=pct_b200.PctBatch(n, 1, item_set=items, seed=1234)
buf=torch.zeros((n,16),dtype=torch.int64,device='cuda')
b.L.pct_debug_set_timer_buffer.argtypes=[C.c_void_p,C.c_void_p]
b.L.pct_debug_set_timer_buffer(b.h, C.c_void_p(buf.data_ptr()))
b.reset()
recs=[]
for t in range(130):
    buf.zero_(); b.step(leaf_idx=b.random_policy(4321,t)); torch.cuda.synchronize()
    if t>=100: recs.append(buf.cpu().numpy().copy())
a=np.concatenate(recs)
pk=a[:,12].astype(np.uint64); cyc=(pk>>np.uint64(24)).astype(np.float64); vis=((pk>>np.uint64(8))&np.uint64(255)).astype(int); nls=(pk&np.uint64(255)).astype(int)
sec=a[:,[3,7,11,13,14]].astype(np.float64)
names=['supports','hull+pip','distrib','persist','next+COM']
print('mean slowest-candidate cycles %.0f ; section means:'%cyc.mean(), dict(zip(names, sec.mean(0).round())))
idx=np.argsort(cyc)[-10:]
for i in idx: print('cycles %8.0f visits %3d lstsq %d | '%(cyc[i],vis[i],nls[i]) + ' '.join('%s %.0f'%(nm,v) for nm,v in zip(names,sec[i])))
