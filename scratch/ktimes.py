import sys, ctypes as C, numpy as np, torch
sys.path.insert(0,'/root/repo')
import pct_b200
items=[(i,j,k) for i in range(1,6) for j in range(1,6) for k in range(1,6)]
n=int(sys.argv[1]) if len(sys.argv)>1 else 4096
setting=int(sys.argv[2]) if len(sys.argv)>2 else 1
b=pct_b200.PctBatch(n, setting, item_set=items, seed=1234)
buf=torch.zeros((n,16),dtype=torch.int64,device='cuda')
b.L.pct_debug_set_timer_buffer.argtypes=[C.c_void_p,C.c_void_p]
b.L.pct_debug_set_timer_buffer(b.h, C.c_void_p(buf.data_ptr()))
b.reset()
acc=[]
for t in range(140):
    buf[:,12]=0; b.step(leaf_idx=b.random_policy(4321,t))
    if t>=100: torch.cuda.synchronize(); acc.append(buf.cpu().numpy().copy().reshape(n,4,4))
a=np.stack(acc).astype(np.float64)  # T,n,4,4
print('n',n,'setting',setting)
t0=a[:,:,0,0].min(axis=1, keepdims=True)
for k,name in enumerate(['apply','cand','feas_emit']):
    st=a[:,:,k,0]-t0; en=a[:,:,k,1]-t0; cyc=a[:,:,k,2]
    print('%-10s start[min %.0f mean %.0f max %.0f] end[mean %.0f max %.0f] us | per-env cycles mean %.0f p50 %.0f p99 %.0f max %.0f'%(name, st.min(1).mean()/1e3, st.mean()/1e3, st.max(1).mean()/1e3, en.mean()/1e3, en.max(1).mean()/1e3, cyc.mean(), np.median(cyc), np.percentile(cyc,99), cyc.max(1).mean()))
# per-group view of the last recorded step
last=a[-1]; t0l=last[:,0,0].min()
G=int(__import__('os').environ.get('PCT_B200_GROUPS','1'))
for g in range(G):
    lo=n*g//G; hi=n*(g+1)//G
    print('group',g,' '.join('%s[%.0f..%.0f]'%(nm,(last[lo:hi,k,0].min()-t0l)/1e3,(last[lo:hi,k,1].max()-t0l)/1e3) for k,nm in enumerate(['apply','cand','feas'])))
pk=np.stack([x[:,3,0] for x in acc]).astype(np.uint64)  # slot 12 = [3][0]
cyc=(pk>>np.uint64(24)).astype(np.float64); vis=((pk>>np.uint64(8))&np.uint64(255)).astype(int); nls=(pk&np.uint64(255)).astype(int)
print('slowest candidate per env: cycles mean %.0f p99 %.0f max %.0f'%(cyc.mean(), np.percentile(cyc,99), cyc.max()))
idx=np.argsort(cyc.ravel())[-12:]
for i in idx: print('   cycles %8.0f visits %3d lstsq %3d'%(cyc.ravel()[i], vis.ravel()[i], nls.ravel()[i]))
