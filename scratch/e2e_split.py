import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
import pct_b200
items=[(i,j,k) for i in range(1,6) for j in range(1,6) for k in range(1,6)]
n=4096
b=pct_b200.PctBatch(n,1,item_set=items,seed=1234)
ol=b.obs_len
pin=lambda shape,dt: torch.empty(shape,dtype=dt,pin_memory=True).numpy()
obs_h=pin((n,ol),torch.float32); rew_h=pin((n,),torch.float32); done_h=pin((n,),torch.uint8); info_h=pin((n,8),torch.int32); idx_h=pin((n,),torch.int32)
b.reset_host(obs_h)
def pol(t):
    nv=(obs_h.reshape(n,-1,9)[:,80:130,8]==1).sum(1)
    idx_h[:]=np.where(nv>0, (np.arange(n)*7+t)%np.maximum(nv,1),0)
for t in range(10): pol(t); b.step_host(obs_h,rew_h,done_h,info_h,leaf_idx=idx_h)
tp=ts=0
for t in range(100):
    t0=time.perf_counter(); pol(t); t1=time.perf_counter(); b.step_host(obs_h,rew_h,done_h,info_h,leaf_idx=idx_h); t2=time.perf_counter()
    tp+=t1-t0; ts+=t2-t1
print('host policy %.3f ms, step_host %.3f ms'%(tp*10, ts*10))
d=torch.empty((n,ol),dtype=torch.float32,device='cuda'); hp=torch.empty((n,ol),dtype=torch.float32,pin_memory=True)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(50): hp.copy_(d,non_blocking=True); torch.cuda.synchronize()
print('D2H 19.3MB pinned: %.3f ms'%((time.perf_counter()-t0)*20))
