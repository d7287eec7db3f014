import sys, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pct_b200
items=[(i,j,k) for i in range(1,6) for j in range(1,6) for k in range(1,6)]
for setting in (1,2):
    b=pct_b200.PctBatch(48, setting, item_set=items, seed=5)
    b.reset()
    for t in range(25): b.step(leaf_idx=b.random_policy(9,t))
    torch.cuda.synchronize()
    b.close()
b=pct_b200.PctBatch(24, 1, container_size=(1.0,1.0,1.0), continuous=True, sample_from_distribution=True, seed=2)
b.reset()
for t in range(15): b.step(leaf_idx=b.random_policy(9,t))
torch.cuda.synchronize()
for l in ("FC","CP","EP","EV"):
    b=pct_b200.PctBatch(8, 1, item_set=items, seed=5, LNES=l); b.reset()
    for t in range(10): b.step(leaf_idx=b.random_policy(9,t))
    torch.cuda.synchronize()
# round 2: shuffle, setting 3, host path (zero-copy + staged), continuous setting 2 / shuffle
b=pct_b200.PctBatch(16, 3, item_set=items, seed=6, shuffle=True); b.reset()
for t in range(12): b.step(leaf_idx=b.random_policy(9,t))
torch.cuda.synchronize(); b.close()
b=pct_b200.PctBatch(16, 2, container_size=(1.0,1.0,1.0), continuous=True, sample_from_distribution=True, seed=3, shuffle=True); b.reset()
for t in range(8): b.step(leaf_idx=b.random_policy(9,t))
torch.cuda.synchronize(); b.close()
n=32
b=pct_b200.PctBatch(n, 1, item_set=items, seed=7)
obs_h=torch.empty((n,b.obs_len),dtype=torch.float32,pin_memory=True).numpy(); rew=torch.empty((n,),dtype=torch.float32,pin_memory=True).numpy()
done=torch.empty((n,),dtype=torch.uint8,pin_memory=True).numpy(); info=torch.empty((n,8),dtype=torch.int32,pin_memory=True).numpy(); idx=np.zeros((n,),dtype=np.int32)
b.reset_host(obs_h)
for t in range(8): b.step_host(obs_h,rew,done,info,leaf_idx=idx)
obs_u=np.empty((n,b.obs_len),dtype=np.float32)  # unpinned: staged path
for t in range(4): b.step_host(obs_u,rew,done,info,leaf_idx=idx)
b.close()
print("sanitizer workload done")
# round 2, later: the opt-in fork-join continuation kernel (device-side piece queue), both domains
os.environ["PCT_B200_WALK"] = "fork"
b=pct_b200.PctBatch(40, 1, item_set=items, seed=8); b.reset()
for t in range(25): b.step(leaf_idx=b.random_policy(9,t))
torch.cuda.synchronize(); b.close()
b=pct_b200.PctBatch(16, 1, container_size=(1.0,1.0,1.0), continuous=True, sample_from_distribution=True, seed=4); b.reset()
for t in range(12): b.step(leaf_idx=b.random_policy(9,t))
torch.cuda.synchronize(); b.close()
del os.environ["PCT_B200_WALK"]
print("sanitizer workload (fork-join) done")
