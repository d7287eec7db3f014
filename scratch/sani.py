import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
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
print("sanitizer workload done")
