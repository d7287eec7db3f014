"""Where does a pct_step_host round trip go?  python scratch/e2e_breakdown.py  (4096 envs, setting 1)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pct_b200
ITEM_SET = [(i, j, k) for i in range(1, 6) for j in range(1, 6) for k in range(1, 6)]
n = 4096
b = pct_b200.PctBatch(n, 1, item_set=ITEM_SET, seed=1234)
ol = b.obs_len
obs_h = torch.empty((n, ol), dtype=torch.float32, pin_memory=True).numpy()
rew_h = torch.empty((n,), dtype=torch.float32, pin_memory=True).numpy()
done_h = torch.empty((n,), dtype=torch.uint8, pin_memory=True).numpy()
info_h = torch.empty((n, 8), dtype=torch.int32, pin_memory=True).numpy()
idx_h = torch.zeros((n,), dtype=torch.int32).pin_memory().numpy()
b.reset_host(obs_h)
for t in range(300):
    b.step_host(obs_h, rew_h, done_h, info_h, leaf_idx=idx_h)
K = 300
t0 = time.perf_counter()
for t in range(K):
    b.step_host(obs_h, rew_h, done_h, info_h, leaf_idx=idx_h)
t1 = time.perf_counter()
print("step_host alone (first-leaf policy, no host policy work): %.1f us / step" % ((t1 - t0) / K * 1e6))
# device-resident step with the same policy, events
idx_d = torch.zeros((n,), dtype=torch.int32, device="cuda")
for t in range(50):
    b.step(leaf_idx=idx_d)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for t in range(K):
    b.step(leaf_idx=idx_d)
ev1.record(); torch.cuda.synchronize()
print("device-resident step (no L2 flush): %.1f us / step" % (ev0.elapsed_time(ev1) / K * 1e3))
t0 = time.perf_counter()
for t in range(K):
    b.step(leaf_idx=idx_d)
    torch.cuda.synchronize()
t1 = time.perf_counter()
print("device-resident step + sync per step (wall): %.1f us / step" % ((t1 - t0) / K * 1e6))
gid = np.arange(n, dtype=np.uint64); GOLD = np.uint64(0x9E3779B97F4A7C15)
def sm64(x):
    with np.errstate(over="ignore"):
        x = x + GOLD; z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))
t0 = time.perf_counter()
for t in range(K):
    nvalid = info_h[:, 5].astype(np.uint64)
    with np.errstate(over="ignore"):
        r = sm64(sm64(np.uint64(4321) ^ (gid * GOLD)) + np.uint64(t))
    idx_h[:] = np.where(nvalid > 0, r % np.maximum(nvalid, np.uint64(1)), 0).astype(np.int32)
t1 = time.perf_counter()
print("numpy host policy alone: %.1f us / step" % ((t1 - t0) / K * 1e6))
