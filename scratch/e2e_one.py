import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import pct_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
setting = int(sys.argv[2]) if len(sys.argv) > 2 else 1
items = [(i, j, k) for i in range(1, 6) for j in range(1, 6) for k in range(1, 6)]
b = pct_b200.PctBatch(n, setting, item_set=items, seed=1234)
ol = b.obs_len
obs_h = torch.empty((n, ol), dtype=torch.float32, pin_memory=True).numpy()
rew_h = torch.empty((n,), dtype=torch.float32, pin_memory=True).numpy()
done_h = torch.empty((n,), dtype=torch.uint8, pin_memory=True).numpy()
info_h = torch.empty((n, 8), dtype=torch.int32, pin_memory=True).numpy()
idx_h = torch.empty((n,), dtype=torch.int32, pin_memory=True).numpy()
rng = np.random.RandomState(0)
def pol():
    nv = info_h[:, 5]
    idx_h[:] = np.where(nv > 0, rng.randint(0, 1 << 30, n) % np.maximum(nv, 1), 0)
b.reset_host(obs_h)
info_h[:, 5] = (obs_h.reshape(n, -1, 9)[:, b.nb:b.nb + b.nl, 8] == 1).sum(1)
for t in range(100):
    pol(); b.step_host(obs_h, rew_h, done_h, info_h, leaf_idx=idx_h)
torch.cuda.synchronize()
K = 400
t0 = time.perf_counter()
for t in range(K):
    pol(); b.step_host(obs_h, rew_h, done_h, info_h, leaf_idx=idx_h)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
t1 = time.perf_counter()
for t in range(K): pol()
tp = time.perf_counter() - t1
print("n=%d s=%d  e2e %.3f ms/step  %.2fM env-steps/s   (host policy %.3f ms)" % (n, setting, dt / K * 1e3, n * K / dt / 1e6, tp / K * 1e3), flush=True)
