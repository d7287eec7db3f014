import sys, torch
sys.path.insert(0, '/root/repo')
import pct_b200
n = 4096
b = pct_b200.PctBatch(n, 1, container_size=(1.0, 1.0, 1.0), continuous=True, sample_from_distribution=True, seed=1234)
b.reset()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for t in range(100):
    b.step(leaf_idx=b.random_policy(99, t))
torch.cuda.synchronize()
K = 300
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
for t in range(K):
    flush.zero_()
    ev[t][0].record()
    b.step(leaf_idx=b.random_policy(99, 100 + t))
    ev[t][1].record()
torch.cuda.synchronize()
ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
print("continuous n=%d: %.4f ms/step -> %.2fM env-steps/s  flags %d" % (n, sum(ms) / K, n * K / sum(ms) / 1e3, int(b._info[:, 1].max())), flush=True)
