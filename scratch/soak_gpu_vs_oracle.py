"""GPU <-> oracle soak (>= 1 M env-steps in the object-alias semantics, the default of both since round 2): N envs on the GPU
(device item generator, device policy keyed by the global env id) next to the threaded C oracle batch with the same generators, compared bit for bit
(float64 observations, reward sums, episode counts) every `chunk` steps.  Test infrastructure / evidence, not product.
python scratch/soak_gpu_vs_oracle.py [--envs 4096] [--steps 300] [--chunk 50] [--settings 1,3,2] [--continuous]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import pct_b200  # noqa: E402
import pct_oracle  # noqa: E402

ITEM_SET = [(i, j, k) for i in range(1, 6) for j in range(1, 6) for k in range(1, 6)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--chunk", type=int, default=50)
    ap.add_argument("--settings", default="1,3,2")
    ap.add_argument("--continuous", action="store_true")
    ap.add_argument("--seed", type=int, default=20260923)
    a = ap.parse_args()
    total, t0 = 0, time.time()
    for setting in [int(s) for s in a.settings.split(",")]:
        n, seed, pseed = a.envs, a.seed + setting, 777
        if a.continuous:
            gpu = pct_b200.PctBatch(n, setting, container_size=(1.0, 1.0, 1.0), continuous=True, sample_from_distribution=True, seed=seed, obs_dtype=torch.float64)
            cpu = pct_oracle.OracleBatchContinuous(n, setting, seed, pseed)
        else:
            gpu = pct_b200.PctBatch(n, setting, item_set=ITEM_SET, seed=seed, obs_dtype=torch.float64)
            cpu = pct_oracle.OracleBatch(n, setting, ITEM_SET, seed, pseed)
        gpu.reset()
        rsum = torch.zeros(n, dtype=torch.float64, device=gpu.device)
        nd = torch.zeros(n, dtype=torch.int64, device=gpu.device)
        flags = 0
        for c0 in range(0, a.steps, a.chunk):
            for t in range(c0, min(a.steps, c0 + a.chunk)):
                obs, r, d, info = gpu.step(leaf_idx=gpu.random_policy(pseed, t))
                rsum += r.double(); nd += d.long()
                flags |= int(info[:, 1].max())
            cpu.run(min(a.steps, c0 + a.chunk) - c0)
            o_ref, r_ref, nd_ref = cpu.get()
            o = obs.cpu().numpy()
            if not np.array_equal(o_ref, o):
                bad = np.unique(np.argwhere(o_ref != o)[:, 0])
                print(json.dumps({"setting": setting, "continuous": a.continuous, "MISMATCH_after_steps": min(a.steps, c0 + a.chunk), "envs": bad[:16].tolist(), "n_bad": int(len(bad))}))
                sys.exit(1)
            assert np.array_equal(nd.cpu().numpy(), nd_ref.astype(np.int64)), "episode counts differ"
        assert flags == 0, "capacity flags %d" % flags
        total += n * a.steps
        print(json.dumps({"setting": setting, "continuous": a.continuous, "envs": n, "steps": a.steps, "env_steps": n * a.steps, "episodes": int(nd.sum()),
                          "bit_exact": True, "oracle_threads": cpu.threads}), flush=True)
        gpu.close(); cpu.close()
    print(json.dumps({"total_env_steps": total, "seconds": time.time() - t0, "result": "all observations, episode counts bit-exact GPU == oracle"}))


if __name__ == "__main__":
    main()
