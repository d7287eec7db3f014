"""Records the reference's own evaluation loop (evaluation_tools.py:7-52: LoadBoxCreator dataset, one env, episode after
episode) on the CONTINUOUS env (pct_envs.PctContinuous0, test mode: items rounded to 3 decimals, C:bin3D.py:84-87) with a
small synthetic dataset -> tests/golden/eval_cont_s{1,2,3}.npz.  Needs /root/reference (build container only).

    python tests/golden/make_eval_golden_continuous.py
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import ref_shim  # noqa: E402
from harness import sequential_eval  # noqa: E402
from pct_oracle import rnd_u64  # noqa: E402

N_TRAJ, TRAJ_LEN = 17, 60


def dataset(setting, seed=4711):
    """items U(0.1, 0.5) with up to 4 decimals for x (the env rounds to 3), 3 for y, z from the five levels or U (setting 2)"""
    d = np.ones((N_TRAJ, TRAJ_LEN, 4 if setting == 3 else 3))
    u = lambda salt, t, k: (rnd_u64(seed + setting + salt, t, k) >> 11) / float(1 << 53)
    for t in range(N_TRAJ):
        for k in range(TRAJ_LEN):
            d[t, k, 0] = round(0.1 + 0.4 * u(1, t, k), 4)
            d[t, k, 1] = round(0.1 + 0.4 * u(2, t, k), 3)
            d[t, k, 2] = round(0.1 + 0.4 * u(3, t, k), 3) if setting == 2 else [0.1, 0.2, 0.3, 0.4, 0.5][rnd_u64(seed ^ 0x77, t, k) % 5]
            if setting == 3:
                d[t, k, 3] = (1 + rnd_u64(seed ^ 0x5555, t, k) % 999) / 1000.0
    return d


def main():
    _, Cm = ref_shim.load_reference()
    for setting in (1, 2, 3):
        data = dataset(setting)
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "set.pt")
            torch.save([t.tolist() for t in data], path)
            env = Cm.PackingContinuous(setting=setting, container_size=[1, 1, 1], item_set=None, data_name=path, load_test_data=True,
                                       internal_node_holder=80, leaf_node_holder=50, shuffle=False, sample_from_distribution=True,
                                       sample_left_bound=0.1, sample_right_bound=0.5)
            rec = sequential_eval(lambda ep: (env, env.reset()), N_TRAJ - 1)
        ratio = np.array([r[0] for r in rec])
        counter = np.array([r[1] for r in rec])
        plen = np.array([len(r[2]) for r in rec])
        flat = np.array([p for r in rec for p in r[2]], dtype=np.float64).reshape(-1, 7)
        out = os.path.join(HERE, "eval_cont_s%d.npz" % setting)
        np.savez_compressed(out, setting=setting, data=data, ratio=ratio, counter=counter, packed_len=plen, packed_flat=flat)
        print(out, os.path.getsize(out), "B  mean ratio %.4f  mean length %.2f" % (ratio.mean(), counter.mean()))


if __name__ == "__main__":
    main()
