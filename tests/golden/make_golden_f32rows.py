"""Golden vectors for the float32 ACTION ROWS the trainer really sends to the continuous env (train_tools.py:66-67: `selected_leaf_node.cpu().numpy()`
of a float32 observation tensor -> pct_envs/PctContinuous0/bin3D.py:151-173), recorded from the UNMODIFIED reference (needs /root/reference).

With float32 rows the reference computes `round(np.float32 - np.float32, 6)` and carries the float32-valued x / y / lx / ly into Space.drop_box, so its
box coordinates hold float32 rounding noise (~3e-8); the product widens float32 rows to float64 and rounds to 6 decimals, which returns the exact
6-decimal values.  The north-star contract for the continuous domain is 1e-6 on coordinates (not bit equality); what must NOT differ is any discrete
outcome: feasibility masks, done flags, counters.  These records pin that: tests/test_f32_rows.py replays them on the oracle (CPU) and on the kernels (GPU).

    python tests/golden/make_golden_f32rows.py        -> tests/golden/f32rows_s{1,2,3}.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import ref_shim  # noqa: E402
from pct_oracle import make_continuous_stream, rnd_u64  # noqa: E402


def record(Cm, setting, seed, env_id, steps):
    stream = make_continuous_stream(seed, env_id, steps + 64, setting)
    env = Cm.PackingContinuous(setting=setting, container_size=[1.0, 1.0, 1.0], item_set=[(0.1, 0.1, 0.1)], internal_node_holder=80, leaf_node_holder=50,
                               shuffle=False, sample_from_distribution=True, sample_left_bound=0.1, sample_right_bound=0.5)
    env.box_creator = ref_shim.make_stream_creator(Cm, [tuple(float(v) for v in (r if setting == 3 else r[:3])) for r in stream])
    env.test = True
    o = env.reset()
    rows, obs, rew, done, counter, ratio = [], [o.copy()], [], [], [], []
    for t in range(steps):
        o32 = o.astype(np.float32).reshape(-1, 9)  # VecPyTorch hands the policy a float32 tensor (envs.py:166-171)
        leaf = o32[80:130]
        nv = int((leaf[:, 8] == 1).sum())
        row = leaf[rnd_u64(seed, env_id, t) % nv].copy() if nv else np.zeros(9, dtype=np.float32)
        assert row.dtype == np.float32
        o, r, d, info = env.step(row)
        rows.append(row); obs.append(o.copy()); rew.append(r); done.append(d)
        counter.append(info["counter"]); ratio.append(info.get("ratio", -1.0))
        if d:
            o = env.reset()
            obs.append(o.copy())
    return dict(stream=stream, rows=np.array(rows, dtype=np.float32), obs=np.array(obs), reward=np.array(rew, dtype=np.float64), done=np.array(done),
                counter=np.array(counter), ratio=np.array(ratio, dtype=np.float64), setting=setting)


def main():
    _, Cm = ref_shim.load_reference()
    for setting in (1, 2, 3):
        rec = record(Cm, setting, 2024 + setting, 3, 260)
        path = os.path.join(HERE, "f32rows_s%d.npz" % setting)
        np.savez_compressed(path, **rec)
        print(path, os.path.getsize(path) // 1024, "KiB", "episodes", int(rec["done"].sum()), flush=True)


if __name__ == "__main__":
    main()
