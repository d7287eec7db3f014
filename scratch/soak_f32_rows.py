"""Live soak for tests/test_f32_rows.py: the unmodified reference continuous env driven with float32 leaf rows (as train_tools.py:66-67 does) next to the
oracle fed the same rows widened to float64, fresh seeds: counts steps, discrete mismatches (mask / done / counter) and the largest coordinate difference.
python scratch/soak_f32_rows.py [minutes]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import ref_shim
from pct_oracle import OracleContinuous, make_continuous_stream, rnd_u64
minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
_, Cm = ref_shim.load_reference()
t_end, seed, steps, bad, maxd, trajs = time.time() + 60 * minutes, 900000, 0, 0, 0.0, 0
while time.time() < t_end:
    seed += 1
    setting = 1 + seed % 3
    stream = make_continuous_stream(seed, 1, 400, setting)
    env = Cm.PackingContinuous(setting=setting, container_size=[1.0, 1.0, 1.0], item_set=[(0.1, 0.1, 0.1)], internal_node_holder=80, leaf_node_holder=50,
                               shuffle=False, sample_from_distribution=True, sample_left_bound=0.1, sample_right_bound=0.5)
    env.box_creator = ref_shim.make_stream_creator(Cm, [tuple(float(v) for v in (r if setting == 3 else r[:3])) for r in stream])
    env.test = True
    orc = OracleContinuous(setting, stream=stream)
    o, q = env.reset(), orc.reset()
    for t in range(150):
        leaf = o.astype(np.float32).reshape(-1, 9)[80:130]
        nv = int((leaf[:, 8] == 1).sum())
        row = leaf[rnd_u64(seed, 1, t) % nv].copy() if nv else np.zeros(9, dtype=np.float32)
        o, r, d, info = env.step(row)
        q, r2, d2, info2 = orc.step(row.astype(np.float64))
        steps += 1
        dd = float(np.abs(o - q).max()); maxd = max(maxd, dd)
        if d != d2 or info["counter"] != info2["counter"] or not np.array_equal(o.reshape(-1, 9)[:, 8], q.reshape(-1, 9)[:, 8]) or dd > 1e-6:
            bad += 1
            print("MISMATCH seed", seed, "setting", setting, "step", t, d, d2, dd, flush=True)
            break
        if d:
            o, q = env.reset(), orc.reset()
    trajs += 1
print("trajectories %d, env-steps %d, discrete mismatches %d, max coordinate difference %.3g" % (trajs, steps, bad, maxd))
