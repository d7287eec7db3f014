"""How often does the reference's object aliasing (oracle alias mode) change a trajectory w.r.t. the snapshot semantics (default mode =
the kernels)?  Oracle-only: both modes side by side on the BASELINE item streams / synthetic policy.  python scratch/alias_rate.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
from harness import make_stream, policy_pick  # noqa: E402
from pct_oracle import OracleContinuous, OracleDiscrete, make_continuous_stream  # noqa: E402


def run(make, n_envs, steps, seed):
    div, total = 0, 0
    for e in range(n_envs):
        a, b = make(e), make(e)
        b.set_alias_mode(True)
        oa, ob = a.reset(), b.reset()
        for t in range(steps):
            _, row = policy_pick(oa, 80, 50, seed, e, t)
            oa, ra, da, _ = a.step(row)
            ob, rb, db, _ = b.step(row)
            total += 1
            if da != db or not np.array_equal(oa, ob):
                div += 1
                print("   diverges: env", e, "step", t, flush=True)
                break  # count trajectories that diverge (first event)
            if da:
                oa, ob = a.reset(), b.reset()
    return div, total


t0 = time.time()
for setting in (1, 3):
    d, n = run(lambda e: OracleDiscrete(setting, stream=make_stream(1234, e, 600, setting)), 1500, 250, 4321)
    print("discrete setting %d: %d diverging trajectories in %d env-steps (%.0f s)" % (setting, d, n, time.time() - t0), flush=True)
for setting in (1, 3):
    d, n = run(lambda e: OracleContinuous(setting, stream=make_continuous_stream(1234, e, 600, setting)), 1000, 250, 4321)
    print("continuous setting %d: %d diverging trajectories in %d env-steps (%.0f s)" % (setting, d, n, time.time() - t0), flush=True)
