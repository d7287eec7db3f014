"""Build-container soak: the C oracle against the UNMODIFIED reference on fresh seeds (needs /root/reference).
    python scratch/soak_oracle_vs_reference.py [minutes]
Discrete (all settings, all LNES schemes, the non-default configurations of tests/harness.CASES) and continuous (unit and non-unit
containers) trajectories under the shared synthetic policy; every observation, reward, done, counter and ratio must be equal."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import ref_shim  # noqa: E402
from harness import CASES, CONT_CASES, ITEM_SET, case_stream, cont_case_stream  # noqa: E402
import make_golden_cases as M  # noqa: E402
from pct_oracle import OracleContinuous, OracleDiscrete  # noqa: E402


def replay(orc, rec, steps):
    o = orc.reset()
    k = 0
    assert np.array_equal(o, rec["obs"][k]); k += 1
    for t in range(steps):
        o, r, d, info = orc.step(rec["rows"][t])
        assert np.array_equal(o, rec["obs"][k]), ("obs", t); k += 1
        assert r == rec["reward"][t] and d == bool(rec["done"][t]) and info["counter"] == rec["counter"][t], ("scalars", t)
        if d:
            assert info["ratio"] == rec["ratio"][t], ("ratio", t)
            o = orc.reset()
            assert np.array_equal(o, rec["obs"][k]), ("reset obs", t); k += 1


def make_oracle(c, seed, env_id):
    if "items" in c:
        return OracleDiscrete(c["setting"], container_size=c["container"], internal_node_holder=c["nb"], leaf_node_holder=c["nl"],
                              size_minimum=min(min(i) for i in c["items"]), stream=case_stream(c, seed, env_id, c["steps"] + 64), lnes=c["lnes"])
    return OracleContinuous(c["setting"], container_size=c["container"], internal_node_holder=c["nb"], leaf_node_holder=c["nl"],
                            size_minimum=c["low"], stream=cont_case_stream(c, seed, env_id, c["steps"] + 64))


def main():
    minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
    D, Cm = ref_shim.load_reference()
    base = dict(CASES)
    for s in (1, 2, 3):
        base["default_s%d" % s] = dict(setting=s, container=(10, 10, 10), items=ITEM_SET, nb=80, nl=50, steps=250, lnes="EMS")
        for l in ("EV", "EP", "CP", "FC"):
            base["%s_s%d" % (l, s)] = dict(setting=s, container=(10, 10, 10), items=ITEM_SET, nb=80, nl=50, steps=80, lnes=l)
    cont = dict(CONT_CASES)
    for s in (1, 2, 3):
        cont["unit_s%d" % s] = dict(setting=s, container=(1.0, 1.0, 1.0), lo=0.1, hi=0.5, low=0.1, nb=80, nl=50, steps=200)
    t0, n, steps_total, seed = time.time(), 0, 0, 100000 + int(time.time()) % 100000
    while time.time() - t0 < minutes * 60:
        for name, c in list(base.items()) + list(cont.items()):
            seed += 1
            try:
                if "items" in c:
                    rec = M.record_case(D, c, seed, n % 7)
                    orc = OracleDiscrete(c["setting"], container_size=c["container"], internal_node_holder=c["nb"], leaf_node_holder=c["nl"],
                                         size_minimum=min(min(i) for i in c["items"]), stream=case_stream(c, seed, n % 7, c["steps"] + 64), lnes=c["lnes"])
                else:
                    rec = M.record_cont_case(Cm, c, seed, n % 7)
                    orc = OracleContinuous(c["setting"], container_size=c["container"], internal_node_holder=c["nb"], leaf_node_holder=c["nl"],
                                           size_minimum=c["low"], stream=cont_case_stream(c, seed, n % 7, c["steps"] + 64))
                try:
                    replay(orc, rec, c["steps"])
                except AssertionError as ex:  # classify: does the oracle's alias mode replay it?
                    mk = make_oracle(c, seed, n % 7)
                    mk.set_alias_mode(True)
                    try:
                        replay(mk, rec, c["steps"])
                        kind = "object aliasing (alias mode replays it)"
                    except AssertionError:
                        kind = "NOT explained by alias mode (LAPACK tie or new)"
                    print("MISMATCH", name, "seed", seed, "env", n % 7, ex.args, "->", kind, flush=True)
            except AssertionError as ex:
                print("MISMATCH", name, "seed", seed, "env", n % 7, ex.args, flush=True)
            n += 1
            steps_total += c["steps"]
            if time.time() - t0 > minutes * 60:
                break
        print("%d trajectories, %d env-steps, %.0f s" % (n, steps_total, time.time() - t0), flush=True)
    print("done: %d trajectories, %d env-steps" % (n, steps_total))


if __name__ == "__main__":
    main()
