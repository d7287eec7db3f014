"""Golden vectors for NON-DEFAULT configurations, recorded from the UNMODIFIED reference (needs /root/reference):
other container / holder sizes and item sets — the configurations tests/test_gpu_discrete_cases.py drives on the GPU —
plus a second continuous trajectory per setting in a non-unit container.  Same record format as make_golden.py; the
configuration travels inside each file.

    python tests/golden/make_golden_cases.py        -> tests/golden/case_*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import ref_shim  # noqa: E402
from harness import CASES, CONT_CASES, case_stream, cont_case_stream, policy_pick  # noqa: E402


def record_case(D, c, seed, env_id):
    stream = case_stream(c, seed, env_id, c["steps"] + 64)
    env = D.PackingDiscrete(setting=c["setting"], container_size=list(c["container"]), item_set=c["items"], internal_node_holder=c["nb"],
                            leaf_node_holder=c["nl"], shuffle=False, LNES=c["lnes"])
    env.box_creator = ref_shim.make_stream_creator(D, [tuple(r) if c["setting"] == 3 else tuple(int(v) for v in r[:3]) for r in stream])
    env.test = True
    o = env.reset()
    rows, obs, rew, done, counter, ratio, ncand = [], [o.copy()], [], [], [], [], []
    for t in range(c["steps"]):
        _, row = policy_pick(o, c["nb"], c["nl"], c.get("pseed", seed), env_id, t)
        o, r, d, info = env.step(row)
        rows.append(row); obs.append(o.copy()); rew.append(r); done.append(d)
        counter.append(info["counter"]); ratio.append(info.get("ratio", -1.0))
        if d:
            o = env.reset()
            obs.append(o.copy())
    return dict(stream=stream, rows=np.array(rows), obs=np.array(obs), reward=np.array(rew), done=np.array(done),
                counter=np.array(counter), ratio=np.array(ratio))


def record_cont_case(Cm, c, seed, env_id):
    stream = cont_case_stream(c, seed, env_id, c["steps"] + 64)
    env = Cm.PackingContinuous(setting=c["setting"], container_size=list(c["container"]), item_set=[(0.1, 0.1, 0.1)], internal_node_holder=c["nb"],
                               leaf_node_holder=c["nl"], shuffle=False, sample_from_distribution=False)
    env.size_minimum = c["low"]
    env.space.low_bound = c["low"]  # = sample_left_bound of the sample_from_distribution configuration (C:bin3D.py:25-27)
    env.box_creator = ref_shim.make_stream_creator(Cm, [tuple(float(v) for v in (r if c["setting"] == 3 else r[:3])) for r in stream])
    env.test = True
    o = env.reset()
    rows, obs, rew, done, counter, ratio = [], [o.copy()], [], [], [], []
    for t in range(c["steps"]):
        _, row = policy_pick(o, c["nb"], c["nl"], seed, env_id, t)
        o, r, d, info = env.step(row)
        rows.append(row); obs.append(o.copy()); rew.append(r); done.append(d)
        counter.append(info["counter"]); ratio.append(info.get("ratio", -1.0))
        if d:
            o = env.reset()
            obs.append(o.copy())
    return dict(stream=stream, rows=np.array(rows), obs=np.array(obs), reward=np.array(rew), done=np.array(done),
                counter=np.array(counter), ratio=np.array(ratio))


def main():
    D, Cm = ref_shim.load_reference()
    only = sys.argv[1:]  # optional: names of the cases to (re-)record
    for name, c in CASES.items():
        if only and name not in only:
            continue
        rec = record_case(D, c, c.get("seed", 517), c.get("env", 1))
        path = os.path.join(HERE, "case_%s.npz" % name)
        np.savez_compressed(path, name=name, **rec)
        print(path, os.path.getsize(path) // 1024, "KiB", "episodes", int(rec["done"].sum()), flush=True)
    for name, c in CONT_CASES.items():
        if only and name not in only:
            continue
        rec = record_cont_case(Cm, c, c.get("seed", 519), c.get("env", 2))
        path = os.path.join(HERE, "ccase_%s.npz" % name)
        np.savez_compressed(path, name=name, **rec)
        print(path, os.path.getsize(path) // 1024, "KiB", "episodes", int(rec["done"].sum()), flush=True)


if __name__ == "__main__":
    main()
