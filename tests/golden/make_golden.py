"""Records golden vectors from the UNMODIFIED reference (run in the build container only; /root/reference is absent
on the GPU box).  Usage:  python tests/golden/make_golden.py

For every (setting, trajectory) it drives pct_envs.PctDiscrete0.PackingDiscrete (under oracle/ref_shim.py) with an
injected item stream and the deterministic test policy, and stores the full float64 observation of every step
INCLUDING the terminal observation of failed placements, rewards, dones, the info dict values, plus, at a few
steps, the internal EMS list and the ordered candidate list (EMSPoint) of the reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import ref_shim  # noqa: E402
from harness import ITEM_SET, make_stream, policy_pick  # noqa: E402
from pct_oracle import make_continuous_stream  # noqa: E402


def record(D, setting, seed, env_id, steps, lnes="EMS"):
    stream = make_stream(seed, env_id, steps + 64, setting)
    env = D.PackingDiscrete(setting=setting, container_size=[10, 10, 10], item_set=ITEM_SET, internal_node_holder=80,
                            leaf_node_holder=50, shuffle=False, LNES=lnes)
    env.box_creator = ref_shim.make_stream_creator(D, [tuple(r) if setting == 3 else tuple(int(v) for v in r[:3]) for r in stream])
    env.test = True
    obs0 = env.reset()
    rows, obs, rew, done, counter, ratio, ems, cand = [], [obs0], [], [], [], [], {}, {}
    o = obs0
    for t in range(steps):
        if t % 25 == 0 and lnes == "EMS":
            ems[t] = np.array([list(map(int, e)) for e in env.space.EMS])
            cand[t] = np.array(env.space.EMSPoint(env.next_box, setting)).reshape(-1, 6)
        _, row = policy_pick(o, 80, 50, seed, env_id, t)
        o, r, d, info = env.step(row)
        rows.append(row); obs.append(o.copy()); rew.append(r); done.append(d)
        counter.append(info["counter"]); ratio.append(info.get("ratio", -1.0))
        if d:
            o = env.reset()
            obs.append(o.copy())  # the observation after the caller's reset follows the terminal one
    out = dict(stream=stream, rows=np.array(rows), obs=np.array(obs), reward=np.array(rew), done=np.array(done),
               counter=np.array(counter), ratio=np.array(ratio))
    for t in ems:
        out["ems_%d" % t] = ems[t]
        out["cand_%d" % t] = cand[t]
    return out


def record_continuous(Cm, setting, seed, env_id, steps):
    stream = make_continuous_stream(seed, env_id, steps + 64, setting)
    env = Cm.PackingContinuous(setting=setting, container_size=[1, 1, 1], item_set=[(0.1, 0.1, 0.1)], internal_node_holder=80,
                               leaf_node_holder=50, shuffle=False, sample_from_distribution=False)
    env.size_minimum = 0.1
    env.space.low_bound = 0.1  # = sample_left_bound of the sample_from_distribution configuration (C:bin3D.py:25-27)
    env.box_creator = ref_shim.make_stream_creator(Cm, [tuple(float(v) for v in (r if setting == 3 else r[:3])) for r in stream])
    env.test = True
    o = env.reset()
    rows, obs, rew, done, counter, ratio = [], [o.copy()], [], [], [], []
    for t in range(steps):
        _, row = policy_pick(o, 80, 50, seed, env_id, t)
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
    for lnes in ("EV", "EP", "CP", "FC"):
        for setting in (1, 2):
            rec = record(D, setting, 31 + setting, 2, 160, lnes)
            path = os.path.join(HERE, "discrete_%s_s%d.npz" % (lnes, setting))
            np.savez_compressed(path, setting=setting, lnes=lnes, **rec)
            print(path, os.path.getsize(path) // 1024, "KiB", "episodes", int(rec["done"].sum()))
    for setting in (1, 2, 3):
        rec = record_continuous(Cm, setting, 4242, 1, 200)
        path = os.path.join(HERE, "continuous_s%d_t0.npz" % setting)
        np.savez_compressed(path, setting=setting, **rec)
        print(path, os.path.getsize(path) // 1024, "KiB", "episodes", int(rec["done"].sum()))
    for setting, steps in ((1, 220), (2, 220), (3, 220)):
        for k, (seed, env_id) in enumerate(((2024, 0), (77, 5))):
            rec = record(D, setting, seed, env_id, steps)
            path = os.path.join(HERE, "discrete_s%d_t%d.npz" % (setting, k))
            np.savez_compressed(path, setting=setting, seed=seed, env_id=env_id, **rec)
            print(path, os.path.getsize(path) // 1024, "KiB", "episodes", int(rec["done"].sum()))


if __name__ == "__main__":
    main()
