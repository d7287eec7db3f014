"""Records the reference's UNMODIFIED heuristic baselines on the CONTINUOUS env (tools.py:217-218 allows LSAH, OnlineBPH
and BR there: heuristic.py LASH :138-226, OnlineBPH :364-424, BR :500-577 driving pct_envs.PctContinuous0.PackingContinuous)
-> tests/golden/heur_cont_s{1,2,3}.npz (item stream + per-episode packed lists).  Needs /root/reference.

    python tests/golden/make_heuristic_golden_continuous.py

Items come from an injected stream with RandomBoxCreator's draw discipline (oracle/ref_shim.make_stream_creator), the way
tests/golden/make_golden.py records the continuous env: same values as the sample_from_distribution configuration
(x, y [, z] = round(U(0.1, 0.5), 3), C:bin3D.py:103-115), known to both sides.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import ref_shim  # noqa: E402
from harness import CONT_ITEM_SET  # noqa: E402
from pct_oracle import make_continuous_stream  # noqa: E402

EPISODES = {"LSAH": 6, "OnlineBPH": 6, "BR": 6}
SEED, ENV_ID, STREAM_LEN = 9191, 3, 600


def main():
    _, Cm = ref_shim.load_reference()
    sys.argv = [sys.argv[0]]
    import heuristic as H  # the reference module, unmodified
    fns = {"LSAH": H.LASH, "OnlineBPH": H.OnlineBPH, "BR": H.BR}

    class Recording(Cm.PackingContinuous):
        def reset(self):
            if hasattr(self, "packed"):
                self.log.append([list(map(float, p)) for p in self.packed])
            return super().reset()

    for setting in (1, 2, 3):
        stream = make_continuous_stream(SEED + setting, ENV_ID, STREAM_LEN, setting)
        rec = {}
        for name, fn in fns.items():
            env = Recording(setting=setting, container_size=[1, 1, 1], item_set=CONT_ITEM_SET, internal_node_holder=80,
                            leaf_node_holder=50, shuffle=False, sample_from_distribution=False)
            env.size_minimum = 0.1
            env.space.low_bound = 0.1  # = sample_left_bound of the sample_from_distribution configuration (C:bin3D.py:25-27)
            env.box_creator = ref_shim.make_stream_creator(Cm, [tuple(float(v) for v in (r if setting == 3 else r[:3])) for r in stream])
            env.test = True
            env.log = []
            with contextlib.redirect_stdout(io.StringIO()):
                fn(env, EPISODES[name])
            eps = env.log[:EPISODES[name]]
            assert len(eps) == EPISODES[name] and env.box_creator.pos < STREAM_LEN
            rec["len_" + name] = np.array([len(e) for e in eps])
            rec["flat_" + name] = np.array([p for e in eps for p in e], dtype=np.float64).reshape(-1, 7)
            print(setting, name, "lengths", rec["len_" + name].tolist(), "draws", env.box_creator.pos, flush=True)
        out = os.path.join(HERE, "heur_cont_s%d.npz" % setting)
        np.savez_compressed(out, setting=setting, stream=stream, **rec)
        print(out, os.path.getsize(out), "B")


if __name__ == "__main__":
    main()
