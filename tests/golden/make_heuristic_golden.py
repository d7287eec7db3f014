"""Records the reference's UNMODIFIED heuristic baselines (heuristic.py: LASH, OnlineBPH, BR, MACS, DBL, heightmap_min) on a
small synthetic dataset -> tests/golden/heur_s{1,2,3}.npz (per-episode packed lists).  Needs /root/reference.

    python tests/golden/make_heuristic_golden.py
"""
import contextlib
import io
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
from harness import ITEM_SET  # noqa: E402
from make_eval_golden import dataset  # noqa: E402

EPISODES = {"LSAH": 6, "OnlineBPH": 6, "BR": 6, "MACS": 3, "DBL": 5, "HM": 5}


def main():
    D, _ = ref_shim.load_reference()
    sys.argv = [sys.argv[0]]
    import heuristic as H  # the reference module, unmodified (imports givenData / tools from /root/reference)
    fns = {"LSAH": H.LASH, "OnlineBPH": H.OnlineBPH, "BR": H.BR, "MACS": H.MACS, "DBL": H.DBL, "HM": H.heightmap_min}

    class Recording(D.PackingDiscrete):
        def reset(self):
            if hasattr(self, "packed"):
                self.log.append([list(p) for p in self.packed])
            return super().reset()

    for setting in (1, 2, 3):
        data = dataset(setting)
        rec = {}
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "set.pt")
            torch.save([t.tolist() for t in data], path)
            for name, fn in fns.items():
                env = Recording(setting=setting, container_size=[10, 10, 10], item_set=ITEM_SET, data_name=path, load_test_data=True,
                                internal_node_holder=80, leaf_node_holder=50)
                env.log = []
                with contextlib.redirect_stdout(io.StringIO()):
                    fn(env, EPISODES[name])
                eps = env.log[:EPISODES[name]]
                assert len(eps) == EPISODES[name]
                rec["len_" + name] = np.array([len(e) for e in eps])
                rec["flat_" + name] = np.array([p for e in eps for p in e], dtype=np.int64).reshape(-1, 7)
                print(setting, name, "lengths", rec["len_" + name].tolist(), flush=True)
        out = os.path.join(HERE, "heur_s%d.npz" % setting)
        np.savez_compressed(out, setting=setting, data=data, **rec)
        print(out, os.path.getsize(out), "B")


if __name__ == "__main__":
    main()
