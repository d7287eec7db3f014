"""Build-container soak: the restated heuristic baselines (oracle/pct_oracle_heuristics.py over the C oracles) against the UNMODIFIED
heuristic.py on fresh random datasets / streams, both domains.   python scratch/soak_heuristics_vs_reference.py [minutes]"""
import contextlib
import io
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import ref_shim  # noqa: E402
import pct_oracle_heuristics as OH  # noqa: E402
from harness import CONT_ITEM_SET, ITEM_SET  # noqa: E402
from pct_oracle import OracleContinuous, OracleDiscrete, make_continuous_stream, rnd_u64  # noqa: E402


def main():
    minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
    D, Cm = ref_shim.load_reference()
    sys.argv = sys.argv[:1]
    import heuristic as H
    fns = {"LSAH": H.LASH, "OnlineBPH": H.OnlineBPH, "BR": H.BR, "DBL": H.DBL, "HM": H.heightmap_min}
    t0, n, bad, seed = time.time(), 0, 0, 500000 + int(time.time()) % 100000
    while time.time() - t0 < minutes * 60:
        for setting in (1, 2, 3):
            seed += 1
            # ---- discrete, dataset mode ----
            data = np.ones((6, 45, 4 if setting == 3 else 3))
            for t in range(6):
                for k in range(45):
                    data[t, k, :3] = ITEM_SET[rnd_u64(seed, t, k) % 125]
                    if setting == 3:
                        data[t, k, 3] = (1 + rnd_u64(seed ^ 0x5555, t, k) % 999) / 1000.0
            with tempfile.TemporaryDirectory() as tmp:
                path = os.path.join(tmp, "set.pt")
                torch.save([t.tolist() for t in data], path)
                for name, fn in fns.items():
                    class Rec(D.PackingDiscrete):
                        def reset(self):
                            if hasattr(self, "packed"):
                                self.log.append([list(p) for p in self.packed])
                            return super().reset()
                    env = Rec(setting=setting, container_size=[10, 10, 10], item_set=ITEM_SET, data_name=path, load_test_data=True)
                    env.log = []
                    with contextlib.redirect_stdout(io.StringIO()):
                        fn(env, 3)
                    rows = []
                    for t in data[1:]:
                        rows += [t if t.shape[1] == 4 else np.concatenate([t, np.ones((len(t), 1))], 1), [[100, 100, 100, 1.0]]]
                    orc = OracleDiscrete(setting, stream=np.concatenate(rows))
                    orc.set_trajectory_length(46)
                    got = [r[2] for r in OH.run_episodes(name, orc, 3, item_set=ITEM_SET)]
                    n += 1
                    if got != env.log[:3]:
                        bad += 1
                        print("MISMATCH discrete", name, "setting", setting, "seed", seed, flush=True)
            # ---- continuous, stream mode ----
            stream = make_continuous_stream(seed, 2, 500, setting)
            for name, fn in (("LSAH", H.LASH), ("OnlineBPH", H.OnlineBPH), ("BR", H.BR)):
                class RecC(Cm.PackingContinuous):
                    def reset(self):
                        if hasattr(self, "packed"):
                            self.log.append([list(map(float, p)) for p in self.packed])
                        return super().reset()
                env = RecC(setting=setting, container_size=[1, 1, 1], item_set=CONT_ITEM_SET, sample_from_distribution=False)
                env.size_minimum = 0.1
                env.space.low_bound = 0.1
                env.box_creator = ref_shim.make_stream_creator(Cm, [tuple(float(v) for v in (r if setting == 3 else r[:3])) for r in stream])
                env.test = True
                env.log = []
                with contextlib.redirect_stdout(io.StringIO()):
                    fn(env, 4)
                got = [r[2] for r in OH.run_episodes(name, OracleContinuous(setting, stream=stream), 4, item_set=CONT_ITEM_SET)]
                n += 1
                if got != env.log[:4]:
                    bad += 1
                    print("MISMATCH continuous", name, "setting", setting, "seed", seed, flush=True)
        print("%d runs, %d mismatches, %.0f s" % (n, bad, time.time() - t0), flush=True)
    print("done: %d runs, %d mismatches" % (n, bad))


if __name__ == "__main__":
    main()
