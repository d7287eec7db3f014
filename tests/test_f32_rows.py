"""float32 ACTION ROWS on the continuous env.  train_tools.py:66-67 sends `selected_leaf_node.cpu().numpy()` — float32 — to
PackingContinuous.step; the reference then computes `round(np.float32 - np.float32, 6)` (pct_envs/PctContinuous0/bin3D.py:151-173) and places the box at
float32-VALUED coordinates (rounding noise ~3e-8 that depends on the numpy version's scalar promotion rules).  The product widens float32 rows to float64
and rounds to 6 decimals, which returns the exact 6-decimal values.  BASELINE.json's contract for this domain is 1e-6 on coordinates; every DISCRETE
outcome — feasibility mask, done, counter — must be identical.  tests/golden/f32rows_s*.npz were recorded from the unmodified reference driven with
float32 rows (tests/golden/make_golden_f32rows.py); here they are replayed on the oracle (CPU) and on the kernels with float32 action tensors (GPU).
scratch/soak_f32_rows.py runs the same comparison live against the reference on fresh seeds."""
import glob
import os

import numpy as np
import pytest

from pct_oracle import OracleContinuous

G = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "f32rows_s*.npz")))
TOL = 1e-6  # BASELINE.json north star: "within 1e-6 on leaf coordinates for the continuous setting"


def _check(o, ref, what):
    assert np.array_equal(o.reshape(-1, 9)[:, 8], ref.reshape(-1, 9)[:, 8]), "%s: valid / feasibility flags differ" % what
    assert np.array_equal(o.reshape(-1, 9)[:, 6:8], ref.reshape(-1, 9)[:, 6:8])
    assert np.abs(o - ref).max() <= TOL, "%s: coordinates differ by %g" % (what, np.abs(o - ref).max())


def test_files_present():
    assert len(G) == 3


@pytest.mark.parametrize("path", G, ids=[os.path.basename(p) for p in G])
def test_oracle_with_widened_float32_rows_follows_the_reference(path):
    g = np.load(path)
    env = OracleContinuous(int(g["setting"]), stream=g["stream"])
    obs, k = g["obs"], 0
    o = env.reset()
    assert np.array_equal(o, obs[k]); k += 1
    for t in range(len(g["rows"])):
        assert g["rows"].dtype == np.float32
        o, r, d, info = env.step(g["rows"][t].astype(np.float64))
        _check(o, obs[k], "step %d" % t); k += 1
        assert d == bool(g["done"][t]) and info["counter"] == g["counter"][t]
        assert abs(r - g["reward"][t]) <= 1e-6
        if d:
            assert abs(info["ratio"] - g["ratio"][t]) <= 1e-6
            o = env.reset()
            assert np.array_equal(o, obs[k]); k += 1
    assert k == len(obs)


@pytest.mark.gpu
@pytest.mark.parametrize("path", G, ids=[os.path.basename(p) for p in G])
def test_gpu_float32_action_rows_follow_the_reference(path):
    torch = pytest.importorskip("torch")
    import pct_b200
    g = np.load(path)
    env = pct_b200.PctBatch(1, int(g["setting"]), container_size=(1.0, 1.0, 1.0), continuous=True, obs_dtype=torch.float64, item_stream=g["stream"][None],
                            size_minimum=0.1, auto_reset=False)
    obs, k = g["obs"], 0
    o = env.reset().cpu().numpy()[0]
    assert np.array_equal(o, obs[k]); k += 1
    for t in range(len(g["rows"])):
        a = torch.from_numpy(g["rows"][t][None].copy()).cuda()
        assert a.dtype == torch.float32
        ob, r, d, info = env.step(actions=a)  # float32 rows: action_f64 = 0 in pct_step
        o = ob.cpu().numpy()[0]
        _check(o, obs[k], "step %d" % t); k += 1
        rec = env.decode_info(info)
        assert bool(d.cpu().numpy()[0]) == bool(g["done"][t]) and rec["counter"][0] == g["counter"][t] and not rec["flags"][0]
        if bool(g["done"][t]):
            o = env.reset().cpu().numpy()[0]
            assert np.array_equal(o, obs[k]); k += 1
    assert k == len(obs)
    env.close()
