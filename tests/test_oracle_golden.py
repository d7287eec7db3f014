"""CPU: the C oracle must reproduce the golden vectors recorded from the unmodified reference (tests/golden/*.npz,
made by tests/golden/make_golden.py).  This is what pins the oracle when /root/reference is not mounted."""
import glob
import os

import numpy as np
import pytest

from pct_oracle import OracleDiscrete

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "discrete_s*.npz")))
GOLD += sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "discrete_[ECF]*_s*.npz")))  # EV / EP / CP / FC schemes


def test_golden_files_present():
    assert len(GOLD) >= 14


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_oracle_replays_reference_trajectory(path):
    g = np.load(path)
    setting = int(g["setting"])
    lnes = str(g["lnes"]) if "lnes" in g.files else "EMS"
    env = OracleDiscrete(setting, stream=g["stream"], lnes=lnes)
    obs = g["obs"]
    k = 0
    o = env.reset()
    assert np.array_equal(o, obs[k]); k += 1
    for t in range(len(g["rows"])):
        if t % 25 == 0 and lnes == "EMS":  # internals pinned at a few steps: EMS list and ordered candidate list of the reference
            assert np.array_equal(env.ems(), g["ems_%d" % t]), "EMS list, step %d" % t
            cand, _ = env.candidates()
            assert np.array_equal(cand, g["cand_%d" % t].reshape(-1, 6)), "candidate order, step %d" % t
        o, r, d, info = env.step(g["rows"][t])
        assert np.array_equal(o, obs[k]), "observation after step %d (done=%s)" % (t, d); k += 1
        assert r == g["reward"][t] and d == bool(g["done"][t]) and info["counter"] == g["counter"][t]
        if d:
            assert info["ratio"] == g["ratio"][t]
            o = env.reset()
            assert np.array_equal(o, obs[k]); k += 1
    assert k == len(obs)


GOLD_C = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "continuous_s*.npz")))


@pytest.mark.parametrize("path", GOLD_C, ids=[os.path.basename(p) for p in GOLD_C])
def test_continuous_oracle_replays_reference_trajectory(path):
    from pct_oracle import OracleContinuous
    g = np.load(path)
    env = OracleContinuous(int(g["setting"]), stream=g["stream"])
    obs, k = g["obs"], 0
    o = env.reset()
    assert np.array_equal(o, obs[k]); k += 1
    for t in range(len(g["rows"])):
        o, r, d, info = env.step(g["rows"][t])
        assert np.array_equal(o, obs[k]), "observation after step %d (done=%s)" % (t, d); k += 1
        assert r == g["reward"][t] and d == bool(g["done"][t]) and info["counter"] == g["counter"][t]
        if d:
            assert info["ratio"] == g["ratio"][t]
            o = env.reset()
            assert np.array_equal(o, obs[k]); k += 1
    assert k == len(obs) and len(GOLD_C) == 3


# ---- non-default configurations (tests/golden/make_golden_cases.py): other containers / item sets / holder sizes -----------
from harness import CASES, CONT_CASES, KNOWN_DIVERGENT  # noqa: E402

GOLD_CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "case_*.npz")))
GOLD_CCASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ccase_*.npz")))


def _replay(env, g):
    obs, k = g["obs"], 0
    o = env.reset()
    assert np.array_equal(o, obs[k]); k += 1
    for t in range(len(g["rows"])):
        o, r, d, info = env.step(g["rows"][t])
        assert np.array_equal(o, obs[k]), "observation after step %d (done=%s)" % (t, d); k += 1
        assert r == g["reward"][t] and d == bool(g["done"][t]) and info["counter"] == g["counter"][t] and "error" not in info
        if d:
            assert info["ratio"] == g["ratio"][t]
            o = env.reset()
            assert np.array_equal(o, obs[k]); k += 1
    assert k == len(obs)


def test_case_files_present():
    assert len(GOLD_CASES) == len(CASES) and len(GOLD_CCASES) == len(CONT_CASES)


@pytest.mark.parametrize("path", GOLD_CASES, ids=[os.path.basename(p) for p in GOLD_CASES])
@pytest.mark.parametrize("alias", [False, True], ids=["default", "alias"])
def test_oracle_replays_reference_on_other_configurations(path, alias):
    g = np.load(path)
    c = CASES[str(g["name"])]
    env = OracleDiscrete(c["setting"], container_size=c["container"], internal_node_holder=c["nb"], leaf_node_holder=c["nl"],
                         size_minimum=min(min(i) for i in c["items"]), stream=g["stream"], lnes=c["lnes"])
    env.set_alias_mode(alias)
    _replay(env, g)


def _divergent(path):
    return any(os.path.basename(path) == "ccase_%s.npz" % n for n in KNOWN_DIVERGENT)


@pytest.mark.parametrize("path", GOLD_CCASES, ids=[os.path.basename(p) for p in GOLD_CCASES])
@pytest.mark.parametrize("alias", [False, True], ids=["default", "alias"])
def test_continuous_oracle_replays_reference_on_other_configurations(path, alias):
    """alias = the oracle reads `up_edges` values that ARE the upper box's own Stack object live, as the reference's Python objects behave
    (DESIGN.md section 3); the default mode holds snapshots like the CUDA kernels.  The two modes differ on ONE record, found by the
    fresh-seed soak after round 1's GPU budget was spent: there the default mode is a known divergence from the reference (strict xfail:
    it flips to a failure the day the kernels and the default mode implement the aliasing)."""
    from pct_oracle import OracleContinuous
    if _divergent(path) and not alias:
        pytest.xfail("known divergence of the snapshot semantics (GPU-equal default) from the reference's object aliasing; alias mode replays it")
    g = np.load(path)
    c = CONT_CASES[str(g["name"])]
    env = OracleContinuous(c["setting"], container_size=c["container"], internal_node_holder=c["nb"], leaf_node_holder=c["nl"],
                           size_minimum=c["low"], stream=g["stream"])
    env.set_alias_mode(alias)
    _replay(env, g)


def test_known_divergence_is_real():
    """the default mode really does NOT replay the alias record (so the xfail above is not hiding a pass)"""
    from pct_oracle import OracleContinuous
    for n in KNOWN_DIVERGENT:
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ccase_%s.npz" % n))
        c = CONT_CASES[n]
        env = OracleContinuous(c["setting"], container_size=c["container"], internal_node_holder=c["nb"], leaf_node_holder=c["nl"],
                               size_minimum=c["low"], stream=g["stream"])
        env.set_alias_mode(False)
        with pytest.raises(AssertionError):
            _replay(env, g)
