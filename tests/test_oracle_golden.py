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
from harness import CASES, CONT_CASES, NEEDS_ALIAS, NEEDS_ALIAS_D  # noqa: E402

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
@pytest.mark.parametrize("alias", [None, False, True], ids=["default", "snapshot", "alias"])
def test_oracle_replays_reference_on_other_configurations(path, alias):
    g = np.load(path)
    c = CASES[str(g["name"])]
    env = OracleDiscrete(c["setting"], container_size=c["container"], internal_node_holder=c["nb"], leaf_node_holder=c["nl"],
                         size_minimum=min(min(i) for i in c["items"]), stream=g["stream"], lnes=c["lnes"])
    if alias is not None:
        env.set_alias_mode(alias)
    if alias is False and str(g["name"]) in NEEDS_ALIAS_D:
        # the parting trajectories: the snapshot semantics of round 1 must NOT reproduce the reference here (sensitivity of the records)
        with pytest.raises(AssertionError):
            _replay(env, g)
        return
    _replay(env, g)


def _divergent(path):
    return any(os.path.basename(path) == "ccase_%s.npz" % n for n in NEEDS_ALIAS)


@pytest.mark.parametrize("path", GOLD_CCASES, ids=[os.path.basename(p) for p in GOLD_CCASES])
@pytest.mark.parametrize("alias", [None, False, True], ids=["default", "snapshot", "alias"])
def test_continuous_oracle_replays_reference_on_other_configurations(path, alias):
    """alias = the oracle reads `up_edges` values that ARE the upper box's own Stack object live, as the reference's Python objects behave
     — the DEFAULT of the oracle and of the CUDA kernels since round 2.  The snapshot mode of round 1 (kept behind
    PCT_ORACLE_ALIAS=0 / PCT_B200_ALIAS=0 for the sensitivity tests) differs on ONE record found by the fresh-seed soak; it is skipped for
    that record here and shown to fail on it in test_snapshot_mode_does_not_replay_the_alias_record."""
    from pct_oracle import OracleContinuous
    if _divergent(path) and alias is False:
        pytest.skip("snapshot semantics do not reproduce this record (that is the point of it): see test_snapshot_mode_does_not_replay_the_alias_record")
    g = np.load(path)
    c = CONT_CASES[str(g["name"])]
    env = OracleContinuous(c["setting"], container_size=c["container"], internal_node_holder=c["nb"], leaf_node_holder=c["nl"],
                           size_minimum=c["low"], stream=g["stream"])
    if alias is not None:
        env.set_alias_mode(alias)
    _replay(env, g)


def test_snapshot_mode_does_not_replay_the_alias_record():
    """the snapshot mode really does NOT replay the alias record (so the default's pass above means something)"""
    from pct_oracle import OracleContinuous
    for n in NEEDS_ALIAS:
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ccase_%s.npz" % n))
        c = CONT_CASES[n]
        env = OracleContinuous(c["setting"], container_size=c["container"], internal_node_holder=c["nb"], leaf_node_holder=c["nl"],
                               size_minimum=c["low"], stream=g["stream"])
        env.set_alias_mode(False)
        with pytest.raises(AssertionError):
            _replay(env, g)
