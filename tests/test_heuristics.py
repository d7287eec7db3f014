"""Heuristic baselines: the CPU restatement (oracle/pct_oracle_heuristics.py) against records of the
reference's unmodified heuristic.py (tests/golden/heur_s*.npz), and the batched CUDA kernel against the restatement."""
import glob
import os

import numpy as np
import pytest

import pct_oracle_heuristics as OH
from harness import CONT_ITEM_SET, ITEM_SET
from pct_oracle import OracleContinuous, OracleDiscrete

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "heur_s*.npz")))
RECORDED = ("LSAH", "OnlineBPH", "BR", "MACS", "DBL", "HM")


def golden(path, name):
    g = np.load(path)
    off = np.concatenate([[0], np.cumsum(g["len_" + name])])
    return int(g["setting"]), g["data"], [g["flat_" + name][off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]


def dataset_stream(data, first, count):
    """LoadBoxCreator order (episode k plays trajectory k+1) as ONE item stream with a sentinel after every trajectory"""
    rows = []
    for k in range(first, first + count):
        t = data[(k + 1) % len(data)]
        t = t if t.shape[1] == 4 else np.concatenate([t, np.ones((len(t), 1))], 1)
        rows += [t, [[100, 100, 100, 1.0]]]
    return np.concatenate(rows)


def test_golden_present():
    assert len(GOLDEN) == 3


@pytest.mark.parametrize("name", RECORDED)
@pytest.mark.parametrize("path", GOLDEN)
def test_restated_heuristics_replay_reference(path, name):
    setting, data, packed = golden(path, name)
    L = data.shape[1] + 1
    env = OracleDiscrete(setting, stream=dataset_stream(data, 0, len(packed) + 1))
    env.set_trajectory_length(L)
    rec = OH.run_episodes(name, env, len(packed), item_set=ITEM_SET)
    assert [r[2] for r in rec] == packed


# ---- continuous domain: LSAH / OnlineBPH / BR (tools.py:217-218) ---------------------------------------------------------
GOLDEN_C = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "heur_cont_s*.npz")))
CONT_NAMES = ("LSAH", "OnlineBPH", "BR")


def golden_c(path, name):
    g = np.load(path)
    off = np.concatenate([[0], np.cumsum(g["len_" + name])])
    return int(g["setting"]), g["stream"], [g["flat_" + name][off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]


def test_golden_continuous_present():
    assert len(GOLDEN_C) == 3


@pytest.mark.parametrize("name", CONT_NAMES)
@pytest.mark.parametrize("path", GOLDEN_C)
def test_restated_heuristics_replay_reference_continuous(path, name):
    """float64 packed lists [x, y, z, lx, ly, lz, 0] equal the unmodified heuristic.py on PackingContinuous, exactly"""
    setting, stream, packed = golden_c(path, name)
    env = OracleContinuous(setting, stream=stream)
    rec = OH.run_episodes(name, env, len(packed), item_set=CONT_ITEM_SET)
    assert [r[2] for r in rec] == packed


def test_continuous_rejects_grid_heuristics():
    env = OracleContinuous(1, stream=np.array([[0.2, 0.2, 0.2, 1.0]]))
    with pytest.raises(ValueError):
        OH.run_episodes("DBL", env, 1)


# ---- GPU: the batched selection kernel ------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", RECORDED)
@pytest.mark.parametrize("path", GOLDEN)
def test_batched_heuristic_replays_reference(path, name):
    """dataset mode, 4 envs sharing the episodes: per-episode packed lists equal the unmodified reference's"""
    from pct_b200.heuristics import run_heuristic
    setting, data, packed = golden(path, name)
    (mean, var, length), rec = run_heuristic(name, setting, len(packed), item_set=ITEM_SET, data=list(data), n_envs=4, return_episodes=True)
    assert rec["packed"] == packed
    ratios = [sum(p[0] * p[1] * p[2] for p in ep) / 1000.0 for ep in packed]
    assert abs(mean - np.mean(ratios)) < 1e-12 and abs(var - np.var(ratios)) < 1e-12 and length == np.mean([len(ep) for ep in packed])


@pytest.mark.gpu
@pytest.mark.parametrize("name,setting,n_envs", [("RANDOM", 1, 6), ("RANDOM", 2, 6), ("LSAH", 1, 24), ("DBL", 3, 12), ("OnlineBPH", 2, 24),
                                                  ("BR", 1, 12), ("HM", 2, 6), ("MACS", 1, 4)])
def test_batched_heuristic_on_the_fly_items(name, setting, n_envs):
    """items from the device generator (RandomBoxCreator mode), one env per oracle env, two episodes each"""
    from harness import make_stream
    from pct_b200.heuristics import run_heuristic
    seed = 77 + setting
    _, rec = run_heuristic(name, setting, 2 * n_envs, item_set=ITEM_SET, n_envs=n_envs, seed=seed, return_episodes=True)
    for e in range(n_envs):
        env = OracleDiscrete(setting, stream=make_stream(seed, e, 400, setting))
        ref = OH.run_episodes(name, env, 2, item_set=ITEM_SET, seed=seed, gid=e)
        for j in range(2):
            assert rec["packed"][e + j * n_envs] == ref[j][2], (name, e, j)
            assert abs(rec["ratio"][e + j * n_envs] - ref[j][0]) < 1e-12


class _FacadeView(object):
    """what heuristic.py touches on the env, routed to the drop-in facade exactly the way heuristic.py calls it"""

    def __init__(self, env):
        self.env, self.container, self.setting = env, tuple(env.bin_size), env.setting

    def ems(self):
        return [[int(v) for v in e] for e in self.env.space.EMS]

    def drop_box_virtual(self, d, lx, ly):
        return self.env.space.drop_box_virtual(list(d), (lx, ly), False, self.env.next_den, self.env.setting, returnH=True)

    def plain(self):
        return np.array(self.env.space.plain)

    @property
    def next_box(self):
        return self.env.next_box


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["LSAH", "HM"])
def test_single_env_facade_serves_the_reference_heuristic_loop(name, tmp_path):
    """env.space.drop_box_virtual / env.space.EMS / env.next_box = [...] / env.step([0, lx, ly]) on PackingDiscrete"""
    import torch
    import pct_b200
    setting, data, packed = golden(GOLDEN[0], name)
    ds = os.path.join(str(tmp_path), "set.pt")
    torch.save([t.tolist() for t in data], ds)
    env = pct_b200.PackingDiscrete(setting=setting, container_size=[10, 10, 10], item_set=ITEM_SET, data_name=ds, load_test_data=True)
    view = _FacadeView(env)
    env.reset()
    state, got = OH.fresh_state(view.container), []
    while len(got) < 2:
        c = OH.choose(name, view, state, ITEM_SET)
        if c is None:
            got.append(env.packed)
            env.reset()
            state = OH.fresh_state(view.container)
            continue
        env.next_box = c[0]
        env.step([0, c[1], c[2]])
        OH.note_placement(state, c)
    assert got == packed[:2]
    feas, hmap = env.space.drop_box_virtual([2, 2, 2], (0, 0), False, 1.0, setting, False, True)
    assert hmap.shape == (10, 10) and hmap[0, 0] == 2 and hmap[5, 5] == 0 and feas
    env.close()
