"""GPU: batched heuristic baselines on the CONTINUOUS env (LSAH / OnlineBPH / BR, tools.py:217-218) — the selection kernel
csrc/pct_heuristics_continuous.cuh behind pct_heuristic_actions_f64, against (a) records of the reference's unmodified
heuristic.py on PackingContinuous (tests/golden/heur_cont_s*.npz) and (b) the CPU restatement on per-env item streams.
"""
import glob
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import pct_oracle_heuristics as OH  # noqa: E402
from harness import CONT_ITEM_SET  # noqa: E402
from pct_oracle import OracleContinuous, make_continuous_stream  # noqa: E402

pytestmark = pytest.mark.gpu

GOLDEN_C = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "heur_cont_s*.npz")))
NAMES = ("LSAH", "OnlineBPH", "BR")


def _golden(path, name):
    g = np.load(path)
    off = np.concatenate([[0], np.cumsum(g["len_" + name])])
    return int(g["setting"]), g["stream"], [g["flat_" + name][off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("path", GOLDEN_C)
def test_batched_continuous_heuristic_replays_reference(path, name):
    """one env on the recorded stream: per-episode packed lists [x, y, z, lx, ly, lz, 0] equal the unmodified reference's, exactly"""
    from pct_b200.heuristics import run_heuristic
    setting, stream, packed = _golden(path, name)
    (mean, var, length), rec = run_heuristic(name, setting, len(packed), container_size=(1.0, 1.0, 1.0), item_set=CONT_ITEM_SET,
                                             continuous=True, item_stream=stream[None], n_envs=1, return_episodes=True)
    assert rec["packed"] == packed
    ratios = [sum(p[0] * p[1] * p[2] for p in ep) for ep in packed]
    assert abs(mean - np.mean(ratios)) < 1e-12 and length == np.mean([len(ep) for ep in packed])


@pytest.mark.parametrize("name,setting,n_envs", [("LSAH", 1, 24), ("LSAH", 2, 12), ("OnlineBPH", 3, 24), ("OnlineBPH", 2, 12), ("BR", 1, 24),
                                                  ("BR", 3, 12)])
def test_batched_continuous_heuristic_on_the_fly_items(name, setting, n_envs):
    """sample_from_distribution items from the device generator (C:bin3D.py:103-115), one oracle env per GPU env, two episodes each"""
    from pct_b200.heuristics import run_heuristic
    seed = 311 + setting
    _, rec = run_heuristic(name, setting, 2 * n_envs, container_size=(1.0, 1.0, 1.0), item_set=CONT_ITEM_SET, continuous=True,
                           sample_from_distribution=True, sample_left_bound=0.1, sample_right_bound=0.5, n_envs=n_envs, seed=seed,
                           return_episodes=True)
    for e in range(n_envs):
        env = OracleContinuous(setting, stream=make_continuous_stream(seed, e, 400, setting))
        ref = OH.run_episodes(name, env, 2, item_set=CONT_ITEM_SET)
        for j in range(2):
            assert rec["packed"][e + j * n_envs] == ref[j][2], (name, e, j)
            assert abs(rec["ratio"][e + j * n_envs] - ref[j][0]) < 1e-12


def test_continuous_rejects_other_heuristics():
    import pct_b200
    b = pct_b200.PctBatch(4, 1, container_size=(1.0, 1.0, 1.0), continuous=True, sample_from_distribution=True, seed=1)
    b.reset()
    with pytest.raises(pct_b200.PctError, match="only LSAH, OnlineBPH, and BR"):
        b.heuristic_actions("DBL")
    with pytest.raises(pct_b200.PctError, match="pct_set_item_set"):
        b.heuristic_actions("BR")
    b.close()


class _FacadeViewC(object):
    """what heuristic.py touches on the env, routed to the drop-in PackingContinuous exactly the way heuristic.py calls it"""

    def __init__(self, env):
        self.env, self.container, self.setting = env, tuple(float(c) for c in env.bin_size), env.setting

    def ems(self):
        return [list(e) for e in self.env.space.EMS]

    def drop_box_virtual(self, d, lx, ly):
        return self.env.space.drop_box_virtual(list(d), (lx, ly), False, self.env.next_den, self.env.setting, returnH=True)

    @property
    def next_box(self):
        return self.env.next_box


@pytest.mark.parametrize("name", ["LSAH", "BR"])
def test_single_env_facade_serves_the_reference_heuristic_loop_continuous(name):
    """env.space.drop_box_virtual(returnH) / env.space.EMS / env.next_box = [...] / env.step([0, lx, ly]) on PackingContinuous"""
    import pct_b200
    setting, stream, packed = _golden(GOLDEN_C[0], name)
    env = pct_b200.PackingContinuous(setting=setting, container_size=[1, 1, 1], item_set=CONT_ITEM_SET, sample_from_distribution=False,
                                     item_stream=stream[None], size_minimum=0.1)
    view = _FacadeViewC(env)
    env.reset()
    state, got = OH.fresh_state(view.container), []
    while len(got) < 2:
        c = OH.choose(name, view, state, CONT_ITEM_SET)
        if c is None:
            got.append(env.packed)
            env.reset()
            state = OH.fresh_state(view.container)
            continue
        env.next_box = c[0]
        env.step([0, c[1], c[2]])
        OH.note_placement(state, c)
    for ep, ref in zip(got, packed[:2]):
        assert len(ep) == len(ref)
        assert np.allclose(np.array(ep), np.array(ref), rtol=0, atol=1e-12)  # `packed` of the facade derives sizes as hi - lo
    env.close()
