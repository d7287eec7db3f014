"""Batched evaluation driver (pct_b200.evaluation) against records of the reference's own evaluation loop
(tests/golden/eval_s*.npz, made by tests/golden/make_eval_golden.py from the unmodified reference)."""
import glob
import os

import numpy as np
import pytest

from harness import ITEM_SET, eval_policy_torch, sequential_eval
from pct_oracle import OracleDiscrete

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "eval_s*.npz")))


def _golden(path):
    g = np.load(path)
    off = np.concatenate([[0], np.cumsum(g["packed_len"])])
    packed = [g["packed_flat"][off[i]:off[i + 1]].tolist() for i in range(len(g["ratio"]))]
    return int(g["setting"]), g["data"], g["ratio"], g["counter"], packed


def test_golden_present():
    assert len(GOLDEN) == 3


@pytest.mark.parametrize("path", GOLDEN)
def test_oracle_replays_reference_evaluation(path):
    """LoadBoxCreator discipline (episode k plays trajectory k+1, then the [100,100,100] sentinel) + env semantics, oracle vs reference"""
    setting, data, ratio, counter, packed = _golden(path)

    def make_env(ep):
        t = data[ep + 1]
        s = np.concatenate([t if t.shape[1] == 4 else np.concatenate([t, np.ones((len(t), 1))], 1), [[100, 100, 100, 1.0]] * 2])
        env = OracleDiscrete(setting, stream=s)
        return env, env.reset()

    rec = sequential_eval(make_env, len(ratio))
    assert [r[1] for r in rec] == counter.tolist()
    assert np.allclose([r[0] for r in rec], ratio, rtol=0, atol=1e-15)
    assert [r[2] for r in rec] == packed


def test_stream_layout():
    from pct_b200.evaluation import _streams
    trajs = [np.full((3 + (i % 2), 3), float(i)) for i in range(8)]  # trajectory i is filled with the value i
    s, traj_len, quota = _streams(trajs, 7, 3)
    assert traj_len == 5 and quota.tolist() == [3, 2, 2] and s.shape == (3, 4 * 5, 4)
    s = s.reshape(3, 4, 5, 4)
    for e in range(3):
        for j in range(4):
            if j < quota[e]:
                k = e + j * 3 + 1
                assert (s[e, j, :len(trajs[k]), :3] == k).all() and (s[e, j, len(trajs[k]):, :3] == 100).all()
            else:
                assert (s[e, j, :, :3] == 100).all()
    assert (s[..., 3] == 1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN)
@pytest.mark.parametrize("n_envs", [1, 7, 24])
def test_batched_evaluation_matches_reference(path, n_envs, tmp_path):
    from pct_b200.evaluation import evaluate_batched
    setting, data, ratio, counter, packed = _golden(path)
    out = evaluate_batched(list(data), setting, policy=eval_policy_torch, item_set=ITEM_SET, n_envs=n_envs, out_dir=str(tmp_path))
    assert out["length"].tolist() == counter.tolist()
    assert np.allclose(out["ratio"], ratio, rtol=0, atol=1e-15)
    assert out["packed"] == packed
    saved = np.load(os.path.join(str(tmp_path), "trajs.npy"), allow_pickle=True)
    assert [list(map(list, ep)) for ep in saved] == packed
    txt = open(os.path.join(str(tmp_path), "result.txt")).read()
    assert txt == "Evaluation using {} episodes\nMean ratio {:.5f}, mean length{:.5f}\n".format(len(ratio), np.mean(ratio), np.mean(counter))


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN[:1] + GOLDEN[2:])
def test_single_env_facade_replays_reference_evaluation(path, tmp_path):
    """the unchanged single-env loop of evaluation_tools.evaluate on the drop-in PackingDiscrete(load_test_data=True)"""
    import torch
    import pct_b200
    setting, data, ratio, counter, packed = _golden(path)
    ds = os.path.join(str(tmp_path), "set.pt")
    torch.save([t.tolist() for t in data], ds)
    env = pct_b200.PackingDiscrete(setting=setting, container_size=[10, 10, 10], item_set=ITEM_SET, data_name=ds, load_test_data=True,
                                   internal_node_holder=80, leaf_node_holder=50, shuffle=False, LNES="EMS")
    rec = sequential_eval(lambda ep: (env, env.reset()), 8)
    assert [r[1] for r in rec] == counter[:8].tolist()
    assert np.allclose([r[0] for r in rec], ratio[:8], rtol=0, atol=1e-15)
    assert [r[2] for r in rec] == packed[:8]
    env.close()


# ---- continuous env (PackingContinuous in test mode: items rounded to 3 decimals, C:bin3D.py:84-87) ------------------------
GOLDEN_C = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "eval_cont_s*.npz")))


def _round3(t):
    """what the continuous env does to a dataset item in test mode: Python round(v, 3) per size; the density column is untouched"""
    out = np.array(t, dtype=np.float64)
    for r in out:
        for i in range(3):
            r[i] = round(float(r[i]), 3)
    return out


def test_golden_continuous_present():
    assert len(GOLDEN_C) == 3


@pytest.mark.parametrize("path", GOLDEN_C)
def test_oracle_replays_reference_evaluation_continuous(path):
    from pct_oracle import OracleContinuous
    setting, data, ratio, counter, packed = _golden(path)

    def make_env(ep):
        t = _round3(data[ep + 1])
        s = np.concatenate([t if t.shape[1] == 4 else np.concatenate([t, np.ones((len(t), 1))], 1), [[100, 100, 100, 1.0]] * 2])
        env = OracleContinuous(setting, stream=s)
        return env, env.reset()

    rec = sequential_eval(make_env, len(ratio))
    assert [r[1] for r in rec] == counter.tolist()
    assert [r[0] for r in rec] == ratio.tolist()
    assert [r[2] for r in rec] == packed
