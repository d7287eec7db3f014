"""CPU: the HOST logic of the product (pct_b200.evaluation.evaluate_batched, pct_b200.heuristics.run_heuristic, the single-env facades'
dataset handling) on an oracle-backed stand-in for PctBatch (tests/fake_batch.py), against the records of the unmodified reference.
What is exercised here is Python only — stream layout, per-env quotas, episode bookkeeping, the 3-decimal rounding of continuous datasets,
packed-list extraction — ; the kernels behind the real PctBatch are checked by the `-m gpu` tests."""
import glob
import importlib
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
from fake_batch import FakeBatch  # noqa: E402
from harness import CONT_ITEM_SET, ITEM_SET, eval_policy_torch, sequential_eval  # noqa: E402

G = os.path.join(os.path.dirname(__file__), "golden")


def _eval_golden(path):
    g = np.load(path)
    off = np.concatenate([[0], np.cumsum(g["packed_len"])])
    packed = [g["packed_flat"][off[i]:off[i + 1]].tolist() for i in range(len(g["ratio"]))]
    return int(g["setting"]), g["data"], g["ratio"], g["counter"], packed


def _heur_golden(path, name, key):
    g = np.load(path)
    off = np.concatenate([[0], np.cumsum(g["len_" + name])])
    return int(g["setting"]), g[key], [g["flat_" + name][off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]


@pytest.fixture
def fake(monkeypatch):
    for mod in ("pct_b200.evaluation", "pct_b200.heuristics", "pct_b200.envs"):
        monkeypatch.setattr(importlib.import_module(mod), "PctBatch", FakeBatch)
    return FakeBatch


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(G, "eval_s*.npz"))))
@pytest.mark.parametrize("n_envs", [1, 7])
def test_evaluate_batched_host_logic_discrete(fake, path, n_envs, tmp_path):
    from pct_b200.evaluation import evaluate_batched
    setting, data, ratio, counter, packed = _eval_golden(path)
    out = evaluate_batched(list(data), setting, policy=eval_policy_torch, item_set=ITEM_SET, n_envs=n_envs, out_dir=str(tmp_path))
    assert out["length"].tolist() == counter.tolist() and out["packed"] == packed
    assert np.allclose(out["ratio"], ratio, rtol=0, atol=1e-15)
    saved = np.load(os.path.join(str(tmp_path), "trajs.npy"), allow_pickle=True)
    assert [list(map(list, ep)) for ep in saved] == packed


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(G, "eval_cont_s*.npz"))))
@pytest.mark.parametrize("n_envs", [1, 5])
def test_evaluate_batched_host_logic_continuous(fake, path, n_envs):
    from pct_b200.evaluation import evaluate_batched
    setting, data, ratio, counter, packed = _eval_golden(path)
    out = evaluate_batched(list(data), setting, policy=eval_policy_torch, container_size=(1.0, 1.0, 1.0), continuous=True, sample_left_bound=0.1,
                           n_envs=n_envs)
    assert out["length"].tolist() == counter.tolist()
    assert out["packed"] == packed
    assert out["ratio"].tolist() == ratio.tolist()


@pytest.mark.parametrize("name", ["LSAH", "OnlineBPH", "BR", "DBL"])
def test_run_heuristic_host_logic_discrete(fake, name):
    from pct_b200.heuristics import run_heuristic
    setting, data, packed = _heur_golden(os.path.join(G, "heur_s1.npz"), name, "data")
    (mean, var, length), rec = run_heuristic(name, setting, len(packed), item_set=ITEM_SET, data=list(data), n_envs=3, return_episodes=True)
    assert rec["packed"] == packed
    ratios = [sum(p[0] * p[1] * p[2] for p in ep) / 1000.0 for ep in packed]
    assert abs(mean - np.mean(ratios)) < 1e-12 and abs(var - np.var(ratios)) < 1e-12 and length == np.mean([len(ep) for ep in packed])


@pytest.mark.parametrize("name", ["LSAH", "OnlineBPH", "BR"])
@pytest.mark.parametrize("setting", [1, 2])
def test_run_heuristic_host_logic_continuous(fake, name, setting):
    from pct_b200.heuristics import run_heuristic
    _, stream, packed = _heur_golden(os.path.join(G, "heur_cont_s%d.npz" % setting), name, "stream")
    (mean, var, length), rec = run_heuristic(name, setting, len(packed), container_size=(1.0, 1.0, 1.0), item_set=CONT_ITEM_SET, continuous=True,
                                             item_stream=stream[None], n_envs=1, return_episodes=True)
    assert rec["packed"] == packed
    with pytest.raises(ValueError):
        run_heuristic("DBL", setting, 1, continuous=True)


def test_facade_dataset_handling_discrete(fake, tmp_path):
    import pct_b200
    setting, data, ratio, counter, packed = _eval_golden(os.path.join(G, "eval_s1.npz"))
    ds = os.path.join(str(tmp_path), "set.pt")
    torch.save([t.tolist() for t in data], ds)
    env = pct_b200.PackingDiscrete(setting=setting, container_size=[10, 10, 10], item_set=ITEM_SET, data_name=ds, load_test_data=True)
    rec = sequential_eval(lambda ep: (env, env.reset()), 6)
    assert [r[1] for r in rec] == counter[:6].tolist() and [r[2] for r in rec] == packed[:6]
    assert np.allclose([r[0] for r in rec], ratio[:6], rtol=0, atol=1e-15)


def test_facade_dataset_handling_continuous(fake, tmp_path):
    import pct_b200
    setting, data, ratio, counter, packed = _eval_golden(os.path.join(G, "eval_cont_s1.npz"))
    ds = os.path.join(str(tmp_path), "set.pt")
    torch.save([t.tolist() for t in data], ds)
    env = pct_b200.PackingContinuous(setting=setting, container_size=[1, 1, 1], item_set=None, data_name=ds, load_test_data=True,
                                     sample_from_distribution=True, sample_left_bound=0.1, sample_right_bound=0.5)
    rec = sequential_eval(lambda ep: (env, env.reset()), 5)
    assert [r[1] for r in rec] == counter[:5].tolist() and [r[2] for r in rec] == packed[:5]
    assert np.allclose([r[0] for r in rec], ratio[:5], rtol=0, atol=1e-12)
