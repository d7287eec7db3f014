"""shuffle=True (D:bin3D.py:114-115, C:bin3D.py:126-127; the reference's training default, tools.py:136).  The reference shuffles the candidate
list with the global numpy RNG — no parity definition —, so the product DEFINES the permutation (include/pct_b200.h, pct_config::shuffle): stable
argsort of the counter-based keys rnd_u64(seed ^ 0x5AFE5EED, global env id, draws << 16 | i).  CPU part: the oracle implements that definition
(checked here against a Python restatement of it), the shuffled list is a permutation of the unshuffled one, and the host surfaces pass the
kwarg through.  GPU part (-m gpu): the kernels against the oracle under the same definition, lock-step."""
import importlib

import numpy as np
import pytest

from harness import ITEM_SET, OracleVec, make_stream
from pct_oracle import OracleContinuous, OracleDiscrete, make_continuous_stream, policy_pick, rnd_u64

SALT = 0x5AFE5EED


def _expected_order(n, seed, gid, draws):
    keys = [rnd_u64(seed ^ SALT, gid, (draws << 16) | i) for i in range(n)]
    return sorted(range(n), key=lambda i: (keys[i], i))


@pytest.mark.parametrize("setting", [1, 2])
def test_oracle_shuffle_is_the_keyed_permutation_of_the_ordered_list(setting):
    seed, gid = 1234, 5
    stream = make_stream(seed, gid, 300, setting)
    a, b = OracleDiscrete(setting, stream=stream), OracleDiscrete(setting, stream=stream)
    b.set_shuffle(seed, gid)
    oa, ob = a.reset(), b.reset()
    moved = 0
    for t in range(60):
        ca, _ = a.candidates()
        cb, _ = b.candidates()
        assert len(ca) == len(cb)
        order = _expected_order(len(ca), seed, gid, b.stream_pos)
        assert np.array_equal(cb, ca[order]), "step %d" % t
        moved += int(order != list(range(len(ca))))
        # drive both with the SAME placement (a leaf of the unshuffled env that is also feasible in the shuffled one: same state, same item)
        _, row = policy_pick(oa, 80, 50, 99, gid, t)
        oa, _, da, _ = a.step(row)
        ob, _, db, _ = b.step(row)
        assert da == db
        if da:
            oa, ob = a.reset(), b.reset()
    assert moved > 40


def test_oracle_shuffle_depends_on_the_global_env_id_only():
    stream = make_stream(7, 0, 100, 2)
    envs = [OracleDiscrete(2, stream=stream) for _ in range(3)]
    envs[0].set_shuffle(7, 10); envs[1].set_shuffle(7, 10); envs[2].set_shuffle(7, 11)
    obs = [e.reset() for e in envs]
    for t in range(3):
        c = [e.candidates()[0] for e in envs]
        assert np.array_equal(c[0], c[1])
        if len(c[0]) > 3:
            assert not np.array_equal(c[0], c[2])
        rows = [policy_pick(o, 80, 50, 1, 0, t)[1] for o in obs[:1]] * 3
        obs = [e.step(rows[0])[0] for e in envs]


def test_continuous_oracle_shuffle_is_a_permutation():
    stream = make_continuous_stream(3, 1, 200, 1)
    a, b = OracleContinuous(1, stream=stream), OracleContinuous(1, stream=stream)
    b.set_shuffle(3, 1)
    a.reset(); b.reset()
    ca, cb = a.candidates()[0], b.candidates()[0]
    assert len(ca) == len(cb) and sorted(map(tuple, ca)) == sorted(map(tuple, cb))


def test_make_vec_envs_honours_args_shuffle(monkeypatch):
    """envs.make_vec_envs forwards args.shuffle (reference default True, tools.py:136) instead of silently forcing False"""
    from fake_batch import FakeBatch
    seen = {}

    class Rec(FakeBatch):
        def __init__(self, *a, **kw):
            seen.update(kw)
            super().__init__(*a, **kw)
    monkeypatch.setattr(importlib.import_module("pct_b200.vec_env"), "PctBatch", Rec)
    import pct_b200

    class Args(object):
        num_processes, setting, container_size, item_size_set = 2, 2, (10, 10, 10), ITEM_SET
        internal_node_holder, leaf_node_holder, seed, shuffle = 80, 50, 4, True
    venv = pct_b200.make_vec_envs(Args())
    assert seen["shuffle"] is True
    o = venv.reset()
    assert tuple(o.shape) == (2, 131 * 9)
    Args.shuffle = False
    pct_b200.make_vec_envs(Args())
    assert seen["shuffle"] is False


# ---------------------------------------------------------------- GPU ----------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("setting", [2, 1, 3])
def test_gpu_shuffle_lockstep_discrete(setting):
    torch = pytest.importorskip("torch")
    import pct_b200
    n, steps, seed, base = 24, 100, 1234, 300
    streams = np.stack([make_stream(seed, base + e, 400, setting) for e in range(n)])
    orc = OracleVec(n, setting, streams)
    for i, e in enumerate(orc.envs):
        e.set_shuffle(seed, base + i)
    gpu = pct_b200.PctBatch(n, setting, item_set=ITEM_SET, obs_dtype=torch.float64, item_stream=streams, seed=seed, env_id_base=base, shuffle=True)
    o_ref, o = orc.reset(), gpu.reset().cpu().numpy()
    for t in range(steps):
        assert np.array_equal(o_ref, o), "setting %d step %d envs %s" % (setting, t, np.unique(np.argwhere(o_ref != o)[:, 0])[:8])
        idx, rows = orc.pick(o_ref, seed, t)
        o_ref, _, d_ref, _ = orc.step(rows)
        ob, _, d, info = gpu.step(leaf_idx=torch.from_numpy(idx).cuda())
        o = ob.cpu().numpy()
        assert np.array_equal(d.cpu().numpy().astype(bool), d_ref)
        assert not gpu.decode_info(info)["flags"].any()
    gpu.close()


@pytest.mark.gpu
def test_gpu_shuffle_off_and_on_differ_only_in_leaf_order():
    """same state, same item: the shuffled observation holds the same internal rows / item row, and — when the cap does not bind — the same SET of leaves"""
    torch = pytest.importorskip("torch")
    import pct_b200
    n, seed = 64, 9
    a = pct_b200.PctBatch(n, 1, item_set=ITEM_SET, obs_dtype=torch.float64, seed=seed)
    b = pct_b200.PctBatch(n, 1, item_set=ITEM_SET, obs_dtype=torch.float64, seed=seed, shuffle=True)
    oa, ob = a.reset().cpu().numpy().reshape(n, 131, 9), b.reset().cpu().numpy().reshape(n, 131, 9)
    assert np.array_equal(oa[:, :80], ob[:, :80]) and np.array_equal(oa[:, 130], ob[:, 130])
    differ = 0
    for e in range(n):
        la, lb = oa[e, 80:130], ob[e, 80:130]
        assert sorted(map(tuple, la)) == sorted(map(tuple, lb))
        differ += int(not np.array_equal(la, lb))
    assert differ > n // 2
    a.close(); b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("setting", [1, 2])
def test_gpu_shuffle_lockstep_continuous(setting):
    torch = pytest.importorskip("torch")
    import pct_b200
    n, steps, seed, base = 12, 80, 21, 40
    streams = np.stack([make_continuous_stream(seed, base + e, 300, setting) for e in range(n)])
    envs = [OracleContinuous(setting, stream=streams[i]) for i in range(n)]
    for i, e in enumerate(envs):
        e.set_shuffle(seed, base + i)
    gpu = pct_b200.PctBatch(n, setting, container_size=(1.0, 1.0, 1.0), continuous=True, obs_dtype=torch.float64, item_stream=streams,
                            size_minimum=0.1, seed=seed, env_id_base=base, shuffle=True)
    o_ref, o = np.stack([e.reset() for e in envs]), gpu.reset().cpu().numpy()
    for t in range(steps):
        assert np.array_equal(o_ref, o), "continuous setting %d step %d" % (setting, t)
        picks = [policy_pick(o_ref[e], 80, 50, seed, e, t) for e in range(n)]
        idx = np.array([p[0] for p in picks], dtype=np.int32)
        nxt = []
        for e, p in zip(envs, picks):
            ob_, _, d_, _ = e.step(p[1])
            nxt.append(e.reset() if d_ else ob_)
        o_ref = np.stack(nxt)
        ob, _, _, info = gpu.step(leaf_idx=torch.from_numpy(idx).cuda())
        o = ob.cpu().numpy()
        assert not gpu.decode_info(info)["flags"].any()
    gpu.close()
