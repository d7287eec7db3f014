"""GPU parity, harder cases: other container / holder sizes (32-bit candidate keys), the 2048-slot set table,
many supports under one box (per-env scratch + lstsq), full-size batches against the threaded oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from harness import ITEM_SET, OracleVec, make_stream  # noqa: E402
from pct_oracle import OracleBatch, OracleDiscrete, rnd_u64  # noqa: E402

pytestmark = pytest.mark.gpu


def _lockstep(setting, n, steps, container, items, nb=80, nl=50, seed=3):
    import pct_b200
    streams = np.zeros((n, 300, 4))
    for e in range(n):
        for d in range(300):
            streams[e, d, :3] = items[rnd_u64(seed, e, d) % len(items)]
            streams[e, d, 3] = max(rnd_u64(seed ^ 0xABCDEF, e, d) >> 11, 1) / float(1 << 53) if setting == 3 else 1.0
    orc = OracleVec(n, setting, streams, nb=nb, nl=nl, container=container)
    gpu = pct_b200.PctBatch(n, setting, container_size=container, item_set=items, internal_node_holder=nb, leaf_node_holder=nl,
                            obs_dtype=torch.float64, item_stream=streams)
    o_ref, o = orc.reset(), gpu.reset().cpu().numpy()
    max_cand = 0
    for t in range(steps):
        assert np.array_equal(o_ref, o), "setting %d step %d envs %s" % (setting, t, np.unique(np.argwhere(o_ref != o)[:, 0])[:5])
        idx, rows = orc.pick(o_ref, seed, t)
        o_ref, r_ref, d_ref, _ = orc.step(rows)
        ob, r, d, info = gpu.step(leaf_idx=torch.from_numpy(idx).cuda())
        o = ob.cpu().numpy()
        inf = gpu.decode_info(info)
        assert not inf["flags"].any()
        max_cand = max(max_cand, int(inf["n_cand"].max()))
        assert np.array_equal(d.cpu().numpy().astype(bool), d_ref)
    return max_cand


@pytest.mark.parametrize("setting", [1, 2])
def test_large_container_32bit_keys(setting):
    items = [(i, j, k) for i in (4, 6, 9) for j in (5, 8) for k in (3, 7, 10)]  # keeps the EMS list under the 128-entry capacity
    _lockstep(setting, 12, 90, (20, 18, 24), items)


def test_other_holder_sizes():
    _lockstep(1, 16, 80, (10, 10, 10), ITEM_SET, nb=60, nl=30)
    _lockstep(2, 16, 60, (10, 10, 10), ITEM_SET, nb=80, nl=64)


def test_big_candidate_sets_use_the_2048_slot_stage():
    # 16^3 bin, small items: hundreds of candidates per step, 16-bit keys; setting 1 exercises the HBM table stage
    items = [(2, 2, 2), (3, 3, 2), (2, 3, 3), (4, 2, 2), (3, 4, 3)]  # EMS stays <= 128, > 306 distinct candidates occur
    m2 = _lockstep(2, 6, 70, (16, 16, 16), items, nb=80, nl=50)
    m1 = _lockstep(1, 6, 70, (16, 16, 16), items, nb=80, nl=50)
    assert m2 > 306 and m1 > 306, (m1, m2)  # > 306 distinct candidates forces the 2048-slot table


def test_many_supports_scratch_and_lstsq_path():
    """16 unit boxes in a 4x4 grid, then a 4x4 box whose centre of mass sits on the common corner of the four
    middle boxes: 16 supports, no direct edge -> the >8-support scratch path and the lstsq split, real and virtual."""
    import pct_b200
    seq = [[1, 1, 1, 1.0]] * 16 + [[4, 4, 1, 1.0]] + [[2, 2, 1, 1.0]] * 6
    stream = np.array([seq], dtype=np.float64)
    for setting in (1,):
        env = pct_b200.PackingDiscrete(setting=setting, container_size=[10, 10, 10], item_set=ITEM_SET, item_stream=stream)
        orc = OracleDiscrete(setting, stream=stream[0])
        o, o_ref = env.reset(), orc.reset()
        acts = [(0, x, y) for x in range(4) for y in range(4)] + [(0, 0, 0)]
        for a in acts:
            assert np.array_equal(o, o_ref)
            o, r, d, info = env.step(a)
            o_ref, r_ref, d_ref, _ = orc.step(np.array(a, dtype=np.float64))
            assert (r, d) == (r_ref, d_ref) and "flags" not in info, (a, info)
        assert np.array_equal(o, o_ref) and orc.n_lstsq > 0
        assert env.packed == orc.packed and len(env.packed) == 17
        env.close()


@pytest.mark.parametrize("setting,n", [(1, 4096), (2, 8192)])
def test_full_size_final_observations(setting, n):
    """BASELINE.json sizes: after K steps of the synthetic policy every one of the n final observations, the reward
    sums and the episode counts must equal the threaded CPU oracle's (trajectories are chaotic: equality of the final
    state certifies every intermediate step)."""
    import pct_b200
    steps, iseed, pseed = 120, 1234, 4321
    gpu = pct_b200.PctBatch(n, setting, item_set=ITEM_SET, seed=iseed, obs_dtype=torch.float64)
    cpu = OracleBatch(n, setting, ITEM_SET, iseed, pseed)
    gpu.reset()
    rsum = torch.zeros(n, dtype=torch.float64, device="cuda")
    nd = torch.zeros(n, dtype=torch.int64, device="cuda")
    flags = torch.zeros(n, dtype=torch.int32, device="cuda")
    for t in range(steps):
        obs, r, d, info = gpu.step(leaf_idx=gpu.random_policy(pseed, t))
        rsum += r.double(); nd += d.long(); flags |= info[:, 1]
    cpu.run(steps)
    o_ref, r_ref, nd_ref = cpu.get()
    assert int(flags.max()) == 0
    assert np.array_equal(nd.cpu().numpy(), nd_ref)
    assert np.array_equal(obs.cpu().numpy(), o_ref)
    assert np.allclose(rsum.cpu().numpy(), r_ref, rtol=0, atol=1e-4)  # GPU rewards are float32 (VecPyTorch contract)


@pytest.mark.parametrize("lnes", ["FC", "EV", "CP", "EP"])
@pytest.mark.parametrize("setting", [1, 2])
def test_other_leaf_expansion_schemes(lnes, setting):
    """EV (as degenerate as in the reference) / EP / CP / FC candidate generators (D:space.py:573-806) against the oracle."""
    import pct_b200
    from pct_oracle import policy_pick
    n, steps, seed = 16, 70, 13
    streams = np.stack([make_stream(seed, e, 300, setting) for e in range(n)])
    orcs = [OracleDiscrete(setting, stream=streams[e], lnes=lnes) for e in range(n)]
    gpu = pct_b200.PctBatch(n, setting, item_set=ITEM_SET, obs_dtype=torch.float64, item_stream=streams, LNES=lnes)
    o_ref = np.stack([o.reset() for o in orcs])
    o = gpu.reset().cpu().numpy()
    for t in range(steps):
        assert np.array_equal(o_ref, o), "%s setting %d step %d envs %s" % (lnes, setting, t, np.unique(np.argwhere(o_ref != o)[:, 0])[:5])
        picks = [policy_pick(o_ref[e], 80, 50, seed, e, t) for e in range(n)]
        rows = np.stack([p[1] for p in picks])
        nxt = []
        for e in range(n):
            ob, r, d, info = orcs[e].step(rows[e])
            nxt.append(orcs[e].reset() if d else ob)
        o_ref = np.stack(nxt)
        ob, r, d, info = gpu.step(actions=torch.from_numpy(rows).cuda())
        o = ob.cpu().numpy()
        assert not gpu.decode_info(info)["flags"].any()
