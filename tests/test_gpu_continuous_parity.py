"""GPU parity, continuous domain (PctContinuous0): CUDA path vs the CPU oracle on identical item streams.
North-star tolerance is 1e-6 on leaf coordinates; this path is bit-exact in float64 (asserted with equality)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from pct_oracle import OracleContinuous, make_continuous_stream, policy_pick  # noqa: E402

pytestmark = pytest.mark.gpu


class _Vec(object):
    def __init__(self, n, setting, streams):
        self.envs = [OracleContinuous(setting, stream=streams[i]) for i in range(n)]
        self.n = n

    def reset(self):
        return np.stack([e.reset() for e in self.envs])

    def step(self, rows):
        out = []
        for e, a in zip(self.envs, rows):
            o, r, d, i = e.step(a)
            if d:
                o = e.reset()
            out.append((o, r, d, i))
        return np.stack([x[0] for x in out]), np.array([x[1] for x in out]), np.array([x[2] for x in out]), [x[3] for x in out]


def _run(setting, n, steps, mode, seed=21):
    import pct_b200
    streams = np.stack([make_continuous_stream(seed, e, 300, setting) for e in range(n)])
    orc = _Vec(n, setting, streams)
    gpu = pct_b200.PctBatch(n, setting, container_size=(1.0, 1.0, 1.0), continuous=True, obs_dtype=torch.float64, item_stream=streams,
                            size_minimum=0.1)
    o_ref, o = orc.reset(), gpu.reset().cpu().numpy()
    for t in range(steps):
        if not np.array_equal(o_ref, o):
            bad = np.argwhere(o_ref != o)
            raise AssertionError("continuous setting %d step %d: %d mismatches, first %s ref %r gpu %r" % (
                setting, t, len(bad), bad[0], o_ref[tuple(bad[0])], o[tuple(bad[0])]))
        picks = [policy_pick(o_ref[e], 80, 50, seed, e, t) for e in range(n)]
        idx = np.array([p[0] for p in picks], dtype=np.int32)
        rows = np.stack([p[1] for p in picks])
        o_ref, r_ref, d_ref, i_ref = orc.step(rows)
        if mode == "idx":
            ob, r, d, info = gpu.step(leaf_idx=torch.from_numpy(idx).cuda())
        else:
            ob, r, d, info = gpu.step(actions=torch.from_numpy(rows).cuda())  # float64 rows
        o = ob.cpu().numpy()
        assert np.array_equal(d.cpu().numpy().astype(bool), d_ref)
        assert np.array_equal(r.cpu().numpy(), r_ref.astype(np.float32))
        inf = gpu.decode_info(info)
        assert not inf["flags"].any(), inf["flags"]
        assert np.array_equal(inf["counter"], np.array([i["counter"] for i in i_ref]))


@pytest.mark.parametrize("setting", [2, 1, 3])
@pytest.mark.parametrize("mode", ["idx", "rows"])
def test_continuous_lockstep(setting, mode):
    _run(setting, 24, 100, mode)


def test_continuous_random_items_run():
    """sample_from_distribution (C:bin3D.py:103-115) with the device generator: items have 3 decimals in [0.1, 0.5], heights in the 5-set."""
    import pct_b200
    b = pct_b200.PctBatch(256, 1, container_size=(1.0, 1.0, 1.0), continuous=True, sample_from_distribution=True, seed=3)
    obs = b.reset()
    for t in range(40):
        obs, r, d, info = b.step(leaf_idx=b.random_policy(9, t))
    nxt = obs.view(256, 131, 9)[:, 130, 3:6].double().cpu().numpy()
    assert nxt.min() > 0.0999 and nxt.max() < 0.5001
    assert not b.decode_info(info)["flags"].any()
    assert np.allclose(nxt * 1000, np.round(nxt * 1000), atol=1e-3)


@pytest.mark.parametrize("setting", [1, 2])
def test_continuous_single_env_facade(setting):
    """PackingContinuous drop-in: float64 observations incl. the terminal one, reset() continues the stream, `packed`."""
    import pct_b200
    seed = 8
    stream = make_continuous_stream(seed, 0, 300, setting)
    orc = OracleContinuous(setting, stream=stream)
    env = pct_b200.PackingContinuous(setting=setting, container_size=[1, 1, 1], item_set=None, sample_from_distribution=False,
                                     item_stream=stream[None], size_minimum=0.1)
    o_ref, o = orc.reset(), env.reset()
    episodes = 0
    for t in range(90):
        assert np.array_equal(o, o_ref), t
        _, row = policy_pick(o_ref, 80, 50, seed, 0, t)
        o_ref, r_ref, d_ref, i_ref = orc.step(row)
        o, r, d, info = env.step(row)
        assert d == d_ref and abs(r - r_ref) < 1e-12 and info["counter"] == i_ref["counter"]
        if d:
            assert np.array_equal(o, o_ref), "terminal observation"
            assert np.allclose(np.array(env.packed)[:, :6], np.array(orc.packed)[:, :6], rtol=0, atol=1e-12)
            o_ref, o = orc.reset(), env.reset()
            episodes += 1
    assert episodes >= 2
