"""GPU parity proper: the CUDA path (through the C ABI) against the CPU oracle, bit-exact, discrete domain."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from harness import ITEM_SET, OracleVec, make_stream  # noqa: E402

pytestmark = pytest.mark.gpu


def _run(setting, n_envs, steps, action_mode, seed=1234, obs_dtype=None):
    import pct_b200
    obs_dtype = obs_dtype or torch.float64
    streams = np.stack([make_stream(seed, e, 400, setting) for e in range(n_envs)])
    orc = OracleVec(n_envs, setting, streams)
    gpu = pct_b200.PctBatch(n_envs, setting, item_set=ITEM_SET, obs_dtype=obs_dtype, item_stream=streams)
    o_ref = orc.reset()
    o_gpu = gpu.reset().cpu().numpy()
    for t in range(steps):
        ref = o_ref if obs_dtype == torch.float64 else o_ref.astype(np.float32)
        if not np.array_equal(ref, o_gpu):
            bad = np.argwhere(ref != o_gpu)
            e = bad[0][0]
            raise AssertionError("obs mismatch setting %d step %d env %d at %s: ref %s gpu %s\nstate %s" % (
                setting, t, e, bad[:8].tolist(), ref[e][bad[:8, 1]], o_gpu[e][bad[:8, 1]], gpu.state(int(e))))
        idx, rows = orc.pick(o_ref, seed, t)
        o_ref, r_ref, d_ref, i_ref = orc.step(rows)
        if action_mode == "idx":
            o, r, d, info = gpu.step(leaf_idx=torch.from_numpy(idx).cuda())
        elif action_mode == "policy":
            o, r, d, info = gpu.step(leaf_idx=gpu.random_policy(seed, t))
        else:
            o, r, d, info = gpu.step(actions=torch.from_numpy(rows.astype(np.float32)).cuda())
        o_gpu = o.cpu().numpy()
        assert np.array_equal(r.cpu().numpy(), r_ref.astype(np.float32)), "reward mismatch step %d" % t
        assert np.array_equal(d.cpu().numpy().astype(bool), d_ref), "done mismatch step %d" % t
        inf = gpu.decode_info(info)
        assert not inf["flags"].any(), "overflow flags %s" % inf["flags"][inf["flags"] != 0]
        for e in range(n_envs):
            assert inf["counter"][e] == i_ref[e]["counter"]
            if d_ref[e]:
                assert inf["ratio"][e] == np.float32(i_ref[e]["ratio"])
    return True


@pytest.mark.parametrize("setting", [2, 1, 3])
@pytest.mark.parametrize("action_mode", ["idx", "rows", "policy"])
def test_lockstep_small(setting, action_mode):
    _run(setting, 32, 120, action_mode)


def test_lockstep_f32_obs():
    _run(1, 16, 60, "idx", obs_dtype=torch.float32)


@pytest.mark.parametrize("setting", [1, 2])
def test_lockstep_many_envs(setting):
    _run(setting, 300, 80, "policy", seed=77)
