"""GPU: the opt-in fork-join continuation kernel (PCT_B200_WALK=fork: pieces of stability walks from a device-side queue, csrc/pct_walkq.cuh, verdict =
AND over a walk's pieces) must give the observations of the default sequential continuation kernel — and so the oracle's — bit for bit, in both
domains, at sizes where the queue's forked region, the helper warps' tickets and the per-step counter reset are all exercised.

The host build of the same source (stab_piece) is pinned against the oracle by tests/test_host_emul_stability.py (routine "fork_join").
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from harness import ITEM_SET  # noqa: E402

pytestmark = pytest.mark.gpu


def _trace(monkeypatch, mode, continuous, setting, n, steps, keep=None):
    import pct_b200
    monkeypatch.setenv("PCT_B200_WALK", mode)
    if keep is not None:
        monkeypatch.setenv("PCT_B200_WALK_KEEP", str(keep))
    if continuous:
        b = pct_b200.PctBatch(n, setting, container_size=(1.0, 1.0, 1.0), continuous=True, sample_from_distribution=True, seed=5)
    else:
        b = pct_b200.PctBatch(n, setting, item_set=ITEM_SET, seed=5)
    out = [b.reset().clone()]
    rew = []
    for t in range(steps):
        obs, r, d, info = b.step(leaf_idx=b.random_policy(9, t))
        out.append(obs.clone())
        rew.append(r.clone())
    assert not b.decode_info(info)["flags"].any()
    res = torch.stack(out).cpu().numpy(), torch.stack(rew).cpu().numpy()
    b.close()
    return res


@pytest.mark.parametrize("continuous,setting,n,steps", [(False, 1, 1500, 120), (False, 3, 600, 100), (True, 1, 400, 80), (False, 1, 1, 150), (True, 1, 1, 120)])
def test_fork_join_walks_equal_sequential_walks(monkeypatch, continuous, setting, n, steps):
    ref = _trace(monkeypatch, "seq", continuous, setting, n, steps)
    for keep in (None, 8):  # default helper count; almost no helpers (forked pieces wait for warps that re-claim)
        got = _trace(monkeypatch, "fork", continuous, setting, n, steps, keep)
        assert np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1])


def test_fork_join_against_the_oracle(monkeypatch):
    from test_gpu_discrete_parity import _run
    monkeypatch.setenv("PCT_B200_WALK", "fork")
    _run(1, 64, 150, "policy", seed=31)
    _run(3, 64, 120, "idx", seed=32)
