"""GPU: device-resident rollout — a CUDA graph of T x (policy -> step) must reproduce eager stepping exactly."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from harness import ITEM_SET  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("setting", [1, 2])
def test_graphed_rollout_equals_eager(setting):
    import pct_b200
    n, T, R = 512, 5, 6
    a = pct_b200.GraphedRollout(pct_b200.PctBatch(n, setting, item_set=ITEM_SET, seed=77), T, policy_seed=5, use_graph=False)
    b = pct_b200.GraphedRollout(pct_b200.PctBatch(n, setting, item_set=ITEM_SET, seed=77), T, policy_seed=5, use_graph=True)
    for r in range(R):
        oa, ra, da, aa = a.run()
        ob, rb, db, ab = b.run()
        torch.cuda.synchronize()
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db) and torch.equal(aa, ab), "rollout %d" % r
    assert b.graph is not None and int(b.t_dev) == R * T
    assert float(db.float().mean()) >= 0.0 and oa.shape == (T + 1, n, 1179)


def test_rollout_with_torch_policy():
    """A policy written in torch ops (argmax over a score of the leaf rows) runs inside the captured graph."""
    import pct_b200

    def policy(obs, t_dev):
        leaf = obs.view(obs.shape[0], 131, 9)[:, 80:130]
        score = leaf[:, :, 8] * (1000.0 - leaf[:, :, 2] * 100.0 - leaf[:, :, 0] - leaf[:, :, 1] * 10.0)  # deepest-bottom-left
        return score.argmax(dim=1)

    n, T = 256, 4
    roll = pct_b200.GraphedRollout(pct_b200.PctBatch(n, 2, item_set=ITEM_SET, seed=3), T, policy=policy)
    tot = 0.0
    for _ in range(10):
        obs, rew, done, act = roll.run()
        tot += float(rew.sum())
    assert tot > 0 and int(done.sum()) >= 0
    # greedy DBL packs tighter than random: every env places at least one box per rollout on average
    assert tot / (10 * T * n) > 0.2
