"""GPU: SURVEY 8(f)-1 with the REAL policy — the reference's unmodified DRL_GAT network (model.py / attention_model.py / graph_encoder.py, loaded by
pct_b200.compat.load_policy_modules from the git-ignored copy baseline/_ref that scratch/install_reference.sh makes; skipped without it) runs
forward + sampling + leaf-index hand-off inside GraphedRollout: T x (policy -> pct_step) captured in ONE CUDA graph, every tensor of
PCTRolloutStorage.insert (storage.py:33-39) written on the device.

The check is against ORACLE envs, not against the rollout itself: the actions the network sampled inside the graph are replayed on CPU oracle envs
with the same item streams — every stored observation, reward and done flag must equal the oracle's —, and the stored log-probabilities / values
must be what the same network returns for the stored (observation, action) pairs (model.evaluate_actions, the call of the PPO / ACKTR update)."""
import os
import types

import numpy as np
import pytest

torch = pytest.importorskip("torch")
from harness import ITEM_SET, OracleVec, make_stream  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = next((p for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference") if os.path.isfile(os.path.join(p, "attention_model.py"))), None)


def _net(setting, dev):
    from pct_b200 import compat
    model, tools = compat.load_policy_modules(REF, strip_asserts=True)
    args = types.SimpleNamespace(embedding_size=64, hidden_size=128, gat_layer_num=1, internal_node_holder=80,
                                 internal_node_length=7 if setting == 3 else 6, leaf_node_holder=50)  # tools.py:148-190 defaults
    torch.manual_seed(99 + setting)
    return model.DRL_GAT(args).to(dev), tools


@pytest.mark.skipif(REF is None, reason="no copy of the reference's policy network (run scratch/install_reference.sh)")
@pytest.mark.parametrize("setting,use_graph", [(1, True), (2, True), (1, False)])
def test_drl_gat_rollout_matches_oracle_envs_under_the_same_actions(setting, use_graph):
    import pct_b200
    n, T, R, seed = 48, 5, 8, 31
    dev = torch.device("cuda", 0)
    net, tools = _net(setting, dev)
    factor = 1.0 / 10.0  # normFactor = 1 / max(container_size) (tools.py:170)
    streams = np.stack([make_stream(seed, e, 600, setting) for e in range(n)])
    batch = pct_b200.PctBatch(n, setting, item_set=ITEM_SET, item_stream=streams)
    roll = pct_b200.GraphedRollout(batch, T, policy=pct_b200.drl_gat_policy(net, tools, 80, 50, factor), use_graph=use_graph)
    orc = OracleVec(n, setting, streams)
    o_ref = orc.reset().astype(np.float32)
    with torch.no_grad():
        for r in range(R):
            obs, rew, done, act = roll.run()
            torch.cuda.synchronize()
            obs_h, rew_h, done_h, act_h = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy().astype(bool), act.cpu().numpy()
            for t in range(T):
                assert np.array_equal(obs_h[t], o_ref), "rollout %d step %d: observation" % (r, t)
                leaf = o_ref.reshape(n, 131, 9)[:, 80:130]
                has_leaf = leaf[:, :, 8].sum(1) > 0  # an env without any feasible leaf: every probability is the 1e-20 floor, any (all-zero) row may come
                assert (leaf[np.arange(n), act_h[t], 8] == 1)[has_leaf].all(), "the network selected a masked leaf"
                rows = leaf[np.arange(n), act_h[t]].astype(np.float64)
                o64, r_ref, d_ref, _ = orc.step(rows)
                o_ref = o64.astype(np.float32)
                assert np.array_equal(rew_h[t], r_ref.astype(np.float32)) and np.array_equal(done_h[t], d_ref), "rollout %d step %d" % (r, t)
            assert np.array_equal(obs_h[T], o_ref)
            # the stored log-probs / values are the network's for the stored (observation, action) pairs
            flat_obs = obs[:-1].reshape(T * n, -1)
            all_nodes, _ = tools.get_leaf_nodes(flat_obs, 80, 50)
            values, logp, _ = net.evaluate_actions(all_nodes, act.reshape(T * n, 1).long(), normFactor=factor)
            assert torch.allclose(logp.reshape(T, n, 1), roll.action_log_probs, atol=1e-5, rtol=1e-4)
            assert torch.allclose(values.reshape(T, n, 1), roll.value_preds, atol=1e-5, rtol=1e-4)
            assert torch.equal(roll.masks[1:, :, 0], 1 - done.float())
    if use_graph:
        assert roll.graph is not None
    batch.close()
