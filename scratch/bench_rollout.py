"""SURVEY 8(f)-1 throughput: rollout collection with the reference's unmodified DRL_GAT policy (baseline/_ref copy, scratch/install_reference.sh).
  (a) GraphedRollout: T x (DRL_GAT forward + sampling -> pct_step) in ONE CUDA graph, storages on the device;
  (b) the reference's collection loop (train_tools.py:63-70) on PctVecEnv: policy on the device, leaf rows through .cpu().numpy(), envs.step(rows),
      observation back, PCTRolloutStorage.insert — one host round trip per step, as the reference does it;
  (c) (b) with leaf INDICES handed over on the device (no numpy hop), still one sync per step for done / infos.
python scratch/bench_rollout.py [--envs 4096] [--T 5] [--rollouts 40]  -> JSON lines (policy+env steps/s)."""
import argparse
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import pct_b200  # noqa: E402
from pct_b200 import compat  # noqa: E402

ITEM_SET = [(i, j, k) for i in range(1, 6) for j in range(1, 6) for k in range(1, 6)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--T", type=int, default=5)  # tools.py --num-steps default
    ap.add_argument("--rollouts", type=int, default=40)
    ap.add_argument("--setting", type=int, default=1)
    a = ap.parse_args()
    ref = next(p for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference") if os.path.isfile(os.path.join(p, "attention_model.py")))
    model, tools = compat.load_policy_modules(ref, strip_asserts=True)
    dev = torch.device("cuda", 0)
    args = types.SimpleNamespace(embedding_size=64, hidden_size=128, gat_layer_num=1, internal_node_holder=80,
                                 internal_node_length=7 if a.setting == 3 else 6, leaf_node_holder=50)
    torch.manual_seed(0)
    net = model.DRL_GAT(args).to(dev)
    n, T, R, factor = a.envs, a.T, a.rollouts, 0.1
    out = {"envs": n, "num_steps": T, "rollouts": R, "setting": a.setting, "policy": "reference DRL_GAT (unmodified), sampling"}
    with torch.no_grad():
        # (a) one CUDA graph per rollout
        roll = pct_b200.GraphedRollout(pct_b200.PctBatch(n, a.setting, item_set=ITEM_SET, seed=1), T, policy=pct_b200.drl_gat_policy(net, tools, 80, 50, factor))
        for _ in range(5):
            roll.run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(R):
            roll.run()
        torch.cuda.synchronize()
        out["graphed_rollout_steps_per_s"] = n * T * R / (time.perf_counter() - t0)
        # policy forward alone (the floor of any rollout with this network)
        obs = roll.obs[0]
        pol = pct_b200.drl_gat_policy(net, tools, 80, 50, factor)
        for _ in range(5):
            pol(obs, None)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(T * R):
            pol(obs, None)
        torch.cuda.synchronize()
        out["policy_forward_only_steps_per_s"] = n * T * R / (time.perf_counter() - t0)
        roll.batch.close()
        # (b) / (c) the reference's loop on the VecEnv surface
        for mode in ("rows_numpy", "idx_device"):
            venv = pct_b200.PctVecEnv(n, a.setting, item_set=ITEM_SET, seed=1)
            obs = venv.reset()
            batchX = torch.arange(n, device=dev)
            storage_obs = torch.zeros((T + 1, n, 131, 9), device=dev)

            def one_rollout(obs):
                for step in range(T):
                    all_nodes, leaf_nodes = tools.get_leaf_nodes(obs, 80, 50)
                    logp, idx, _, _ = net(all_nodes, normFactor=factor)
                    if mode == "rows_numpy":
                        rows = leaf_nodes[batchX, idx.squeeze()]
                        obs, reward, done, infos = venv.step(rows.cpu().numpy())
                    else:
                        obs, reward, done, infos = venv.step(idx.squeeze().to(torch.int32))
                    storage_obs[step + 1].copy_(obs.view(n, 131, 9))
                return obs
            for _ in range(3):
                obs = one_rollout(obs)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(R):
                obs = one_rollout(obs)
            torch.cuda.synchronize()
            out["vec_env_loop_%s_steps_per_s" % mode] = n * T * R / (time.perf_counter() - t0)
            venv.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
