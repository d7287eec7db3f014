"""Device-resident n-step rollout: policy -> env.step, T times, with no host round trip.

The reference collects rollouts as
    logp, idx, ... = policy(all_nodes); leaf = leaf_nodes[batchX, idx]; obs, r, d, infos = envs.step(leaf.cpu().numpy())
    pct_rollout.insert(...)                                             (train_tools.py:63-70, storage.py:33-39)
i.e. one device->host->device trip per step.  Here the selected leaf INDEX goes straight into pct_step, every observation is
written by the kernel directly into a (T+1, N, obs_len) storage tensor (the layout of PCTRolloutStorage.obs), and the whole
T-step sequence is captured once in a CUDA graph and replayed with a single launch.
"""
import torch


class GraphedRollout(object):
    """policy: callable(obs (N, obs_len) float32 CUDA tensor, t_dev int64 CUDA scalar tensor) -> int32 CUDA tensor (N,) of leaf indices,
    made of capturable torch ops (e.g. a DRL_GAT forward + argmax / multinomial); None = the built-in uniform-valid-leaf kernel."""

    def __init__(self, batch, num_steps, policy=None, policy_seed=0, use_graph=True):
        self.batch, self.T, self.policy, self.seed = batch, int(num_steps), policy, int(policy_seed)
        n, dev = batch.n_envs, batch.device
        self.obs = torch.zeros((self.T + 1, n, batch.obs_len), dtype=batch.obs_dtype, device=dev)
        self.rewards = torch.zeros((self.T, n), dtype=torch.float32, device=dev)
        self.dones = torch.zeros((self.T, n), dtype=torch.uint8, device=dev)
        self.actions = torch.zeros((self.T, n), dtype=torch.int32, device=dev)
        self.t_dev = torch.zeros((), dtype=torch.int64, device=dev)
        self.graph = None
        self.use_graph = use_graph
        self._started = False

    def start(self):
        self.batch.reset(out=self.obs[0])
        self.t_dev.zero_()
        self._started = True

    def _steps(self):
        b = self.batch
        for t in range(self.T):
            if self.policy is None:
                idx = b.random_policy_dev(self.seed, self.t_dev, out=self.actions[t])
            else:
                idx = self.policy(self.obs[t], self.t_dev).to(torch.int32)
                self.actions[t].copy_(idx)
                idx = self.actions[t]
            _, r, d, _ = b.step(leaf_idx=idx, out=self.obs[t + 1])
            self.rewards[t].copy_(r)
            self.dones[t].copy_(d)
            self.t_dev += 1

    def run(self):
        """One T-step rollout continuing from the last observation; returns (obs, rewards, dones, actions) storages."""
        if not self._started:
            self.start()
        else:
            self.obs[0].copy_(self.obs[self.T])  # storage.after_update (storage.py:41-42)
        if not self.use_graph:
            self._steps()
        else:
            if self.graph is None:
                # start() already launched every kernel once (first launches set kernel attributes, which cannot be captured)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._steps()
            self.graph.replay()
        return self.obs, self.rewards, self.dones, self.actions
