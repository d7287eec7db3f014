"""TEST INFRASTRUCTURE — a stand-in for pct_b200.PctBatch backed by the CPU oracle, for the `-m "not gpu"` tests of the HOST logic
(stream layout, quotas, episode bookkeeping, rounding, packed-list extraction in pct_b200.evaluation / pct_b200.heuristics / the
single-env facades).  It has PctBatch's call surface and return types (torch tensors, raw pct_step_info records) but every environment
computation is the oracle's.  Tests monkeypatch it in place of PctBatch; the product never imports it and has no CPU path."""
import numpy as np
import torch

import pct_oracle_heuristics as OH
from pct_oracle import OracleContinuous, OracleDiscrete


class FakeBatch(object):
    def __init__(self, n_envs, setting, container_size=(10, 10, 10), item_set=None, internal_node_holder=80, leaf_node_holder=50,
                 continuous=False, obs_dtype=torch.float32, seed=0, env_id_base=0, device=0, sample_from_distribution=False,
                 sample_left_bound=None, sample_right_bound=None, item_stream=None, size_minimum=None, auto_reset=True, LNES="EMS", shuffle=False):
        self.n_envs, self.setting, self.continuous = int(n_envs), int(setting), bool(continuous)
        self.nb, self.nl = int(internal_node_holder), int(leaf_node_holder)
        self.obs_len = (self.nb + self.nl + 1) * 9
        self.obs_dtype, self.device, self.auto_reset = obs_dtype, torch.device("cpu"), auto_reset
        self.container_size, self.item_set = tuple(container_size), item_set
        if size_minimum is None:
            size_minimum = sample_left_bound if (continuous and sample_from_distribution) else (float(np.min(np.array(item_set))) if item_set is not None else 1.0)
        if item_stream is not None:
            stream = np.asarray(item_stream, dtype=np.float64)
            assert stream.shape[0] == self.n_envs
            self._streams = [np.ascontiguousarray(s) for s in stream]
        else:  # the oracle's restatement of the device item generators, keyed by the global env index
            self._streams = [None] * self.n_envs
            if continuous and sample_from_distribution:
                if sample_left_bound is None:  # PctBatch / tools.get_args :178-181
                    sample_left_bound, sample_right_bound = 0.1 * min(container_size), 0.5 * min(container_size)
                    if size_minimum is None or size_minimum == 1.0:
                        size_minimum = sample_left_bound
        self.size_minimum = size_minimum
        if continuous:
            self.envs = [OracleContinuous(setting, container_size=container_size, internal_node_holder=self.nb, leaf_node_holder=self.nl,
                                          size_minimum=size_minimum, stream=s) for s in self._streams]
        else:
            self.envs = [OracleDiscrete(setting, container_size=container_size, internal_node_holder=self.nb, leaf_node_holder=self.nl,
                                        size_minimum=size_minimum, stream=s, lnes=LNES) for s in self._streams]
        if shuffle:
            for i, e in enumerate(self.envs):
                e.set_shuffle(seed, env_id_base + i)
        if item_stream is None:
            for i, e in enumerate(self.envs):
                if continuous and sample_from_distribution:
                    e.set_random_sample(seed, env_id_base + i, sample_left_bound, sample_right_bound)
                elif continuous:
                    raise ValueError("FakeBatch: continuous item_set mode needs an explicit stream")
                else:
                    e.set_random_items(item_set, seed, env_id_base + i)
        self._obs64 = np.zeros((self.n_envs, self.obs_len))
        self._hstate = [OH.fresh_state(self.envs[0].container) for _ in range(self.n_envs)]
        self._ep = np.zeros((self.n_envs, 2))  # reward sum, length of the running episode

    def set_trajectory_length(self, n):
        for e in self.envs:
            e.set_trajectory_length(n)

    def _out(self):
        return torch.from_numpy(self._obs64.astype(np.float64 if self.obs_dtype == torch.float64 else np.float32))

    def reset(self, out=None):
        for i, e in enumerate(self.envs):
            self._obs64[i] = e.reset()
            self._hstate[i] = OH.fresh_state(e.container)
        self._ep[:] = 0
        return self._out()

    def step(self, actions=None, leaf_idx=None, out=None):
        rew = np.zeros(self.n_envs, dtype=np.float32)
        done = np.zeros(self.n_envs, dtype=np.uint8)
        info = np.zeros((self.n_envs, 8), dtype=np.int32)
        for i, e in enumerate(self.envs):
            if leaf_idx is not None:
                k = int(leaf_idx[i])
                leaf = self._obs64[i].reshape(-1, 9)[self.nb:self.nb + self.nl]
                row = leaf[k].copy() if 0 <= k < int((leaf[:, 8] == 1).sum()) else np.zeros(9)
            else:
                row = np.asarray(actions[i], dtype=np.float64)
            o, r, d, inf = e.step(row)
            rew[i], done[i], info[i, 0] = r, d, inf["counter"]
            self._ep[i] += (r, 1)
            if d:
                info[i, 2:4] = np.array([inf["ratio"], self._ep[i, 0]], dtype=np.float32).view(np.int32)
                info[i, 4] = int(self._ep[i, 1])
                self._ep[i] = 0
                if self.auto_reset:
                    o = e.reset()
                    self._hstate[i] = OH.fresh_state(e.container)
            info[i, 5] = int((o.reshape(-1, 9)[self.nb:self.nb + self.nl, 8] == 1).sum())
            self._obs64[i] = o
        return self._out(), torch.from_numpy(rew), torch.from_numpy(done), torch.from_numpy(info)

    def heuristic_actions(self, name, seed=0, t=0, out=None):
        rows = np.zeros((self.n_envs, 9))
        for i, e in enumerate(self.envs):
            c = OH.choose(name, e, self._hstate[i], self.item_set, seed, i, t)
            rows[i] = OH.action_row(c, e.container, self.continuous)
            if c is not None:
                OH.note_placement(self._hstate[i], c)
        return torch.from_numpy(rows if self.continuous else rows.astype(np.float32))

    def query_placement(self, env, dims, lx, ly, density=1.0, want_map=False):
        """Space.drop_box_virtual(returnH / returnMap) for one env, like PctBatch.query_placement"""
        e = self.envs[env]
        if self.continuous:
            ok, mh = e.drop_box_virtual(dims, lx, ly)
            return (ok, mh, None) if want_map else (ok, mh)
        ok, mh = e.drop_box_virtual(dims, lx, ly)
        if not want_map:
            return ok, mh
        hm = e.plain().copy()
        hm[int(lx):int(lx) + int(dims[0]), int(ly):int(ly) + int(dims[1])] = mh + int(dims[2])
        return ok, mh, hm

    def state(self, env):
        e = self.envs[env]
        o = self._obs64[env].reshape(-1, 9)
        n = len(e.packed)
        boxes = np.concatenate([o[:n, :6], np.ones((n, 1))], 1) if n else np.zeros((0, 7))
        nb = e.next_box
        return dict(n_boxes=n, boxes=boxes, ems=np.asarray(e.ems(), dtype=np.float64), next_box=list(nb), next_den=e.next_den, flags=0)

    @staticmethod
    def check_flags(flags, ignore=0, what="pct_step"):
        from pct_b200.batch import PctBatch
        return PctBatch.check_flags(flags, ignore=ignore, what=what)

    @staticmethod
    def decode_info(info_cpu):
        a = info_cpu.cpu().numpy() if hasattr(info_cpu, "cpu") else np.asarray(info_cpu)
        f = a.view(np.float32)
        return dict(counter=a[:, 0], flags=a[:, 1], ratio=f[:, 2], ep_reward=f[:, 3], ep_len=a[:, 4], n_leaf=a[:, 5], n_cand=a[:, 6], n_ems=a[:, 7])

    def close(self):
        self.envs = []
