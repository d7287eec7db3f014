"""PctVecEnv — the reference's vectorised-environment surface on top of the CUDA batch.

Replaces `VecPyTorch(ShmemVecEnv([make_env(...)] * N))` (reference: envs.py:75-116,159-182,
wrapper/shmem_vec_env.py:20-156, wrapper/vec_env.py:29-139): same methods, same return types, same
auto-reset semantics, but observations never leave the GPU and there are no worker processes.

    obs                      torch.float32 (N, 1179) on the device      (VecPyTorch.reset/step_wait, envs.py:166-182)
    reward                   torch.float32 (N, 1) on the CPU             (envs.py:181)
    done                     numpy bool (N,)                             (shmem_vec_env.py:81)
    infos                    tuple of N dicts; finished envs carry {'counter','ratio','reward','episode':{'r','l','t'}}
                             (D:bin3D.py:163-164, wrapper/monitor.py:58-77; consumed at train_tools.py:72-79)
"""
import time

import numpy as np
import torch

from .batch import PctBatch


class _BoxSpace(object):
    """Stand-in for gym.spaces.Box (D:bin3D.py:41-42) when gym is not installed."""

    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)


def _make_box(low, high, shape):
    try:
        import gym  # noqa
        return gym.spaces.Box(low=low, high=high, shape=shape)
    except Exception:
        return _BoxSpace(low, high, shape)


class PctVecEnv(object):
    """N PCT environments stepped together on one GPU; duck-types wrapper.vec_env.VecEnv."""

    closed = False
    viewer = None
    metadata = {"render.modes": []}

    def __init__(self, num_envs, setting, container_size=(10, 10, 10), item_set=None, internal_node_holder=80,
                 leaf_node_holder=50, continuous=False, device=0, seed=0, env_id_base=0, sample_from_distribution=False,
                 sample_left_bound=None, sample_right_bound=None, item_stream=None, LNES="EMS", shuffle=False, **_ignored):
        if shuffle:
            raise NotImplementedError("shuffle=True draws from the global numpy RNG in the reference and has no parity definition")
        self.batch = PctBatch(num_envs, setting, container_size=container_size, item_set=item_set,
                              internal_node_holder=internal_node_holder, leaf_node_holder=leaf_node_holder, continuous=continuous,
                              obs_dtype=torch.float32, seed=seed, env_id_base=env_id_base, device=device,
                              sample_from_distribution=sample_from_distribution, sample_left_bound=sample_left_bound,
                              sample_right_bound=sample_right_bound, item_stream=item_stream, LNES=LNES)
        self.num_envs = int(num_envs)
        self.device = self.batch.device
        self.observation_space = _make_box(0.0, float(container_size[2]), (self.batch.obs_len,))
        self.action_space = None  # the reference env never defines one (read at envs.py:107)
        self._tstart = time.time()
        self._pending = None

    # ---- VecEnv API (wrapper/vec_env.py:48-108) ----
    def reset(self):
        return self.batch.reset()

    def step_async(self, actions):
        """actions: (N, 9) leaf rows (torch tensor or numpy, any float dtype) — what train_tools.py:66-67 passes —
        or a (N,) / (N,1) integer tensor of leaf indices (fast path, no row decode)."""
        if isinstance(actions, np.ndarray):
            actions = torch.from_numpy(np.ascontiguousarray(actions))
        if actions.dtype in (torch.int32, torch.int64) and actions.numel() == self.num_envs:
            self._pending = ("idx", actions.reshape(-1).to(self.device, non_blocking=True))
        else:
            self._pending = ("rows", actions.reshape(self.num_envs, -1)[:, :9].to(self.device, non_blocking=True))

    def step_wait(self):
        kind, a = self._pending
        self._pending = None
        if kind == "idx":
            obs, rew, done, info = self.batch.step(leaf_idx=a)
        else:
            if a.shape[1] < 9:  # 6-float rows (evaluation_tools.py:24)
                a = torch.cat([a, torch.zeros((a.shape[0], 9 - a.shape[1]), dtype=a.dtype, device=a.device)], dim=1)
            obs, rew, done, info = self.batch.step(actions=a)
        done_h = done.cpu().numpy().astype(bool)
        reward = rew.detach().cpu().unsqueeze(1)
        infos = self._infos(info, done_h)
        return obs, reward, done_h, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def _infos(self, info, done_h):
        rec = PctBatch.decode_info(info)
        out = []
        now = round(time.time() - self._tstart, 6)
        for e in range(self.num_envs):
            d = {"counter": int(rec["counter"][e])}
            if done_h[e]:
                ratio = float(rec["ratio"][e])
                d.update(ratio=ratio, reward=ratio * 10,
                         episode={"r": round(float(rec["ep_reward"][e]), 6), "l": int(rec["ep_len"][e]), "t": now})
            if rec["flags"][e]:
                d["flags"] = int(rec["flags"][e])
            out.append(d)
        return tuple(out)

    def close(self):
        if not self.closed:
            self.batch.close()
            self.closed = True

    @property
    def unwrapped(self):
        return self

    def get_images(self):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError
