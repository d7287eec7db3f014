"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

from pct_oracle import OracleDiscrete, policy_pick, rnd_u64  # noqa: F401  (oracle/ is on sys.path via conftest)

ITEM_SET = [(i, j, k) for i in range(1, 6) for j in range(1, 6) for k in range(1, 6)]  # givenData.py:7-14
# the same 125 sizes scaled to the unit container of the continuous domain (givenData.py:4 `container_size = [1,1,1]`)
CONT_ITEM_SET = [(round(0.1 * i, 1), round(0.1 * j, 1), round(0.1 * k, 1)) for i in range(1, 6) for j in range(1, 6) for k in range(1, 6)]


def make_stream(seed, env, n, setting):
    """Deterministic per-env draw sequence (x,y,z,density) over the 125-item set."""
    s = np.zeros((n, 4))
    for d in range(n):
        s[d, :3] = ITEM_SET[rnd_u64(seed, env, d) % 125]
        s[d, 3] = max((rnd_u64(seed ^ 0xABCDEF, env, d) >> 11), 1) / float(1 << 53) if setting == 3 else 1.0
    return s


class OracleVec(object):
    """N independent oracle envs with the ShmemVecEnv worker semantics (auto-reset on done,
    wrapper/shmem_vec_env.py:139-143) and the shared deterministic policy."""

    def __init__(self, n, setting, streams, nb=80, nl=50, container=(10, 10, 10)):
        self.envs = [OracleDiscrete(setting, container_size=container, internal_node_holder=nb, leaf_node_holder=nl, stream=streams[i])
                     for i in range(n)]
        self.n, self.nb, self.nl = n, nb, nl

    def reset(self):
        return np.stack([e.reset() for e in self.envs])

    def step(self, actions):
        obs, rew, done, infos = [], [], [], []
        for e, a in zip(self.envs, actions):
            o, r, d, i = e.step(a)
            if d:
                o = e.reset()
            obs.append(o); rew.append(r); done.append(d); infos.append(i)
        return np.stack(obs), np.array(rew), np.array(done), infos

    def pick(self, obs, seed, t, env_id_base=0):
        idx, rows = [], []
        for e in range(self.n):
            k, row = policy_pick(obs[e], self.nb, self.nl, seed, env_id_base + e, t)
            idx.append(k); rows.append(row)
        return np.array(idx, dtype=np.int32), np.stack(rows)


# ---- batched-evaluation parity helpers (evaluation_tools.py:7-52 restated as a plain loop) -------------------------------
def eval_policy_np(obs, nb=80, nl=50):
    """deterministic test policy: a fixed function of (#valid leaves, #valid internal rows)"""
    o = obs.reshape(nb + nl + 1, 9)
    nv, nbx = int((o[nb:nb + nl, 8] == 1).sum()), int((o[:nb, 8] == 1).sum())
    return (7 * nv + 3 * nbx) % max(nv, 1)


def eval_policy_torch(obs, batch):
    import torch
    o = obs.reshape(obs.shape[0], batch.nb + batch.nl + 1, 9)
    nv, nbx = (o[:, batch.nb:batch.nb + batch.nl, 8] == 1).sum(1), (o[:, :batch.nb, 8] == 1).sum(1)
    return ((7 * nv + 3 * nbx) % torch.clamp(nv, min=1)).to(torch.int32)


def sequential_eval(make_env, episodes, nb=80, nl=50):
    """the reference's evaluation loop on any single env with its call surface (reference env or oracle):
    per episode -> (ratio, counter, packed)"""
    out = []
    for ep in range(episodes):
        env, obs = make_env(ep)
        while True:
            k = eval_policy_np(obs, nb, nl)
            row = obs.reshape(nb + nl + 1, 9)[nb + k].copy()
            items = env.packed
            obs, _, done, info = env.step(row)
            if done:
                out.append((float(info["ratio"]), int(info["counter"]), [list(p) for p in items]))
                break
    return out
