"""the reference's OWN CPU vector env — ShmemVecEnv(context='fork') over Monitor-less PackingDiscrete / PackingContinuous
workers (wrapper/shmem_vec_env.py:20-156, envs.py:27-73,107-108), unmodified, imported from the git-ignored copy baseline/_ref that
scratch/install_reference.sh makes — timed on THIS box's host cores with the same synthetic inputs as the GPU arm: uniform items over the 125
sizes (the env's own RandomBoxCreator), shuffle=False, random-valid-leaf policy in the parent on the returned numpy observations.

    python scratch/shmem_baseline.py [--procs 1,8,64,128] [--setting 1] [--continuous] [--steps 200] [--warmup 20]   -> one JSON line per process count

The product never imports this; it is the measurement the BASELINE metric names ("vs CPU shmem_vec_env")."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
os.environ.setdefault("PCT_REFERENCE_ROOT", REF if os.path.isdir(REF) else "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402
import ref_shim  # noqa: E402  (gym stub + np.float alias; test infrastructure, used here as the loader only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", default="1,8,64")
    ap.add_argument("--setting", type=int, default=1)
    ap.add_argument("--continuous", action="store_true")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    a = ap.parse_args()
    D, Cm = ref_shim.load_reference()
    from wrapper.shmem_vec_env import ShmemVecEnv
    items = [(i, j, k) for i in range(1, 6) for j in range(1, 6) for k in range(1, 6)]

    def thunk(rank):
        def _t():
            if a.continuous:
                env = Cm.PackingContinuous(setting=a.setting, container_size=[1.0, 1.0, 1.0], item_set=[(0.1 * i, 0.1 * j, 0.1 * k) for i, j, k in items],
                                           internal_node_holder=80, leaf_node_holder=50, shuffle=False, sample_from_distribution=True,
                                           sample_left_bound=0.1, sample_right_bound=0.5)
            else:
                env = D.PackingDiscrete(setting=a.setting, container_size=[10, 10, 10], item_set=items, internal_node_holder=80, leaf_node_holder=50,
                                        shuffle=False, LNES="EMS")
            env.seed(100 + rank)
            return env
        return _t

    probe = thunk(0)()
    for n in [int(x) for x in a.procs.split(",")]:
        venv = ShmemVecEnv([thunk(r) for r in range(n)], [probe.observation_space, None], context="fork")
        obs = venv.reset()
        rng = np.random.default_rng(0)

        def policy(o):
            leaf = np.asarray(o).reshape(n, -1, 9)[:, 80:130]
            nv = (leaf[:, :, 8] == 1).sum(1)
            k = (rng.integers(0, 1 << 30, n) % np.maximum(nv, 1))
            rows = leaf[np.arange(n), k].copy()
            rows[nv == 0] = 0
            return rows
        for _ in range(a.warmup):
            obs, _, _, _ = venv.step(policy(obs))
        t0 = time.perf_counter()
        for _ in range(a.steps):
            obs, _, _, _ = venv.step(policy(obs))
        dt = time.perf_counter() - t0
        venv.close()
        print(json.dumps({"impl": "reference ShmemVecEnv (unmodified, context=fork)", "procs": n, "host_cores": os.cpu_count(), "setting": a.setting,
                          "continuous": a.continuous, "steps": a.steps, "env_steps_per_s": n * a.steps / dt, "seconds": dt}), flush=True)


if __name__ == "__main__":
    main()
