#!/usr/bin/env python
"""bench.py — env-steps/s of the batched PCT step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--setting S] [--envs-per-gpu E]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE batched environment step over all envs of a rank: the synthetic uniform-valid-leaf policy
kernel + the PCT step kernel (action -> placement -> EMS update -> candidate leaves -> feasibility ->
observation / reward / done, with auto-reset).  Workload at N=1 = BASELINE.json configs[1]: setting 1
discrete, 10x10x10 bin, items {1..5}^3, 80 internal / 50 leaf holders, EMS scheme, 4096 envs; weak scaling
(4096 envs per GPU, env streams keyed by the GLOBAL env index).

`value`    : device-timed (CUDA events, L2 flushed between steps, max over ranks), inputs resident in HBM.
`e2e`      : the same metric through the C-ABI host-buffer call (pct_step_host): actions come from pinned
             host memory, observation / reward / done / info are copied back to pinned host memory every
             step and the policy runs on the host from that observation.
`roofline` : HBM roofline of the dominant kernel (pct_feas_emit_kernel), algorithmic bytes per launch
              / its mean launch duration measured here with CUDA events.
`cpu_baseline` / `--impl reference`: the CPU restatement of the reference env (oracle/, C, pthreads over
             envs like the reference's ShmemVecEnv workers) on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ITEM_SET = [(i, j, k) for i in range(1, 6) for j in range(1, 6) for k in range(1, 6)]  # givenData.py:7-14
ITEM_SEED, POLICY_SEED = 1234, 4321
METRIC = "env-steps/s (batched PCT step)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--setting", type=int, default=1)
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between steps (diagnostic only)")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--continuous", action="store_true", help="BASELINE config 4: PctContinuous, sample_from_distribution, bin 1x1x1")
    return ap.parse_args()


def workload_name(a, n_gpus):
    if a.continuous:
        return "setting %d continuous (sample_from_distribution U(0.1,0.5)), bin 1x1x1, 80 internal / 50 leaf, EMS, %d envs/GPU x %d GPU" % (
            a.setting, a.envs_per_gpu, n_gpus)
    return "setting %d discrete, bin 10x10x10, items 1-5, 80 internal / 50 leaf, EMS, %d envs/GPU x %d GPU" % (
        a.setting, a.envs_per_gpu, n_gpus)


# ------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx, self.rows, self.stop_flag, self.proc = gpu_index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu summary (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------------------
def cpu_port_rate(setting, n_envs, warm, steps, threads=None, continuous=False):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pct_oracle  # the ONLY place bench.py touches oracle/: the CPU baseline / reference arm
    if continuous:  # BASELINE config 4: PackingContinuous, sample_from_distribution U(0.1, 0.5), unit container
        b = pct_oracle.OracleBatchContinuous(n_envs, setting, ITEM_SEED, POLICY_SEED, threads=threads)
    else:
        b = pct_oracle.OracleBatch(n_envs, setting, ITEM_SET, ITEM_SEED, POLICY_SEED, threads=threads)
    if warm:
        b.run(warm)
    dt = b.run(steps)
    cores = b.threads
    b.close()
    return n_envs * steps / dt, dt, cores


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = a.envs_per_gpu * a.gpus  # whole-job workload of the GPU arm, stepped by all host threads
    # bounded sample: the CPU steps at most 8192 envs per vector step
    n_s = min(n, 8192)
    rate, dt, cores = cpu_port_rate(a.setting, n_s, a.warmup, a.steps, continuous=a.continuous)
    line = {"metric": METRIC, "value": rate, "unit": "env-steps/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if a.continuous else "int32+f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(a, a.gpus), "note": "reference's CPU env path: C restatement (oracle/, kind=port; the "
                       "reference itself is pure Python and does not travel to the GPU box), pthreads over envs like ShmemVecEnv workers"},
            "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
                             "sample": "%d envs x %d vector steps (+%d warm-up)" % (n_s, a.steps, a.warmup)},
            "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
def run_ours(a):
    import numpy as np
    import torch
    import pct_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n = a.envs_per_gpu
    K, W = a.steps, max(a.warmup, 3)

    if a.continuous:
        batch = pct_b200.PctBatch(n, a.setting, container_size=(1.0, 1.0, 1.0), continuous=True, sample_from_distribution=True,
                                  seed=ITEM_SEED, env_id_base=rank * n, device=local)
    else:
        batch = pct_b200.PctBatch(n, a.setting, item_set=ITEM_SET, seed=ITEM_SEED, env_id_base=rank * n, device=local)
    launches0 = batch.kernel_launches
    flush = None if a.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    obs = batch.reset()
    for t in range(W):
        idx = batch.random_policy(POLICY_SEED, t)
        batch.step(leaf_idx=idx)
    torch.cuda.synchronize()

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    stats = torch.zeros((4,), dtype=torch.float64, device=dev)  # sums of counter, n_leaf, n_cand, n_ems
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    l0 = batch.kernel_launches
    t_wall0 = time.perf_counter()
    for t in range(K):
        if flush is not None:
            flush.zero_()  # L2 flush (256 MiB > 126 MB L2), outside the timed interval of the step
        ev[t][0].record()
        idx = batch.random_policy(POLICY_SEED, W + t)
        ev[t][1].record()
        _, _, _, info = batch.step(leaf_idx=idx)
        ev[t][2].record()
        if t % 16 == 0:
            stats += torch.stack([info[:, 0].double().mean(), info[:, 5].double().mean(), info[:, 6].double().mean(), info[:, 7].double().mean()])
    barrier()
    wall = time.perf_counter() - t_wall0
    launches = batch.kernel_launches - l0
    # Per-kernel durations for the roofline: a second pass over the same number of steps with CUDA events between the three
    # kernels (pct_profile_enable).  Events between the kernels serialise them, i.e. this pass runs WITHOUT the overlapped
    # launch mode the timed region above uses (programmatic dependent launch + per-env hand-over flags), so a kernel's duration
    # here is its stand-alone duration; the step time of the timed region is shorter than their sum.
    kms, ksteps = {}, 0
    if not a.continuous:
        batch.profile(True)
        for t in range(K):
            if flush is not None:
                flush.zero_()
            batch.step(leaf_idx=batch.random_policy(POLICY_SEED, W + K + t))
        torch.cuda.synchronize()
        kms, ksteps = batch.profile_read()
        batch.profile(False)
    clocks = sampler.finish() if sampler else None
    per_step = sorted(ev[t][0].elapsed_time(ev[t][2]) for t in range(K))
    step_ms = sum(per_step)
    kern_ms = sum(ev[t][1].elapsed_time(ev[t][2]) for t in range(K))
    tt = torch.tensor([step_ms, kern_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    step_ms, kern_ms = float(tt[0]), float(tt[1])
    nsamp = len(range(0, K, 16))
    mean_boxes, mean_leaf, mean_cand, mean_ems = [float(x) / nsamp for x in stats.cpu()]
    value = world * n * K / (step_ms * 1e-3)

    # ---- e2e: host buffers through pct_step_host, host policy on the returned observation ----
    Ke, We = max(3, min(a.e2e_steps, K)), 5
    ol = batch.obs_len
    obs_h = torch.empty((n, ol), dtype=torch.float32, pin_memory=True).numpy()
    rew_h = torch.empty((n,), dtype=torch.float32, pin_memory=True).numpy()
    done_h = torch.empty((n,), dtype=torch.uint8, pin_memory=True).numpy()
    info_h = torch.empty((n, 8), dtype=torch.int32, pin_memory=True).numpy()
    idx_h = torch.empty((n,), dtype=torch.int32, pin_memory=True).numpy()
    gid = (np.arange(n, dtype=np.uint64) + np.uint64(rank * n))
    GOLD = np.uint64(0x9E3779B97F4A7C15)

    def sm64(x):
        with np.errstate(over="ignore"):
            x = x + GOLD
            z = x
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))

    def host_policy(t):
        # uniform choice among the valid leaves, on the HOST from the step's returned records: the number of valid leaf rows
        # is pct_step_info.n_leaf (== the count of 1s in column 8 of the leaf rows of the returned observation)
        nvalid = info_h[:, 5].astype(np.uint64)
        with np.errstate(over="ignore"):
            r = sm64(sm64(np.uint64(POLICY_SEED) ^ (gid * GOLD)) + np.uint64(t))
        idx_h[:] = np.where(nvalid > 0, r % np.maximum(nvalid, np.uint64(1)), 0).astype(np.int32)

    batch.reset_host(obs_h)
    info_h[:, 5] = (obs_h.reshape(n, -1, 9)[:, batch.nb:batch.nb + batch.nl, 8] == 1).sum(1)  # first step: count from the observation
    for t in range(We):
        host_policy(t)
        batch.step_host(obs_h, rew_h, done_h, info_h, leaf_idx=idx_h)
    barrier()
    t0 = time.perf_counter()
    for t in range(Ke):
        host_policy(We + t)
        batch.step_host(obs_h, rew_h, done_h, info_h, leaf_idx=idx_h)
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    te = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n * Ke / float(te[0])
    h2d = idx_h.nbytes
    d2h = obs_h.nbytes + rew_h.nbytes + done_h.nbytes + info_h.nbytes

    if rank == 0:
        peak, peak_src = measured_peak()
        # algorithmic bytes per env-step.  Dominant kernel = pct_feas_emit_kernel (K3): it reads the
        # hot record, the staged loads / polygons and the candidate list, writes the observation, the leaf list and info.
        hot, prefix = 3584, 2576  # sizeof(DEnvHot), HOT_PREFIX
        stab = a.setting != 2
        edges = mean_boxes if stab else 0.0
        b_k3 = hot + (32 * edges + 16 * 0.3 * edges if stab else 0) + 2 * mean_cand + ol * 4 + 12 * mean_leaf + 16
        b_step = 2 * hot + prefix + 2 * 2 * mean_cand + b_k3 + (2 * 32 * edges if stab else 0) + 4 + 4 + 1 + 32
        if ksteps:
            k3_ms = kms["feas_emit"] / ksteps
            ach = b_k3 * n / (k3_ms * 1e-3) / 1e9
        else:
            k3_ms = None
            ach = b_step * n / (kern_ms / K * 1e-3) / 1e9
        # SURVEY.md 8(d) / BASELINE.md 3.5 figure for the WHOLE step, independent of this implementation's record layout:
        # B_step = 5593 + 24 N + 48 E bytes (action, item, boxes, EMS before / after, height map, observation, reward / done)
        b_survey = 5593.0 + 24.0 * mean_boxes + 48.0 * mean_ems
        ach_survey = b_survey * n / (kern_ms / K * 1e-3) / 1e9  # per GPU: n envs of this rank over this rank's kernel time
        traffic = ncu_traffic()
        line = {"metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": step_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int16/int32 geometry + f64 stability, f32 observations", "data": "synthetic",
                "config": {"workload": workload_name(a, world), "items": "device counter-based RNG seed %d, uniform over 125 sizes" % ITEM_SEED,
                           "policy": "uniform over valid leaves (device kernel)", "launch_mode": os.environ.get("PCT_B200_OVERLAP", "1") != "0" and
                           "overlapped (PDL + per-env flags)" or "back-to-back kernels", "l2": "not flushed (diagnostic)" if a.no_flush else
                           "flushed between steps (256 MiB memset outside the timed interval)",
                           "mean_boxes": mean_boxes, "mean_ems": mean_ems, "mean_valid_leaves": mean_leaf, "mean_candidates": mean_cand,
                           "switches": {k: v for k, v in sorted(os.environ.items()) if k.startswith("PCT_B200_")}},
                "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "steps": Ke, "path": ("pct_step_host (C ABI, pinned host buffers, zero-copy: kernels write the observation into the mapped host buffer)"
                                 if os.environ.get("PCT_B200_HOST_ZEROCOPY", "0") != "0" else
                                 "pct_step_host (C ABI, pinned host buffers, 4 pipelined env ranges)") + " + numpy policy on the host from the returned step records"},
                "gpu_launches": int(launches), "kernel_ms_per_step": kern_ms / K,
                "ms_per_step_p50": per_step[K // 2], "ms_per_step_p99": per_step[min(K - 1, int(K * 0.99))], "wall_s_timed_loop": wall,
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "traffic": traffic.get("dram_bytes_per_launch") if traffic else None, "kernel": "pct_feas_emit_kernel",
                             "peak_source": peak_src, "algorithmic_bytes_per_env_kernel": b_k3, "algorithmic_bytes_per_env_step": b_step,
                             "kernel_ms": k3_ms, "all_kernels_ms": {k: v / ksteps for k, v in kms.items()} if ksteps else None,
                             "kernel_timing": "second pass of %d steps right after the timed region with CUDA events between the three kernels; "
                                              "the events serialise the kernels, the timed region itself runs them overlapped (programmatic "
                                              "dependent launch + per-env hand-over flags), so ms_per_step < sum of these" % ksteps,
                             "step_fraction_of_peak": b_step * n / (kern_ms / K * 1e-3) / 1e9 / peak,
                             "survey_formula": {"bytes_per_env_step": b_survey, "achieved": ach_survey, "frac": ach_survey / peak,
                                                "note": "SURVEY 8(d): (5593 + 24 N + 48 E) B x env-steps/s of one GPU / peak, whole step"}},
                "clocks": clocks}
        if world == 1 and not a.skip_cpu:
            try:
                rate, dt, cores = cpu_port_rate(a.setting, 2048, 20, 300, continuous=a.continuous)
                line["cpu_baseline"] = {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                        "sample": "2048 envs x 300 vector steps (+20 warm-up), same items / policy"}
            except Exception as ex:  # the oracle is test infrastructure; never fail the GPU number on it
                line["cpu_baseline"] = {"value": None, "error": str(ex)}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
