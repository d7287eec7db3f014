"""Worker of tests/test_gpu_multi.py (one process per GPU, NCCL): every rank steps its shard of N global envs for T steps with the keyed policy, the
observation buffers are all-gathered (the path's one optional collective), and rank 0 compares the gathered batch with ONE device stepping all N envs:
env by env, step by step.  Prints MULTI_GPU_OK on success."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pct_b200  # noqa: E402

ITEM_SET = [(i, j, k) for i in range(1, 6) for j in range(1, 6) for k in range(1, 6)]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_total, T, setting, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 55
    base, count = pct_b200.shard_range(n_total, world, rank)
    shard = pct_b200.PctBatch(count, setting, item_set=ITEM_SET, seed=seed, env_id_base=base, device=local)
    full = pct_b200.PctBatch(n_total, setting, item_set=ITEM_SET, seed=seed, device=local) if rank == 0 else None
    obs = shard.reset()
    ref = full.reset() if full else None
    ok = True
    for t in range(T):
        gathered = pct_b200.gather_observations(obs, world)
        if rank == 0 and not torch.equal(gathered, ref):
            bad = (gathered != ref).any(dim=1).nonzero().flatten()[:8].tolist()
            print("step %d: gathered shards differ from the single-device batch at envs %s" % (t, bad), flush=True)
            ok = False
            break
        obs, _, _, _ = shard.step(leaf_idx=shard.random_policy(seed, t))
        if full:
            ref, _, _, _ = full.step(leaf_idx=full.random_policy(seed, t))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    if rank == 0 and ok:
        print("MULTI_GPU_OK world=%d envs=%d steps=%d setting=%d" % (world, n_total, T, setting), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag[0]) else 1)


if __name__ == "__main__":
    main()
