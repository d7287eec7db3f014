"""CPU, world_size 2 over gloo: the N>1 host path — shard ranges, global-index keyed item streams, observation gather."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from harness import rnd_u64


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pct_b200
    base, count = pct_b200.shard_range(n_total, world, rank)
    # a rank's "observation" = the first item draw of each of its envs, computed from the GLOBAL env index
    obs = torch.tensor([[float(rnd_u64(1234, base + e, 0) % 125), float(base + e)] for e in range(count)])
    full = pct_b200.gather_observations(obs, world)
    q.put((rank, base, count, full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_and_are_disjoint():
    import pct_b200
    for n, w in ((65536, 8), (4096, 2), (10, 4), (7, 3)):
        spans = [pct_b200.shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (b0, c0), (b1, _) in zip(spans, spans[1:]):
            assert b0 + c0 == b1


def test_two_rank_gather_matches_single_process():
    world, n_total, port = 2, 12, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    single = np.array([[float(rnd_u64(1234, e, 0) % 125), float(e)] for e in range(n_total)])
    for rank, base, count, full in res:
        assert np.array_equal(full, single), "rank %d sees a different global batch" % rank
