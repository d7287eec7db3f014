"""GPU, >= 2 devices: the N > 1 path on real GPUs.  (1) one process, two devices: 2 x N/2 envs keyed by the global env index equal
1 x N envs env by env; (2) two processes over NCCL (python -m torch.distributed.run, 127.0.0.1): the shards' observation buffers all-gathered
(pct_b200.gather_observations, the one optional collective of the path) equal the single-device batch at every step.  Skipped with fewer than 2 GPUs —
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu` runs it."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
from harness import ITEM_SET  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need2():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


@pytest.mark.parametrize("setting", [1, 2])
def test_two_device_shards_equal_one_device_batch(setting):
    _need2()
    import pct_b200
    n, T, seed = 256, 60, 9
    a = pct_b200.PctBatch(n // 2, setting, item_set=ITEM_SET, seed=seed, env_id_base=0, device=0)
    b = pct_b200.PctBatch(n // 2, setting, item_set=ITEM_SET, seed=seed, env_id_base=n // 2, device=1)
    f = pct_b200.PctBatch(n, setting, item_set=ITEM_SET, seed=seed, device=0)
    oa, ob, of = a.reset(), b.reset(), f.reset()
    for t in range(T):
        assert torch.equal(torch.cat([oa.cpu(), ob.cpu()]), of.cpu()), "step %d" % t
        oa, ra, da, _ = a.step(leaf_idx=a.random_policy(seed, t))
        ob, rb, db, _ = b.step(leaf_idx=b.random_policy(seed, t))
        of, rf, df, _ = f.step(leaf_idx=f.random_policy(seed, t))
        assert torch.equal(torch.cat([ra.cpu(), rb.cpu()]), rf.cpu()) and torch.equal(torch.cat([da.cpu(), db.cpu()]), df.cpu())
    for x in (a, b, f):
        x.close()


def test_two_ranks_nccl_gather_equals_one_device():
    _need2()
    port = 29600 + os.getpid() % 300
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py"), "512", "40", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "MULTI_GPU_OK" in out.stdout, (out.stdout[-2000:], out.stderr[-2000:])
