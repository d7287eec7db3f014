#!/bin/bash
# Round 2, 2-GPU call: multi-GPU parity tests on real GPUs + bench at N = 2 (weak scaling line with config 5's per-GPU load and the timed all-gather)
O=gpurun_out/r2_multi2; mkdir -p $O
export PCT_B200_LIB=$PWD/scratch/variants/lib_final.so
nvidia-smi -L > $O/gpus.txt
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -4 $O/tests.log | tee -a $O/summary.txt
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 100 ) > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?" | tee -a $O/summary.txt
( timeout 300 python bench.py --gpus 1 --steps 300 --warmup 100 --skip-configs --skip-cpu ) > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench n1 rc=$?" | tee -a $O/summary.txt
python - <<'PY' | tee -a gpurun_out/r2_multi2/summary.txt
import json
for f in ("bench_n1", "bench_n2"):
    try:
        j = [json.loads(l) for l in open("gpurun_out/r2_multi2/%s.json" % f) if l.startswith("{")][0]
        print(f, "n_gpus", j["n_gpus"], "value %.2fM" % (j["value"] / 1e6), "ms/step %.3f" % j["ms_per_step"], "e2e %.2fM" % (j["e2e"]["value"] / 1e6), "allgather", j.get("allgather"),
              "config5", {k: ("%.2fM" % (v["value"] / 1e6)) for k, v in (j.get("configs") or {}).items()})
    except Exception as ex:
        print(f, "ERR", ex)
PY
