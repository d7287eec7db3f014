#!/bin/bash
# Round 2, 8-GPU call: bench at N = 8 (+ 4): weak scaling, config 5 (65 536 envs), all-gather, and the CPU arm on the same workload
O=gpurun_out/r2_multi8; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
for N in 8 4; do
  ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 300 --warmup 100 ) > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench n$N rc=$?" | tee -a $O/summary.txt
done
( timeout 300 python bench.py --impl reference --gpus 8 --steps 20 --warmup 5 ) > $O/bench_reference_n8.json 2>&1; echo "ref n8 rc=$?" | tee -a $O/summary.txt
python - <<'PY' | tee -a gpurun_out/r2_multi8/summary.txt
import json
for f in ("bench_n8", "bench_n4", "bench_reference_n8"):
    try:
        j = [json.loads(l) for l in open("gpurun_out/r2_multi8/%s.json" % f) if l.startswith("{")][0]
        print(f, "n_gpus", j["n_gpus"], "value %.2fM" % (j["value"] / 1e6), "ms/step %.3f" % j["ms_per_step"], "e2e %.2fM" % (j["e2e"]["value"] / 1e6), "allgather", j.get("allgather"),
              "config5", {k: ("%.2fM" % (v["value"] / 1e6)) for k, v in (j.get("configs") or {}).items()})
    except Exception as ex:
        print(f, "ERR", ex)
PY
