"""collect_final.py [DIR] [COMMIT-NOTE] — turn the outputs of scratch/r2_final.sh (gpurun_out/r2_final/) into the tracked evidence under profiles/:
the bench lines, the ncu --set full summaries of the five discrete kernels and of the continuous kernels, traffic.json, the launch list with the
kernels' shares of the step, and the test / smoke tails."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else os.path.join(ROOT, "gpurun_out", "r2_final")
note = sys.argv[2] if len(sys.argv) > 2 else subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
P = os.path.join(ROOT, "profiles")

TRAFFIC_ONLY = "--traffic-only" in sys.argv  # on the GPU box, between the ncu captures and the bench runs: bench.py then reports the traffic of THIS build
if not TRAFFIC_ONLY and "--no-bench" not in sys.argv:  # --no-bench: the run was cut before its benches (keep the bench lines of the previous complete run)
    for a, b in (("bench_default.json", "r2_final_bench.json"), ("bench_driver_flags.json", "r2_final_bench_driver_flags.json"),
                 ("bench_reference.json", "r2_final_bench_reference.json")):
        line = [l for l in open(os.path.join(src, a)).read().splitlines() if l.startswith("{")][-1]
        json.loads(line)
        open(os.path.join(P, b), "w").write(line + "\n")

traffic = {"commit": note, "file": "r2_head_pct_*.txt", "kernels": {}}
for k in ("pct_apply", "pct_candidates", "pct_walk_light", "pct_walk_kernel", "pct_emit"):
    rep = os.path.join(src, "head_%s.ncu-rep" % k)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scratch", "ncu_summary.py"), rep, "30"], capture_output=True, text=True).stdout
    open(os.path.join(P, "r2_head_%s.txt" % k), "w").write("# ncu --set full --clock-control none --import-source on, one launch of the default bench (4096 envs, setting 1), %s\n" % note + out)
    for l in out.splitlines():
        if "TRAFFIC_JSON" in l:
            d = json.loads(l.split("TRAFFIC_JSON", 1)[1])
            name = re.sub(r"<.*|\(.*", "", d.pop("kernel")).split("::")[-1].replace("void ", "").strip()
            d["note"] = "4096 envs, setting 1, one launch, ncu --set full (caches flushed per replay)"
            traffic["kernels"][name] = d
traffic["step_total_dram_bytes"] = sum(v["dram_bytes_per_launch"] for v in traffic["kernels"].values())
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
if TRAFFIC_ONLY:
    print(json.dumps({k: v["dram_bytes_per_launch"] for k, v in traffic["kernels"].items()}), traffic["step_total_dram_bytes"])
    sys.exit(0)

out = subprocess.run([sys.executable, os.path.join(ROOT, "scratch", "ncu_summary.py"), os.path.join(src, "cont_head.ncu-rep"), "20"], capture_output=True, text=True).stdout
open(os.path.join(P, "r2_cont_head_final.txt"), "w").write("# ncu --set full, five consecutive launches of the continuous kernels (bench.py --continuous, 4096 envs, setting 2), %s\n" % note + out)

# launch list: per kernel name count, mean us, share of the listed time
rows = [r for r in csv.reader(l for l in open(os.path.join(src, "launches.csv")) if not l.startswith("==")) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
acc = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = re.sub(r"<.*|\(.*", "", r[ki]).split("::")[-1].replace("void ", "").strip()
    a = acc.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
unit = rows[1][hdr.index("Metric Unit")] if "Metric Unit" in hdr else "ns"
scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1e-3)
tot = sum(a[1] for a in acc.values())
with open(os.path.join(P, "r2_final_launch_shares.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, launches 500..569 of `bench.py --steps 3 --warmup 60 ...` (%s); serialised, cold caches:\n"
            "# compare the SHARES with bench.py's roofline.all_kernels_ms, not the absolute times\n" % note)
    f.write("%-34s %6s %10s %8s\n" % ("kernel", "n", "mean us", "share"))
    for name, (n, t) in acc.items():
        f.write("%-34s %6d %10.2f %7.1f%%\n" % (name, n, t * scale / n, 100.0 * t / tot))
shutil.copy(os.path.join(src, "launches.csv"), os.path.join(P, "r2_final_launches.csv"))
with open(os.path.join(P, "r2_final_tests.txt"), "w") as f:
    f.write("# scratch/r2_final.sh on one B200 (%s): pytest -m gpu, smoke(), return codes\n" % note)
    f.write(open(os.path.join(src, "summary.txt")).read().split("total ")[0])
print(open(os.path.join(P, "r2_final_launch_shares.txt")).read())
print(json.dumps({k: v["dram_bytes_per_launch"] for k, v in traffic["kernels"].items()}), traffic["step_total_dram_bytes"])
