timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3800 gpurun_out/bench_default.json
timeout 200 python bench.py --skip-cpu --steps 300 --warmup 100 --e2e-steps 20 --setting 2 --envs-per-gpu 8192 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config3 value %.2fM ms %.4f' % (d['value']/1e6, d['ms_per_step']))"
timeout 200 python bench.py --skip-cpu --steps 300 --warmup 100 --continuous 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config4 value %.2fM ms %.4f' % (d['value']/1e6, d['ms_per_step']))"
