for v in base w1m16 w1m12 w4m4; do
  if [ $v = base ]; then unset PCT_B200_LIB; else export PCT_B200_LIB=/root/repo/scratch/variants/lib_$v.so; fi
  timeout 200 python bench.py --skip-cpu --steps 400 --warmup 100 --e2e-steps 20 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'value %.2fM' % (d['value']/1e6), 'ms %.4f' % d['ms_per_step'], d['roofline']['all_kernels_ms'])
"
done
export PCT_B200_LIB=/root/repo/scratch/variants/lib_w1m16.so
timeout 300 python -m pytest tests/test_gpu_discrete_parity.py -x -q 2>&1 | tail -2
