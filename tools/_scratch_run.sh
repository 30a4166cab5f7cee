timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_networks.py -x -q 2>&1 | tail -4
for i in 1 2 3; do
for lib in libb21_prev.so libb21.so; do
  echo "== $lib"
  B21_LIB=$PWD/brats21_b200/$lib python bench.py --no-cpu-baseline --no-train --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print(d['ms_per_step'], round(r['frac'],4), {k:(round(v['ms'],1),round(v['tflops'])) for k,v in r['by_kernel'].items()})"
done
done
