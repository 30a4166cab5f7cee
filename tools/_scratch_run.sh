timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_networks.py -x -q 2>&1 | tail -3
for i in 1 2; do
for sw in 0 1; do
  echo "== B21_UPSAMPLE_WIDE=$sw"
  B21_UPSAMPLE_WIDE=$sw python bench.py --no-cpu-baseline --no-train --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['hbm']['by_kernel']['upsample2x']
print(d['ms_per_step'], round(r['ms'],2), round(r['gbs']), round(r['frac'],3))"
done
done
