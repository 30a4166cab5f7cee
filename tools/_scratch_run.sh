for sb in 9 18 12 9; do
  echo "== sw-batch $sb"
  python bench.py --no-cpu-baseline --no-train --steps 5 --warmup 3 --sw-batch $sb 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print(d['ms_per_step'], d['e2e']['ms_per_step'], round(r['frac'],4), d['gpu_launches'])"
done
