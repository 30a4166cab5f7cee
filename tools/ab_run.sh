#!/bin/bash
# same-box A/B of one environment switch (default: the first-conv kernel):  bash tools/ab_run.sh <tag> [VAR]
tag=${1:-ab}
var=${2:-B21_INPUT_CONV}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Warning\|warnings.warn" | tail -150 > gpurun_out/${tag}_gputest.log
tail -8 gpurun_out/${tag}_gputest.log
for i in 1 2; do
  env $var=0 python bench.py --no-cpu-baseline --no-train --steps 5 --warmup 3 > gpurun_out/${tag}_off_$i.json 2> gpurun_out/${tag}_off_$i.err
  env $var=1 python bench.py --no-cpu-baseline --no-train --steps 5 --warmup 3 > gpurun_out/${tag}_on_$i.json 2> gpurun_out/${tag}_on_$i.err
done
for f in gpurun_out/${tag}_o*.json; do echo $f; python - $f <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(d["ms_per_step"], round(r["frac"], 4), {k: (round(v["ms"], 1), round(v["tflops"])) for k, v in r["by_kernel"].items()})
except Exception as e:
    print("unreadable:", e)
PY
done
