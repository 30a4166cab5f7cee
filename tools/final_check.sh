#!/bin/bash
# What the driver runs at round end, in one gpurun call: GPU tests, smoke(), the default bench line.
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Warning\|warnings.warn" | tail -30 > gpurun_out/final_gputest.log; tail -2 gpurun_out/final_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], round(d['roofline']['frac'], 4), d['clocks'])
print('train', d['train']['value'], d['train']['ms_per_step'], d['train'].get('roofline', {}).get('frac'))
PY
