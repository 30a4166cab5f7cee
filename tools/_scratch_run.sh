timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Warning\|warnings.warn" | tail -60 > gpurun_out/r02aa_gputest.log; tail -2 gpurun_out/r02aa_gputest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02aa_bench.json 2> gpurun_out/r02aa_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02aa_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d.get('clocks'))
t=d.get('train',{})
print({k:t[k] for k in t if k in ('value','ms_per_step','e2e','conv_roofline_frac','conv_tflops')} or list(t.keys()))
PY
