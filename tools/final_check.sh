timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Warning\|warnings.warn" | tail -30 > gpurun_out/r02ae_gputest.log; tail -2 gpurun_out/r02ae_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-300
