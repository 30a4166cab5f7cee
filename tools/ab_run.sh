python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02t_gputest.log
for i in 1 2; do
  B21_LIB=$PWD/brats21_b200/libb21_prev.so python bench.py --no-cpu-baseline --no-train --steps 5 --warmup 3 > gpurun_out/r02t_ab_prev_$i.json 2> gpurun_out/r02t_ab_prev_$i.err
  python bench.py --no-cpu-baseline --no-train --steps 5 --warmup 3 > gpurun_out/r02t_ab_new_$i.json 2> gpurun_out/r02t_ab_new_$i.err
done
tail -3 gpurun_out/r02t_gputest.log; for f in gpurun_out/r02t_ab_*.json; do echo $f; cut -c1-220 $f; done
