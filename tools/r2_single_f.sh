#!/bin/bash
# Round 2, final 1-GPU run: sanitizers, whole GPU suite, smoke, 8f rows at size, sweeps, ncu launch list + full captures,
# the default bench line and the other single-GPU workloads.
cd "$(dirname "$0")/.."
REPO=$(pwd); OUT=$REPO/gpurun_out/r2f; mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1; nproc > $OUT/host.txt; free -g >> $OUT/host.txt
B=$REPO/icde2019-gpu-join_b200/bin/bench
(cd /tmp && timeout 300 $B -b 7 -a HJC -R 1048576 -S 1048576 --parallel-gen) > $OUT/driver_small.log 2>&1; echo "exit $?" >> $OUT/driver_small.log
(cd /tmp && timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 $B -b 7 -a HJC -R 3000000 -S 5000000 --parallel-gen) > $OUT/memcheck.log 2>&1; echo "exit $?" >> $OUT/memcheck.log
(cd /tmp && timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 $B -b 7 -a HJC -R 2200000 -S 2400000 -s 1.0 --payload rowid) > $OUT/racecheck.log 2>&1; echo "exit $?" >> $OUT/racecheck.log
tail -3 $OUT/memcheck.log $OUT/racecheck.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "exit $?" >> $OUT/smoke.log; tail -2 $OUT/smoke.log
timeout 600 python tools/next_rows_bench.py late > $OUT/next_rows_late.log 2>&1; echo "exit $?" >> $OUT/next_rows_late.log; tail -4 $OUT/next_rows_late.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.log 2>&1; echo "exit $?" >> $OUT/bench.log
tail -2 $OUT/bench.log | cut -c1-1200
for w in A zipf0.5 zipf1.0; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --no-ref-cuda --no-cfg5 > $OUT/bench_$w.log 2>&1; echo "exit $?" >> $OUT/bench_$w.log
  tail -2 $OUT/bench_$w.log | cut -c1-400
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/ncu_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-cfg5 --no-materialize > $OUT/ncu_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scatter_kernel -s 4 -c 4 -f -o $OUT/prof_scatter \
    python tools/sweep.py --what none --reps 1 > $OUT/ncu_scatter.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"join_kernel|hist_kernel" -s 3 -c 3 -f -o $OUT/prof_join_hist \
    python tools/sweep.py --what none --reps 1 > $OUT/ncu_join.log 2>&1
for k in scatter join_hist; do
  ncu -i $OUT/prof_$k.ncu-rep --page raw --csv > $OUT/ncu_raw_$k.csv 2>/dev/null
done
python tools/ncu_source_summary.py $OUT/prof_scatter.ncu-rep 'scatter_kernel.*\(bool\)1' 0 > $OUT/ncu_source_scatter_pass1.txt 2>&1
python tools/ncu_source_summary.py $OUT/prof_scatter.ncu-rep 'scatter_kernel.*\(bool\)0' 0 > $OUT/ncu_source_scatter_pass2.txt 2>&1
python tools/ncu_source_summary.py $OUT/prof_join_hist.ncu-rep join_kernel 0 > $OUT/ncu_source_join.txt 2>&1
python tools/ncu_source_summary.py $OUT/prof_join_hist.ncu-rep hist_kernel 0 > $OUT/ncu_source_hist.txt 2>&1
rm -f $OUT/*.ncu-rep
ls -la $OUT
echo done
