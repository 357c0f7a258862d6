#!/bin/bash
# One gpurun call: correctness first, then numbers, then profiles.  Everything lands in gpurun_out/.
# usage: tools/gpu_round.sh [stages]   stages default: "sanity tests full bench sweep ncu"
cd "$(dirname "$0")/.."
REPO=$(pwd)
OUT=$REPO/gpurun_out
mkdir -p $OUT
STAGES=${1:-"sanity tests full bench sweep ncu"}
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
nproc > $OUT/host.txt; free -g >> $OUT/host.txt
has() { [[ " $STAGES " == *" $1 "* ]]; }

if has sanity; then
  (cd /tmp && timeout 300 $REPO/icde2019-gpu-join_b200/bin/bench -b 7 -a HJC -R 1048576 -S 1048576 --parallel-gen) > $OUT/driver_small.log 2>&1
  echo "exit $?" >> $OUT/driver_small.log
  (cd /tmp && timeout 300 $REPO/icde2019-gpu-join_b200/bin/bench -b 7 -a HJC -R 1000000 -S 4000000 -s 1.0 --payload rowid) > $OUT/driver_zipf.log 2>&1
  echo "exit $?" >> $OUT/driver_zipf.log
  (cd /tmp && timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 $REPO/icde2019-gpu-join_b200/bin/bench -b 7 -a HJC -R 300000 -S 700000 --parallel-gen) > $OUT/memcheck.log 2>&1
  echo "exit $?" >> $OUT/memcheck.log
  (cd /tmp && timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 $REPO/icde2019-gpu-join_b200/bin/bench -b 7 -a HJC -R 100000 -S 200000 --parallel-gen) > $OUT/racecheck.log 2>&1
  echo "exit $?" >> $OUT/racecheck.log
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
  echo "exit $?" >> $OUT/smoke.log
fi
if has tests; then
  timeout 2400 python -m pytest tests/test_gpu_join.py tests/test_gpu_reference_differential.py -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
  echo "exit $?" >> $OUT/pytest_gpu.log
fi
if has full; then
  timeout 2400 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/pytest_full.log 2>&1
  echo "exit $?" >> $OUT/pytest_full.log
fi
if has bench; then
  timeout 1500 python bench.py --steps 10 --warmup 3 > $OUT/bench.log 2>&1
  echo "exit $?" >> $OUT/bench.log
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.log 2>&1
fi
if has sweep; then
  timeout 1200 python tools/sweep.py > $OUT/sweep.log 2>&1
  echo "exit $?" >> $OUT/sweep.log
  timeout 600 python tools/sweep.py --n 16777216 --nS 268435456 --what join > $OUT/sweep_A.log 2>&1
  echo "exit $?" >> $OUT/sweep_A.log
fi
if has ncu; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > $OUT/ncu_launches_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scatter_kernel -s 4 -c 4 -f -o $OUT/prof_scatter \
      python tools/sweep.py --what none --reps 1 > $OUT/ncu_scatter.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"join_kernel|hist_kernel" -s 3 -c 3 -f -o $OUT/prof_join_hist \
      python tools/sweep.py --what none --reps 1 > $OUT/ncu_join.log 2>&1
fi
ls -la $OUT > $OUT/listing.txt
echo done
