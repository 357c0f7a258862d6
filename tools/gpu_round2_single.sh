#!/bin/bash
# Round-2 queue, one GPU: first run of everything written after round 1's GPU budget was spent.
#   1. the late-materialisation join (gj_join_aggregate_late) and the non-partitioned baseline
#      (gj_join_aggregate_nopart), tests gated by GJ_RUN_UNVERIFIED; then the crossover sweep
#   2. the whole GPU suite + smoke + the default bench (no regression)
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
GJ_RUN_UNVERIFIED=1 timeout 900 python -m pytest tests/test_gpu_unverified.py -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/r2_pytest_late.log 2>&1
echo "exit $?" >> $OUT/r2_pytest_late.log; tail -5 $OUT/r2_pytest_late.log
timeout 600 python tools/nopart_crossover.py > $OUT/r2_nopart_crossover.log 2>&1; echo "exit $?" >> $OUT/r2_nopart_crossover.log
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/r2_pytest_gpu.log 2>&1
echo "exit $?" >> $OUT/r2_pytest_gpu.log; tail -3 $OUT/r2_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2_smoke.log 2>&1; echo "exit $?" >> $OUT/r2_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/r2_bench.log 2>&1; echo "exit $?" >> $OUT/r2_bench.log
tail -2 $OUT/r2_bench.log | cut -c1-600
echo done
