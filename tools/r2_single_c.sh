#!/bin/bash
# Round 2, run C (1 GPU): changed tests, boundary proof, crossover with the perfect array, the 8f rows at size, smoke, full bench.
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2c; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "exit $?" >> $OUT/pytest_gpu.log; tail -8 $OUT/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "exit $?" >> $OUT/smoke.log; tail -2 $OUT/smoke.log
timeout 600 python tools/nopart_crossover.py --max-log2 24 > $OUT/nopart_crossover.log 2>&1; echo "exit $?" >> $OUT/nopart_crossover.log
tail -10 $OUT/nopart_crossover.log | cut -c1-330
timeout 900 python tools/next_rows_bench.py all > $OUT/next_rows_bench.log 2>&1; echo "exit $?" >> $OUT/next_rows_bench.log
tail -8 $OUT/next_rows_bench.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.log 2>&1; echo "exit $?" >> $OUT/bench.log
tail -2 $OUT/bench.log | cut -c1-3000
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/bench_reference.log 2>&1; echo "exit $?" >> $OUT/bench_reference.log
tail -2 $OUT/bench_reference.log | cut -c1-1200
timeout 300 python bench.py --workload small --steps 20 --warmup 5 --no-cfg5 > $OUT/bench_small.log 2>&1; echo "exit $?" >> $OUT/bench_small.log
tail -2 $OUT/bench_small.log | cut -c1-1500
echo done
