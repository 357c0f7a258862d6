#!/bin/bash
# Round 2, run A (1 GPU): first execution of the SURVEY 8f rows (late materialisation, non-partitioned
# baselines incl. perfect array, streamed probe side), the warp-aggregated materialising join, the whole
# GPU suite, the non-partitioned crossover sweep, and the default bench.
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2a; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_next_rows.py -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/pytest_next_rows.log 2>&1
echo "exit $?" >> $OUT/pytest_next_rows.log; tail -15 $OUT/pytest_next_rows.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider --deselect tests/test_gpu_next_rows.py > $OUT/pytest_gpu.log 2>&1
echo "exit $?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
timeout 600 python tools/nopart_crossover.py --max-log2 26 > $OUT/nopart_crossover.log 2>&1; echo "exit $?" >> $OUT/nopart_crossover.log
tail -12 $OUT/nopart_crossover.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-ref-cuda > $OUT/bench.log 2>&1; echo "exit $?" >> $OUT/bench.log
tail -2 $OUT/bench.log | cut -c1-900
echo done
