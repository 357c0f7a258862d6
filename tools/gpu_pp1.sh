#!/bin/bash
# 1-GPU check of the sharded "partition, then push" pipeline: parity tests, kernel timings, no-regression bench.
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_join.py -m gpu -q --timeout 600 -p no:cacheprovider -x > $OUT/pp1_pytest.log 2>&1; echo "exit $?" >> $OUT/pp1_pytest.log
timeout 600 python tools/pp_probe.py > $OUT/pp1_probe.log 2>&1; echo "exit $?" >> $OUT/pp1_probe.log
timeout 600 python tools/pp_probe.py --n 250000000 --B 16 --reps 2 > $OUT/pp1_probe_cfg5.log 2>&1; echo "exit $?" >> $OUT/pp1_probe_cfg5.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda > $OUT/pp1_bench.log 2>&1; echo "exit $?" >> $OUT/pp1_bench.log
tail -3 $OUT/pp1_pytest.log; tail -4 $OUT/pp1_probe.log
echo done
