#!/bin/bash
# Round 2, run B (1 GPU): the streamed pcp exchange with virtual ranks on one GPU, then the whole suite.
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2b; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_join.py -m gpu -q --timeout 600 -p no:cacheprovider -k "pcp" -x > $OUT/pytest_pcp.log 2>&1
echo "exit $?" >> $OUT/pytest_pcp.log; tail -30 $OUT/pytest_pcp.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "exit $?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
echo done
