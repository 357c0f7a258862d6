#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2q; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_join.py -m gpu -q --timeout 600 -p no:cacheprovider -k "pcp" -x > $OUT/pytest_pcp.log 2>&1
echo "exit $?" >> $OUT/pytest_pcp.log; tail -30 $OUT/pytest_pcp.log
