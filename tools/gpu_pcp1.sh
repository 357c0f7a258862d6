#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_join.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "pcp" > $OUT/pcp1_pytest.log 2>&1; echo "exit $?" >> $OUT/pcp1_pytest.log
tail -25 $OUT/pcp1_pytest.log
echo done
