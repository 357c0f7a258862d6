#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "pcp" > $OUT/pcp_pytest_multi$N.log 2>&1; echo "exit $?" >> $OUT/pcp_pytest_multi$N.log
tail -5 $OUT/pcp_pytest_multi$N.log
shift
bash tools/gpu_pp3.sh $N "$@"
