#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_join.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "pp_ or virtual or peer_store" > $OUT/pp4_pytest.log 2>&1; echo "exit $?" >> $OUT/pp4_pytest.log
tail -3 $OUT/pp4_pytest.log
bash tools/gpu_pp3.sh 2 "pp:" "pp:pp_tile16k=1" "pp:pp_tile16k=1,pass1_bits=10" "pp:pass1_bits=10" "p2p:"
