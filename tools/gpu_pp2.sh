#!/bin/bash
# N-GPU check of the sharded pipeline: parity on real peers, then bench lines per shuffle mode.
# usage: tools/gpu_pp2.sh N "modes" [extra bench args]
cd "$(dirname "$0")/.."
N=${1:-2}; MODES=${2:-"pp"}; shift 2
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/pp_topo$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 -p no:cacheprovider -x > $OUT/pp_pytest_multi$N.log 2>&1; echo "exit $?" >> $OUT/pp_pytest_multi$N.log
tail -3 $OUT/pp_pytest_multi$N.log
for m in $MODES; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 8 --warmup 3 --shuffle $m "$@" > $OUT/pp_bench_${N}gpu_$m.log 2>&1
  echo "exit $?" >> $OUT/pp_bench_${N}gpu_$m.log
  tail -2 $OUT/pp_bench_${N}gpu_$m.log | cut -c1-3000
done
echo done
