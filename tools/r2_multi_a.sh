#!/bin/bash
# Round 2, multi-GPU run A (N GPUs, default 2): parity of every exchange mode on real GPUs, then the streamed
# pcp exchange under different copy-kernel shapes (CTAs x ring depth) and stage counts.
N=${1:-2}
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2m_${N}; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 800 -p no:cacheprovider > $OUT/pytest_multi.log 2>&1
echo "exit $?" >> $OUT/pytest_multi.log; tail -5 $OUT/pytest_multi.log
run() {   # name, extra args
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 10 --warmup 3 --no-cfg5 "$@" > $OUT/bench_$name.log 2>&1
  echo "exit $?" >> $OUT/bench_$name.log
  python - $OUT/bench_$name.log $name <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l)
        print(sys.argv[2], round(d["value"] / 1e9, 1), "G/s", round(d["ms_per_step"], 3), "ms", d["roofline"].get("local_phases_ms"),
              "nvlink", d["shuffle"].get("nvlink_out_GBs_per_gpu"), d["shuffle"].get("trace_ms_rank0"))
        break
else:
    print(sys.argv[2], "NO LINE:", open(sys.argv[1]).read()[-600:])
PY
}
run default || true
GJ_TRACE=1 run trace
run st11 --pcp-stages 1,1
run st28 --pcp-stages 2,8
run st48 --pcp-stages 4,8
run g24r --opt shuffle_grid=24 --opt pcp_ring=1
run g16r --opt shuffle_grid=16 --opt pcp_ring=1
run g32r --opt shuffle_grid=32 --opt pcp_ring=1
run g48r --opt shuffle_grid=48 --opt pcp_ring=1
run g24r_st48 --opt shuffle_grid=24 --opt pcp_ring=1 --pcp-stages 4,8
run g74 --opt shuffle_grid=74
echo done
