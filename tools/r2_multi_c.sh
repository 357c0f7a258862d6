#!/bin/bash
# Round 2, 8-GPU run: parity of every exchange mode on 8 real GPUs, streamed pcp (defaults, timeline, a few shapes),
# NVLink byte counters around the full default line (with the config-5 sub-record).
N=${1:-8}
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2m_${N}c; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 800 -p no:cacheprovider > $OUT/pytest_multi.log 2>&1
echo "exit $?" >> $OUT/pytest_multi.log; tail -5 $OUT/pytest_multi.log
run() {   # name, extra args
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 10 --warmup 3 "$@" > $OUT/bench_$name.log 2>&1
  echo "exit $?" >> $OUT/bench_$name.log
  python - $OUT/bench_$name.log $name <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l)
        ph = d["roofline"].get("local_phases_ms") or {}
        print(sys.argv[2], round(d["value"] / 1e9, 1), "G/s", round(d["ms_per_step"], 3), "ms", {k: round(v, 2) for k, v in ph.items()},
              "nvlink", round(d["shuffle"].get("nvlink_out_GBs_per_gpu") or 0), d["shuffle"].get("trace_ms_rank0"),
              "cfg5", (d.get("config5") or {}).get("ms_per_step"), (d.get("config5") or {}).get("speedup_vs_1gpu"))
        break
else:
    print(sys.argv[2], "NO LINE:", open(sys.argv[1]).read()[-900:])
PY
}
run default --no-cfg5
GJ_TRACE=1 run trace --no-cfg5
run st_2,2 --no-cfg5 --pcp-stages 2,2
run st_4,8 --no-cfg5 --pcp-stages 4,8
run ctas_32 --no-cfg5 --opt pcp_copy_ctas=32
run ctas_48 --no-cfg5 --opt pcp_copy_ctas=48
run ctas_148 --no-cfg5 --opt pcp_copy_ctas=0
nvidia-smi nvlink -gt d -i 0 > $OUT/nvlink_before.txt 2>&1
run full
nvidia-smi nvlink -gt d -i 0 > $OUT/nvlink_after.txt 2>&1
run cfg5_st48 --workload cfg5 --pcp-stages 4,8 --steps 3
echo done
