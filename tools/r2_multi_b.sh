#!/bin/bash
# Round 2, multi-GPU run B (N GPUs): streamed pcp exchange, sweep around the defaults (copy CTAs, stages), NVLink
# byte counters around one run, then the full default bench line (with the config-5 sub-record).
N=${1:-2}; MODE=${2:-sweep}
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2m_${N}b; mkdir -p $OUT
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
if [ "$MODE" = sweep ]; then
  run default --no-cfg5
  GJ_TRACE=1 run trace --no-cfg5
  for st in 1,1 1,2 1,4 2,2 2,8 4,4; do run st_$st --no-cfg5 --pcp-stages $st; done
  for c in 12 16 32 48; do run ctas_$c --no-cfg5 --opt pcp_copy_ctas=$c; done
  run ctas_148 --no-cfg5 --opt pcp_copy_ctas=0
fi
nvidia-smi nvlink -gt d > $OUT/nvlink_before.txt 2>&1
run full
nvidia-smi nvlink -gt d > $OUT/nvlink_after.txt 2>&1
echo done
