#!/bin/bash
# Bench lines on N GPUs for a list of "mode:opts" specs (opts = comma separated: name=value engine options,
# no-overlap, workload=W, steps=K).  usage: gpurun --gpus N -- 'bash tools/gpu_multi_bench.sh N "pcp:" "p2p:steps=5,workload=cfg5"'
cd "$(dirname "$0")/.."
N=${1:-2}; shift
OUT=gpurun_out; mkdir -p $OUT
i=0
for spec in "$@"; do
  i=$((i+1))
  m=${spec%%:*}; o=${spec#*:}
  args=""
  IFS=',' read -ra KV <<< "$o"
  for kv in "${KV[@]}"; do
    case "$kv" in
      "") ;;
      no-overlap) args="$args --no-overlap" ;;
      workload=*) args="$args --workload ${kv#workload=}" ;;
      steps=*) args="$args --steps ${kv#steps=}" ;;
      *) args="$args --opt $kv" ;;
    esac
  done
  f=$OUT/mb_${N}gpu_${i}_$m.log
  echo "### $spec" > $f
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$i \
      bench.py --gpus $N --steps 6 --warmup 3 --shuffle $m $args >> $f 2>&1
  echo "exit $?" >> $f
  python - "$f" "$spec" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l)
        print(sys.argv[2], "| G tuples/s", round(d["value"] / 1e9, 1), "| ms", round(d["ms_per_step"], 3), "|", d["shuffle"].get("scatter_kernel_ms"),
              d["shuffle"].get("nvlink_out_GBs_per_gpu"), d["roofline"].get("local_phases_ms"), d["shuffle"].get("host_ms_rank0"))
PY
done
echo done
