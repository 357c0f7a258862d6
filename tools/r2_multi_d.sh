#!/bin/bash
# N GPUs: pcp virtual-shard tests, real multi-GPU parity, then bench variants given as arguments "name:args" ...
N=${1:-2}; shift
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2m_${N}d; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_join.py -m gpu -q --timeout 600 -p no:cacheprovider -k "pcp" -x > $OUT/pytest_pcp.log 2>&1
echo "exit $?" >> $OUT/pytest_pcp.log; tail -4 $OUT/pytest_pcp.log
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 800 -p no:cacheprovider -k "pcp" > $OUT/pytest_multi.log 2>&1
echo "exit $?" >> $OUT/pytest_multi.log; tail -4 $OUT/pytest_multi.log
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
    print(sys.argv[2], "NO LINE:", open(sys.argv[1]).read()[-1200:])
PY
}
for spec in "$@"; do
  name=${spec%%:*}; args=${spec#*:}
  if [ "$name" = trace ]; then GJ_TRACE=1 run $name $args; else run $name $args; fi
done
echo done
