#!/bin/bash
# Round 2, last 8-GPU run: parity of the peer-histogram variant on 8 real GPUs, the full default line, the same with
# --pcp-peer-hist, and config 5 with a 512-way source pass (pass1_bits = 9).
N=${1:-8}
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2m_${N}f; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 500 -p no:cacheprovider -k "peer-hist" > $OUT/pytest_multi.log 2>&1
echo "exit $?" >> $OUT/pytest_multi.log; tail -3 $OUT/pytest_multi.log
run() {   # name, extra args
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 10 --warmup 3 "$@" > $OUT/bench_$name.log 2>&1
  echo "exit $?" >> $OUT/bench_$name.log
  python - $OUT/bench_$name.log $name <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l)
        ph = d["roofline"].get("local_phases_ms") or {}
        print(sys.argv[2], round(d["value"] / 1e9, 1), "G/s", round(d["ms_per_step"], 3), "ms", {k: round(v, 2) for k, v in ph.items()},
              "nvlink", round(d["shuffle"].get("nvlink_out_GBs_per_gpu") or 0),
              "cfg5", (d.get("config5") or {}).get("ms_per_step"), (d.get("config5") or {}).get("speedup_vs_1gpu"))
        break
else:
    print(sys.argv[2], "NO LINE:", open(sys.argv[1]).read()[-1200:])
PY
}
run full
run peerhist --pcp-peer-hist
run cfg5_p9 --workload cfg5 --opt pass1_bits=9 --steps 3
run cfg5_p9_peerhist --workload cfg5 --opt pass1_bits=9 --steps 3 --pcp-peer-hist
echo done
