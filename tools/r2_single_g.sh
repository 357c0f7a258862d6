#!/bin/bash
# 1 GPU: join tests + workloads B, A, Zipf 1.0 (short) after a join-kernel change; optional racecheck.
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2g; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_join.py tests/test_gpu_next_rows.py tests/test_gpu_fullsize.py -m gpu -q --timeout 900 -p no:cacheprovider -x > $OUT/pytest.log 2>&1
echo "exit $?" >> $OUT/pytest.log; tail -4 $OUT/pytest.log
for w in B A zipf1.0; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-ref-cuda --no-cfg5 --no-cpu-baseline > $OUT/bench_$w.log 2>&1; echo "exit $?" >> $OUT/bench_$w.log
  python - $OUT/bench_$w.log $w <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); r = d["roofline"]
        print(sys.argv[2], round(d["value"] / 1e9, 2), "G/s", round(d["ms_per_step"], 3), "ms", {k: round(v, 3) for k, v in r["per_phase_ms"].items()},
              "mat", round((d.get("materialize") or {}).get("ms_per_step", 0), 3), round((d.get("materialize") or {}).get("join_ms", 0), 3))
        break
else:
    print(sys.argv[2], "NO LINE", open(sys.argv[1]).read()[-500:])
PY
done
B=icde2019-gpu-join_b200/bin/bench
(cd /tmp && timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 $OLDPWD/$B -b 7 -a HJC -R 2200000 -S 2400000 -s 1.0 --payload rowid) > $OUT/racecheck.log 2>&1; echo "exit $?" >> $OUT/racecheck.log
grep -c "Race reported" $OUT/racecheck.log; tail -2 $OUT/racecheck.log
echo done
