#!/bin/bash
# Copy the judged summaries of the latest gpurun into profiles/<name>/ (text only; .ncu-rep stay in gpurun_out/).
set -e
cd "$(dirname "$0")/.."
D=profiles/$1
mkdir -p $D
G=gpurun_out
for f in bench.log bench_reference.log sweep.log sweep_A.log pytest_gpu.log pytest_full.log pytest_multi.log pytest_multi8.log memcheck.log racecheck.log smoke.log driver_small.log \
         bench_2gpu_p2p.log bench_4gpu_p2p.log bench_8gpu_p2p.log bench_8gpu_nccl.log bench_2gpu_nccl.log topo8.txt nvidia-smi.txt host.txt; do
  [ -f $G/$f ] && cp $G/$f $D/ || true
done
[ -f $G/launches.csv ] && cp $G/launches.csv $D/ncu_launches_bench_steps2.csv
for k in scatter join_hist; do
  [ -f $G/prof_$k.ncu-rep ] && ncu -i $G/prof_$k.ncu-rep --page raw --csv 2>/dev/null > $D/ncu_raw_$k.csv || true
done
if [ -f $G/prof_scatter.ncu-rep ]; then
  python tools/ncu_source_summary.py $G/prof_scatter.ncu-rep 'scatter_kernel.*\(bool\)1' 0 > $D/ncu_source_scatter_pass1.txt || true
  python tools/ncu_source_summary.py $G/prof_scatter.ncu-rep 'scatter_kernel.*\(bool\)0' 0 > $D/ncu_source_scatter_pass2.txt || true
fi
if [ -f $G/prof_join_hist.ncu-rep ]; then
  python tools/ncu_source_summary.py $G/prof_join_hist.ncu-rep join_kernel 0 > $D/ncu_source_join.txt || true
  python tools/ncu_source_summary.py $G/prof_join_hist.ncu-rep hist_kernel 0 > $D/ncu_source_hist.txt || true
fi
ls $D
