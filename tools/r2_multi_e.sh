#!/bin/bash
# Round 2, 2-GPU run E: skewed multi-GPU test + pcp parity, the exchange kernels of one process driving two GPUs
# under ncu (pcp_copy_kernel with NVLink counters), the 2-GPU bench line with the config-5 record.
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2m_2e; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 800 -p no:cacheprovider -k "skew or pcp" > $OUT/pytest_multi.log 2>&1
echo "exit $?" >> $OUT/pytest_multi.log; tail -6 $OUT/pytest_multi.log
timeout 300 python tools/ncu_pcp_two_devices.py > $OUT/two_devices_check.log 2>&1; echo "exit $?" >> $OUT/two_devices_check.log; tail -4 $OUT/two_devices_check.log
ncu --query-metrics 2>/dev/null | grep -o -i "^nvl[rt]x__bytes[a-z_]*" | sort -u > $OUT/nvl_metric_names.txt
NVL=$(grep -E "^nvl(rx|tx)__bytes(_data_user)?$" $OUT/nvl_metric_names.txt | sed 's/$/.sum/' | paste -sd, -)
echo "nvlink metrics: $NVL"
timeout 600 ncu --kernel-name regex:pcp_copy_kernel --launch-count 4 --clock-control none \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_aperture_peer.sum,lts__t_sectors_aperture_peer_op_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active${NVL:+,$NVL} \
    --csv --log-file $OUT/ncu_pcp_copy_metrics.csv python tools/ncu_pcp_two_devices.py --copy-only --no-check > $OUT/ncu_pcp_copy.log 2>&1
echo "exit $?" >> $OUT/ncu_pcp_copy.log; tail -3 $OUT/ncu_pcp_copy.log; head -30 $OUT/ncu_pcp_copy_metrics.csv | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:pcp_copy_kernel --launch-count 1 -f -o $OUT/prof_pcp_copy \
    python tools/ncu_pcp_two_devices.py --copy-only --no-check > $OUT/ncu_pcp_copy_full.log 2>&1
echo "exit $?" >> $OUT/ncu_pcp_copy_full.log
ncu -i $OUT/prof_pcp_copy.ncu-rep --page raw --csv > $OUT/ncu_raw_pcp_copy.csv 2>/dev/null
rm -f $OUT/prof_pcp_copy.ncu-rep
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_full.log 2>&1; echo "exit $?" >> $OUT/bench_full.log
tail -2 $OUT/bench_full.log | cut -c1-600
echo done
