#!/bin/bash
# Round-2 queue for the multi-GPU exchange (what round 1 could not measure: GPU budget).
# usage: gpurun --gpus N -- 'bash tools/gpu_round2_multi.sh N'     (N = 2, 4 or 8)
#   1. parity of every exchange mode on N real GPUs
#   2. pcp / pp / p2p on workload B (weak) and, at N = 8, config 5 (strong)
#   3. why kernels running under the pcp copy kernel are slowed 2.6-3.8x at 8 GPUs (1.5x at 2 GPUs)
#      (source pass of S 1.25 ms instead of 0.48; receiver pass of R 2.26 ms instead of ~0.6).  Hypotheses and
#      the knob that tests each:
#        H1 the copy's streaming traffic + incoming peer writes evict the scatter's partially written lines
#           from L2 (the 8-byte-store scatter relies on L2 merging)            -> pcp_l2_hint=1
#        H2 occupancy: the 512x16 scatter variants fill the register file exactly (2 CTAs x 512 threads x 64
#           registers = 64 K), so ONE resident copy CTA per SM halves their occupancy (the hist kernel's 128 KB
#           of shared memory still fits next to the copy's 68 KB)              -> shuffle_grid=74 / 32 / 16,
#           with pcp_ring=1 (12-slot ring, 10 loads in flight per CTA) when the copy runs on few SMs
#        H3 page-walk contention: a 512-way scatter touches 512 x 2 MB pages (TLB reach 128 pages) while the
#           NVLink ingress of 7 peers translates too                           -> pass1_bits=8 / 10 (256- / 1024-way)
#      HBM bandwidth itself is not the limit: copy traffic is ~1.5 TB/s of 6.5.
cd "$(dirname "$0")/.."
N=${1:-8}
OUT=gpurun_out; mkdir -p $OUT
GJ_RUN_UNVERIFIED=1 timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/r2_pytest_multi$N.log 2>&1
echo "exit $?" >> $OUT/r2_pytest_multi$N.log; tail -3 $OUT/r2_pytest_multi$N.log
export GJ_TRACE=1   # per-stage GPU timeline of pcp in every bench line (shuffle.trace_ms_rank0)
SPECS=("pcp:steps=5" "pcp2:steps=5" "pcp:steps=5,pcp_l2_hint=1" "pcp:steps=5,shuffle_grid=74" "pcp:steps=5,shuffle_grid=32,pcp_ring=1" "pcp:steps=5,shuffle_grid=16,pcp_ring=1" "pcp:steps=5,shuffle_grid=296" "pcp:steps=5,pass1_bits=10" "pp:steps=5" "p2p:steps=5")
if [ "$N" = "8" ]; then SPECS+=("pcp:steps=5,workload=cfg5" "pcp2:steps=5,workload=cfg5" "pcp:steps=5,workload=cfg5,pass1_bits=9" "p2p:steps=5,workload=cfg5"); fi
bash tools/gpu_multi_bench.sh $N "${SPECS[@]}"
