/*
 * gpujoin.h -- C ABI of libgpujoin.so: a B200-native (sm_100a) radix hash-join engine.
 *
 * Drop-in boundary for the reference's in-GPU join path (psiul/ICDE2019-GPU-Join).  Every entry
 * point names the reference interface it replaces (file:line into the reference's src/).
 * Plain pointers and sizes only; no C++/torch types.  All functions return 0 on success or a
 * negative gj_status; gj_last_error() then describes the failure.  Nothing in here ever calls
 * exit() (the reference's CHK_ERROR does, common.h:132-141) and there is no CPU fallback: when
 * no CUDA device is usable every compute entry point fails with GJ_ERR_CUDA.
 *
 * Data model (reference: hash_join_clustered_probe.cu:802, join-primitives.cu:1582):
 *   a relation is two device (or host, *_host variants) arrays of n int32: keys and payloads
 *   ("columnar", as the reference passes R/Pr and S/Ps).  n < 2^32 - 2^20 per relation.
 *   Join = inner equi-join on the full 32-bit key, all pairs (N:M allowed).
 *   Aggregate = { matches, checksum } with checksum = SUM over result pairs of
 *   (int64)Pr*(int64)Ps mod 2^64; its low 32 bits are the int32 the reference prints as
 *   "%d results" (hash_join_clustered_probe.cu:984-986, join-primitives.cu:1073,1092).
 */
#ifndef GPUJOIN_H
#define GPUJOIN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GJ_VERSION 1

typedef enum gj_status {
    GJ_OK = 0,
    GJ_ERR_ARG = -1,      /* bad argument (null pointer, size over capacity, bad option) */
    GJ_ERR_CUDA = -2,     /* a CUDA runtime call failed / no device */
    GJ_ERR_NOMEM = -3,    /* device or host allocation failed */
    GJ_ERR_STATE = -4     /* call sequence error (e.g. result queried before a join) */
} gj_status;

typedef struct gj_ctx gj_ctx;

/* Device-side timings of the last call, CUDA events on the engine's stream (milliseconds).
 * Same window as the reference's t1..t2 (hash_join_clustered_probe.cu:954-980): inputs resident
 * in HBM, no allocation inside.  h2d_ms/d2h_ms are only filled by the *_host entry points. */
typedef struct gj_timings {
    float hist_ms;       /* radix histograms of R and S + offset scan + work planning */
    float part_ms;       /* all radix scatter passes of R and S */
    float join_ms;       /* per-partition build + probe + final reduction */
    float total_ms;      /* hist + part + join, one event pair around everything */
    float h2d_ms;        /* host->device copies (host entry points), else 0 */
    float wall_ms;       /* host steady_clock around the whole call, including the final sync */
    float pass_ms[4];    /* each scatter launch: build pass 1, build pass 2, probe pass 1, probe pass 2
                            (pass-2 slots are 0 for a single-pass plan) */
    uint32_t radix_bits;     /* total radix bits B used (2^B partitions) */
    uint32_t pass1_bits;     /* bits of the first pass (== radix_bits when single pass) */
    uint32_t pass2_bits;     /* bits of the second pass (0 when single pass) */
    uint32_t pass3_bits;     /* bits of the third pass (0 unless the build side exceeds 2^28 tuples) */
    uint32_t kernel_launches;/* kernels launched inside the timed window */
} gj_timings;

/* ---- lifetime -------------------------------------------------------------------------------
 * Replaces the 16+ cudaMalloc calls the reference issues per join and never frees
 * (hash_join_clustered_probe.cu:832-872): all scratch is allocated once, here, for relations of
 * up to max_R / max_S tuples, and released by gj_destroy. */
int gj_create(gj_ctx** out, int device, uint64_t max_R, uint64_t max_S);
void gj_destroy(gj_ctx* ctx);
const char* gj_last_error(void);
int gj_version(void);

/* Run on a caller-owned cudaStream_t (passed as void*); NULL = the engine's own stream.
 * Reference: the single explicit stream streams_R[1], hash_join_clustered_probe.cu:809-811. */
int gj_set_stream(gj_ctx* ctx, void* cuda_stream);

/* Tuning knobs; the reference fixes these at compile time (common.h:51-71).  Names:
 *   "radix_bits"   total radix bits B, 0 = choose from |build side| (default)
 *   "pass1_bits"   bits of the first pass, 0 = choose
 *   "scatter_cfg"  scatter kernel shape variant (0 = default)
 *   "unit_tuples"  probe-side work-unit size in tuples (skew splitting), 0 = default
 *   "scatter_cfg1"/"scatter_cfg2" per-pass variants; 255 = measured best per fan-out (default)
 *   "join_cfg"     join kernel shape variant (0 = default)
 *   "gpu_bits"     number of key bits above the radix field consumed by the multi-GPU shuffle
 *   "shuffle_grid" persistent CTA count of the peer-store scatter (0 = one tile per CTA)
 *   "pp_out", "pp_tile16k"  sharded pipeline, see gj_pp_begin
 *   "pcp_copy_ctas" CTAs of the pcp copy kernel (default 24).  Up to half the SMs: deep ring (12 slots, 10 bulk loads in
 *                  flight per CTA, 196 KB of shared memory), the CTAs fill their SMs and the kernels next to the copy
 *                  size their grids to the remaining SMs; more, or 0 = one per SM: 4-slot ring sharing the SMs
 *   "pcp_timeout_ms" bound of the receiver's wait for a peer's stage flag (default 5000)
 *   "nopart_max"   gj_join_aggregate takes the non-partitioned path (gj_join_aggregate_nopart) when the
 *                  smaller relation has at most this many tuples and no radix plan is forced (default 2^21, from
 *                  the measured crossover, tools/nopart_crossover.py / profiles/r2_a; 0 = never) */
int gj_set_option(gj_ctx* ctx, const char* name, int64_t value);
int gj_get_option(gj_ctx* ctx, const char* name, int64_t* value);

/* ---- the operator --------------------------------------------------------------------------
 * gj_join_aggregate replaces outOfGPU_Join1_payload's aggregate run
 * (hash_join_clustered_probe.cu:944-991: prepare_Relation_payload x2, decompose_chains,
 * join_partitioned_aggregate) with inputs already on the device.  d_* are device pointers. */
int gj_join_aggregate(gj_ctx* ctx, const int32_t* d_Rk, const int32_t* d_Rp, uint64_t nR,
                      const int32_t* d_Sk, const int32_t* d_Sp, uint64_t nS,
                      uint64_t* matches, uint64_t* checksum, gj_timings* t);

/* Same join over packed device tuples {key,payload} (int32 pairs, 8-byte aligned), the layout
 * the multi-GPU shuffle delivers. */
int gj_join_aggregate_tuples(gj_ctx* ctx, const void* d_Rtup, uint64_t nR, const void* d_Stup,
                             uint64_t nS, uint64_t* matches, uint64_t* checksum, gj_timings* t);

/* End-to-end variant: h_* are HOST pointers (pinned or pageable); the H2D copies are part of the
 * call, chunked and overlapped with the radix histograms.  Replaces the cudaMemcpy H2D block +
 * aggregate run of outOfGPU_Join1_payload (hash_join_clustered_probe.cu:874-877, 944-991). */
int gj_join_aggregate_host(gj_ctx* ctx, const int32_t* h_Rk, const int32_t* h_Rp, uint64_t nR,
                           const int32_t* h_Sk, const int32_t* h_Sp, uint64_t nS,
                           uint64_t* matches, uint64_t* checksum, gj_timings* t);

/* Out-of-HBM probe side (replaces outOfGPU_Join3_payload,
 * hash_join_clustered_probe.cu:1684-1984, the reference's PCIe-streaming mode): R (the build side,
 * nR <= max_R) is copied and partitioned once and stays resident; S streams from HOST memory in chunks of
 * chunk_tuples (<= max_S / 2: double buffer), the copy of chunk i+1 running under the partitioning and
 * join of chunk i.  nS is unbounded.  timings: hist_ms = build side (copy + partition), join_ms = all
 * probe chunks, h2d_ms = until the last byte arrived. */
int gj_join_aggregate_stream_host(gj_ctx* ctx, const int32_t* h_Rk, const int32_t* h_Rp, uint64_t nR,
                                  const int32_t* h_Sk, const int32_t* h_Sp, uint64_t nS,
                                  uint64_t chunk_tuples, uint64_t* matches, uint64_t* checksum,
                                  gj_timings* t);

/* Materialising join: replaces join_partitioned_results (join-primitives.cu:1107-1416) and the
 * first run of outOfGPU_Join1_payload (hash_join_clustered_probe.cu:883-940).  Writes result
 * pairs (Pr, Ps) to the device columns d_out_Rp/d_out_Sp in unspecified order; at most `cap`
 * pairs are written, *n_pairs is always the exact result size (the reference's ring buffer
 * overwrites itself instead, join-primitives.cu:1097-1099).  matches/checksum as above. */
int gj_join_materialize(gj_ctx* ctx, const int32_t* d_Rk, const int32_t* d_Rp, uint64_t nR,
                        const int32_t* d_Sk, const int32_t* d_Sp, uint64_t nS, int32_t* d_out_Rp,
                        int32_t* d_out_Sp, uint64_t cap, uint64_t* n_pairs, uint64_t* checksum,
                        gj_timings* t);

/* Late-materialisation join (replaces join_partitioned_varpayload,
 * join-primitives.cu:1420-1557, and its driver outOfGPU_Join_payload_var,
 * hash_join_clustered_probe.cu:542-708): the payload columns d_Rid / d_Sid hold ROW IDS into
 * column-major side tables, value of column z for row id i = d_Dr[z * stride_r + i] (the reference's
 * Dr[pval + z*rel_size], :1531-1536).  Every result pair adds the cols_r values of its R row and the
 * cols_s values of its S row to the aggregate: *sum = that total as int64 mod 2^64, whose low 32 bits are
 * the int32 the reference accumulates (:1460, 1552).  Row ids must lie in [0, stride); at most 64 columns
 * per side.  Same pipeline as gj_join_aggregate; the gathers hit L2 when the side tables fit its 126 MB. */
int gj_join_aggregate_late(gj_ctx* ctx, const int32_t* d_Rk, const int32_t* d_Rid, uint64_t nR,
                           const int32_t* d_Sk, const int32_t* d_Sid, uint64_t nS,
                           const int32_t* d_Dr, uint32_t cols_r, uint64_t stride_r,
                           const int32_t* d_Ds, uint32_t cols_s, uint64_t stride_s,
                           uint64_t* matches, uint64_t* sum, gj_timings* t);

/* Non-partitioned baseline (replaces build_ht_chains / chains_probing,
 * join-primitives.cu:681-742): one chained hash table over the whole build side in global memory, probed
 * straight from the probe columns -- no radix pass, 3 launches.  Same result as gj_join_aggregate.  The
 * comparison point for the partitioned path: on B200 a build side whose table (8 B of heads + 4 B of links
 * + 8 B of columns per tuple) stays in the 126 MB L2 is the regime where it wins (config 1).
 * timings: hist_ms = table clear + build, join_ms = probe. */
int gj_join_aggregate_nopart(gj_ctx* ctx, const int32_t* d_Rk, const int32_t* d_Rp, uint64_t nR,
                             const int32_t* d_Sk, const int32_t* d_Sp, uint64_t nS,
                             uint64_t* matches, uint64_t* checksum, gj_timings* t);

/* Perfect-array variant of the non-partitioned baseline (replaces build_perfect_array /
 * probe_perfect_array, join-primitives.cu:628-668): the key itself addresses a table of key_range slots,
 * slot = key - key_min; no hash, no chains, no key compare.  Precondition (checked on the device, GJ_ERR_ARG
 * when violated): the keys of the build side (the smaller relation; R when equal) are unique and lie in
 * [key_min, key_min + key_range).  A slot is a 64-bit word {1, payload}, so every int32 payload survives
 * (the reference stores payload + 1 in an int32, which drops the payload -1).  timings as for _nopart. */
int gj_join_aggregate_perfect(gj_ctx* ctx, const int32_t* d_Rk, const int32_t* d_Rp, uint64_t nR,
                              const int32_t* d_Sk, const int32_t* d_Sp, uint64_t nS,
                              int32_t key_min, uint64_t key_range,
                              uint64_t* matches, uint64_t* checksum, gj_timings* t);

/* ---- the partitioner on its own --------------------------------------------------------------
 * Replaces prepare_Relation_payload (join-primitives.cu:1582-1613: init_metadata_double,
 * partition_pass_one, compute_bucket_info, partition_pass_two).  Partitions one relation on its
 * low `radix_bits` key bits (identity hash, first_bit 0, like common.h:45-47 /
 * hash_join_clustered_probe.cu:813) into 2^radix_bits CONTIGUOUS partitions (no bucket chains).
 * radix_bits = 0 lets the engine choose from n.  Results stay in the context until the next call
 * on the same slot (slot 0 = build side buffers, 1 = probe side buffers):
 *   *d_tuples   device pointer to n packed {key,payload} pairs, partition p occupying
 *               [offsets[p], offsets[p+1])
 *   *d_offsets  device pointer to 2^B + 1 uint32 offsets
 * Order inside a partition is unspecified. */
int gj_partition(gj_ctx* ctx, int slot, const int32_t* d_keys, const int32_t* d_pays, uint64_t n,
                 uint32_t radix_bits, const void** d_tuples, const uint32_t** d_offsets,
                 uint32_t* radix_bits_used, gj_timings* t);

/* ---- multi-GPU shuffle step (no reference counterpart) ----------------
 * Splits one relation by destination GPU d = (key >> gpu_shift) & (n_gpus-1) (n_gpus a power of
 * two <= 256).  Output: packed tuples grouped by destination in d_out_tuples (n tuples) and the
 * per-destination counts in h_counts[n_gpus] (host).  The caller exchanges counts and ships the
 * groups (NCCL all-to-all or peer stores), then runs gj_join_aggregate_tuples on what it
 * received with option "gpu_bits" = log2(n_gpus). */
int gj_shuffle_split(gj_ctx* ctx, const int32_t* d_keys, const int32_t* d_pays, uint64_t n,
                     uint32_t n_gpus, uint32_t gpu_shift, void* d_out_tuples, uint64_t* h_counts);

/* Peer-store variant: d_peer_bases[g] is a device pointer (local, or a peer / IPC mapping over
 * NVLink) to GPU g's receive buffer of packed tuples and h_peer_offsets[g] the first tuple slot
 * this GPU may write there.  The scatter kernel stores straight into the destinations from its
 * shared-memory staging, so partitioning and the all-to-all are one kernel.  Counts as above. */
int gj_shuffle_scatter_peers(gj_ctx* ctx, const int32_t* d_keys, const int32_t* d_pays,
                             uint64_t n, uint32_t n_gpus, uint32_t gpu_shift,
                             void* const* d_peer_bases, const uint64_t* h_peer_offsets);

/* Asynchronous peer-store scatter of relation `which` (0 = R, 1 = S) on `cuda_stream`: returns
 * without synchronising; per-relation cursor/base buffers let both relations be in flight.
 * With option "shuffle_grid" = k > 0 the kernel runs as k persistent CTAs that loop over the
 * tiles, leaving SM resources to kernels running concurrently on other streams.
 * gj_shuffle_scatter_ms waits for that kernel and returns its duration. */
int gj_shuffle_scatter_peers_async(gj_ctx* ctx, int which, const int32_t* d_keys, const int32_t* d_pays,
                                   uint64_t n, uint32_t n_gpus, uint32_t gpu_shift,
                                   void* const* d_peer_bases, const uint64_t* h_peer_offsets,
                                   void* cuda_stream);
int gj_shuffle_scatter_ms(gj_ctx* ctx, int which, float* ms);
/* Device-to-device (also peer / IPC-mapped) copy on a stream: runs on the copy engines, so the
 * "dma" shuffle variant moves tuples over NVLink without occupying SMs. */
int gj_memcpy_d2d_async(void* dst, const void* src, uint64_t bytes, void* cuda_stream);

/* Staged form of gj_join_aggregate_tuples for overlapping with the shuffle: begin fixes the plan
 * from (nR, nS); partition(side) enqueues histogram + scan + radix passes of one relation's
 * received packed tuples on a stream of the caller's choice (the two sides may use different
 * streams; they serialise only on the shared first-pass buffer); join enqueues unit planning +
 * the join after both sides; finish synchronises and returns the aggregate.  Options
 * "radix_bits" / "gpu_bits" as for gj_join_aggregate_tuples. */
int gj_stage_begin(gj_ctx* ctx, uint64_t nR, uint64_t nS, void* cuda_stream);
int gj_stage_partition(gj_ctx* ctx, int side, const void* d_tuples, void* cuda_stream);
int gj_stage_join(gj_ctx* ctx, void* cuda_stream);
int gj_stage_finish(gj_ctx* ctx, uint64_t* matches, uint64_t* checksum);
/* Durations of the local scatter launches of the last staged join: R pass 1, R pass 2, S pass 1,
 * S pass 2 (CUDA events on the streams they ran on; 0 where a pass did not run). */
int gj_stage_pass_ms(gj_ctx* ctx, float pass_ms[4]);

/* ---- sharded "partition, then push" pipeline (multi-GPU, alternative to gj_pcp_*)
 * Radix field = [gpu bits | local bits] (destination = (key >> local_bits) & (n_gpus-1), as above).
 * Every GPU partitions its OWN shard on all gpu+local bits in two passes; the second pass stores
 * its runs straight into the destination GPU's final partition buffer (local or peer-mapped over
 * NVLink), so the all-to-all IS the last radix pass and the receiver joins what arrives without
 * another pass over it.  Per relation (`which` 0 = R, 1 = S), each on a stream of the caller's:
 *   gj_pp_local  histogram + first pass (local) + this shard's fine histogram into d_fine_hist
 *                (2^(gpu bits + local bits) uint32, device);
 *   -- the caller all-gathers the fine histograms of all ranks into d_all_hist
 *      ([n_gpus][2^(gpu bits + local bits)] uint32, rank-major) on the same stream --
 *   gj_pp_push   derives this rank's write cursors from d_all_hist (every rank computes the same
 *                layout, no further exchange) and runs the pushing pass; peer_bases[g] = GPU g's
 *                partition buffer for this relation (16-byte aligned, cap_tuples + 16 tuples);
 *   -- the caller makes sure every rank's push has completed (e.g. a 1-element all-reduce) --
 *   gj_pp_join   unit planning + build/probe over this GPU's received partitions;
 *   gj_pp_finish synchronises; returns the local aggregate, the tuples this GPU received and
 *                (phase_ms[5], optional) the device time of local R, push R, local S, push S, join.
 * gj_pp_finish fails with GJ_ERR_ARG when some destination would have overflowed cap_tuples (then
 * nothing was pushed on any rank).  Options: "pass1_bits" (first-pass bits, 0 = the larger half of
 * the field), "pp_out" (pushed runs leave the SM as 8-byte stores (0) or TMA bulk stores (1,
 * default)), "pp_tile16k" (16 K-tuple push tiles (1, default) or 4-8 K (0): longer runs cross
 * NVLink faster). */
int gj_pp_begin(gj_ctx* ctx, uint64_t n_R_global, uint64_t n_S_global, uint32_t n_gpus, uint32_t rank,
                uint32_t local_bits, void* cuda_stream);
int gj_pp_local(gj_ctx* ctx, int which, const int32_t* d_keys, const int32_t* d_pays, uint64_t n,
                uint32_t* d_fine_hist, void* cuda_stream);
int gj_pp_push(gj_ctx* ctx, int which, const uint32_t* d_all_hist, void* const* peer_bases,
               uint64_t cap_tuples, uint64_t n, void* cuda_stream);
int gj_pp_join(gj_ctx* ctx, const void* d_own_R, const void* d_own_S, uint64_t cap_R, uint64_t cap_S,
               void* cuda_stream);
int gj_pp_finish(gj_ctx* ctx, uint64_t* matches, uint64_t* checksum, uint64_t* n_local_R,
                 uint64_t* n_local_S, float* phase_ms);
int gj_pp_plan(gj_ctx* ctx, uint32_t* pass1_bits, uint32_t* pass2_bits);

/* ---- sharded "partition, copy, partition" pipeline (multi-GPU, the default exchange) ---------
 * NVLink moves long runs far better than short ones, so here only WHOLE first-pass partitions
 * cross it, and they cross it as a STREAM: in ascending first-pass partition on every GPU, in
 * n_stages groups, each followed by a flag store into every peer; the receiver partitions and joins a
 * group while the later ones are still in flight.  The relation that builds (the globally smaller one,
 * R when equal) goes first.  Per relation (`which` 0 = R, 1 = S), each on streams of the caller's:
 *   gj_pcp_hist  this shard's histogram on [gpu bits | top local bits] (2^(g + bl) uint32 into
 *                d_coarse_hist); the caller all-gathers those into d_all_hist ([n_gpus][2^(g + bl)]);
 *   gj_pcp_part  layout + first radix pass: remote chunks into the context's stage buffer, this GPU's
 *                own chunks straight into its receive buffer d_own (never copied);
 *   gj_pcp_copy  n_stages x (TMA bulk-copy kernel: every remote chunk of the group to its slot in the
 *                destination's receive buffer peer_bases[g] (16-byte aligned, cap_tuples + 16 tuples),
 *                its histogram warps counting every piece by the receiver-side radix bits while it sits
 *                in shared memory; then those counts and this source's flag word into every GPU's
 *                control block peer_ctrl[g]);
 *   gj_pcp_recv  per stage: wait until every source's flag arrived (bounded by option
 *                "pcp_timeout_ms", default 5000: then GJ_ERR_STATE from gj_pcp_finish, never a hang),
 *                sum the delivered counts (the receiver never re-reads the data to count it), LAST
 *                radix pass over the group in d_own; for the probing relation also the join of the
 *                group, and at the end the local {matches, checksum} into d_result_out (2 x uint64 on
 *                the device, optional: the input of the caller's all-reduce).  Call it on a stream other
 *                than gj_pcp_copy's so that receiving overlaps this GPU's own sending; the building
 *                relation must be received first.
 *   gj_pcp_finish synchronises; local aggregate, tuples received, phase_ms[7] = part R, copy R,
 *                recv R, part S, copy S, recv S, tail (last byte of the probing relation landed ->
 *                last join done); plan_bits[3] = gpu bits, source-side local bits, receiver-side bits.
 * Control block: gj_pcp_ctrl_bytes(n_gpus) bytes per GPU, 16-byte aligned, zeroed once (gj_malloc_device +
 * gj_memset_device): stage flags [relation][stage][source] uint32 -- a flag holds the join number (epoch)
 * of the last completed stage, so it never needs resetting -- followed by the delivered fine histograms
 * [relation][source][2^local_bits] uint32 and the coarse histograms + flags of gj_pcp_hist_exchange.  Option "pcp_copy_ctas": see gj_set_option.
 * n + 2^(g + bl) must not exceed the context capacity (one spare stage slot per chunk). */
int gj_pcp_begin(gj_ctx* ctx, uint64_t n_R_global, uint64_t n_S_global, uint32_t n_gpus, uint32_t rank,
                 uint32_t local_bits, void* cuda_stream);
int gj_pcp_plan(gj_ctx* ctx, uint32_t plan_bits[3]);
int gj_pcp_hist(gj_ctx* ctx, int which, const int32_t* d_keys, uint64_t n, uint32_t* d_coarse_hist,
                void* cuda_stream);
int gj_pcp_part(gj_ctx* ctx, int which, const int32_t* d_keys, const int32_t* d_pays,
                const uint32_t* d_all_hist, void* d_own, uint64_t cap_tuples, void* cuda_stream);
int gj_pcp_copy(gj_ctx* ctx, int which, void* const* peer_bases, void* const* peer_ctrl, uint32_t n_stages,
                void* cuda_stream);
int gj_pcp_recv(gj_ctx* ctx, int which, const void* d_own, const void* d_ctrl, uint64_t cap_tuples,
                void* d_result_out, void* cuda_stream);
uint64_t gj_pcp_ctrl_bytes(uint32_t n_gpus);
/* Optional replacement of the caller's all-gather between gj_pcp_hist and gj_pcp_part: pushes this shard's coarse
 * histogram into every GPU's control block, waits (bounded, "pcp_timeout_ms") for every source's, and compacts them
 * into d_all_hist ([n_gpus][2^(g + bl)] uint32).  Stream-ordered, no collective library involved. */
int gj_pcp_hist_exchange(gj_ctx* ctx, int which, const uint32_t* d_coarse_hist, void* const* peer_ctrl,
                         const void* d_ctrl, uint32_t* d_all_hist, void* cuda_stream);
int gj_pcp_finish(gj_ctx* ctx, uint64_t* matches, uint64_t* checksum, uint64_t* n_local_R,
                  uint64_t* n_local_S, float* phase_ms, uint32_t* plan_bits);

/* CUDA IPC plumbing for the peer-store variant when every GPU is driven by its own process:
 * export a gj_malloc_device allocation as a 64-byte handle, open a peer's handle (peer access is
 * enabled lazily), close it again. */
int gj_ipc_export(void* d_ptr, char handle[64]);
int gj_ipc_open(const char handle[64], void** d_ptr);
int gj_ipc_close(void* d_ptr);
/* One process driving several GPUs (tests, ncu): let kernels on `device` dereference `peer`'s memory. */
int gj_enable_peer_access(int device, int peer);

/* Per-destination histogram only (what ranks exchange before gj_shuffle_scatter_peers). */
int gj_shuffle_count(gj_ctx* ctx, const int32_t* d_keys, uint64_t n, uint32_t n_gpus,
                     uint32_t gpu_shift, uint64_t* h_counts);

/* ---- synthetic relations on the device ---------------------
 * rows [row_begin, row_begin+n_rows) of a relation of n_total unique keys: key = pi_seed(row), a
 * seeded bijection on [0, n_total) (cycle-walking Feistel network); payload = mix(key, seed)
 * so that the checksum of a join is a function of the key set only. */
int gj_generate_unique(gj_ctx* ctx, int32_t* d_keys, int32_t* d_pays, uint64_t row_begin,
                       uint64_t n_rows, uint64_t n_total, uint32_t seed, uint32_t pay_seed);

/* Host helpers mirroring the device generator bit for bit (used to derive known answers). */
uint32_t gj_bijection(uint64_t row, uint64_t n_total, uint32_t seed);
int32_t gj_payload_of_key(uint32_t key, uint32_t pay_seed);

/* ---- device memory convenience (so a plain-C host can drive the engine without cudart) ------ */
int gj_device_count(int* n);
int gj_malloc_device(void** p, uint64_t bytes);
int gj_free_device(void* p);
int gj_malloc_pinned(void** p, uint64_t bytes);
int gj_free_pinned(void* p);
int gj_memcpy_h2d(void* d, const void* h, uint64_t bytes);
int gj_memcpy_d2h(void* h, const void* d, uint64_t bytes);
int gj_memset_device(void* d, int value, uint64_t bytes);
int gj_device_synchronize(void);
/* L2 flush helper for benchmarks: writes a scratch buffer larger than L2. */
int gj_flush_l2(gj_ctx* ctx);
/* Number of kernels this library has launched since it was loaded (bench.py's gpu_launches). */
uint64_t gj_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GPUJOIN_H */
