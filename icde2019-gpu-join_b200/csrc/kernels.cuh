// kernels.cuh -- device code of libgpujoin.so (sm_100a only).
//
// Replaces the reference's GPU primitives (file:line into /root/reference/src):
//   hist_kernel + scan_lookback_kernel + plan_kernel
//        -> init_metadata_double (join-primitives.cu:577-618) and compute_bucket_info (:294-312):
//           the reference discovers partition sizes while scattering (bucket chains); here an
//           exact key-only histogram of ALL radix bits is taken once (4 B/tuple) and scanned, so
//           every pass writes contiguous partitions.
//   scatter_kernel
//        -> partition_pass_one (:58-283) and partition_pass_two (:338-535): per-tile shared
//           memory histogram, one global ticket per (tile, digit), tuples reordered in shared
//           memory and written out as contiguous runs (key and payload travel together).
//   join_kernel
//        -> decompose_chains (:843-874, probe-side splitting becomes the unit list written by
//           plan_kernel), join_partitioned_aggregate (:885-1095) and join_partitioned_results
//           (:1107-1416); <..., LATE>: join_partitioned_varpayload (:1420-1557); fed by a TMA
//           bulk-copy ring (cp.async.bulk + mbarrier).
//   np_build_kernel / np_probe_kernel, np_build_perfect_kernel / np_probe_perfect_kernel
//        -> build_ht_chains / chains_probing (:681-742), build_perfect_array / probe_perfect_array
//           (:628-668): the non-partitioned joins.
//   subhist_tiles_kernel, pp_cursor_kernel, scatter_kernel<..., PUSH>   (section 3c)
//   pcp_layout_kernel, pcp_copy_kernel, pcp_self_hist_kernel, pcp_push_kernel, pcp_wait_kernel,
//   pcp_sum_hist_kernel, pcp_hist_push_kernel, pcp_hist_gather_kernel   (section 3d)
//        -> no reference counterpart (the reference is single-GPU, hash_join_clustered_probe.cu
//           :1001,1685 only select a device): the multi-GPU exchange, SURVEY.md section 8e.
//
// Data layout in HBM: inputs are columnar int32 keys / payloads (the reference's R/Pr, S/Ps);
// between passes and into the join tuples are packed {key,payload} 8-byte pairs (tup_t) so one
// 8-byte access moves a whole tuple.  Partition p of a relation is tuples[off[p] .. off[p+1]).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gj {

typedef uint2 tup_t;  // .x = key bits, .y = payload bits

constexpr int MAX_RADIX_BITS = 16;    // fine histogram: 128 KB of smem = 2^15 u32 or 2^16 packed u16 counters
constexpr int MAX_PASS_BITS = 8;      // fan-out per scatter pass <= 256
constexpr int NB_MAX = 1 << MAX_PASS_BITS;
constexpr uint32_t CUR1_STRIDE = 1;   // words between first-pass cursors (measured: padding to one cursor per 128-byte line
                                      // is SLOWER, 0.71 -> 0.58 of HBM peak: a warp's 32 adjacent tickets coalesce into one L2 request)
constexpr uint32_t EMPTY32 = 0xFFFFFFFFu;
constexpr uint32_t EMPTY16 = 0xFFFFu;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// ------------------------------------------------------------------------------------------
// 1. Radix histogram (keys only).  digit = (key >> shift) & (2^bits - 1).
//    Persistent grid (one 1024-thread CTA per SM), the 2^bits counters live in dynamic shared
//    memory (128 KB -- only possible with Blackwell's 227 KB per CTA), 16-byte loads, counters
//    flushed with one global reduction per non-empty bin per CTA.
//    P16 = false: 32-bit counters, bits <= 15.  P16 = true (bits == 16): two 16-bit counters per
//    word; the add that lifts a field to 0x8000 moves those 0x8000 counts to the global
//    histogram at once, so a field never carries into its neighbour (that would take another
//    32767 adds landing between that thread's add and its subtract).
//    Algorithmic bytes: 4 per tuple (columnar) -- the packed-tuple variant reads 8.
// ------------------------------------------------------------------------------------------
template <bool P16>
__device__ __forceinline__ void hist_add(uint32_t* sh, uint32_t d, uint32_t* ghist) {
    if (!P16) {
        atomicAdd(&sh[d], 1u);
    } else {
        const uint32_t sft = (d & 1u) << 4;
        const uint32_t old = atomicAdd(&sh[d >> 1], 1u << sft);
        if (((old >> sft) & 0xFFFFu) == 0x7FFFu) {
            atomicAdd(&ghist[d], 0x8000u);
            atomicSub(&sh[d >> 1], 0x8000u << sft);
        }
    }
}

template <bool PACKED, bool P16>
__global__ void __launch_bounds__(1024, 1)
hist_kernel(const void* __restrict__ in, uint32_t n, uint32_t shift, uint32_t bits,
            uint32_t* __restrict__ ghist, const uint32_t* __restrict__ lo_dev = nullptr,
            const uint32_t* __restrict__ hi_dev = nullptr) {
    extern __shared__ uint32_t sh_hist[];
    // sharded pipelines: the slots [*lo_dev, *hi_dev) of `in` (what this GPU received for a group of
    // first-pass partitions) -- known on the device only
    uint32_t first = 0;
    if (lo_dev) { first = *lo_dev; n = *hi_dev - first; }
    const uint32_t nb = 1u << bits, mask = nb - 1u;
    const uint32_t nwords = P16 ? nb >> 1 : nb;
    for (uint32_t i = threadIdx.x; i < nwords; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();

    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t gsz = gridDim.x * blockDim.x;
#define GJ_HADD(key) hist_add<P16>(sh_hist, ((uint32_t)(key) >> shift) & mask, ghist)
    if (!PACKED) {
        const int32_t* keys = (const int32_t*)in + first;
        uint32_t head = (uint32_t)(((16u - (uint32_t)((size_t)keys & 15u)) & 15u) >> 2);
        if (head > n) head = n;
        const uint32_t nvec = (n - head) >> 2;
        const int4* v = (const int4*)(keys + head);
        uint32_t i = gtid;
        // 4 independent 16-byte loads in flight per thread
        for (; i + 3 * gsz < nvec; i += 4 * gsz) {
            int4 k0 = __ldg(v + i), k1 = __ldg(v + i + gsz), k2 = __ldg(v + i + 2 * gsz),
                 k3 = __ldg(v + i + 3 * gsz);
#define GJ_H4(k) GJ_HADD((k).x); GJ_HADD((k).y); GJ_HADD((k).z); GJ_HADD((k).w);
            GJ_H4(k0) GJ_H4(k1) GJ_H4(k2) GJ_H4(k3)
        }
        for (; i < nvec; i += gsz) {
            int4 k0 = __ldg(v + i);
            GJ_H4(k0)
        }
#undef GJ_H4
        // unaligned head and the < 4 element tail
        const uint32_t tail0 = head + (nvec << 2);
        if (gtid < head) GJ_HADD(keys[gtid]);
        if (tail0 + gtid < n) GJ_HADD(keys[tail0 + gtid]);
    } else {
        const tup_t* tp = (const tup_t*)in + first;
        uint32_t head = (uint32_t)(((size_t)tp & 15u) ? 1u : 0u);
        if (head > n) head = n;
        const uint32_t nvec = (n - head) >> 1;
        const uint4* v = (const uint4*)(tp + head);
        uint32_t i = gtid;
        for (; i + 3 * gsz < nvec; i += 4 * gsz) {
            uint4 k0 = __ldg(v + i), k1 = __ldg(v + i + gsz), k2 = __ldg(v + i + 2 * gsz),
                  k3 = __ldg(v + i + 3 * gsz);
#define GJ_H2(k) GJ_HADD((k).x); GJ_HADD((k).z);
            GJ_H2(k0) GJ_H2(k1) GJ_H2(k2) GJ_H2(k3)
        }
        for (; i < nvec; i += gsz) {
            uint4 k0 = __ldg(v + i);
            GJ_H2(k0)
        }
#undef GJ_H2
        const uint32_t tail0 = head + (nvec << 1);
        if (gtid < head) GJ_HADD(tp[gtid].x);
        if (tail0 + gtid < n) GJ_HADD(tp[tail0 + gtid].x);
    }
#undef GJ_HADD
    __syncthreads();
    if (!P16) {
        for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) {
            const uint32_t c = sh_hist[i];
            if (c) atomicAdd(&ghist[i], c);
        }
    } else {
        for (uint32_t i = threadIdx.x; i < nwords; i += blockDim.x) {
            const uint32_t w = sh_hist[i];
            if (w & 0xFFFFu) atomicAdd(&ghist[2 * i], w & 0xFFFFu);
            if (w >> 16) atomicAdd(&ghist[2 * i + 1], w >> 16);
        }
    }
}

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk async copies (TMA engine, SASS UBLKCP)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// named barrier over a subset of the CTA's warps (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes: multiple of 16, 16 B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// 2. Exclusive prefix sums: single-pass chained scan with decoupled look-back.
//    blockIdx.y selects the sequence: 0/1 = fine histogram of the build/probe relation -> offsets,
//    2 = join units per partition (computed on the fly from the two histograms) -> unit bases.
//    Tile = 256 threads x 8 values.  Descriptor word = (status << 32) | value, status 0 = not
//    ready, 1 = tile aggregate, 2 = inclusive prefix; written and read as one 64-bit access, so
//    no fence is needed between flag and value.  Tile ids come from an atomic ticket so a
//    waiting tile's predecessors are always already running.
// ------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256, SCAN_IPT = 8, SCAN_TILE = SCAN_THREADS * SCAN_IPT;

enum { SCAN_PLAIN = 0, SCAN_UNITS = 1, SCAN_TILES = 2 };
struct ScanSeq {
    const uint32_t* in;           // PLAIN: nb values; UNITS: build counts; TILES: nb + 1 offsets
    const uint32_t* in2;          // UNITS: probe counts
    uint32_t* out;                // nb + 1 exclusive prefix sums
    unsigned long long* desc;     // one word per tile, zeroed
    uint32_t* ticket;             // zeroed
    uint32_t mode;                // SCAN_*
    uint32_t param;               // UNITS: probe tuples per unit; TILES: tuples per scatter tile
    uint32_t param2;              // TILES: parent order (tile_perm), 0 = identity
    uint32_t lo = 0, hi = 0;      // UNITS: only partitions in [lo, hi) produce units (hi == 0: all)
};
struct ScanArgs {
    ScanSeq seq[3];
    uint32_t nb;
    uint32_t seq_base;            // blockIdx.y + seq_base selects the sequence
};

__device__ __forceinline__ uint32_t units_of(uint32_t n_bld, uint32_t n_prb, uint32_t unit) {
    return (n_bld && n_prb) ? (n_prb + unit - 1) / unit : 0u;
}
// Order in which the pushing pass of the sharded pipeline visits its first-pass partitions: the
// top g bits of a partition id name the destination GPU, so position k takes partition
// (k mod 2^g, k div 2^g) -- consecutive positions go to different destinations (no NVLink ingress
// hot spot) while all tiles of one partition stay adjacent (their runs complete cache lines in L2
// together).  pm = (b1 << 8) | g, 0 = identity.
__host__ __device__ __forceinline__ uint32_t tile_perm(uint32_t k, uint32_t pm) {
    if (!pm) return k;
    const uint32_t g = pm & 0xFFu, b1 = pm >> 8;
    return ((k & ((1u << g) - 1u)) << (b1 - g)) | (k >> g);
}
// scatter tiles of a parent partition [lo, hi): tiles start on even slots
__device__ __forceinline__ uint32_t tiles_of(uint32_t lo, uint32_t hi, uint32_t tile) {
    return hi > lo ? (hi - (lo & ~1u) + tile - 1) / tile : 0u;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_lookback_kernel(ScanArgs a) {
    __shared__ uint32_t s_tile, s_prefix, s_warp[SCAN_THREADS / 32];
    const uint32_t which = blockIdx.y + a.seq_base;
    const ScanSeq r = a.seq[which];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(r.ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * SCAN_TILE + tid * SCAN_IPT;

    uint32_t v[SCAN_IPT], tsum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_IPT; ++j) {
        v[j] = 0;
        if (base + j < a.nb) {
            if (r.mode == SCAN_PLAIN) v[j] = r.in[base + j];
            else if (r.mode == SCAN_UNITS) v[j] = (r.hi == 0u || (base + j >= r.lo && base + j < r.hi)) ? units_of(r.in[base + j], r.in2[base + j], r.param) : 0u;
            else { const uint32_t c = tile_perm(base + j, r.param2); v[j] = tiles_of(r.in[c], r.in[c + 1], r.param); }
        }
        tsum += v[j];
    }
    uint32_t incl = warp_incl_scan(tsum, lane);
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t woff = 0, agg = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        uint32_t x = s_warp[w];
        if (w < wid) woff += x;
        agg += x;
    }
    const uint32_t texcl = incl - tsum + woff;

    if (wid == 0) {
        volatile unsigned long long* desc = r.desc;
        uint32_t prefix = 0;
        if (tile == 0) {
            if (lane == 0) desc[0] = (2ull << 32) | agg;
        } else {
            if (lane == 0) desc[tile] = (1ull << 32) | agg;
            // warp-wide look-back window over tiles tile-1, tile-2, ...
            int32_t top = (int32_t)tile - 1;
            for (;;) {
                const int32_t idx = top - lane;
                unsigned long long w = (2ull << 32);  // virtual tiles before 0: prefix 0
                if (idx >= 0) {
                    do { w = desc[idx]; } while ((w >> 32) == 0ull);
                }
                const uint32_t st = (uint32_t)(w >> 32), val = (uint32_t)w;
                const uint32_t done = __ballot_sync(0xffffffffu, st == 2u);
                const int first = __ffs(done) - 1;  // nearest tile holding an inclusive prefix
                uint32_t contrib = (first < 0 || lane <= first) ? val : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                prefix += contrib;
                if (first >= 0) break;
                top -= 32;
            }
            if (lane == 0) desc[tile] = (2ull << 32) | (uint32_t)(prefix + agg);
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    uint32_t run = s_prefix + texcl;
#pragma unroll
    for (int j = 0; j < SCAN_IPT; ++j) {
        if (base + j < a.nb) r.out[base + j] = run;
        run += v[j];
        if (base + j + 1 == a.nb) r.out[a.nb] = run;
    }
}

// ------------------------------------------------------------------------------------------
// 3. Work planning (fully parallel): scatter cursors, pass-2 tile descriptors, and the join's
//    unit list (replaces decompose_chains, join-primitives.cu:843-874: probe partitions longer
//    than `unit` tuples are cut into units that different CTAs join against the same build
//    partition).
//    Tile descriptor {a0, lo, hi, cursor_base}: the tile covers slots [a0, a0+T) of the
//    first-pass output, a0 EVEN (16-byte aligned tuple pairs), of which [lo, hi) belong to this
//    tile's first-pass partition.  Unit descriptor {probe_begin, probe_end, build_begin,
//    build_end}.
// ------------------------------------------------------------------------------------------
constexpr int PLAN_THREADS = 256;

struct PlanRel {
    const uint32_t* off;   // nb + 1 fine offsets
    uint32_t* cur1;        // 2^b1 cursors of the first pass (unused when single pass)
    uint32_t* cur2;        // nb cursors of the last pass
    uint4* tiles;          // pass-2 tile descriptors
    uint32_t* num_tiles;
};
struct PlanArgs {
    PlanRel rel[2];        // [0] build side, [1] probe side
    uint32_t nrel;         // relations whose cursors / tile descriptors are prepared (0, 1 or 2)
    uint32_t with_units;   // also write the join's unit list (needs both relations' offsets)
    uint32_t b1, b2;       // b2 == 0: single pass
    uint32_t tile;         // tuples per pass-2 scatter tile
    uint32_t unit;         // probe tuples per join unit
    const uint32_t* unit_base;   // nb + 1 (exclusive scan of units per partition)
    uint4* units;
    // staged receiver of the sharded pipeline: only first-pass partitions [j_lo, j_hi) get cursors,
    // tiles and units (j_hi == 0: everything)
    uint32_t j_lo, j_hi;
};

__global__ void __launch_bounds__(PLAN_THREADS)
plan_kernel(PlanArgs a) {
    // 16-byte aligned and padded: the compiler reads s_warp with 128-bit loads, which otherwise also touch the
    // neighbouring array's last word (harmless, but racecheck reports it)
    __shared__ __align__(16) uint32_t s_tp[NB_MAX + 4], s_a0[NB_MAX], s_warp[PLAN_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t gtid = blockIdx.x * PLAN_THREADS + tid, gsz = gridDim.x * PLAN_THREADS;
    const uint32_t B = a.b1 + a.b2, nb = 1u << B, n1 = 1u << a.b1;
    const uint32_t j_lo = a.j_hi ? a.j_lo : 0u, j_hi = a.j_hi ? a.j_hi : n1;
    const uint32_t p_lo = a.j_hi ? (j_lo << a.b2) : 0u, p_hi = a.j_hi ? (j_hi << a.b2) : nb;
    for (uint32_t r = 0; r < a.nrel; ++r) {
        const PlanRel R = a.rel[r];
        for (uint32_t p = p_lo + gtid; p < p_hi; p += gsz) R.cur2[p] = R.off[p];
        if (a.b2) {
            // every CTA recomputes the (<= 256 entry) tile prefix in shared memory
            uint32_t lo = 0, hi = 0, a0 = 0, tiles = 0;
            if (tid < n1) {
                lo = R.off[tid << a.b2]; hi = R.off[(tid + 1) << a.b2];
                a0 = lo & ~1u;
                tiles = (hi > lo && tid >= j_lo && tid < j_hi) ? (hi - a0 + a.tile - 1) / a.tile : 0u;
                if (blockIdx.x == 0) R.cur1[tid * CUR1_STRIDE] = lo;
            }
            __syncthreads();   // previous relation's readers of s_tp are done
            uint32_t incl = warp_incl_scan(tiles, lane);
            if (lane == 31) s_warp[wid] = incl;
            __syncthreads();
            uint32_t woff = 0, tot = 0;
#pragma unroll
            for (uint32_t w = 0; w < PLAN_THREADS / 32; ++w) {
                const uint32_t x = s_warp[w];
                if (w < wid) woff += x;
                tot += x;
            }
            s_tp[tid] = incl - tiles + woff;
            s_a0[tid] = a0;
            if (tid == 0) s_tp[NB_MAX] = tot;
            __syncthreads();
            if (gtid == 0) *R.num_tiles = tot;
            for (uint32_t t = gtid; t < tot; t += gsz) {
                uint32_t l = 0, h = n1;   // largest parent with s_tp[parent] <= t and tiles > 0
                while (h - l > 1) {
                    const uint32_t m = (l + h) >> 1;
                    if (s_tp[m] <= t) l = m; else h = m;
                }
                const uint32_t plo = R.off[l << a.b2], phi = R.off[(l + 1) << a.b2];
                const uint32_t ta = s_a0[l] + (t - s_tp[l]) * a.tile;
                R.tiles[t] = make_uint4(ta, max(ta, plo), min(ta + a.tile, phi), l << a.b2);
            }
        }
    }
    if (!a.with_units) return;
    const uint32_t* offB = a.rel[0].off;
    const uint32_t* offP = a.rel[1].off;
    for (uint32_t p = p_lo + gtid; p < p_hi; p += gsz) {
        const uint32_t bb = offB[p], be = offB[p + 1], lo = offP[p], hi = offP[p + 1];
        if (be > bb && hi > lo) {
            uint32_t at = a.unit_base[p];
            for (uint32_t s = lo; s < hi; s += a.unit) a.units[at++] = make_uint4(s, min(hi, s + a.unit), bb, be);
        }
    }
}

// ------------------------------------------------------------------------------------------
// 3b. Third-pass support (more than 16 radix bits: build sides beyond 2^28 tuples).  After two
//     passes on the top 16 bits the data is grouped into 65536 second-level partitions; the third
//     pass splits each of them on the low b3 bits.  sub_hist_kernel counts those bits per
//     partition (one CTA per partition, plain stores: every counter is written, nothing needs
//     zeroing); tiles3_kernel writes the third pass's tile descriptors from the scanned tile
//     counts.  Extra traffic of the third level: 8 B/tuple (count) + 16 B/tuple (scatter).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sub_hist_kernel(const tup_t* __restrict__ data, const uint32_t* __restrict__ off2, uint32_t nparents,
                uint32_t b3, uint32_t* __restrict__ ghist3) {
    __shared__ uint32_t sh[32];
    const uint32_t m3 = (1u << b3) - 1u;
    for (uint32_t p = blockIdx.x; p < nparents; p += gridDim.x) {
        if (threadIdx.x < 32) sh[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t lo = off2[p], hi = off2[p + 1];
        uint32_t i = lo + threadIdx.x;
        for (; i + 3 * 256 < hi; i += 4 * 256) {
            const tup_t t0 = __ldg(data + i), t1 = __ldg(data + i + 256), t2 = __ldg(data + i + 512), t3 = __ldg(data + i + 768);
            atomicAdd(&sh[t0.x & m3], 1u); atomicAdd(&sh[t1.x & m3], 1u);
            atomicAdd(&sh[t2.x & m3], 1u); atomicAdd(&sh[t3.x & m3], 1u);
        }
        for (; i < hi; i += 256) atomicAdd(&sh[__ldg(data + i).x & m3], 1u);
        __syncthreads();
        if (threadIdx.x <= m3) ghist3[((size_t)p << b3) + threadIdx.x] = sh[threadIdx.x];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
tiles3_kernel(const uint32_t* __restrict__ off2, const uint32_t* __restrict__ tile_prefix, uint32_t nparents,
              uint32_t tile, uint32_t b3, uint4* __restrict__ tiles, uint32_t perm = 0) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nparents; k += gridDim.x * blockDim.x) {
        const uint32_t p = tile_perm(k, perm);     // tile_prefix is indexed by position, offsets by partition
        const uint32_t lo = off2[p], hi = off2[p + 1];
        uint32_t at = tile_prefix[k];
        for (uint32_t a0 = lo & ~1u; hi > lo && a0 < hi; a0 += tile)
            tiles[at++] = make_uint4(a0, max(a0, lo), min(a0 + tile, hi), p << b3);
    }
}

// ------------------------------------------------------------------------------------------
// 3c. Sharded "partition, then push" pipeline (multi-GPU, no reference counterpart).
//     Every GPU partitions its own shard on ALL radix bits [gpu bits | local bits]: pass 1 on
//     the high b1 bits (local), then the fine counts of the low b2 bits inside every first-pass
//     partition (subhist_tiles_kernel, one pass-2 tile per step, 8 B/tuple read), an all-gather
//     of those fine histograms (host: NCCL), and pass 2 whose runs go straight into the
//     DESTINATION GPU's final partition buffer over NVLink (scatter_kernel<..., PUSH>): the
//     all-to-all is the last radix pass and the receiver joins what arrives without another pass.
//     pp_cursor_kernel turns the gathered histograms into this GPU's write cursors: destination
//     d lays partition p out at  sum_{p' < p} C[d][p']  (C = counts summed over all sources) and
//     source r writes its share at  + sum_{s < r} H[s][d][p]  -- every rank computes the same
//     layout from the same numbers, no further communication.  The CTA of d == rank also writes
//     this GPU's own partition counts / offsets for the join's unit planning.
// ------------------------------------------------------------------------------------------
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
subhist_tiles_kernel(const tup_t* __restrict__ data, const uint4* __restrict__ tiles,
                     const uint32_t* __restrict__ num_tiles, uint32_t bits, uint32_t tile_tuples,
                     uint32_t* __restrict__ fine_hist) {
    __shared__ uint32_t sh[1024];
    const uint32_t nb = 1u << bits, mask = nb - 1u, nt = *num_tiles;
    for (uint32_t t = blockIdx.x; t < nt; t += gridDim.x) {
        for (uint32_t i = threadIdx.x; i < nb; i += THREADS) sh[i] = 0;
        __syncthreads();
        const uint4 td = __ldg(tiles + t);
        const uint32_t a0 = td.x, lo = td.y, hi = td.z;
        const uint4* v = reinterpret_cast<const uint4*>(data + a0);   // a0 even, buffer 16-byte aligned
        const bool full = (lo == a0) && (hi - a0 == tile_tuples);
        const uint32_t npairs = tile_tuples >> 1;
        uint32_t pi = threadIdx.x;
        if (full) {   // four independent 16-byte loads in flight per thread
            for (; pi + 3u * THREADS < npairs; pi += 4u * THREADS) {
                const uint4 x0 = __ldg(v + pi), x1 = __ldg(v + pi + THREADS), x2 = __ldg(v + pi + 2 * THREADS),
                            x3 = __ldg(v + pi + 3 * THREADS);
                atomicAdd(&sh[x0.x & mask], 1u); atomicAdd(&sh[x0.z & mask], 1u);
                atomicAdd(&sh[x1.x & mask], 1u); atomicAdd(&sh[x1.z & mask], 1u);
                atomicAdd(&sh[x2.x & mask], 1u); atomicAdd(&sh[x2.z & mask], 1u);
                atomicAdd(&sh[x3.x & mask], 1u); atomicAdd(&sh[x3.z & mask], 1u);
            }
        }
        for (; pi < npairs; pi += THREADS) {
            const uint32_t s0 = a0 + 2u * pi;
            if (full || (s0 >= lo && s0 + 1u < hi)) {
                const uint4 x = __ldg(v + pi);
                atomicAdd(&sh[x.x & mask], 1u);
                atomicAdd(&sh[x.z & mask], 1u);
            } else {   // tile edge: never touch a tuple outside [lo, hi)
                if (s0 >= lo && s0 < hi) atomicAdd(&sh[__ldg(data + s0).x & mask], 1u);
                if (s0 + 1u >= lo && s0 + 1u < hi) atomicAdd(&sh[__ldg(data + s0 + 1u).x & mask], 1u);
            }
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nb; i += THREADS) {
            const uint32_t c = sh[i];
            if (c) atomicAdd(&fine_hist[(size_t)td.w + i], c);
        }
        __syncthreads();
    }
}

constexpr int PPC_THREADS = 1024, PPC_IPT = 4;
// status words: [0] abort (some destination would overflow its buffer), [1] tuples this GPU receives
__global__ void __launch_bounds__(PPC_THREADS)
pp_cursor_kernel(const uint32_t* __restrict__ all_hist, uint32_t n_gpus, uint32_t rank, uint32_t local_bits,
                 uint32_t cap_tuples, uint32_t* __restrict__ cur_fine, uint32_t* __restrict__ loc_cnt,
                 uint32_t* __restrict__ loc_off, uint32_t* __restrict__ status) {
    __shared__ uint32_t s_warp[PPC_THREADS / 32];
    __shared__ uint32_t s_run;
    const uint32_t d = blockIdx.x, tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t np = 1u << local_bits;
    const size_t nq = (size_t)n_gpus << local_bits;
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (uint32_t base = 0; base < np; base += PPC_THREADS * PPC_IPT) {
        uint32_t c[PPC_IPT], pre[PPC_IPT], tsum = 0;
#pragma unroll
        for (int j = 0; j < PPC_IPT; ++j) {
            const uint32_t p = base + tid * PPC_IPT + j;
            c[j] = 0; pre[j] = 0;
            if (p < np) {
                const size_t q = ((size_t)d << local_bits) + p;
                for (uint32_t s = 0; s < n_gpus; ++s) {
                    const uint32_t h = __ldg(all_hist + (size_t)s * nq + q);
                    c[j] += h;
                    if (s < rank) pre[j] += h;
                }
            }
            tsum += c[j];
        }
        const uint32_t incl = warp_incl_scan(tsum, lane);
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        uint32_t woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < PPC_THREADS / 32; ++w) {
            const uint32_t x = s_warp[w];
            if ((uint32_t)w < wid) woff += x;
            tot += x;
        }
        uint32_t run = s_run + woff + incl - tsum;
#pragma unroll
        for (int j = 0; j < PPC_IPT; ++j) {
            const uint32_t p = base + tid * PPC_IPT + j;
            if (p < np) {
                cur_fine[((size_t)d << local_bits) + p] = run + pre[j];
                if (d == rank) { loc_cnt[p] = c[j]; loc_off[p] = run; }
            }
            run += c[j];
        }
        __syncthreads();              // everyone has read s_run / s_warp
        if (tid == 0) s_run += tot;
        __syncthreads();
    }
    const uint32_t total = s_run;
    // +16: bulk copies and 16-byte loads of the join round partition ends out to tuple pairs
    const bool overflow = (unsigned long long)total + 16ull > (unsigned long long)cap_tuples;
    if (overflow && tid == 0) atomicExch(&status[0], 1u);
    if (d == rank) {
        if (tid == 0) { status[1] = total; loc_off[np] = overflow ? 0u : total; }
        if (overflow)   // nothing will arrive: give the join an empty relation
            for (uint32_t p = tid; p < np; p += PPC_THREADS) { loc_cnt[p] = 0; loc_off[p] = 0; }
    }
}

// ------------------------------------------------------------------------------------------
// 3d. Sharded "partition, copy, partition" pipeline (pcp).  NVLink moves long runs far better than
//     short ones (profiles/README.md), so here only WHOLE first-pass partitions cross it: the
//     source scatters its shard on [gpu bits | top bL local bits] (<= 1024 chunks), a copy
//     kernel moves every chunk into the destination's first-pass layout with TMA bulk copies
//     (global -> shared -> peer global, no register traffic), and the receiver runs the last
//     radix pass + join.  The coarse histograms are all-gathered BEFORE the source pass:
//     pcp_layout_kernel derives, for chunk c = (destination d, first-pass partition j), where it
//     lands at d -- first-pass partition j of d starts at sum_{j'<j} C[d][j'], source r writes at
//     + sum_{s<r} H[s][c] -- and where the source stages it: every REMOTE chunk gets its count + 1
//     stage slots and starts on the slot whose 16-byte phase equals its destination's, so head
//     tuple, bulk body and tail tuple line up on both sides.  The chunks a GPU keeps (d == rank) are
//     never staged or copied: the source pass stores them straight into this GPU's receive buffer.
//     The exchange is a STREAM: chunks travel in ascending first-pass partition j on every GPU
//     (tile_perm), the copy is issued in stages (groups of j), each followed by a flag store into
//     every peer (pcp_push_kernel); the receiver waits per stage (pcp_wait_kernel) and runs the last
//     radix pass + join over the partitions of that stage while the later stages are still crossing
//     NVLink.  The receiver never re-reads what arrived to count it: the copy kernel's histogram
//     warps count every piece by the receiver-side radix bits while it sits in shared memory, and
//     pcp_push_kernel delivers those counts with the flag.
//     status: [0] abort (a destination would overflow) [1] tuples this GPU receives [2] pieces
//             [3] a wait timed out
// ------------------------------------------------------------------------------------------
constexpr int PCP_MAX_CHUNKS = 1024;
constexpr int PCP_MAX_STAGES = 64;
constexpr int MAX_RADIX_BITS_CTRL = 16;
constexpr uint32_t PCP_PIECE = 4096;     // tuples per bulk copy (32 KB), even
// per-GPU control block (peer-mapped): stage flags [relation][stage][source] uint32, then the fine
// histograms the sources deliver, fine_in [relation][source][2^B] uint32
__host__ __device__ __forceinline__ size_t pcp_ctrl_flag_bytes(uint32_t n_gpus) { return (size_t)2 * PCP_MAX_STAGES * n_gpus * sizeof(uint32_t); }
__host__ __device__ __forceinline__ size_t pcp_ctrl_fine_bytes(uint32_t n_gpus, uint32_t B) { return (((size_t)2 * n_gpus) << B) * sizeof(uint32_t); }
// ... then (at the offset for the largest B) the coarse histograms the sources deliver before the source pass,
// coarse_in [relation][source][PCP_MAX_CHUNKS] uint32, and their flags [relation][source]
__host__ __device__ __forceinline__ size_t pcp_ctrl_coarse_off(uint32_t n_gpus) { return pcp_ctrl_flag_bytes(n_gpus) + pcp_ctrl_fine_bytes(n_gpus, MAX_RADIX_BITS_CTRL); }
__host__ __device__ __forceinline__ size_t pcp_ctrl_hflag_off(uint32_t n_gpus) { return pcp_ctrl_coarse_off(n_gpus) + (size_t)2 * n_gpus * PCP_MAX_CHUNKS * sizeof(uint32_t); }
__host__ __device__ __forceinline__ size_t pcp_ctrl_bytes(uint32_t n_gpus) { return pcp_ctrl_hflag_off(n_gpus) + (size_t)2 * n_gpus * sizeof(uint32_t) + 64; }
struct PcpTables {                       // device arrays of n1 (+1) entries
    uint32_t* cur;                       // pass-1 cursors (consumed by the scatter), relative to dig_base[c]
    tup_t** dig_base;                    // pass-1 output base of chunk c: the stage buffer, or this GPU's receive buffer
    uint32_t* src_start;                 // remote chunks: first slot of chunk c in the source's stage buffer
    uint32_t* dst_start;                 // first slot of this source's share at the destination
    uint32_t* cnt;                       // tuples of chunk c in this shard
    uint32_t* piece_prefix;              // [n1 + 1], indexed by POSITION k (chunk tile_perm(k)); own chunks: no pieces
    uint32_t* recv_off;                  // [2^bl + 1] this GPU's receive layout: first-pass partition j starts at recv_off[j]
    uint32_t* status;
};

__device__ __forceinline__ uint32_t pcp_pieces(uint32_t cnt, uint32_t phase) {
    return cnt ? max(1u, (cnt - phase + PCP_PIECE - 1u) / PCP_PIECE) : 0u;
}

// one CTA of 1024 threads; thread t owns chunk t (partition order) and position t (copy order)
__global__ void __launch_bounds__(PCP_MAX_CHUNKS)
pcp_layout_kernel(const uint32_t* __restrict__ all_hist, uint32_t n_gpus, uint32_t rank, uint32_t b1, uint32_t bl,
                  uint32_t cap_tuples, uint32_t perm, tup_t* stage, tup_t* own, PcpTables t) {
    __shared__ uint32_t s_a[PCP_MAX_CHUNKS], s_ph[PCP_MAX_CHUNKS], s_cn[PCP_MAX_CHUNKS];
    __shared__ uint32_t s_warp[3][PCP_MAX_CHUNKS / 32];
    __shared__ uint32_t s_over;
    const uint32_t c = threadIdx.x, lane = c & 31u, wid = c >> 5, n1 = 1u << b1;
    const uint32_t d = c >> bl;
    const bool mine_stays = (d == rank);
    if (c == 0) s_over = 0;
    uint32_t tot = 0, pre = 0, mine = 0;
    if (c < n1)
        for (uint32_t s = 0; s < n_gpus; ++s) {
            const uint32_t h = all_hist[(size_t)s * n1 + c];
            tot += h;
            if (s < rank) pre += h;
            if (s == rank) mine = h;
        }
    // two block-wide exclusive scans in chunk order: destination totals, stage regions (count + 1, remote chunks only)
    const uint32_t region_sz = (c < n1 && !mine_stays) ? mine + 1u : 0u;
    uint32_t i0 = warp_incl_scan(tot, lane), i1 = warp_incl_scan(region_sz, lane);
    if (lane == 31) { s_warp[0][wid] = i0; s_warp[1][wid] = i1; }
    __syncthreads();
    uint32_t w0 = 0, w1 = 0;
    for (uint32_t w = 0; w < wid; ++w) { w0 += s_warp[0][w]; w1 += s_warp[1][w]; }
    const uint32_t ex_tot = i0 - tot + w0, region = i1 - region_sz + w1;
    s_a[c] = ex_tot;
    __syncthreads();
    const uint32_t dbase = s_a[min(d << bl, (uint32_t)PCP_MAX_CHUNKS - 1u)];   // first chunk of this destination
    const uint32_t dst = ex_tot - dbase + pre;
    const uint32_t src = region + ((region ^ dst) & 1u);       // same 16-byte phase as the destination slot
    const bool last_of_dest = (c < n1) && (((c + 1u) & ((1u << bl) - 1u)) == 0u);
    uint32_t dtot = 0;
    if (last_of_dest) {
        dtot = ex_tot + tot - dbase;
        // +16: bulk copies and 16-byte loads of the join round partition ends out to tuple pairs
        if ((unsigned long long)dtot + 16ull > (unsigned long long)cap_tuples) { atomicExch(&t.status[0], 1u); s_over = 1u; }
    }
    __syncthreads();
    const bool over = s_over != 0u;       // some destination overflows: nothing is sent, every receiver sees an empty relation
    if (c < n1) {
        t.cur[c] = mine_stays ? dst : src;
        t.dig_base[c] = mine_stays ? own : stage;
        t.src_start[c] = src; t.dst_start[c] = dst; t.cnt[c] = mine;
        if (mine_stays) {
            t.recv_off[c & ((1u << bl) - 1u)] = over ? 0u : ex_tot - dbase;
            if (last_of_dest) { t.recv_off[1u << bl] = over ? 0u : dtot; t.status[1] = over ? 0u : dtot; }
        }
    }
    s_ph[c] = src & 1u; s_cn[c] = (c < n1 && !mine_stays) ? mine : 0u;
    __syncthreads();
    // pieces in COPY order: position k takes chunk tile_perm(k)
    const uint32_t ck = tile_perm(c, perm);
    const uint32_t pieces = (c < n1) ? pcp_pieces(s_cn[ck], s_ph[ck]) : 0u;
    const uint32_t i2 = warp_incl_scan(pieces, lane);
    if (lane == 31) s_warp[2][wid] = i2;
    __syncthreads();
    uint32_t w2 = 0, all = 0;
    for (uint32_t w = 0; w < PCP_MAX_CHUNKS / 32; ++w) { const uint32_t x = s_warp[2][w]; if (w < wid) w2 += x; all += x; }
    if (c < n1) t.piece_prefix[c] = i2 - pieces + w2;
    if (c == 0) { t.piece_prefix[n1] = all; t.status[2] = all; }
}

struct PcpCopyArgs {
    const tup_t* stage;                  // first-pass output of this shard (remote chunks)
    tup_t* const* peer_bases;            // [n_gpus] receive buffers (mapped over NVLink)
    PcpTables t;
    uint32_t b1, bl, perm;
    uint32_t pos_lo, pos_hi;             // this launch moves the chunks at copy positions [pos_lo, pos_hi)
    uint32_t b2;                         // receiver-side radix bits
    uint32_t* fine;                      // this source's fine histogram [2^(b1 + b2)], index (chunk << b2) | low key bits
};

constexpr int PCP_HIST_WARPS = 4;
constexpr int PCP_COPY_THREADS = 32 * (2 + PCP_HIST_WARPS);
constexpr uint32_t PCP_BAR_HIST = 2;     // named barrier of the histogram warps

// Warp-specialised copy pipeline over a ring of NS slots of PCP_PIECE tuples; this CTA's pieces are those
// of the stage taken round robin.
//   warp 0 (one lane)  LOADER: per piece, address arithmetic (the chunk's table entries stay in registers
//                      while consecutive pieces belong to the same chunk), wait for the slot to be free,
//                      bulk load global -> shared (completion on the slot's "full" mbarrier).  Never waits
//                      for a load: up to NS loads are in flight per CTA.
//   warp 1 (one lane)  STORER: waits for "full", issues the bulk store shared -> peer global, and frees a
//                      slot once the store has READ it (bulk async-group accounting, one group per piece).
//   warps 2..5         HISTOGRAM: while a piece sits in shared memory they count its tuples by the
//                      receiver-side radix bits -- the FINE histogram the receiver needs for its last pass,
//                      taken here for free (no HBM traffic) instead of by a second read of everything that
//                      arrived (8 B/tuple at the receiver).  Counts gather per chunk in shared memory and go
//                      to `fine` when the CTA moves to another chunk.
// A slot is free again when the storer and the four histogram warps have arrived on its "free" mbarrier.
// NS = 3 on every SM, or NS = 6 (192 KB of shared memory) on a few SMs that then run nothing else, so the
// radix passes and the join next to it keep the other SMs whole.
template <int NS>
__global__ void __launch_bounds__(PCP_COPY_THREADS)
pcp_copy_kernel(PcpCopyArgs a) {
    static_assert(NS >= 2, "ring too small");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    tup_t* ring = reinterpret_cast<tup_t*>(smem_raw);                                   // [NS][PCP_PIECE]
    uint32_t* s_prefix = reinterpret_cast<uint32_t*>(smem_raw + (size_t)NS * PCP_PIECE * sizeof(tup_t));   // [n1 + 1]
    __shared__ uint64_t s_full[NS], s_free[NS];
    __shared__ tup_t* s_dst[NS];
    __shared__ uint32_t s_bytes[NS], s_chunk[NS];
    __shared__ uint32_t s_hist[1u << 10];       // receiver-side pass: <= 10 bits
    const uint32_t n1 = 1u << a.b1, tid = threadIdx.x, wid = tid >> 5;
    if (a.t.status[0]) return;
    for (uint32_t i = tid; i <= n1; i += PCP_COPY_THREADS) s_prefix[i] = a.t.piece_prefix[i];
    for (uint32_t i = tid; i < (1u << a.b2); i += PCP_COPY_THREADS) s_hist[i] = 0;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_free[s], 1 + PCP_HIST_WARPS); }
        fence_mbar_init();
    }
    __syncthreads();
    const uint32_t k0 = s_prefix[a.pos_lo], total = s_prefix[min(a.pos_hi, n1)];
    const uint32_t first = k0 + blockIdx.x;
    const uint32_t mine = first < total ? (total - first + gridDim.x - 1) / gridDim.x : 0u;   // pieces of this CTA
    const uint32_t mask2 = (1u << a.b2) - 1u;
    if (wid >= 2) {
        // ---------------- histogram warps
        const uint32_t ht = tid - 64, HT = 32 * PCP_HIST_WARPS;
        uint32_t cur = 0xFFFFFFFFu;
        for (uint32_t i = 0; i < mine; ++i) {
            const uint32_t slot = i % NS;
            if ((tid & 31u) == 0) mbar_wait(&s_full[slot], (i / NS) & 1u);     // one poller per warp
            __syncwarp();
            const uint32_t c = s_chunk[slot], m = s_bytes[slot] / (uint32_t)sizeof(tup_t);
            if (c != cur) {
                if (cur != 0xFFFFFFFFu) {
                    named_bar_sync(PCP_BAR_HIST, HT);          // every count of the previous chunk is in
                    for (uint32_t x = ht; x <= mask2; x += HT) {
                        const uint32_t v = s_hist[x];
                        if (v) { atomicAdd(&a.fine[((size_t)cur << a.b2) + x], v); s_hist[x] = 0; }
                    }
                    named_bar_sync(PCP_BAR_HIST, HT);          // zeroed before anybody counts again
                }
                cur = c;
            }
            const tup_t* p = ring + (size_t)slot * PCP_PIECE;
#pragma unroll 4
            for (uint32_t x = ht; x < m; x += HT) atomicAdd(&s_hist[p[x].x & mask2], 1u);
            __syncwarp();
            if ((tid & 31u) == 0) mbar_arrive(&s_free[slot]);
        }
        if (cur != 0xFFFFFFFFu) {
            named_bar_sync(PCP_BAR_HIST, HT);
            for (uint32_t x = ht; x <= mask2; x += HT) {
                const uint32_t v = s_hist[x];
                if (v) atomicAdd(&a.fine[((size_t)cur << a.b2) + x], v);
            }
        }
        return;
    }
    if ((tid & 31u) != 0) return;
    if (wid == 1) {
        // ---------------- storer lane
        for (uint32_t i = 0; i < mine; ++i) {
            const uint32_t slot = i % NS;
            mbar_wait(&s_full[slot], (i / NS) & 1u);
            if (s_bytes[slot]) bulk_s2g(s_dst[slot], ring + (size_t)slot * PCP_PIECE, s_bytes[slot]);
            bulk_commit();
            if (i) {   // all groups but the newest have been read: piece i - 1 has left its slot
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                mbar_arrive(&s_free[(i - 1) % NS]);
            }
        }
        if (mine) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            mbar_arrive(&s_free[(mine - 1) % NS]);
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the stores have completed, not just been read
        __threadfence_system();
        return;
    }
    // ---------------- loader lane
    uint32_t pos = a.pos_lo, c = 0, src0 = 0, dst0 = 0, cnt = 0, cached = 0xFFFFFFFFu;
    tup_t* dbase = nullptr;
    uint32_t i = 0;
    for (uint32_t k = first; k < total; k += gridDim.x, ++i) {
        while (s_prefix[pos + 1] <= k) ++pos;           // the position holding piece k (k < total <= prefix of the stage's end)
        if (pos != cached) {
            cached = pos;
            c = tile_perm(pos, a.perm);
            src0 = a.t.src_start[c]; dst0 = a.t.dst_start[c]; cnt = a.t.cnt[c];
            dbase = a.peer_bases[c >> a.bl];
        }
        const uint32_t slice = k - s_prefix[pos];
        const uint32_t phase = src0 & 1u;     // == dst0 & 1 by construction
        const tup_t* src = a.stage + src0;
        tup_t* dst = dbase + dst0;
        if (slice == 0 && phase) {            // odd first slot: plain 8-byte copy, counted here
            const tup_t t = *src;
            *dst = t;
            atomicAdd(&a.fine[((size_t)c << a.b2) + (t.x & mask2)], 1u);
        }
        const uint32_t body0 = phase + slice * PCP_PIECE;         // even slot on both sides
        uint32_t m = (cnt > body0) ? min(PCP_PIECE, cnt - body0) : 0u;
        if (m & 1u) {                         // odd tail (last piece only)
            const tup_t t = src[body0 + m - 1u];
            dst[body0 + m - 1u] = t;
            atomicAdd(&a.fine[((size_t)c << a.b2) + (t.x & mask2)], 1u);
            --m;
        }
        const uint32_t slot = i % NS;
        if (i >= (uint32_t)NS) mbar_wait(&s_free[slot], ((i / NS) - 1u) & 1u);   // stored and counted
        s_dst[slot] = dst + body0;
        s_bytes[slot] = m * (uint32_t)sizeof(tup_t);
        s_chunk[slot] = c;
        mbar_arrive_expect_tx(&s_full[slot], m * (uint32_t)sizeof(tup_t));
        if (m) bulk_g2s(ring + (size_t)slot * PCP_PIECE, src + body0, m * (uint32_t)sizeof(tup_t), &s_full[slot]);
    }
    __threadfence_system();                   // the plain head / tail stores
}

// Fine histogram of the chunks this GPU KEEPS (its source pass stored them straight into its receive
// buffer; they never pass through the copy kernel): one CTA per (first-pass partition j, slice of
// PCP_SELF_SLICE tuples), counts by the receiver-side radix bits into `fine`.  8 B/tuple over 1/G of the shard.
constexpr uint32_t PCP_SELF_SLICE = 16384;
__global__ void __launch_bounds__(256)
pcp_self_hist_kernel(const tup_t* __restrict__ own, PcpTables t, uint32_t rank, uint32_t bl, uint32_t b2,
                     uint32_t* __restrict__ fine) {
    __shared__ uint32_t s_hist[1u << 10];
    if (t.status[0]) return;
    const uint32_t mask2 = (1u << b2) - 1u;
    const uint32_t c = (rank << bl) | blockIdx.y;
    const uint32_t cnt = t.cnt[c], base = t.dst_start[c];
    for (uint32_t s0 = blockIdx.x * PCP_SELF_SLICE; s0 < cnt; s0 += gridDim.x * PCP_SELF_SLICE) {
        for (uint32_t x = threadIdx.x; x <= mask2; x += 256) s_hist[x] = 0;
        __syncthreads();
        const uint32_t hi = min(cnt, s0 + PCP_SELF_SLICE);
        for (uint32_t i = s0 + threadIdx.x; i < hi; i += 256) atomicAdd(&s_hist[__ldg(&own[base + i].x) & mask2], 1u);
        __syncthreads();
        for (uint32_t x = threadIdx.x; x <= mask2; x += 256) {
            const uint32_t v = s_hist[x];
            if (v) atomicAdd(&fine[((size_t)c << b2) + x], v);
        }
        __syncthreads();
    }
}

// After a copy stage: hand every destination d the fine counts of what this source sent it in the stage
// (fine[(d << B) + p] for the stage's partitions p, into d's table fine_in[source = rank]) and then tell it
// that the stage has landed.  Runs on the copy's stream, i.e. after the copy kernel (and all its bulk
// stores) completed.  A flag word holds the epoch (join number) of the last completed stage; control
// block of one GPU: flags [relation][stage][source], then fine_in [relation][source][2^B].
// One CTA per destination (this GPU included: its own chunks' counts come from pcp_self_hist_kernel).
__global__ void __launch_bounds__(256)
pcp_push_kernel(const uint32_t* __restrict__ fine, unsigned char* const* __restrict__ peer_ctrl, uint32_t n_gpus,
                uint32_t rank, uint32_t which, uint32_t stage, uint32_t epoch, uint32_t B, uint32_t p_lo, uint32_t p_hi,
                const uint32_t* __restrict__ status) {
    const uint32_t d = blockIdx.x;
    unsigned char* ctrl = peer_ctrl[d];
    uint32_t* flags = reinterpret_cast<uint32_t*>(ctrl);
    uint32_t* fine_in = reinterpret_cast<uint32_t*>(ctrl + pcp_ctrl_flag_bytes(n_gpus)) + (((size_t)which * n_gpus + rank) << B);
    if (!status[0]) {   // (aborted exchange: the tables keep the previous join's counts -- sums within the same buffers' capacity;
                        //  whatever the receiver then computes is discarded, gj_pcp_finish returns the overflow error)
        const uint32_t* src = fine + ((size_t)d << B);
        for (uint32_t p = p_lo + threadIdx.x; p < p_hi; p += 256) fine_in[p] = src[p];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        volatile uint32_t* f = flags + ((size_t)which * PCP_MAX_STAGES + stage) * n_gpus + rank;
        *f = epoch;
    }
}

// Receiver: counts of the stage's partitions = sum over the sources' tables.
__global__ void __launch_bounds__(256)
pcp_sum_hist_kernel(const unsigned char* __restrict__ ctrl, uint32_t n_gpus, uint32_t which, uint32_t B, uint32_t p_lo,
                    uint32_t p_hi, uint32_t* __restrict__ ghist) {
    const uint32_t* fine_in = reinterpret_cast<const uint32_t*>(ctrl + pcp_ctrl_flag_bytes(n_gpus)) + (((size_t)which * n_gpus) << B);
    for (uint32_t p = p_lo + blockIdx.x * 256 + threadIdx.x; p < p_hi; p += gridDim.x * 256) {
        uint32_t v = 0;
        for (uint32_t s = 0; s < n_gpus; ++s) v += fine_in[((size_t)s << B) + p];
        ghist[p] = v;
    }
}

// Exchange of the coarse histograms WITHOUT a collective (option "pcp_peer_hist"): every source writes its 2^b1
// counters into every GPU's control block and raises a flag; pcp_hist_gather_kernel waits for all sources (bounded)
// and compacts them into the dense [n_gpus][2^b1] array the layout kernel reads.  ~15 us against ~110 us for an
// 8-rank NCCL all-gather of 2 KB -- on the critical path of the building relation.
__global__ void __launch_bounds__(256)
pcp_hist_push_kernel(const uint32_t* __restrict__ coarse, unsigned char* const* __restrict__ peer_ctrl, uint32_t n_gpus,
                     uint32_t rank, uint32_t which, uint32_t n1, uint32_t epoch) {
    unsigned char* ctrl = peer_ctrl[blockIdx.x];
    uint32_t* dst = reinterpret_cast<uint32_t*>(ctrl + pcp_ctrl_coarse_off(n_gpus)) + ((size_t)which * n_gpus + rank) * PCP_MAX_CHUNKS;
    for (uint32_t c = threadIdx.x; c < n1; c += 256) dst[c] = coarse[c];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        volatile uint32_t* f = reinterpret_cast<uint32_t*>(ctrl + pcp_ctrl_hflag_off(n_gpus)) + (size_t)which * n_gpus + rank;
        *f = epoch;
    }
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Receiver side: spin until every source has signalled `slot` for this epoch (bounded: after
// timeout_ns the kernel gives up and raises status[3], so a lost peer becomes an error, not a hang).
__global__ void pcp_wait_kernel(const uint32_t* __restrict__ flags, uint32_t n_gpus, uint32_t rank, uint32_t slot,
                                uint32_t epoch, unsigned long long timeout_ns, uint32_t* __restrict__ status) {
    const uint32_t s = threadIdx.x;
    (void)rank;                                 // this GPU's own counts arrive through its own push kernel: wait for it too
    if (s >= n_gpus) return;
    const volatile uint32_t* f = flags + (size_t)slot * n_gpus + s;
    const unsigned long long t0 = global_timer_ns();
    while ((int32_t)(*f - epoch) < 0) {
        if (global_timer_ns() - t0 > timeout_ns) { atomicExch(&status[3], 1u); break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

__global__ void __launch_bounds__(256)
pcp_hist_gather_kernel(const unsigned char* __restrict__ ctrl, uint32_t n_gpus, uint32_t which, uint32_t n1, uint32_t epoch,
                       unsigned long long timeout_ns, uint32_t* __restrict__ all_hist, uint32_t* __restrict__ status) {
    const uint32_t s = blockIdx.x;       // one CTA per source
    const volatile uint32_t* f = reinterpret_cast<const uint32_t*>(ctrl + pcp_ctrl_hflag_off(n_gpus)) + (size_t)which * n_gpus + s;
    if (threadIdx.x == 0) {
        const unsigned long long t0 = global_timer_ns();
        while ((int32_t)(*f - epoch) < 0) {
            if (global_timer_ns() - t0 > timeout_ns) { atomicExch(&status[3], 1u); break; }
            __nanosleep(100);
        }
        __threadfence_system();
    }
    __syncthreads();
    const volatile uint32_t* src = reinterpret_cast<const uint32_t*>(ctrl + pcp_ctrl_coarse_off(n_gpus)) + ((size_t)which * n_gpus + s) * PCP_MAX_CHUNKS;
    for (uint32_t c = threadIdx.x; c < n1; c += 256) all_hist[(size_t)s * n1 + c] = src[c];
}

// ------------------------------------------------------------------------------------------
// 4. Radix scatter pass.  One tile of THREADS*IPT tuples per CTA:
//      load (16-byte loads, registers) -> shared-memory histogram that also yields each tuple's
//      rank inside its digit -> block scan of the 2^bits counts + ONE global ticket per
//      non-empty digit (atomicAdd on the partition cursor; partitioning needs no stable order,
//      so no inter-tile dependency chain exists at all) -> tuples permuted into shared memory
//      grouped by digit -> written out so each digit's tuples form one contiguous run in HBM
//      (avg run = tile/fanout tuples x 8 B).
//    Algorithmic bytes: 16 per tuple (8 read + 8 written).
//    MODE 0: rank returned by the histogram atomic, kept in registers.
//    MODE 1: count first, second shared atomic on a per-digit cursor yields the slot.
//    OUT 0:  every thread stores tuples of consecutive tile slots (8-byte stores, one run per
//            group of lanes).
//    OUT 1:  TMA: each digit's run is laid out in shared memory with the same 16-byte phase as
//            its destination and leaves the SM as ONE bulk async copy (cp.async.bulk, UBLKCP)
//            issued by the thread that owns the digit; odd head/tail tuples use plain stores.
//    Pass 1 (tiles == nullptr): tiles cover the whole input.  Pass 2: tile descriptors send each
//    CTA to a chunk of ONE first-pass partition; cursors are the fine (2^B) cursors.
//    With `dst_bases` the output base pointer is chosen per digit (multi-GPU: peer receive
//    buffers mapped over NVLink) -- the all-to-all is the scatter itself.
// ------------------------------------------------------------------------------------------
struct ScatterArgs {
    const int32_t* in_keys;      // columnar input (COLUMNAR)
    const int32_t* in_pays;
    const tup_t* in_tup;         // packed input (!COLUMNAR), 16-byte aligned
    tup_t* out;                  // packed output, 16-byte aligned
    tup_t* const* dst_bases;     // optional per-digit output bases (device array of 2^bits pointers)
    uint32_t n;
    uint32_t shift, bits;
    uint32_t* cursors;
    uint32_t cursor_stride;      // words between consecutive cursors
    uint32_t ntiles;             // pass 1: number of tiles (the grid may be smaller: CTAs loop)
    const uint4* tiles;          // pass 2 only
    const uint32_t* num_tiles;   // pass 2 only
    // PUSH (last pass of the sharded "partition, then push" pipeline): the output base is chosen
    // per TILE -- the tile's first-pass partition belongs to one destination GPU, whose final
    // partition buffer (local or mapped over NVLink) receives the runs.
    tup_t* const* part_bases;    // [n_dest] final partition buffers
    uint32_t part_shift;         // destination = cursor base >> part_shift
    uint32_t n_dest;
    const uint32_t* abort_flag;  // non-zero: a destination would overflow, nothing is written
};

template <int THREADS, int IPT, int MODE, int OUT, bool COLUMNAR, int MINB, bool PERSIST = false, int NBT = NB_MAX, bool PUSH = false>
__global__ void __launch_bounds__(THREADS, MINB)
scatter_kernel(ScatterArgs a) {
    constexpr uint32_t T = THREADS * IPT;
    static_assert(THREADS >= NBT, "one thread per digit in the scan step");
    static_assert(IPT % 4 == 0, "vector loads");
    static_assert(!PUSH || (!COLUMNAR && !PERSIST), "the push pass reads first-pass output, one tile per CTA");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    tup_t* tile = reinterpret_cast<tup_t*>(smem_raw);   // T (+ 2*NBT padding slots for OUT 1)
    __shared__ uint32_t s_hist[NBT];                    // NBT = largest fan-out of this variant (256, 512 or 1024)
    __shared__ uint32_t s_lbase[MODE == 0 ? NBT : 1];
    __shared__ tup_t* s_dst[NBT];
    __shared__ uint32_t s_warp[NBT / 32];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t nb = 1u << a.bits, mask = nb - 1u;

    // One tile per CTA (the grid covers all tiles).  PERSIST: a smaller grid loops over the tiles
    // -- the multi-GPU peer scatter can leave SM resources to concurrently running kernels.  (As a
    // run-time loop in every variant it cost the one-tile launches ~1.5 %, hence the template.)
    const uint32_t ntiles_total = (a.tiles == nullptr) ? a.ntiles : *a.num_tiles;
    const uint32_t first_tile = blockIdx.x;
    if (PUSH) {   // (the tile list is already ordered destination-interleaved, see tile_perm)
        if (blockIdx.x >= ntiles_total || *a.abort_flag) return;
    } else if (a.abort_flag != nullptr && *a.abort_flag) return;   // sharded source pass: a destination would overflow
    for (uint32_t tile_id = first_tile; tile_id < ntiles_total; tile_id += gridDim.x) {
    // ---- which slots does this tile cover ----
    uint32_t a0, lo, hi, cbase;
    const tup_t* in_tup = a.in_tup;
    if (a.tiles == nullptr) {
        // caller-provided packed tuples may be only 8-byte aligned: step back one tuple so that
        // slot pairs are 16-byte aligned, and treat slot 0 as not ours
        const uint32_t mis = (!COLUMNAR && ((size_t)in_tup & 8u)) ? 1u : 0u;
        in_tup -= mis;
        const unsigned long long s = (unsigned long long)tile_id * T;
        const unsigned long long n_eff = (unsigned long long)a.n + mis;
        a0 = (uint32_t)s; lo = max(a0, mis);
        hi = (s < n_eff) ? a0 + (uint32_t)min((unsigned long long)T, n_eff - s) : a0;
        cbase = 0;
    } else {
        const uint4 td = __ldg(a.tiles + tile_id);
        a0 = td.x; lo = td.y; hi = td.z; cbase = td.w;
    }
    if (hi <= lo) { if (PERSIST) continue; else return; }
    if (tid < NBT) s_hist[tid] = 0;
    __syncthreads();
    const bool full = (lo == a0) && (hi - a0 == T);

    // ---- load: item j of this thread is slot a0 + slot_of(j) ----
    uint32_t key[IPT], pay[IPT];
    bool vec = true;
    if (COLUMNAR) {
        vec = full && ((((size_t)(a.in_keys + a0) | (size_t)(a.in_pays + a0)) & 15u) == 0);
        if (vec) {
            const int4* kv = reinterpret_cast<const int4*>(a.in_keys + a0);
            const int4* pv = reinterpret_cast<const int4*>(a.in_pays + a0);
#pragma unroll
            for (int j = 0; j < IPT / 4; ++j) {
                const int4 k = __ldg(kv + j * THREADS + tid);
                key[4 * j] = k.x; key[4 * j + 1] = k.y; key[4 * j + 2] = k.z; key[4 * j + 3] = k.w;
            }
#pragma unroll
            for (int j = 0; j < IPT / 4; ++j) {
                const int4 p = __ldg(pv + j * THREADS + tid);
                pay[4 * j] = p.x; pay[4 * j + 1] = p.y; pay[4 * j + 2] = p.z; pay[4 * j + 3] = p.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                const uint32_t i = a0 + j * THREADS + tid;
                if (i >= lo && i < hi) {
                    key[j] = (uint32_t)__ldg(a.in_keys + i);
                    pay[j] = (uint32_t)__ldg(a.in_pays + i);
                } else { key[j] = 0; pay[j] = 0; }
            }
        }
    } else {
        // a0 is even and the buffer 16-byte aligned: one 16-byte load = two tuples
        const uint4* tv = reinterpret_cast<const uint4*>(in_tup + a0);
#pragma unroll
        for (int j = 0; j < IPT / 2; ++j) {
            const uint32_t pi = j * THREADS + tid;
            const uint32_t s0 = a0 + 2 * pi;
            uint4 t = make_uint4(0, 0, 0, 0);
            if (full || (s0 >= lo && s0 + 1 < hi)) t = __ldg(tv + pi);
            else {   // tile edge: never touch a tuple outside [lo, hi)
                if (s0 >= lo && s0 < hi) { const tup_t e = __ldg(in_tup + s0); t.x = e.x; t.y = e.y; }
                if (s0 + 1 >= lo && s0 + 1 < hi) { const tup_t e = __ldg(in_tup + s0 + 1); t.z = e.x; t.w = e.y; }
            }
            key[2 * j] = t.x; pay[2 * j] = t.y; key[2 * j + 1] = t.z; pay[2 * j + 1] = t.w;
        }
    }
    // slot index of item j inside [a0, a0+T)
#define GJ_SLOT(j) (COLUMNAR ? (vec ? 4u * (((uint32_t)(j) >> 2) * THREADS + tid) + ((uint32_t)(j) & 3u)   \
                                    : (uint32_t)(j) * THREADS + tid)                                       \
                             : 2u * (((uint32_t)(j) >> 1) * THREADS + tid) + ((uint32_t)(j) & 1u))
#define GJ_VALID(j) (full || (a0 + GJ_SLOT(j) >= lo && a0 + GJ_SLOT(j) < hi))

    // ---- per-tile histogram (+ rank) ----
    uint32_t rk[MODE == 0 ? IPT : 1];
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        if (GJ_VALID(j)) {
            const uint32_t d = (key[j] >> a.shift) & mask;
            if (MODE == 0) rk[j] = atomicAdd(&s_hist[d], 1u);
            else atomicAdd(&s_hist[d], 1u);
        }
    }
    __syncthreads();

    // ---- scan of the digit counts, global tickets ----
    // OUT 1 pads every digit's region to an even number of slots (+ room for a 1-slot phase
    // shift) so that region starts are 16-byte aligned in shared memory.
    uint32_t cnt = 0, sz = 0, incl = 0;
    if (tid < NBT) {
        cnt = (tid < nb) ? s_hist[tid] : 0u;
        sz = (OUT == 1) ? ((cnt + 2u) & ~1u) : cnt;
        incl = warp_incl_scan(sz, lane);
        if (lane == 31) s_warp[wid] = incl;
    }
    __syncthreads();
    uint32_t gb = 0, sbeg = 0;
    tup_t* dbase = nullptr;
    if (tid < nb) {
        uint32_t woff = 0;
#pragma unroll
        for (uint32_t w = 0; w < NBT / 32; ++w)
            if (w < wid) woff += s_warp[w];
        const uint32_t excl = incl - sz + woff;
        if (cnt) gb = atomicAdd(&a.cursors[(size_t)(cbase + tid) * a.cursor_stride], cnt);
        dbase = PUSH ? a.part_bases[cbase >> a.part_shift] : (a.dst_bases ? a.dst_bases[tid] : a.out);
        sbeg = (OUT == 1) ? excl + (gb & 1u) : excl;   // same 16-byte phase as the destination
        // tile slot i (i >= sbeg for this digit) goes to dbase[gb + (i - sbeg)]
        s_dst[tid] = reinterpret_cast<tup_t*>(reinterpret_cast<unsigned long long>(dbase) +
                                              ((long long)gb - (long long)sbeg) * (long long)sizeof(tup_t));
        if (MODE == 0) s_lbase[tid] = sbeg; else s_hist[tid] = sbeg;
    }
    __syncthreads();

    // ---- permute into shared memory, grouped by digit ----
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        if (GJ_VALID(j)) {
            const uint32_t d = (key[j] >> a.shift) & mask;
            uint32_t pos;
            if (MODE == 0) pos = s_lbase[d] + rk[j];
            else pos = atomicAdd(&s_hist[d], 1u);
            tile[pos] = make_uint2(key[j], pay[j]);
        }
    }
#undef GJ_VALID
#undef GJ_SLOT
    if (OUT == 1) fence_proxy_async();   // generic-proxy writes -> visible to the bulk copy engine
    __syncthreads();

    if (OUT == 0) {
        // ---- write out: consecutive threads -> consecutive slots of the same run ----
        const uint32_t count = hi - lo;
#pragma unroll 4
        for (uint32_t i = tid; i < count; i += THREADS) {
            const tup_t t = tile[i];
            const uint32_t d = (t.x >> a.shift) & mask;
            s_dst[d][i] = t;
        }
    } else {
        // ---- one bulk copy per digit run ----
        if (tid < nb && cnt) {
            const tup_t* src = tile + sbeg;
            tup_t* dst = dbase + gb;
            uint32_t n = cnt;
            if (gb & 1u) { *dst = *src; ++src; ++dst; --n; }
            const uint32_t body = n & ~1u;
            if (body) bulk_s2g(dst, src, body * (uint32_t)sizeof(tup_t));
            if (n & 1u) dst[n - 1] = src[n - 1];
            bulk_commit();
            bulk_wait_read0();   // shared memory must stay valid until the engine has read it
        }
    }
    if (!PERSIST) break;
    }   // tile loop (the barrier at the top of the next iteration orders the reuse of shared memory)
}

// ------------------------------------------------------------------------------------------
// 5. Per-partition hash join: persistent CTAs (one per SM), blocks of 8 consecutive units dealt round
//    robin, fed by TMA bulk copies into two shared-memory rings.
//    A unit {probe range, build partition} is processed in steps of (build chunk <= CAP tuples)
//    x (probe chunk <= U tuples).  A dedicated LOADER WARP (one lane) runs an iterator ahead of
//    the consumer warps and issues bulk async copies (cp.async.bulk, completion on "full"
//    mbarriers; slots come back through "empty" mbarriers; the consumers synchronise among
//    themselves with a named barrier, so the ~1.3k-cycle issue path is off their critical
//    path -- as thread 0's side job it made 31 warps wait at the post-build barrier, ncu):
//    every step's probe chunk goes into
//    the S ring (NS slots); a build chunk goes into the R ring (NR slots) only when it differs
//    from the previous step's -- a build partition that is probed by many chunks / units
//    (large or skewed probe sides) is loaded and hashed ONCE.  All HBM traffic of the join is
//    asynchronous bulk traffic: no load instructions, no register staging, latency hidden by
//    the ring depth.
//    Per step the CTA waits on the slot's mbarrier; on a new build chunk it builds the chained
//    table in place over the staged tuples (atomicExch on the head, 16-bit next links: index
//    chains need no sentinel key and are N:M safe; a MULTI bit in the head marks buckets with
//    more than one entry so the common single-entry probe never reads a link); then probes with
//    the staged probe tuples (full 32-bit key compare, 64-bit per-thread accumulators).  Heads
//    carry a 15-bit version = build chunk number, so the table is never cleared: 2 barriers per
//    new build chunk, 1 per further probe chunk.
//    Hash = xor-fold of the key bits above the radix (+GPU) field: the identity on dense keys
//    (the reference's choice, common.h:45-47), a real hash otherwise.
//    Build partitions larger than CAP simply produce more steps (the reference's block-nested
//    branch, join-primitives.cu:929-1003).
//    MATERIALIZE: result pairs are staged per CTA in shared memory and flushed with ONE global
//    reservation per flush as coalesced column writes; pairs beyond `cap` are counted, not
//    written.  Algorithmic bytes: 8 per input tuple (+ 8 per result pair when materialising).
// ------------------------------------------------------------------------------------------
struct JoinArgs {
    const tup_t* bld; const tup_t* prb;     // partitioned tuples (16-byte aligned, +2 slack)
    const uint4* units; const uint32_t* num_units;
    uint32_t hash_shift;
    uint32_t unit_block;          // consecutive units a CTA takes at a time (1 or JOIN_UNIT_BLOCK)
    unsigned long long* result;   // [0] matches [1] checksum [2] pairs reserved (materialise)
    int32_t* out_bld_pay; int32_t* out_prb_pay; unsigned long long cap;
    // LATE (late materialisation, join_partitioned_varpayload, join-primitives.cu:1420-1557): payloads are row
    // ids into column-major side tables; a result pair adds cols[z * stride + id] of both sides
    const int32_t* bld_cols; const int32_t* prb_cols;
    uint32_t ncols_bld, ncols_prb;
    unsigned long long stride_bld, stride_prb;
};

// what one result pair contributes to the aggregate: the payload product (reference
// join-primitives.cu:1073) or, LATE, the gathered side-table values (:1531-1536)
template <bool LATE>
__device__ __forceinline__ unsigned long long pair_value(const JoinArgs& a, uint32_t bld_pay, uint32_t prb_pay) {
    if (!LATE) return (unsigned long long)((long long)(int32_t)bld_pay * (long long)(int32_t)prb_pay);
    long long v = 0;
    for (uint32_t z = 0; z < a.ncols_bld; ++z) v += (long long)__ldg(a.bld_cols + (size_t)z * a.stride_bld + bld_pay);
    for (uint32_t z = 0; z < a.ncols_prb; ++z) v += (long long)__ldg(a.prb_cols + (size_t)z * a.stride_prb + prb_pay);
    return (unsigned long long)v;
}

constexpr int JOIN_STAGE_PAIRS = 4096;   // staged result pairs per CTA (materialise): 32 KB
constexpr uint32_t STEP_DONE = 0xFFFFFFFFu;
constexpr uint32_t JOIN_UNIT_BLOCK = 8;   // consecutive units a CTA takes at a time when probe partitions span several units
// head word = MULTI(1) | version(15) | entry index(16).  A head is live only if its version equals
// the current build chunk's, so the table is never cleared between chunks.
constexpr uint32_t HEAD_MULTI = 0x80000000u;   // bucket holds more than one entry
constexpr uint32_t HEAD_VER_MASK = 0x7FFF0000u;

template <int CAP, int U, int NR, int NS, bool MATERIALIZE>
struct JoinSmem {
    static constexpr size_t rbuf_bytes = (size_t)(CAP + 2) * sizeof(tup_t);
    static constexpr size_t sbuf_bytes = (size_t)(U + 2) * sizeof(tup_t);
    static constexpr size_t off_s = rbuf_bytes * NR;
    static constexpr size_t off_head = off_s + sbuf_bytes * NS;
    static constexpr size_t off_next = off_head + (size_t)CAP * 4;
    static constexpr size_t off_out = off_next + (size_t)CAP * 2;
    static constexpr size_t off_hdr = off_out + (MATERIALIZE ? (size_t)JOIN_STAGE_PAIRS * 8 : 0);
    static constexpr size_t off_bar = off_hdr + (size_t)NS * 32;
    static constexpr size_t total = off_bar + (size_t)(NS + NR) * 16;   // full + empty barriers
};

template <int THREADS, int CAP, int U, int NR, int NS, bool MATERIALIZE, bool OPTIMISTIC, bool LATE = false>
__global__ void __launch_bounds__(THREADS, 1)
join_kernel(JoinArgs a) {
    static_assert(!(LATE && MATERIALIZE), "late materialisation aggregates");
    using L = JoinSmem<CAP, U, NR, NS, MATERIALIZE>;
    static_assert(CAP < 0xFFFF, "16-bit entry indices");
    constexpr uint32_t NC = THREADS - 32;      // consumer threads: warps 0 .. THREADS/32-2
    constexpr uint32_t BAR_C = 1;              // named barrier of the consumer warps
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* head = reinterpret_cast<uint32_t*>(smem_raw + L::off_head);   // [CAP]
    uint16_t* next = reinterpret_cast<uint16_t*>(smem_raw + L::off_next);   // [CAP]
    int32_t* s_out_b = reinterpret_cast<int32_t*>(smem_raw + L::off_out);
    int32_t* s_out_p = s_out_b + JOIN_STAGE_PAIRS;
    uint4* s_hdr = reinterpret_cast<uint4*>(smem_raw + L::off_hdr);           // [NS][2]
    uint64_t* s_sfull = reinterpret_cast<uint64_t*>(smem_raw + L::off_bar);   // [NS]
    uint64_t* s_rfull = s_sfull + NS;                                         // [NR]
    uint64_t* s_sempty = s_rfull + NR;                                        // [NS]
    uint64_t* s_rempty = s_sempty + NS;                                       // [NR]
    __shared__ uint32_t s_cnt, s_flush[2];
    __shared__ unsigned long long s_base;
    __shared__ unsigned long long s_red[2][THREADS / 32];

    const uint32_t tid = threadIdx.x;
    const uint32_t nunits = *a.num_units;
    unsigned long long matches = 0, sum = 0;
    uint32_t mat_round = 0;                    // materialise: rounds so far (uniform over the consumers)

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&s_sfull[s], 1); mbar_init(&s_sempty[s], 1); }
        for (int s = 0; s < NR; ++s) { mbar_init(&s_rfull[s], 1); mbar_init(&s_rempty[s], 1); }
        fence_mbar_init();
        if (MATERIALIZE) { s_cnt = 0; s_flush[0] = s_flush[1] = 0; }
    }
    // version 0 is never used by a chunk (versions are 1..0x7FFF), so an all-zero table is empty
    for (uint32_t i = tid; i < (uint32_t)CAP / 4; i += THREADS)
        reinterpret_cast<uint4*>(head)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();

    if (tid >= NC) {
        // ================= loader warp: one lane walks (unit, build chunk, probe chunk) ahead of
        // the consumers and feeds the rings; it never touches a tuple =================
        if (tid == NC) {
            // This CTA's units: blocks of a.unit_block consecutive units, dealt round robin.  With probe partitions
            // of several units (workload A: 8) the host picks blocks of 8: further units of the same partition
            // then share ONE load + build of its build chunk (unit-by-unit striding rebuilt it in every CTA), while
            // a hot partition's hundreds of units still spread over all CTAs (whole contiguous ranges per CTA
            // left the hot ranges to a few: measured slower under Zipf).  With one unit per partition (workload B)
            // blocks of 1 = plain striding: all CTAs stream one contiguous window of both relations.
            const uint32_t ub = a.unit_block;
            auto next_unit = [&](uint32_t u) -> uint32_t {
                const uint32_t v = u + 1u;
                return (v % ub) ? v : (u / ub + gridDim.x) * ub;
            };
            uint32_t it_u = blockIdx.x * ub;
            const uint32_t it_end = nunits;
            bool it_valid = it_u < it_end;
            uint4 it_d = make_uint4(0, 0, 0, 0), it_dn = make_uint4(0, 0, 0, 0);
            if (it_valid) {
                it_d = __ldg(a.units + it_u);
                if (next_unit(it_u) < it_end) it_dn = __ldg(a.units + next_unit(it_u));
            }
            uint32_t it_rc = it_d.z, it_sc = it_d.x;
            uint32_t chunks = 0;          // build chunks issued so far
            bool newc = true;             // the next step starts a new build chunk
            for (uint32_t p = 0;; ++p) {
                const uint32_t sslot = p % NS;
                if (p >= (uint32_t)NS) mbar_wait(&s_sempty[sslot], ((p / NS) - 1u) & 1u);
                if (!it_valid) {
                    s_hdr[2 * sslot] = make_uint4(STEP_DONE, 0, 0, 0);
                    mbar_arrive_expect_tx(&s_sfull[sslot], 0);
                    break;
                }
                const uint32_t nr = min((uint32_t)CAP, it_d.w - it_rc), ns = min((uint32_t)U, it_d.y - it_sc);
                const uint32_t chunk = newc ? chunks : chunks - 1u;
                const uint32_t rc = it_rc, sc = it_sc;
                const bool this_new = newc;
                // advance: probe chunks innermost, then build chunks, then this CTA's next unit; the
                // step is the last of its build chunk iff the next step loads a different one
                bool lastc = false;
                it_sc += U;
                if (it_sc >= it_d.y) {
                    it_sc = it_d.x;
                    it_rc += CAP;
                    lastc = true;
                    if (it_rc >= it_d.w) {
                        it_u = next_unit(it_u);
                        it_valid = it_u < it_end;
                        const uint4 prev = it_d;
                        it_d = it_dn;                               // prefetched one unit ahead
                        it_rc = it_d.z; it_sc = it_d.x;
                        if (it_valid && next_unit(it_u) < it_end) it_dn = __ldg(a.units + next_unit(it_u));
                        // consecutive units of one (hot) partition share a single-chunk build side
                        if (it_valid && it_d.z == prev.z && it_d.w == prev.w && prev.w - prev.z <= (uint32_t)CAP) lastc = false;
                    }
                }
                newc = lastc;
                const uint32_t rskip = rc & 1u, sskip = sc & 1u;
                s_hdr[2 * sslot] = make_uint4(ns, sskip, (this_new ? 1u : 0u) | (lastc ? 2u : 0u), chunk);
                s_hdr[2 * sslot + 1] = make_uint4(nr, rskip, 0, 0);
                if (this_new) {
                    const uint32_t rslot = chunk % NR;
                    // the slot's previous tenant (chunk - NR) must have been released
                    if (chunk >= (uint32_t)NR) mbar_wait(&s_rempty[rslot], ((chunk / NR) - 1u) & 1u);
                    const uint32_t rbytes = ((nr + rskip + 1u) & ~1u) * (uint32_t)sizeof(tup_t);
                    mbar_arrive_expect_tx(&s_rfull[rslot], rbytes);
                    bulk_g2s(smem_raw + L::rbuf_bytes * rslot, a.bld + (rc - rskip), rbytes, &s_rfull[rslot]);
                    ++chunks;
                }
                const uint32_t sbytes = ((ns + sskip + 1u) & ~1u) * (uint32_t)sizeof(tup_t);
                mbar_arrive_expect_tx(&s_sfull[sslot], sbytes);
                bulk_g2s(smem_raw + L::off_s + L::sbuf_bytes * sslot, a.prb + (sc - sskip), sbytes, &s_sfull[sslot]);
            }
        }
    } else {
        // ================= consumer warps =================
        for (uint32_t k = 0;; ++k) {
            const uint32_t sslot = k % NS;
            mbar_wait(&s_sfull[sslot], (k / NS) & 1u);
            const uint4 h0 = s_hdr[2 * sslot];
            if (h0.x == STEP_DONE) break;
            const uint4 h1 = s_hdr[2 * sslot + 1];
            const uint32_t ns = h0.x, nr = h1.x, chunk = h0.w, rslot = chunk % NR;
            const bool newc = (h0.z & 1u) != 0;
            const tup_t* rbuf = reinterpret_cast<const tup_t*>(smem_raw + L::rbuf_bytes * rslot) + h1.y;
            const tup_t* sbuf = reinterpret_cast<const tup_t*>(smem_raw + L::off_s + L::sbuf_bytes * sslot) + h0.y;
            const uint32_t hb = 32u - __clz(max(nr, 32u) - 1u);   // ceil(log2(nr)), >= 5
            const uint32_t hmask = (1u << hb) - 1u;
            const uint32_t ver = ((chunk % 0x7FFFu) + 1u) << 16;  // 1..0x7FFF in bits 16..30

            if (newc) {
                if (chunk && (chunk % 0x7FFFu) == 0u) {   // version wrap: wipe stale heads (rare)
                    for (uint32_t i = tid; i < (uint32_t)CAP / 4; i += NC)
                        reinterpret_cast<uint4*>(head)[i] = make_uint4(0u, 0u, 0u, 0u);
                    named_bar_sync(BAR_C, NC);
                }
                mbar_wait(&s_rfull[rslot], (chunk / NR) & 1u);
                constexpr int KB = (CAP + NC - 1) / NC;
                uint32_t bk[KB];
#pragma unroll
                for (int q = 0; q < KB; ++q) {
                    const uint32_t i = q * NC + tid;
                    bk[q] = (i < nr) ? rbuf[i].x : 0u;
                }
                if (OPTIMISTIC) {
                    // Optimistic build: plain stores, last writer wins; a thread whose entry
                    // survived owns a single-entry bucket; losers (collisions, duplicate keys)
                    // are chained in with the atomic path afterwards.  (Measured: slower than
                    // the atomic build on B200 -- the extra barrier costs more than the atomics.)
                    uint32_t hq[KB];
#pragma unroll
                    for (int q = 0; q < KB; ++q) {
                        const uint32_t i = q * NC + tid;
                        const uint32_t kk = bk[q] >> a.hash_shift;
                        hq[q] = (kk ^ (kk >> hb)) & hmask;
                        if (i < nr) head[hq[q]] = ver | i;
                    }
                    named_bar_sync(BAR_C, NC);
                    uint32_t lost = 0;
#pragma unroll
                    for (int q = 0; q < KB; ++q) {
                        const uint32_t i = q * NC + tid;
                        if (i < nr) {
                            if (head[hq[q]] == (ver | i)) next[i] = (uint16_t)0xFFFFu;
                            else lost |= 1u << q;
                        }
                    }
                    named_bar_sync(BAR_C, NC);
#pragma unroll
                    for (int q = 0; q < KB; ++q) {
                        if (lost & (1u << q)) {
                            const uint32_t i = q * NC + tid;
                            uint32_t* hp = &head[hq[q]];
                            const uint32_t old = atomicExch(hp, ver | i);
                            next[i] = (uint16_t)old;      // the winner or an earlier loser: always live
                            atomicOr(hp, HEAD_MULTI);
                        }
                    }
                    named_bar_sync(BAR_C, NC);
                } else {
#pragma unroll
                    for (int q = 0; q < KB; ++q) {
                        const uint32_t i = q * NC + tid;
                        if (i < nr) {
                            const uint32_t kk = bk[q] >> a.hash_shift;
                            uint32_t* hp = &head[(kk ^ (kk >> hb)) & hmask];
                            const uint32_t old = atomicExch(hp, ver | i);
                            const bool live = (old & HEAD_VER_MASK) == ver;
                            next[i] = live ? (uint16_t)old : (uint16_t)0xFFFFu;
                            if (live) atomicOr(hp, HEAD_MULTI);
                        }
                    }
                    named_bar_sync(BAR_C, NC);
                }
            }
            if (!MATERIALIZE) {
                // straight-line probe of up to KP tuples per thread: all shared-memory loads of one
                // kind are issued back to back; only multi-entry buckets take the chain loop
                constexpr int KP = (U + NC - 1) / NC;
                tup_t t[KP], r[KP];
                uint32_t w[KP];
                uint32_t m32 = 0;
#pragma unroll
                for (int q = 0; q < KP; ++q) {
                    const uint32_t j = q * NC + tid;
                    t[q] = (j < ns) ? sbuf[j] : make_uint2(0u, 0u);
                }
#pragma unroll
                for (int q = 0; q < KP; ++q) {
                    const uint32_t j = q * NC + tid;
                    const uint32_t kk = t[q].x >> a.hash_shift;
                    w[q] = (j < ns) ? head[(kk ^ (kk >> hb)) & hmask] : 0u;
                }
#pragma unroll
                for (int q = 0; q < KP; ++q) {
                    const bool live = (w[q] & HEAD_VER_MASK) == ver;
                    r[q] = rbuf[live ? (w[q] & 0xFFFFu) : 0u];
                }
                if (LATE) {
                    // the head-of-bucket hits of all KP tuples gather together: per side-table column the
                    // thread issues its (up to KP) independent loads back to back before it adds them up, so
                    // KP random HBM accesses are in flight per thread instead of one (the gathers, not the
                    // join, are this kernel's cost: 4 x 32-byte sectors per result pair with 2 + 2 columns)
                    bool hit[KP];
#pragma unroll
                    for (int q = 0; q < KP; ++q) hit[q] = (w[q] & HEAD_VER_MASK) == ver && r[q].x == t[q].x;
                    for (uint32_t z = 0; z < a.ncols_bld; ++z) {
                        int32_t v[KP];
#pragma unroll
                        for (int q = 0; q < KP; ++q) v[q] = hit[q] ? __ldg(a.bld_cols + (size_t)z * a.stride_bld + r[q].y) : 0;
#pragma unroll
                        for (int q = 0; q < KP; ++q) sum += (unsigned long long)(long long)v[q];
                    }
                    for (uint32_t z = 0; z < a.ncols_prb; ++z) {
                        int32_t v[KP];
#pragma unroll
                        for (int q = 0; q < KP; ++q) v[q] = hit[q] ? __ldg(a.prb_cols + (size_t)z * a.stride_prb + t[q].y) : 0;
#pragma unroll
                        for (int q = 0; q < KP; ++q) sum += (unsigned long long)(long long)v[q];
                    }
                }
#pragma unroll
                for (int q = 0; q < KP; ++q) {
                    if ((w[q] & HEAD_VER_MASK) == ver) {
                        if (r[q].x == t[q].x) {
                            ++m32;
                            if (!LATE) sum += pair_value<false>(a, r[q].y, t[q].y);
                        }
                        if (w[q] & HEAD_MULTI) {
                            for (uint32_t i = next[w[q] & 0xFFFFu]; i != 0xFFFFu; i = next[i]) {
                                const tup_t rr = rbuf[i];
                                if (rr.x == t[q].x) {
                                    ++m32;
                                    sum += pair_value<LATE>(a, rr.y, t[q].y);
                                }
                            }
                        }
                    }
                }
                matches += m32;
            }
            const uint32_t rounds = (ns + NC - 1) / NC;
            for (uint32_t q = 0; MATERIALIZE && q < rounds; ++q, ++mat_round) {
                // One probe tuple per thread and round.  The chain walk is warp-uniform: every iteration
                // the warp ballots its hits and ONE lane reserves staging slots for all of them
                // (north_star (2); the reference does the same per 16-pair warp buffer,
                // join-primitives.cu:1212-1262).
                const uint32_t j = q * NC + tid;
                tup_t t = make_uint2(0u, 0u);
                uint32_t w = 0, i = 0xFFFFu;
                if (j < ns) {
                    t = sbuf[j];
                    const uint32_t kk = t.x >> a.hash_shift;
                    w = head[(kk ^ (kk >> hb)) & hmask];
                    if ((w & HEAD_VER_MASK) == ver) i = w & 0xFFFFu;
                }
                const uint32_t lane = tid & 31u;
                while (__any_sync(0xffffffffu, i != 0xFFFFu)) {
                    tup_t r = make_uint2(0u, 0u);
                    bool hit = false;
                    if (i != 0xFFFFu) {
                        r = rbuf[i];
                        hit = (r.x == t.x);
                        i = (w & HEAD_MULTI) ? (uint32_t)next[i] : 0xFFFFu;
                    }
                    const uint32_t hits = __ballot_sync(0xffffffffu, hit);
                    if (hits) {
                        const int leader = __ffs(hits) - 1;
                        uint32_t base = 0;
                        if ((int)lane == leader) base = atomicAdd(&s_cnt, (uint32_t)__popc(hits));
                        base = __shfl_sync(0xffffffffu, base, leader);
                        if (hit) {
                            ++matches;
                            sum += pair_value<false>(a, r.y, t.y);
                            const uint32_t pos = base + (uint32_t)__popc(hits & ((1u << lane) - 1u));
                            if (pos < (uint32_t)JOIN_STAGE_PAIRS) {
                                s_out_b[pos] = (int32_t)r.y;
                                s_out_p[pos] = (int32_t)t.y;
                            } else {   // staging full inside a round (N:M bursts): rare direct path
                                const unsigned long long g = atomicAdd(&a.result[2], 1ull);
                                if (g < a.cap) {
                                    a.out_bld_pay[g] = (int32_t)r.y;
                                    a.out_prb_pay[g] = (int32_t)t.y;
                                }
                            }
                        }
                    }
                }
                named_bar_sync(BAR_C, NC);
                // Flush decision: must be the SAME in every consumer thread (the flush below contains
                // barriers).  s_cnt itself may already be growing again (a fast warp is in the next
                // round), so threads do not look at it: thread 0 alone samples it and publishes the
                // verdict one round ahead in s_flush[parity]; the barrier above orders its write before
                // everybody's read.  With <= 1 match per probe tuple the staging area never overflows
                // (flush once more than STAGE - 2 NC pairs may be staged); beyond that the direct path
                // above keeps the result exact.
                const bool do_flush = s_flush[mat_round & 1u] != 0u;
                if (do_flush) {
                    const uint32_t c = min(s_cnt, (uint32_t)JOIN_STAGE_PAIRS);   // stable: every consumer is between the barriers
                    if (tid == 0) s_base = atomicAdd(&a.result[2], (unsigned long long)c);
                    named_bar_sync(BAR_C, NC);
                    const unsigned long long g0 = s_base;
                    for (uint32_t x = tid; x < c; x += NC) {
                        if (g0 + x < a.cap) {
                            a.out_bld_pay[g0 + x] = s_out_b[x];
                            a.out_prb_pay[g0 + x] = s_out_p[x];
                        }
                    }
                    named_bar_sync(BAR_C, NC);
                    if (tid == 0) { s_cnt = 0; s_flush[(mat_round + 1u) & 1u] = 0u; }
                    named_bar_sync(BAR_C, NC);
                } else if (tid == 0) {
                    s_flush[(mat_round + 1u) & 1u] = (s_cnt + 2u * NC > (uint32_t)JOIN_STAGE_PAIRS) ? 1u : 0u;
                }
            }
            named_bar_sync(BAR_C, NC);       // every consumer is done with this step's slots (and the table)
            if (tid == 0) {                  // hand the slots back to the loader
                mbar_arrive(&s_sempty[sslot]);
                if (h0.z & 2u) mbar_arrive(&s_rempty[rslot]);
            }
        }
    }

    __syncthreads();
    if (MATERIALIZE) {
        const uint32_t c = min(s_cnt, (uint32_t)JOIN_STAGE_PAIRS);
        if (c) {
            if (tid == 0) s_base = atomicAdd(&a.result[2], (unsigned long long)c);
            __syncthreads();
            const unsigned long long g0 = s_base;
            for (uint32_t i = tid; i < c; i += THREADS) {
                if (g0 + i < a.cap) {
                    a.out_bld_pay[g0 + i] = s_out_b[i];
                    a.out_prb_pay[g0 + i] = s_out_p[i];
                }
            }
        }
    }
    // block reduction of the two 64-bit accumulators
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        matches += __shfl_xor_sync(0xffffffffu, matches, o);
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if ((tid & 31u) == 0) { s_red[0][tid >> 5] = matches; s_red[1][tid >> 5] = sum; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long m = 0, s = 0;
        for (int w = 0; w < THREADS / 32; ++w) { m += s_red[0][w]; s += s_red[1][w]; }
        if (m) atomicAdd(&a.result[0], m);
        if (s) atomicAdd(&a.result[1], s);
    }
}

// ------------------------------------------------------------------------------------------
// 5b. Non-partitioned baseline (reference build_ht_chains / chains_probing,
//     join-primitives.cu:681-742): one chained hash table over the whole build side in global
//     memory -- heads[h] = index + 1 of the newest entry (0 = empty, so a memset clears it),
//     next[i] = its predecessor -- probed straight from the probe columns.  No radix pass at
//     all: 3 launches; on B200 a build side of a few million tuples keeps heads + columns in the
//     126 MB L2, which is the regime where this beats partitioning (config 1).
//     Hash = xor-fold to hb bits: the identity on dense keys below 2^hb (the reference masks).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t np_hash(uint32_t key, uint32_t hb) { return (key ^ (key >> hb)) & ((1u << hb) - 1u); }

__global__ void __launch_bounds__(256)
np_build_kernel(const int32_t* __restrict__ keys, uint32_t n, uint32_t hb, uint32_t* __restrict__ heads,
                uint32_t* __restrict__ next) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t old = atomicExch(&heads[np_hash((uint32_t)__ldg(keys + i), hb)], i + 1u);
        next[i] = old;
    }
}

__global__ void __launch_bounds__(256)
np_probe_kernel(const int32_t* __restrict__ bk, const int32_t* __restrict__ bp, const uint32_t* __restrict__ heads,
                const uint32_t* __restrict__ next, uint32_t hb, const int32_t* __restrict__ pk,
                const int32_t* __restrict__ pp, uint32_t n, unsigned long long* __restrict__ result) {
    __shared__ unsigned long long s_red[2][8];
    unsigned long long matches = 0, sum = 0;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int32_t key = __ldg(pk + j), pay = __ldg(pp + j);
        for (uint32_t e = heads[np_hash((uint32_t)key, hb)]; e != 0u; e = next[e - 1u]) {
            if (bk[e - 1u] == key) {
                ++matches;
                sum += (unsigned long long)((long long)bp[e - 1u] * (long long)pay);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        matches += __shfl_xor_sync(0xffffffffu, matches, o);
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if ((threadIdx.x & 31u) == 0) { s_red[0][threadIdx.x >> 5] = matches; s_red[1][threadIdx.x >> 5] = sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long m = 0, s = 0;
        for (int w = 0; w < 8; ++w) { m += s_red[0][w]; s += s_red[1][w]; }
        if (m) atomicAdd(&result[0], m);
        if (s) atomicAdd(&result[1], s);
    }
}

// 5c. Perfect-array variant (reference build_perfect_array / probe_perfect_array,
//     join-primitives.cu:628-668): build keys are unique and lie in [key_min, key_min + range), so
//     the key itself addresses a table slot -- no hash, no chain, no key compare.  A slot is the
//     64-bit word (1 << 32) | payload, 0 = empty (the reference stores payload + 1 in an int32 and
//     so loses the payload -1).  The build checks its own precondition: status[0] counts keys out
//     of range, status[1] duplicates (atomicExch returned a live slot).
__global__ void __launch_bounds__(256)
np_build_perfect_kernel(const int32_t* __restrict__ keys, const int32_t* __restrict__ pays, uint32_t n, int32_t key_min,
                        uint32_t range, unsigned long long* __restrict__ slots, uint32_t* __restrict__ status) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t k = (uint32_t)__ldg(keys + i) - (uint32_t)key_min;
        if (k >= range) { atomicAdd(&status[0], 1u); continue; }
        const unsigned long long old = atomicExch(&slots[k], (1ull << 32) | (uint32_t)__ldg(pays + i));
        if (old) atomicAdd(&status[1], 1u);
    }
}

__global__ void __launch_bounds__(256)
np_probe_perfect_kernel(const unsigned long long* __restrict__ slots, int32_t key_min, uint32_t range,
                        const int32_t* __restrict__ pk, const int32_t* __restrict__ pp, uint32_t n,
                        unsigned long long* __restrict__ result) {
    __shared__ unsigned long long s_red[2][8];
    unsigned long long matches = 0, sum = 0;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint32_t k = (uint32_t)__ldg(pk + j) - (uint32_t)key_min;
        if (k < range) {
            const unsigned long long sl = __ldg(slots + k);
            if (sl) {
                ++matches;
                sum += (unsigned long long)((long long)(int32_t)(uint32_t)sl * (long long)__ldg(pp + j));
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        matches += __shfl_xor_sync(0xffffffffu, matches, o);
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if ((threadIdx.x & 31u) == 0) { s_red[0][threadIdx.x >> 5] = matches; s_red[1][threadIdx.x >> 5] = sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long m = 0, s = 0;
        for (int w = 0; w < 8; ++w) { m += s_red[0][w]; s += s_red[1][w]; }
        if (m) atomicAdd(&result[0], m);
        if (s) atomicAdd(&result[1], s);
    }
}

// ------------------------------------------------------------------------------------------
// 6. Synthetic unique relations: key = seeded bijection of the row id.
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// Cycle-walking 4-round balanced Feistel network on 2*half bits, half = ceil(log2(n)/2).
__host__ __device__ inline uint32_t bijection(uint64_t row, uint64_t n_total, uint32_t seed) {
    uint32_t bits = 2;
    while (bits < 32 && (1ull << bits) < n_total) ++bits;
    const uint32_t half = (bits + 1) >> 1;
    const uint32_t hmask = (1u << half) - 1u;
    uint64_t x = row;
    do {
        uint32_t L = (uint32_t)(x >> half) & hmask, R = (uint32_t)x & hmask;
        for (uint32_t r = 0; r < 4; ++r) {
            const uint32_t f = mix32(R + seed * 0x9E3779B9u + r * 0x85EBCA6Bu) & hmask;
            const uint32_t nl = R;
            R = L ^ f;
            L = nl;
        }
        x = ((uint64_t)L << half) | R;
    } while (x >= n_total);
    return (uint32_t)x;
}

__host__ __device__ __forceinline__ int32_t payload_of_key(uint32_t key, uint32_t pay_seed) {
    return (int32_t)mix32(key ^ (pay_seed * 0xC2B2AE35u + 0x27D4EB2Fu));
}

__global__ void generate_unique_kernel(int32_t* keys, int32_t* pays, uint64_t row_begin,
                                       uint64_t n_rows, uint64_t n_total, uint32_t seed,
                                       uint32_t pay_seed) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_rows;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t k = bijection(row_begin + i, n_total, seed);
        keys[i] = (int32_t)k;
        pays[i] = payload_of_key(k, pay_seed);
    }
}

// L2 flush for benchmarks: streaming write over a buffer larger than L2.
__global__ void flush_kernel(uint4* buf, size_t nvec, uint32_t v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec;
         i += (size_t)gridDim.x * blockDim.x)
        buf[i] = make_uint4(v, v, v, v);
}

}  // namespace gj
