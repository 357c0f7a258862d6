// api.cu -- host side of libgpujoin.so: the C ABI declared in include/gpujoin.h.
//
// Shape follows the reference's in-GPU operator outOfGPU_Join1_payload
// (/root/reference/src/hash_join_clustered_probe.cu:802-994): allocate device scratch, run
// partition(R), partition(S), join, read back the aggregate -- but allocation happens once in
// gj_create, all work is enqueued on one stream without host round trips, timing uses CUDA
// events, and errors are returned instead of exit()ed (common.h:132-141).
#include "../../include/gpujoin.h"
#include "kernels.cuh"

#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>

using namespace gj;

// ------------------------------------------------------------------------------------------
// errors, launch accounting
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(GJ_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,            \
                        cudaGetErrorString(e_));                                              \
    } while (0)
#define LAUNCHED()                                                                            \
    do {                                                                                      \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                   \
        ++ctx->launches;                                                                      \
        CK(cudaGetLastError());                                                               \
    } while (0)

// ------------------------------------------------------------------------------------------
// kernel variant tables
// ------------------------------------------------------------------------------------------
typedef void (*scatter_fn)(ScatterArgs);
struct ScatterCfg { int threads, ipt, mode, out; scatter_fn col, packed, col_persist; };
#define GJ_SC(T, I, M, O, B) { T, I, M, O, scatter_kernel<T, I, M, O, true, B>, scatter_kernel<T, I, M, O, false, B>, \
                               scatter_kernel<T, I, M, O, true, B, true> }
// The winners of the round-1 sweeps over 16 shapes (profiles/r1_final/sweep.log) plus one structurally different
// fallback; the losers (8-tuple-per-thread tiles, 1024-thread CTAs, 5-6 CTAs/SM) were 3-25 % slower and are gone.
static const ScatterCfg kScatter[] = {
    GJ_SC(256, 16, 1, 0, 4),   // 0 first pass: two shared atomics, 8-byte stores, 4 CTAs/SM (0.79 of HBM peak at 7 bits)
    GJ_SC(256, 16, 0, 0, 3),   // 1 fallback: rank kept in registers (one shared atomic per tuple), 3 CTAs/SM
    GJ_SC(256, 16, 1, 1, 4),   // 2 TMA bulk-copy output, 4 K-tuple tiles: last pass at <= 7 bits, peer-store shuffle
    GJ_SC(512, 16, 1, 1, 2),   // 3 TMA bulk-copy output, 8 K-tuple tiles: last pass at 8 bits
};
static const int kNumScatter = (int)(sizeof(kScatter) / sizeof(kScatter[0]));
static size_t scatter_smem(const ScatterCfg& c) {
    return ((size_t)c.threads * c.ipt + (c.out ? 2 * NB_MAX : 0)) * sizeof(tup_t);
}

// Sharded "partition, then push" pipeline: first pass (columnar input, local) and last pass
// (packed input, runs pushed into the destination GPU's partition buffer) by fan-out.
struct PPCfg { int threads, ipt, out, nbt; scatter_fn fn; };
static size_t pp_smem(const PPCfg& c) { return ((size_t)c.threads * c.ipt + (c.out ? 2 * c.nbt : 0)) * sizeof(tup_t); }
static const PPCfg kPPFirst[] = {   // indexed by max(bits, 8) - 8
    { 256, 16, 0, 256, scatter_kernel<256, 16, 1, 0, true, 4> },
    { 512, 16, 0, 512, scatter_kernel<512, 16, 1, 0, true, 2, false, 512> },
    { 1024, 8, 0, 1024, scatter_kernel<1024, 8, 1, 0, true, 1, false, 1024> },
};
static const PPCfg kPPPush[2][3] = {   // [8-byte stores | TMA bulk stores][max(bits, 8) - 8]
    { { 256, 16, 0, 256, scatter_kernel<256, 16, 1, 0, false, 4, false, 256, true> },
      { 512, 16, 0, 512, scatter_kernel<512, 16, 1, 0, false, 2, false, 512, true> },
      { 1024, 8, 0, 1024, scatter_kernel<1024, 8, 1, 0, false, 1, false, 1024, true> } },
    { { 512, 16, 1, 256, scatter_kernel<512, 16, 1, 1, false, 2, false, 256, true> },
      { 512, 16, 1, 512, scatter_kernel<512, 16, 1, 1, false, 2, false, 512, true> },
      { 1024, 8, 1, 1024, scatter_kernel<1024, 8, 1, 1, false, 1, false, 1024, true> } },
};
// 16 K-tuple tiles (one 1024-thread CTA per SM): twice / four times longer runs per destination
// partition -- NVLink moves long runs far better than short ones (2 GPUs, TMA stores: 128 B runs
// 422 GB/s, 16 KB runs 689 GB/s)
static const PPCfg kPPPushBig[2][3] = {
    { { 1024, 16, 0, 256, scatter_kernel<1024, 16, 1, 0, false, 1, false, 256, true> },
      { 1024, 16, 0, 512, scatter_kernel<1024, 16, 1, 0, false, 1, false, 512, true> },
      { 1024, 16, 0, 1024, scatter_kernel<1024, 16, 1, 0, false, 1, false, 1024, true> } },
    { { 1024, 16, 1, 256, scatter_kernel<1024, 16, 1, 1, false, 1, false, 256, true> },
      { 1024, 16, 1, 512, scatter_kernel<1024, 16, 1, 1, false, 1, false, 512, true> },
      { 1024, 16, 1, 1024, scatter_kernel<1024, 16, 1, 1, false, 1, false, 1024, true> } },
};
// pcp: last radix pass at the receiver (packed input, tiles from plan_kernel, TMA output); PERSISTENT grids: the
// number of tiles of a stage is known on the device only, and the grid is sized to the SMs the copy kernel leaves
static const PPCfg kPcpLast[4] = {
    { 256, 16, 1, 256, scatter_kernel<256, 16, 1, 1, false, 4, true> },                      // <= 7 bits (tile 4096, as scatter_cfg2)
    { 512, 16, 1, 256, scatter_kernel<512, 16, 1, 1, false, 2, true> },                      // 8 bits
    { 512, 16, 1, 512, scatter_kernel<512, 16, 1, 1, false, 2, true, 512> },                 // 9 bits
    { 1024, 8, 1, 1024, scatter_kernel<1024, 8, 1, 1, false, 1, true, 1024> },               // 10 bits
};
constexpr int PCP_NS = 3;        // ring slots of 32 KB of the copy kernel (1 load in flight per CTA)
constexpr int PCP_NS_DEEP = 6;   // deep ring (4 loads of 32 KB in flight per CTA): for running the copy on a FEW SMs only, so
                                 // that the radix passes next to it keep their full occupancy (option "pcp_copy_ctas")
static size_t pcp_copy_smem(int ns = PCP_NS) { return (size_t)ns * PCP_PIECE * sizeof(tup_t) + (PCP_MAX_CHUNKS + 4) * sizeof(uint32_t); }
constexpr uint32_t PP_MAX_PASS_BITS = 10;
constexpr uint32_t PP_MAX_BITS = 2 * PP_MAX_PASS_BITS;

typedef void (*join_fn)(JoinArgs);
struct JoinCfg { int threads, cap, u; join_fn agg, mat; size_t smem_agg, smem_mat; };
// <threads, build chunk, probe chunk, R ring slots, S ring slots>
#define GJ_JC(T, C, U, NR, NS, OPT) { T, C, U, join_kernel<T, C, U, NR, NS, false, OPT>, join_kernel<T, C, U, NR, NS, true, OPT>, \
                                 JoinSmem<C, U, NR, NS, false>::total, JoinSmem<C, U, NR, NS, true>::total }
// The winner of round 1's sweep over seven shapes (rings of 3 + 3 slots, atomic build) plus the small-ring shape
// the materialising join needs (its 32 KB of pair staging does not fit next to 3 + 3 rings) and, as a structurally
// different fallback, the optimistic (store-then-verify) build, measured 3 % slower on B200.
static const JoinCfg kJoin[] = {
    GJ_JC(1024, 4096, 4096, 3, 3, false),  // 0 default: two whole steps (build + probe chunk) in flight
    GJ_JC(1024, 4096, 4096, 2, 2, false),  // 1 small rings: the materialising join
    GJ_JC(1024, 4096, 4096, 3, 3, true),   // 2 optimistic (atomic-free when collision-free) build
};
static const int kNumJoin = (int)(sizeof(kJoin) / sizeof(kJoin[0]));
// late-materialisation aggregate (payload = row id, side-table gathers on a match): the default shape
static const join_fn kJoinLate = join_kernel<1024, 4096, 4096, 3, 3, false, false, true>;
struct LateCols {   // column-major side tables: value of column z for row id i = cols[z * stride + i]
    const int32_t* cols[2] = {nullptr, nullptr};   // [0] of the user's R, [1] of the user's S
    uint32_t ncols[2] = {0, 0};
    uint64_t stride[2] = {0, 0};
};

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
struct RelMeta {
    uint32_t* ghist;             // 2^15 (zeroed per call)
    unsigned long long* desc;    // scan descriptors (zeroed per call)
    uint32_t* ticket;            // scan ticket (zeroed per call)
    uint32_t* off;               // 2^15 + 1
    uint32_t* cur1;              // 256
    uint32_t* cur2;              // 2^15
    uint4* tiles;                // pass-2 tile descriptors
    uint32_t* num_tiles;         // (zeroed per call)
};

constexpr uint32_t FINE_MAX = 1u << MAX_RADIX_BITS;
constexpr uint32_t MAX3_RADIX_BITS = 21;   // three passes: 8 + 8 + up to 5 bits
constexpr uint32_t NB3_MAX = 1u << MAX3_RADIX_BITS;
constexpr uint32_t SCAN_TILES_MAX = FINE_MAX / SCAN_TILE;
constexpr int N_EVENTS = 64;

struct Plan { uint32_t B = 0, b1 = 0, b2 = 0, b3 = 0; };   // b3 != 0: three passes, b1 + b2 == 16

struct gj_ctx {
    int device = 0, sm_count = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    uint64_t maxR = 0, maxS = 0;
    tup_t* out[2] = {nullptr, nullptr};   // partitioned tuples of R (slot 0) / S (slot 1)
    tup_t* scratch = nullptr;             // first-pass output, max(maxR, maxS) tuples
    unsigned char* zero_block = nullptr;
    size_t zero_bytes = 0;
    unsigned char* meta_block = nullptr;
    RelMeta meta[2];
    uint4* units = nullptr;
    uint64_t units_cap = 0, tiles_cap = 0;
    uint32_t* unit_base = nullptr;            // 2^15 + 1, unit_base[nb] = number of units
    unsigned long long* unit_desc = nullptr;  // scan descriptors of the unit sequence (zeroed)
    uint32_t* unit_ticket = nullptr;          // scan ticket of the unit sequence (zeroed)
    uint4* tiles_block = nullptr;
    unsigned long long* result = nullptr;
    unsigned long long* h_result = nullptr;   // pinned
    tup_t** d_dst_bases = nullptr;            // 256 pointers (shuffle)
    cudaEvent_t ev[5] = {};
    cudaEvent_t pev[2][3] = {};   // per scatter launch: [role][before p1, after p1, after p2]
    cudaEvent_t sev[2][2] = {};   // peer-scatter kernel of relation 0/1: before, after
    cudaEvent_t stage_ev[4] = {}; // staged pipeline: partition done [side 0/1], scratch free, join done
    // staged (multi-GPU overlap) pipeline state
    struct { Plan pl; int role_of_side[2]; uint64_t n_side[2]; bool active, scratch_used; } stage = {};
    uint32_t* shuf_cur[2] = {nullptr, nullptr};      // device: per-destination cursors of relation 0/1
    tup_t** shuf_bases[2] = {nullptr, nullptr};      // device: per-destination base pointers
    unsigned char* h_shuf = nullptr;                 // pinned staging for the two above
    // third-pass state (allocated on first use: build sides beyond 2^28 tuples)
    struct P3 {
        uint32_t* ghist3[2] = {nullptr, nullptr};       // 2^MAX3 counters per relation (fully written by sub_hist)
        uint32_t* off3[2] = {nullptr, nullptr};         // 2^MAX3 + 1 fine offsets
        uint32_t* cur3[2] = {nullptr, nullptr};
        uint32_t* tile_prefix[2] = {nullptr, nullptr};  // 65536 + 1
        uint4* tiles[2] = {nullptr, nullptr};
        uint32_t* unit_base = nullptr;
        unsigned char* zero = nullptr; size_t zero_bytes = 0;
        unsigned long long* desc[5] = {};               // off3 R, off3 S, units, tiles R, tiles S
        uint32_t* ticket[5] = {};
        unsigned char* block = nullptr;
    } p3;
    // sharded "partition, then push" pipeline (gj_pp_*), allocated on first use
    struct PP {
        bool active = false;
        uint32_t G = 0, rank = 0, g = 0, B = 0, Btot = 0, b1 = 0, b2 = 0, out = 0, big = 0;
        int role_of_side[2] = {0, 1};
        uint64_t n_glob[2] = {0, 0};
        unsigned char* block = nullptr;
        uint32_t* cur_fine[2] = {nullptr, nullptr};      // 2^PP_MAX_BITS write cursors (this source's)
        uint32_t* tile_prefix[2] = {nullptr, nullptr};   // 2^b1 + 1, last = number of pass-2 tiles
        tup_t** bases[2] = {nullptr, nullptr};           // destination partition buffers
        unsigned char* zero[2] = {nullptr, nullptr};     // zeroed per call: tile-scan descriptor, ticket, status
        size_t zero_bytes = 0;
        unsigned long long* tdesc[2] = {nullptr, nullptr};
        uint32_t* tticket[2] = {nullptr, nullptr};
        uint32_t* status[2] = {nullptr, nullptr};        // [0] abort, [1] tuples received
        unsigned char* h_pin = nullptr;                  // pinned: bases staging [2][256 ptrs] + status read-back [2][4]
        cudaEvent_t ev[2][4] = {};                       // per relation: local begin, local end, push begin, push end
        cudaEvent_t jev[2] = {};                         // join begin, end
    } pp;
    // sharded "partition, copy, partition" pipeline (gj_pcp_*), allocated on first use
    struct PCP {
        bool active = false;
        uint32_t G = 0, rank = 0, g = 0, B = 0, bl = 0, b1 = 0, b2 = 0;
        int role_of_side[2] = {0, 1};
        uint64_t n_glob[2] = {0, 0}, n_loc[2] = {0, 0};
        unsigned char* block = nullptr;
        PcpTables tab[2];
        tup_t** bases[2] = {nullptr, nullptr};
        unsigned char** ctrl_ptrs[2] = {nullptr, nullptr};   // device: every GPU's control block (flags + fine_in; local or mapped)
        uint32_t* fine[2] = {nullptr, nullptr};         // this source's fine histograms [2^(g + B)] (zeroed per join)
        int first = 0;                       // the relation that builds (travels first)
        uint32_t epoch = 0;                  // join number: the value the stage flags take
        uint32_t stages[2] = {0, 0};
        bool parted[2] = {false, false}, recv_done[2] = {false, false};
        unsigned char* h_pin = nullptr;      // pinned: bases + flag pointers staging [2][2][256 ptrs] + status read-back [2][4]
        cudaEvent_t ev[2][6] = {};           // per relation: part begin/end, copy begin/end, recv begin/end
        cudaEvent_t jev[2] = {};
    } pcp;
    // non-partitioned baseline: chained table in global memory (allocated on first use)
    struct NP { uint32_t* heads = nullptr; uint32_t* next = nullptr; uint64_t heads_cap = 0, next_cap = 0;
                unsigned long long* slots = nullptr; uint64_t slots_cap = 0; } np;   // slots: perfect array
    unsigned char* zero_role[2] = {nullptr, nullptr};
    unsigned char* zero_common = nullptr;
    size_t zero_role_bytes = 0, zero_common_bytes = 0;
    cudaEvent_t cev[N_EVENTS] = {};
    int32_t* d_in[4] = {nullptr, nullptr, nullptr, nullptr};   // host-entry staging Rk,Rp,Sk,Sp
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;
    uint32_t launches = 0;
    float last_shuffle_ms = 0.f;
    // options
    int64_t opt_radix_bits = 0, opt_pass1_bits = 0, opt_scatter_cfg1 = 255, opt_scatter_cfg2 = 255,
            opt_join_cfg = 0, opt_unit = 0, opt_gpu_bits = 0, opt_part_target = 4096,
            opt_join_grid = 0, opt_h2d_chunk = 8u << 20, opt_shuffle_grid = 0, opt_pp_out = 1, opt_pp_tile16k = 1, opt_nopart_max = 1 << 21, opt_pcp_copy_ctas = 24, opt_pcp_timeout_ms = 5000;
    bool attrs_set = false;
};

struct Rel {   // one input relation as handed to the pipeline
    const int32_t* keys = nullptr;
    const int32_t* pays = nullptr;
    const tup_t* tup = nullptr;     // packed alternative
    uint64_t n = 0;
    int slot = 0;                   // 0 = user's R, 1 = user's S
};

static int set_func_attrs(gj_ctx* ctx) {
    if (ctx->attrs_set) return GJ_OK;
    const int hist_smem = 4 << 15;   // 2^15 u32 counters or 2^16 packed u16 counters
    CK(cudaFuncSetAttribute(hist_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, hist_smem));
    CK(cudaFuncSetAttribute(hist_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, hist_smem));
    CK(cudaFuncSetAttribute(hist_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, hist_smem));
    CK(cudaFuncSetAttribute(hist_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, hist_smem));
    for (int i = 0; i < kNumScatter; ++i) {
        const int bytes = (int)scatter_smem(kScatter[i]);
        CK(cudaFuncSetAttribute(kScatter[i].col, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        CK(cudaFuncSetAttribute(kScatter[i].packed, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        CK(cudaFuncSetAttribute(kScatter[i].col_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    }
    for (const PPCfg& c : kPPFirst) CK(cudaFuncSetAttribute(c.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp_smem(c)));
    for (const auto& row : kPPPush)
        for (const PPCfg& c : row) CK(cudaFuncSetAttribute(c.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp_smem(c)));
    for (const PPCfg& c : kPcpLast) CK(cudaFuncSetAttribute(c.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp_smem(c)));
    CK(cudaFuncSetAttribute(pcp_copy_kernel<PCP_NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pcp_copy_smem()));
    CK(cudaFuncSetAttribute(pcp_copy_kernel<PCP_NS_DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pcp_copy_smem(PCP_NS_DEEP)));
    for (const auto& row : kPPPushBig)
        for (const PPCfg& c : row) CK(cudaFuncSetAttribute(c.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp_smem(c)));
    for (int i = 0; i < kNumJoin; ++i) {
        CK(cudaFuncSetAttribute(kJoin[i].agg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kJoin[i].smem_agg));
        if (i == 0) CK(cudaFuncSetAttribute(kJoinLate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kJoin[0].smem_agg));
        if (kJoin[i].smem_mat <= (size_t)227 * 1024)
            CK(cudaFuncSetAttribute(kJoin[i].mat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kJoin[i].smem_mat));
    }
    ctx->attrs_set = true;
    return GJ_OK;
}

extern "C" const char* gj_last_error(void) { return g_err; }
extern "C" int gj_version(void) { return GJ_VERSION; }
extern "C" uint64_t gj_kernel_launch_count(void) { return g_launches.load(); }

extern "C" void gj_destroy(gj_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->out[0]); cudaFree(ctx->out[1]); cudaFree(ctx->scratch);
    cudaFree(ctx->zero_block); cudaFree(ctx->meta_block); cudaFree(ctx->units);
    cudaFree(ctx->d_dst_bases); cudaFree(ctx->flush_buf); cudaFree(ctx->tiles_block);
    cudaFree(ctx->p3.block); cudaFree(ctx->p3.zero);
    cudaFree(ctx->pp.block);
    cudaFree(ctx->pcp.block);
    cudaFree(ctx->np.heads); cudaFree(ctx->np.next); cudaFree(ctx->np.slots);
    if (ctx->pcp.h_pin) cudaFreeHost(ctx->pcp.h_pin);
    for (auto& r : ctx->pcp.ev) for (auto& e : r) if (e) cudaEventDestroy(e);
    for (auto& e : ctx->pcp.jev) if (e) cudaEventDestroy(e);
    if (ctx->pp.h_pin) cudaFreeHost(ctx->pp.h_pin);
    for (auto& r : ctx->pp.ev) for (auto& e : r) if (e) cudaEventDestroy(e);
    for (auto& e : ctx->pp.jev) if (e) cudaEventDestroy(e);
    for (int i = 0; i < 4; ++i) cudaFree(ctx->d_in[i]);
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    for (auto& r : ctx->pev) for (auto& e : r) if (e) cudaEventDestroy(e);
    for (auto& r : ctx->sev) for (auto& e : r) if (e) cudaEventDestroy(e);
    for (auto& e : ctx->stage_ev) if (e) cudaEventDestroy(e);
    cudaFree(ctx->shuf_cur[0]); cudaFree(ctx->shuf_bases[0]);
    if (ctx->h_shuf) cudaFreeHost(ctx->h_shuf);
    for (auto& e : ctx->cev) if (e) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

static int create_impl(gj_ctx* ctx, int device, uint64_t max_R, uint64_t max_S) {
    const uint64_t LIM = 0xFFFFFFFFull - (1ull << 20);
    if (max_R > LIM || max_S > LIM) return fail(GJ_ERR_ARG, "relations are limited to %llu tuples", (unsigned long long)LIM);
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(GJ_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    CK(cudaSetDevice(device));
    ctx->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(GJ_ERR_CUDA, "device %d is sm_%d%d; libgpujoin is built for sm_100a only", device, prop.major, prop.minor);
    ctx->sm_count = prop.multiProcessorCount;
    ctx->maxR = std::max<uint64_t>(max_R, 1);
    ctx->maxS = std::max<uint64_t>(max_S, 1);
    CK(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    for (auto& e : ctx->ev) CK(cudaEventCreate(&e));
    for (auto& r : ctx->pev) for (auto& e : r) CK(cudaEventCreate(&e));
    for (auto& r : ctx->sev) for (auto& e : r) CK(cudaEventCreate(&e));
    for (auto& e : ctx->stage_ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : ctx->cev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));

    const uint64_t mx = std::max(ctx->maxR, ctx->maxS);
    // +16 tuples of slack: bulk copies and 16-byte loads round ranges out to tuple pairs
    if (cudaMalloc(&ctx->out[0], (ctx->maxR + 16) * sizeof(tup_t)) != cudaSuccess ||
        cudaMalloc(&ctx->out[1], (ctx->maxS + 16) * sizeof(tup_t)) != cudaSuccess ||
        cudaMalloc(&ctx->scratch, (mx + 16) * sizeof(tup_t)) != cudaSuccess) {
        cudaGetLastError();
        return fail(GJ_ERR_NOMEM, "cudaMalloc of %.2f GB partition buffers failed", (ctx->maxR + ctx->maxS + mx) * 8e-9);
    }
    // zeroed-per-call block: [role 0 | role 1 | common]; role = fine histogram, scan descriptors,
    // scan ticket, tile count; common = unit-scan descriptors + ticket, result.  The staged
    // (multi-GPU) pipeline zeroes the three regions independently.
    const size_t r_hist = 0;
    const size_t r_desc = r_hist + FINE_MAX * sizeof(uint32_t);
    const size_t r_cnt = r_desc + SCAN_TILES_MAX * sizeof(unsigned long long);
    const size_t role_bytes = r_cnt + 16;
    const size_t c_desc = 0;
    const size_t c_cnt = c_desc + SCAN_TILES_MAX * sizeof(unsigned long long);
    const size_t c_res = c_cnt + 16;
    const size_t common_bytes = c_res + 4 * sizeof(unsigned long long);
    ctx->zero_role_bytes = role_bytes; ctx->zero_common_bytes = common_bytes;
    ctx->zero_bytes = 2 * role_bytes + common_bytes;
    CK(cudaMalloc(&ctx->zero_block, ctx->zero_bytes));
    ctx->zero_role[0] = ctx->zero_block; ctx->zero_role[1] = ctx->zero_block + role_bytes;
    ctx->zero_common = ctx->zero_block + 2 * role_bytes;
    // persistent metadata block
    size_t mb = 0;
    const size_t o_off = mb;    mb += 2 * (FINE_MAX + 4) * sizeof(uint32_t);
    const size_t o_cur1 = mb;   mb += 2 * NB_MAX * CUR1_STRIDE * sizeof(uint32_t);
    const size_t o_cur2 = mb;   mb += 2 * FINE_MAX * sizeof(uint32_t);
    const size_t o_ub = mb;     mb += (FINE_MAX + 4) * sizeof(uint32_t);
    CK(cudaMalloc(&ctx->meta_block, mb));
    // pass-2 tile descriptors: smallest tile is 2048 tuples, one extra tile per first-pass partition
    ctx->tiles_cap = mx / 2048 + (1u << PP_MAX_PASS_BITS) + 16;
    CK(cudaMalloc(&ctx->tiles_block, 2 * ctx->tiles_cap * sizeof(uint4)));
    for (int r = 0; r < 2; ++r) {
        RelMeta& m = ctx->meta[r];
        m.ghist = reinterpret_cast<uint32_t*>(ctx->zero_role[r] + r_hist);
        m.desc = reinterpret_cast<unsigned long long*>(ctx->zero_role[r] + r_desc);
        m.ticket = reinterpret_cast<uint32_t*>(ctx->zero_role[r] + r_cnt);
        m.num_tiles = reinterpret_cast<uint32_t*>(ctx->zero_role[r] + r_cnt) + 1;
        m.off = reinterpret_cast<uint32_t*>(ctx->meta_block + o_off) + (size_t)r * (FINE_MAX + 4);
        m.cur1 = reinterpret_cast<uint32_t*>(ctx->meta_block + o_cur1) + (size_t)r * NB_MAX * CUR1_STRIDE;
        m.cur2 = reinterpret_cast<uint32_t*>(ctx->meta_block + o_cur2) + (size_t)r * FINE_MAX;
        m.tiles = ctx->tiles_block + (size_t)r * ctx->tiles_cap;
    }
    ctx->unit_desc = reinterpret_cast<unsigned long long*>(ctx->zero_common + c_desc);
    ctx->unit_ticket = reinterpret_cast<uint32_t*>(ctx->zero_common + c_cnt);
    ctx->result = reinterpret_cast<unsigned long long*>(ctx->zero_common + c_res);
    ctx->unit_base = reinterpret_cast<uint32_t*>(ctx->meta_block + o_ub);
    // multi-GPU shuffle: per-relation cursors and destination bases (+ pinned staging)
    CK(cudaMalloc(&ctx->shuf_cur[0], 2 * NB_MAX * sizeof(uint32_t)));
    ctx->shuf_cur[1] = ctx->shuf_cur[0] + NB_MAX;
    CK(cudaMalloc(&ctx->shuf_bases[0], 2 * NB_MAX * sizeof(tup_t*)));
    ctx->shuf_bases[1] = ctx->shuf_bases[0] + NB_MAX;
    CK(cudaHostAlloc(&ctx->h_shuf, 2 * NB_MAX * (sizeof(uint32_t) + sizeof(void*)), cudaHostAllocDefault));
    // unit list: probe side cut every >= 1024 tuples, plus one per partition
    ctx->units_cap = mx / 1024 + FINE_MAX + 16;
    CK(cudaMalloc(&ctx->units, ctx->units_cap * sizeof(uint4)));
    CK(cudaMalloc(&ctx->d_dst_bases, NB_MAX * sizeof(tup_t*)));
    CK(cudaHostAlloc(&ctx->h_result, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
    return set_func_attrs(ctx);
}

extern "C" int gj_create(gj_ctx** out, int device, uint64_t max_R, uint64_t max_S) {
    if (!out) return fail(GJ_ERR_ARG, "gj_create: out is NULL");
    *out = nullptr;
    gj_ctx* ctx = new (std::nothrow) gj_ctx();
    if (!ctx) return fail(GJ_ERR_NOMEM, "host allocation failed");
    const int rc = create_impl(ctx, device, max_R, max_S);
    if (rc != GJ_OK) {
        char keep[sizeof(g_err)];
        memcpy(keep, g_err, sizeof(keep));
        gj_destroy(ctx);
        memcpy(g_err, keep, sizeof(keep));
        return rc;
    }
    *out = ctx;
    return GJ_OK;
}

extern "C" int gj_set_stream(gj_ctx* ctx, void* s) {
    if (!ctx) return fail(GJ_ERR_ARG, "ctx is NULL");
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return GJ_OK;
}

static int64_t* option_slot(gj_ctx* ctx, const char* name) {
    struct { const char* n; int64_t* p; } tab[] = {
        {"radix_bits", &ctx->opt_radix_bits}, {"pass1_bits", &ctx->opt_pass1_bits},
        {"scatter_cfg1", &ctx->opt_scatter_cfg1}, {"scatter_cfg2", &ctx->opt_scatter_cfg2},
        {"join_cfg", &ctx->opt_join_cfg}, {"unit_tuples", &ctx->opt_unit},
        {"gpu_bits", &ctx->opt_gpu_bits}, {"part_target", &ctx->opt_part_target},
        {"join_grid", &ctx->opt_join_grid}, {"h2d_chunk", &ctx->opt_h2d_chunk},
        {"shuffle_grid", &ctx->opt_shuffle_grid}, {"pp_out", &ctx->opt_pp_out}, {"pp_tile16k", &ctx->opt_pp_tile16k}, {"nopart_max", &ctx->opt_nopart_max}, {"pcp_copy_ctas", &ctx->opt_pcp_copy_ctas}, {"pcp_timeout_ms", &ctx->opt_pcp_timeout_ms},
    };
    for (auto& t : tab) if (!strcmp(t.n, name)) return t.p;
    return nullptr;
}

extern "C" int gj_set_option(gj_ctx* ctx, const char* name, int64_t v) {
    if (!ctx || !name) return fail(GJ_ERR_ARG, "gj_set_option: NULL argument");
    if (!strcmp(name, "scatter_cfg")) {
        if (v < 0 || (v >= kNumScatter && v != 255)) return fail(GJ_ERR_ARG, "scatter_cfg %lld out of range [0,%d) (255 = auto)", (long long)v, kNumScatter);
        ctx->opt_scatter_cfg1 = ctx->opt_scatter_cfg2 = v;
        return GJ_OK;
    }
    int64_t* p = option_slot(ctx, name);
    if (!p) return fail(GJ_ERR_ARG, "unknown option '%s'", name);
    if (v < 0) return fail(GJ_ERR_ARG, "option '%s' must be >= 0", name);
    if ((p == &ctx->opt_scatter_cfg1 || p == &ctx->opt_scatter_cfg2) && v >= kNumScatter && v != 255)
        return fail(GJ_ERR_ARG, "%s %lld out of range [0,%d)", name, (long long)v, kNumScatter);
    if (p == &ctx->opt_join_cfg && v >= kNumJoin) return fail(GJ_ERR_ARG, "join_cfg out of range [0,%d)", kNumJoin);
    if (p == &ctx->opt_radix_bits && v > MAX3_RADIX_BITS) return fail(GJ_ERR_ARG, "radix_bits <= %d", (int)MAX3_RADIX_BITS);
    if (p == &ctx->opt_pass1_bits && v > PP_MAX_PASS_BITS) return fail(GJ_ERR_ARG, "pass1_bits <= %d", (int)PP_MAX_PASS_BITS);
    if (p == &ctx->opt_pp_out && v > 1) return fail(GJ_ERR_ARG, "pp_out is 0 (8-byte stores) or 1 (TMA bulk stores)");
    if (p == &ctx->opt_pp_tile16k && v > 1) return fail(GJ_ERR_ARG, "pp_tile16k is 0 or 1");
    if (p == &ctx->opt_unit && v && (v < 1024 || v > (1 << 20))) return fail(GJ_ERR_ARG, "unit_tuples must be 0 (default) or in [1024, 2^20]");
    if ((p == &ctx->opt_join_grid || p == &ctx->opt_shuffle_grid || p == &ctx->opt_pcp_copy_ctas) && v > 65535) return fail(GJ_ERR_ARG, "%s <= 65535 CTAs", name);
    if (p == &ctx->opt_nopart_max && v > (1ll << 31)) return fail(GJ_ERR_ARG, "nopart_max <= 2^31");
    if (p == &ctx->opt_gpu_bits && v > 8) return fail(GJ_ERR_ARG, "gpu_bits <= 8");
    if (p == &ctx->opt_part_target && v < 32) return fail(GJ_ERR_ARG, "part_target >= 32");
    if (p == &ctx->opt_h2d_chunk && v < 4096) return fail(GJ_ERR_ARG, "h2d_chunk >= 4096");
    *p = v;
    return GJ_OK;
}

extern "C" int gj_get_option(gj_ctx* ctx, const char* name, int64_t* v) {
    if (!ctx || !name || !v) return fail(GJ_ERR_ARG, "gj_get_option: NULL argument");
    if (!strcmp(name, "scatter_cfg")) { *v = ctx->opt_scatter_cfg1; return GJ_OK; }
    if (!strcmp(name, "num_scatter_cfgs")) { *v = kNumScatter; return GJ_OK; }
    if (!strcmp(name, "num_join_cfgs")) { *v = kNumJoin; return GJ_OK; }
    if (!strcmp(name, "sm_count")) { *v = ctx->sm_count; return GJ_OK; }
    if (!strcmp(name, "last_shuffle_us")) { *v = (int64_t)(ctx->last_shuffle_ms * 1000.f); return GJ_OK; }
    int64_t* p = option_slot(ctx, name);
    if (!p) return fail(GJ_ERR_ARG, "unknown option '%s'", name);
    *v = *p;
    return GJ_OK;
}

// ------------------------------------------------------------------------------------------
// planning: how many radix bits, how they split over passes
// (the reference freezes log_parts1 = 8, log_parts2 = 5 at compile time, common.h:51-52)
// ------------------------------------------------------------------------------------------

static Plan choose_plan(const gj_ctx* ctx, uint64_t n_build, uint32_t forced_bits, bool allow3 = false) {
    Plan p;
    uint32_t B = forced_bits ? forced_bits : (uint32_t)ctx->opt_radix_bits;
    const uint32_t cap = allow3 ? MAX3_RADIX_BITS : (uint32_t)MAX_RADIX_BITS;
    if (!B) {
        const uint64_t target = (uint64_t)ctx->opt_part_target;
        while (B < cap && (n_build >> B) > target) ++B;
    }
    B = std::min<uint32_t>(B, cap);
    if (B > (uint32_t)MAX_RADIX_BITS) {   // third pass on the low bits of every second-level partition
        p.B = B; p.b1 = 8; p.b2 = 8; p.b3 = B - 16;
        return p;
    }
    if (B <= (uint32_t)MAX_PASS_BITS) { p.b1 = B; p.b2 = 0; }
    else {
        p.b1 = ctx->opt_pass1_bits ? (uint32_t)ctx->opt_pass1_bits : B / 2;   // measured: the smaller fan-out first
        p.b1 = std::min<uint32_t>(std::max<uint32_t>(p.b1, B - MAX_PASS_BITS), MAX_PASS_BITS);
        p.b2 = B - p.b1;
    }
    p.B = B;
    return p;
}

// ------------------------------------------------------------------------------------------
// enqueue helpers (no host synchronisation inside)
// ------------------------------------------------------------------------------------------
static int enqueue_hist(gj_ctx* ctx, cudaStream_t s, const void* in, bool packed, uint64_t n,
                        uint32_t shift, uint32_t bits, uint32_t* ghist, int threads = 1024,
                        const uint32_t* lo_dev = nullptr, const uint32_t* hi_dev = nullptr,   // the slot range lives on the device, n is its upper bound
                        int max_ctas = 0) {
    if (!n) return GJ_OK;
    const uint64_t per_cta = (uint64_t)threads * 16;
    const uint64_t sms = max_ctas > 0 ? (uint64_t)max_ctas : (uint64_t)ctx->sm_count;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(sms, (n + per_cta - 1) / per_cta));
    const bool p16 = bits > 15;   // two 16-bit counters per word
    const size_t smem = p16 ? (size_t)2 << bits : (size_t)4 << bits;
    if (packed) {
        if (p16) hist_kernel<true, true><<<grid, threads, smem, s>>>(in, (uint32_t)n, shift, bits, ghist, lo_dev, hi_dev);
        else hist_kernel<true, false><<<grid, threads, smem, s>>>(in, (uint32_t)n, shift, bits, ghist, lo_dev, hi_dev);
    } else {
        if (p16) hist_kernel<false, true><<<grid, threads, smem, s>>>(in, (uint32_t)n, shift, bits, ghist, lo_dev, hi_dev);
        else hist_kernel<false, false><<<grid, threads, smem, s>>>(in, (uint32_t)n, shift, bits, ghist, lo_dev, hi_dev);
    }
    LAUNCHED();
    return GJ_OK;
}

// scatter kernel variants: explicit option or, at 255, the measured best per fan-out
// (profiles/: 8-byte stores win in the first pass, TMA bulk-copy runs win in the second)
static const ScatterCfg& scatter_cfg1(const gj_ctx* ctx) {
    return kScatter[ctx->opt_scatter_cfg1 == 255 ? 0 : ctx->opt_scatter_cfg1];
}
static const ScatterCfg& scatter_cfg2(const gj_ctx* ctx, uint32_t b2) {
    return kScatter[ctx->opt_scatter_cfg2 == 255 ? (b2 <= 7 ? 2 : 3) : ctx->opt_scatter_cfg2];
}

// peer-store shuffle: TMA bulk stores of the (long) per-destination runs -- measured over NVLink on
// 2 GPUs: 689 GB/s against 581 GB/s with 8-byte stores
static const ScatterCfg& scatter_cfg_shuffle(const gj_ctx* ctx) {
    return kScatter[ctx->opt_scatter_cfg1 == 255 ? 2 : ctx->opt_scatter_cfg1];
}

static uint32_t unit_tuples(const gj_ctx* ctx) { return ctx->opt_unit ? (uint32_t)ctx->opt_unit : 8192u; }

// scan of the fine histogram(s) of roles [first, first+nrel) and, for a join, of the unit counts
// nrel == 0 with with_units: only the unit sequence (staged pipeline: both histograms are final)
static int enqueue_scan(gj_ctx* ctx, cudaStream_t s, int first, uint32_t nrel, uint32_t nb, bool with_units) {
    ScanArgs a;
    for (uint32_t r = 0; r < 2; ++r) {
        const RelMeta& m = ctx->meta[r < nrel ? first + (int)r : first];
        a.seq[r].in = m.ghist; a.seq[r].in2 = nullptr; a.seq[r].out = m.off; a.seq[r].desc = m.desc; a.seq[r].ticket = m.ticket;
        a.seq[r].mode = SCAN_PLAIN; a.seq[r].param = 0; a.seq[r].param2 = 0;
    }
    a.seq[2].in = ctx->meta[0].ghist; a.seq[2].in2 = ctx->meta[1].ghist; a.seq[2].out = ctx->unit_base;
    a.seq[2].desc = ctx->unit_desc; a.seq[2].ticket = ctx->unit_ticket;
    a.seq[2].mode = SCAN_UNITS; a.seq[2].param = unit_tuples(ctx); a.seq[2].param2 = 0;
    a.nb = nb;
    a.seq_base = nrel == 0 ? 2 : 0;   // nrel == 0: unit sequence only
    dim3 grid((nb + SCAN_TILE - 1) / SCAN_TILE, nrel == 0 ? 1 : (with_units ? 3 : nrel));
    scan_lookback_kernel<<<grid, SCAN_THREADS, 0, s>>>(a);
    LAUNCHED();
    return GJ_OK;
}

static int enqueue_plan(gj_ctx* ctx, cudaStream_t s, int first, uint32_t nrel, const Plan& pl, bool with_units) {
    PlanArgs a;
    for (uint32_t r = 0; r < 2; ++r) {
        const RelMeta& m = ctx->meta[r < nrel ? first + (int)r : first];
        a.rel[r].off = m.off; a.rel[r].cur1 = m.cur1; a.rel[r].cur2 = m.cur2; a.rel[r].tiles = m.tiles; a.rel[r].num_tiles = m.num_tiles;
    }
    a.nrel = nrel; a.with_units = with_units ? 1u : 0u; a.b1 = pl.b1; a.b2 = pl.b2;
    a.j_lo = a.j_hi = 0;
    if (nrel == 0) for (uint32_t r = 0; r < 2; ++r) a.rel[r].off = ctx->meta[r].off;   // units only
    const ScatterCfg& c2 = scatter_cfg2(ctx, pl.b2);
    a.tile = (uint32_t)(c2.threads * c2.ipt);
    a.unit = unit_tuples(ctx);
    a.unit_base = ctx->unit_base; a.units = ctx->units;
    const uint32_t nb = 1u << pl.B;
    const int grid = (int)std::max<uint32_t>(1, std::min<uint32_t>(128, (nb + PLAN_THREADS - 1) / PLAN_THREADS + (pl.b2 ? 32 : 0)));
    plan_kernel<<<grid, PLAN_THREADS, 0, s>>>(a);
    LAUNCHED();
    return GJ_OK;
}

// all scatter passes of one relation; role indexes ctx->meta, dst is the final buffer
static int enqueue_scatter(gj_ctx* ctx, cudaStream_t s, const Rel& rel, int role, const Plan& pl, tup_t* dst) {
    if (!rel.n) return GJ_OK;
    const RelMeta& m = ctx->meta[role];
    const ScatterCfg& c1 = scatter_cfg1(ctx);
    const uint32_t T1 = (uint32_t)(c1.threads * c1.ipt);
    ScatterArgs a;
    memset(&a, 0, sizeof(a));
    a.in_keys = rel.keys; a.in_pays = rel.pays; a.in_tup = rel.tup;
    a.n = (uint32_t)rel.n;
    a.out = pl.b2 ? ctx->scratch : dst;
    a.shift = pl.b2; a.bits = pl.b1;
    a.cursors = pl.b2 ? m.cur1 : m.cur2;
    a.cursor_stride = pl.b2 ? CUR1_STRIDE : 1;
    const uint32_t grid1 = (uint32_t)((rel.n + (rel.tup ? 1 : 0) + T1 - 1) / T1);   // +1: alignment shift of packed input
    a.ntiles = grid1;
    CK(cudaEventRecord(ctx->pev[role][0], s));
    (rel.tup ? c1.packed : c1.col)<<<grid1, c1.threads, scatter_smem(c1), s>>>(a);
    LAUNCHED();
    CK(cudaEventRecord(ctx->pev[role][1], s));
    if (pl.b2) {
        const ScatterCfg& c2 = scatter_cfg2(ctx, pl.b2);
        const uint32_t T2 = (uint32_t)(c2.threads * c2.ipt);
        ScatterArgs b;
        memset(&b, 0, sizeof(b));
        b.in_tup = ctx->scratch; b.out = dst; b.n = (uint32_t)rel.n;
        b.shift = 0; b.bits = pl.b2;
        b.cursors = m.cur2; b.cursor_stride = 1; b.tiles = m.tiles; b.num_tiles = m.num_tiles;
        const uint32_t grid2 = (uint32_t)(rel.n / T2) + (1u << pl.b1) + 2;   // upper bound on tiles
        c2.packed<<<grid2, c2.threads, scatter_smem(c2), s>>>(b);
        LAUNCHED();
        CK(cudaEventRecord(ctx->pev[role][2], s));
    }
    return GJ_OK;
}

static int fill_pass_times(gj_ctx* ctx, gj_timings* t, const Plan& pl, int nroles, int first_role) {
    if (!t) return GJ_OK;
    for (int k = 0; k < nroles; ++k) {
        const int role = first_role + k;
        CK(cudaEventElapsedTime(&t->pass_ms[2 * k], ctx->pev[role][0], ctx->pev[role][1]));
        if (pl.b2) CK(cudaEventElapsedTime(&t->pass_ms[2 * k + 1], ctx->pev[role][1], ctx->pev[role][2]));
    }
    return GJ_OK;
}

static int enqueue_join(gj_ctx* ctx, cudaStream_t s, const tup_t* bld, const tup_t* prb, const Plan& pl,
                        uint64_t n_bld, uint64_t n_prb, bool mat, int32_t* out_b, int32_t* out_p, uint64_t cap,
                        const uint32_t* num_units = nullptr, int gpu_bits = -1, const LateCols* late = nullptr,
                        bool bld_is_S = false) {
    (void)n_bld;
    int cfg = late ? 0 : (int)ctx->opt_join_cfg;
    if (mat && kJoin[cfg].smem_mat > (size_t)227 * 1024) cfg = 1;   // pair staging needs 32 KB: smaller rings
    const JoinCfg& jc = kJoin[cfg];
    JoinArgs a;
    a.bld = bld; a.prb = prb;
    a.units = ctx->units; a.num_units = num_units ? num_units : ctx->unit_base + (1u << pl.B);
    a.hash_shift = pl.B + (gpu_bits >= 0 ? (uint32_t)gpu_bits : (uint32_t)ctx->opt_gpu_bits);
    a.result = ctx->result;
    a.out_bld_pay = out_b; a.out_prb_pay = out_p; a.cap = cap;
    a.bld_cols = a.prb_cols = nullptr; a.ncols_bld = a.ncols_prb = 0; a.stride_bld = a.stride_prb = 0;
    if (late) {
        const int b = bld_is_S ? 1 : 0, p = 1 - b;
        a.bld_cols = late->cols[b]; a.ncols_bld = late->ncols[b]; a.stride_bld = late->stride[b];
        a.prb_cols = late->cols[p]; a.ncols_prb = late->ncols[p]; a.stride_prb = late->stride[p];
    }
    const size_t smem = mat ? jc.smem_mat : jc.smem_agg;
    uint64_t grid = (uint64_t)ctx->sm_count;   // persistent: one CTA per SM owns the whole shared memory
    if (ctx->opt_join_grid) grid = (uint64_t)ctx->opt_join_grid;
    const uint64_t max_units = n_prb / unit_tuples(ctx) + (1ull << pl.B);
    // probe partitions of several units (measured on workload A: 8 units each): blocks of consecutive units per CTA
    a.unit_block = (n_prb / unit_tuples(ctx)) > (3ull << pl.B) / 2 ? JOIN_UNIT_BLOCK : 1u;
    grid = std::max<uint64_t>(1, std::min(grid, (max_units + a.unit_block - 1) / a.unit_block));
    (late ? kJoinLate : (mat ? jc.mat : jc.agg))<<<(uint32_t)grid, jc.threads, smem, s>>>(a);
    LAUNCHED();
    return GJ_OK;
}

static void fill_plan(gj_timings* t, const Plan& pl) {
    if (!t) return;
    t->radix_bits = pl.B; t->pass1_bits = pl.b1; t->pass2_bits = pl.b2; t->pass3_bits = pl.b3;
}

// ------------------------------------------------------------------------------------------
// three-pass partitioning (more than 16 radix bits)
// ------------------------------------------------------------------------------------------
static int ensure_pass3(gj_ctx* ctx) {
    gj_ctx::P3& q = ctx->p3;
    if (q.block) return GJ_OK;
    const uint64_t mx = std::max(ctx->maxR, ctx->maxS);
    const size_t tiles_cap = mx / 2048 + FINE_MAX + 16;
    size_t b = 0;
    size_t o_h[2], o_o[2], o_c[2], o_tp[2], o_t[2];
    for (int r = 0; r < 2; ++r) {
        o_h[r] = b; b += (size_t)NB3_MAX * 4;
        o_o[r] = b; b += ((size_t)NB3_MAX + 4) * 4;
        o_c[r] = b; b += (size_t)NB3_MAX * 4;
        o_tp[r] = b; b += ((size_t)FINE_MAX + 4) * 4;
        o_t[r] = b; b += tiles_cap * sizeof(uint4);
    }
    const size_t o_ub = b; b += ((size_t)NB3_MAX + 4) * 4;
    if (cudaMalloc(&q.block, b) != cudaSuccess) { cudaGetLastError(); return fail(GJ_ERR_NOMEM, "third-pass metadata (%.1f MB)", b * 1e-6); }
    for (int r = 0; r < 2; ++r) {
        q.ghist3[r] = reinterpret_cast<uint32_t*>(q.block + o_h[r]);
        q.off3[r] = reinterpret_cast<uint32_t*>(q.block + o_o[r]);
        q.cur3[r] = reinterpret_cast<uint32_t*>(q.block + o_c[r]);
        q.tile_prefix[r] = reinterpret_cast<uint32_t*>(q.block + o_tp[r]);
        q.tiles[r] = reinterpret_cast<uint4*>(q.block + o_t[r]);
    }
    q.unit_base = reinterpret_cast<uint32_t*>(q.block + o_ub);
    // zeroed per call: scan descriptors + tickets of the five third-level scans
    const size_t per = (size_t)(NB3_MAX / SCAN_TILE) * sizeof(unsigned long long) + 16;
    q.zero_bytes = 5 * per;
    CK(cudaMalloc(&q.zero, q.zero_bytes));
    for (int i = 0; i < 5; ++i) {
        q.desc[i] = reinterpret_cast<unsigned long long*>(q.zero + i * per);
        q.ticket[i] = reinterpret_cast<uint32_t*>(q.zero + i * per + (size_t)(NB3_MAX / SCAN_TILE) * sizeof(unsigned long long));
    }
    // the unit list must cover 2^21 partitions
    const uint64_t need = mx / 1024 + NB3_MAX + 16;
    if (ctx->units_cap < need) {
        CK(cudaFree(ctx->units));
        ctx->units = nullptr;
        CK(cudaMalloc(&ctx->units, need * sizeof(uint4)));
        ctx->units_cap = need;
    }
    return GJ_OK;
}

static int enqueue_scan_one(gj_ctx* ctx, cudaStream_t s, const ScanSeq& seq, uint32_t nb) {
    ScanArgs a;
    a.seq[0] = a.seq[1] = a.seq[2] = seq;
    a.nb = nb; a.seq_base = 0;
    dim3 grid((nb + SCAN_TILE - 1) / SCAN_TILE, 1);
    scan_lookback_kernel<<<grid, SCAN_THREADS, 0, s>>>(a);
    LAUNCHED();
    return GJ_OK;
}

// passes 1-3 of one relation; on return its tuples sit in `dst` grouped into 2^B partitions and
// p3.off3[role] holds the offsets.  Level-2 (16-bit) histogram/offsets/cursors/tiles of both
// relations were prepared by the caller with the two-pass machinery.
static int enqueue_partition3(gj_ctx* ctx, cudaStream_t s, const Rel& rel, int role, const Plan& pl, tup_t* dst) {
    gj_ctx::P3& q = ctx->p3;
    const RelMeta& m = ctx->meta[role];
    const uint32_t nb3 = 1u << pl.B;
    // pass 1: input -> dst on the top 8 bits; pass 2: dst -> scratch on the next 8
    {
        const ScatterCfg& c1 = scatter_cfg1(ctx);
        const uint32_t T1 = (uint32_t)(c1.threads * c1.ipt);
        ScatterArgs a;
        memset(&a, 0, sizeof(a));
        a.in_keys = rel.keys; a.in_pays = rel.pays; a.in_tup = rel.tup; a.n = (uint32_t)rel.n;
        a.out = dst; a.shift = pl.b3 + pl.b2; a.bits = pl.b1;
        a.cursors = m.cur1; a.cursor_stride = CUR1_STRIDE;
        a.ntiles = (uint32_t)((rel.n + (rel.tup ? 1 : 0) + T1 - 1) / T1);
        (rel.tup ? c1.packed : c1.col)<<<a.ntiles, c1.threads, scatter_smem(c1), s>>>(a);
        LAUNCHED();
    }
    const ScatterCfg& c2 = scatter_cfg2(ctx, pl.b2);
    const uint32_t T2 = (uint32_t)(c2.threads * c2.ipt);
    {
        ScatterArgs b;
        memset(&b, 0, sizeof(b));
        b.in_tup = dst; b.out = ctx->scratch; b.n = (uint32_t)rel.n;
        b.shift = pl.b3; b.bits = pl.b2;
        b.cursors = m.cur2; b.cursor_stride = 1; b.tiles = m.tiles; b.num_tiles = m.num_tiles;
        c2.packed<<<(uint32_t)(rel.n / T2) + (1u << pl.b1) + 2, c2.threads, scatter_smem(c2), s>>>(b);
        LAUNCHED();
    }
    // third level: count the low bits per second-level partition, scan, tile descriptors, scatter
    sub_hist_kernel<<<ctx->sm_count * 8, 256, 0, s>>>(ctx->scratch, m.off, FINE_MAX, pl.b3, q.ghist3[role]);
    LAUNCHED();
    int rc;
    ScanSeq so;
    so.in = q.ghist3[role]; so.in2 = nullptr; so.out = q.off3[role]; so.desc = q.desc[role]; so.ticket = q.ticket[role];
    so.mode = SCAN_PLAIN; so.param = 0; so.param2 = 0;
    if ((rc = enqueue_scan_one(ctx, s, so, nb3))) return rc;
    ScanSeq st;
    st.in = m.off; st.in2 = nullptr; st.out = q.tile_prefix[role]; st.desc = q.desc[3 + role]; st.ticket = q.ticket[3 + role];
    st.mode = SCAN_TILES; st.param = T2; st.param2 = 0;
    if ((rc = enqueue_scan_one(ctx, s, st, FINE_MAX))) return rc;
    tiles3_kernel<<<FINE_MAX / 256, 256, 0, s>>>(m.off, q.tile_prefix[role], FINE_MAX, T2, pl.b3, q.tiles[role]);
    LAUNCHED();
    CK(cudaMemcpyAsync(q.cur3[role], q.off3[role], (size_t)nb3 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    {
        ScatterArgs c;
        memset(&c, 0, sizeof(c));
        c.in_tup = ctx->scratch; c.out = dst; c.n = (uint32_t)rel.n;
        c.shift = 0; c.bits = pl.b3;
        c.cursors = q.cur3[role]; c.cursor_stride = 1; c.tiles = q.tiles[role]; c.num_tiles = q.tile_prefix[role] + FINE_MAX;
        c2.packed<<<(uint32_t)(rel.n / T2) + FINE_MAX + 2, c2.threads, scatter_smem(c2), s>>>(c);
        LAUNCHED();
    }
    return GJ_OK;
}

// unit list over the 2^B third-level partitions (both relations partitioned)
static int enqueue_units3(gj_ctx* ctx, cudaStream_t s, const Plan& pl) {
    gj_ctx::P3& q = ctx->p3;
    const uint32_t nb3 = 1u << pl.B;
    ScanSeq su;
    su.in = q.ghist3[0]; su.in2 = q.ghist3[1]; su.out = q.unit_base; su.desc = q.desc[2]; su.ticket = q.ticket[2];
    su.mode = SCAN_UNITS; su.param = unit_tuples(ctx); su.param2 = 0;
    int rc;
    if ((rc = enqueue_scan_one(ctx, s, su, nb3))) return rc;
    PlanArgs a;
    memset(&a, 0, sizeof(a));
    for (int r = 0; r < 2; ++r) { a.rel[r].off = q.off3[r]; a.rel[r].cur2 = q.cur3[r]; }
    a.nrel = 0; a.with_units = 1; a.b1 = pl.B; a.b2 = 0;
    a.tile = 4096; a.unit = unit_tuples(ctx);
    a.unit_base = q.unit_base; a.units = ctx->units;
    plan_kernel<<<128, PLAN_THREADS, 0, s>>>(a);
    LAUNCHED();
    return GJ_OK;
}

// ------------------------------------------------------------------------------------------
// the join pipeline
// ------------------------------------------------------------------------------------------
static int check_caps(gj_ctx* ctx, uint64_t nR, uint64_t nS) {
    if (!ctx) return fail(GJ_ERR_ARG, "ctx is NULL");
    if (nR > ctx->maxR || nS > ctx->maxS)
        return fail(GJ_ERR_ARG, "relation sizes (%llu, %llu) exceed the context capacity (%llu, %llu)",
                    (unsigned long long)nR, (unsigned long long)nS, (unsigned long long)ctx->maxR, (unsigned long long)ctx->maxS);
    return GJ_OK;
}

static int run_join(gj_ctx* ctx, Rel R, Rel S, bool mat, int32_t* out_Rp, int32_t* out_Sp, uint64_t cap,
                    uint64_t* matches, uint64_t* checksum, uint64_t* n_pairs, gj_timings* t,
                    const LateCols* late = nullptr) {
    const auto w0 = std::chrono::steady_clock::now();
    if (t) memset(t, 0, sizeof(*t));
    if (matches) *matches = 0;
    if (checksum) *checksum = 0;
    if (n_pairs) *n_pairs = 0;
    CK(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    if (R.n == 0 || S.n == 0) return GJ_OK;
    cudaStream_t s = ctx->stream;
    const bool swap = R.n > S.n;   // build on the smaller relation
    const Rel& bld = swap ? S : R;
    const Rel& prb = swap ? R : S;
    const Plan pl = choose_plan(ctx, bld.n, 0, true);
    fill_plan(t, pl);
    int rc;
    if (pl.b3 && (rc = ensure_pass3(ctx))) return rc;

    CK(cudaMemsetAsync(ctx->zero_block, 0, ctx->zero_bytes, s));
    if (pl.b3) CK(cudaMemsetAsync(ctx->p3.zero, 0, ctx->p3.zero_bytes, s));
    CK(cudaEventRecord(ctx->ev[0], s));
    const uint32_t* num_units = nullptr;
    if (!pl.b3) {
        if ((rc = enqueue_hist(ctx, s, bld.tup ? (const void*)bld.tup : (const void*)bld.keys, bld.tup != nullptr, bld.n, 0, pl.B, ctx->meta[0].ghist))) return rc;
        if ((rc = enqueue_hist(ctx, s, prb.tup ? (const void*)prb.tup : (const void*)prb.keys, prb.tup != nullptr, prb.n, 0, pl.B, ctx->meta[1].ghist))) return rc;
        if ((rc = enqueue_scan(ctx, s, 0, 2, 1u << pl.B, true))) return rc;
        if ((rc = enqueue_plan(ctx, s, 0, 2, pl, true))) return rc;
        CK(cudaEventRecord(ctx->ev[1], s));
        if ((rc = enqueue_scatter(ctx, s, bld, 0, pl, ctx->out[bld.slot]))) return rc;
        if ((rc = enqueue_scatter(ctx, s, prb, 1, pl, ctx->out[prb.slot]))) return rc;
    } else {
        // second-level (top 16 bits) histogram, offsets, cursors and pass-2 tiles with the two-pass machinery
        Plan p2; p2.B = 16; p2.b1 = pl.b1; p2.b2 = pl.b2;
        if ((rc = enqueue_hist(ctx, s, bld.tup ? (const void*)bld.tup : (const void*)bld.keys, bld.tup != nullptr, bld.n, pl.b3, 16, ctx->meta[0].ghist))) return rc;
        if ((rc = enqueue_hist(ctx, s, prb.tup ? (const void*)prb.tup : (const void*)prb.keys, prb.tup != nullptr, prb.n, pl.b3, 16, ctx->meta[1].ghist))) return rc;
        if ((rc = enqueue_scan(ctx, s, 0, 2, FINE_MAX, false))) return rc;
        if ((rc = enqueue_plan(ctx, s, 0, 2, p2, false))) return rc;
        CK(cudaEventRecord(ctx->ev[1], s));
        if ((rc = enqueue_partition3(ctx, s, bld, 0, pl, ctx->out[bld.slot]))) return rc;
        if ((rc = enqueue_partition3(ctx, s, prb, 1, pl, ctx->out[prb.slot]))) return rc;
        if ((rc = enqueue_units3(ctx, s, pl))) return rc;
        num_units = ctx->p3.unit_base + (1u << pl.B);
    }
    CK(cudaEventRecord(ctx->ev[2], s));
    if ((rc = enqueue_join(ctx, s, ctx->out[bld.slot], ctx->out[prb.slot], pl, bld.n, prb.n, mat,
                           swap ? out_Sp : out_Rp, swap ? out_Rp : out_Sp, cap, num_units, -1, late, swap))) return rc;
    CK(cudaEventRecord(ctx->ev[3], s));
    CK(cudaMemcpyAsync(ctx->h_result, ctx->result, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (matches) *matches = ctx->h_result[0];
    if (checksum) *checksum = ctx->h_result[1];
    if (n_pairs) *n_pairs = ctx->h_result[2];
    if (t) {
        CK(cudaEventElapsedTime(&t->hist_ms, ctx->ev[0], ctx->ev[1]));
        CK(cudaEventElapsedTime(&t->part_ms, ctx->ev[1], ctx->ev[2]));
        CK(cudaEventElapsedTime(&t->join_ms, ctx->ev[2], ctx->ev[3]));
        CK(cudaEventElapsedTime(&t->total_ms, ctx->ev[0], ctx->ev[3]));
        if (!pl.b3 && (rc = fill_pass_times(ctx, t, pl, 2, 0))) return rc;
        t->kernel_launches = ctx->launches;
        t->wall_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - w0).count();
    }
    return GJ_OK;
}

extern "C" int gj_join_aggregate_nopart(gj_ctx* ctx, const int32_t* d_Rk, const int32_t* d_Rp, uint64_t nR,
                                        const int32_t* d_Sk, const int32_t* d_Sp, uint64_t nS,
                                        uint64_t* matches, uint64_t* checksum, gj_timings* t);

extern "C" int gj_join_aggregate(gj_ctx* ctx, const int32_t* d_Rk, const int32_t* d_Rp, uint64_t nR,
                                 const int32_t* d_Sk, const int32_t* d_Sp, uint64_t nS,
                                 uint64_t* matches, uint64_t* checksum, gj_timings* t) {
    int rc = check_caps(ctx, nR, nS);
    if (rc) return rc;
    // small build sides: the L2-resident global hash table beats partitioning (option "nopart_max", 0 = never; the
    // default 2^21 comes from the measured crossover, profiles/r2_a/nopart_crossover.log: 2^20 x 2^20 in 41 us against
    // 68 us partitioned, break-even between 2^22 and 2^23).  A forced radix plan ("radix_bits") keeps the partitioned path.
    if (ctx->opt_nopart_max && !ctx->opt_radix_bits && std::min(nR, nS) && std::min(nR, nS) <= (uint64_t)ctx->opt_nopart_max)
        return gj_join_aggregate_nopart(ctx, d_Rk, d_Rp, nR, d_Sk, d_Sp, nS, matches, checksum, t);
    if ((nR && (!d_Rk || !d_Rp)) || (nS && (!d_Sk || !d_Sp))) return fail(GJ_ERR_ARG, "NULL input column");
    Rel R, S;
    R.keys = d_Rk; R.pays = d_Rp; R.n = nR; R.slot = 0;
    S.keys = d_Sk; S.pays = d_Sp; S.n = nS; S.slot = 1;
    return run_join(ctx, R, S, false, nullptr, nullptr, 0, matches, checksum, nullptr, t);
}

extern "C" int gj_join_aggregate_tuples(gj_ctx* ctx, const void* d_Rtup, uint64_t nR, const void* d_Stup,
                                        uint64_t nS, uint64_t* matches, uint64_t* checksum, gj_timings* t) {
    int rc = check_caps(ctx, nR, nS);
    if (rc) return rc;
    if ((nR && !d_Rtup) || (nS && !d_Stup)) return fail(GJ_ERR_ARG, "NULL input");
    if (((size_t)d_Rtup | (size_t)d_Stup) & 7u) return fail(GJ_ERR_ARG, "packed tuples must be 8-byte aligned");
    Rel R, S;
    R.tup = (const tup_t*)d_Rtup; R.n = nR; R.slot = 0;
    S.tup = (const tup_t*)d_Stup; S.n = nS; S.slot = 1;
    return run_join(ctx, R, S, false, nullptr, nullptr, 0, matches, checksum, nullptr, t);
}

extern "C" int gj_join_materialize(gj_ctx* ctx, const int32_t* d_Rk, const int32_t* d_Rp, uint64_t nR,
                                   const int32_t* d_Sk, const int32_t* d_Sp, uint64_t nS, int32_t* d_out_Rp,
                                   int32_t* d_out_Sp, uint64_t cap, uint64_t* n_pairs, uint64_t* checksum,
                                   gj_timings* t) {
    int rc = check_caps(ctx, nR, nS);
    if (rc) return rc;
    if ((nR && (!d_Rk || !d_Rp)) || (nS && (!d_Sk || !d_Sp))) return fail(GJ_ERR_ARG, "NULL input column");
    if (cap && (!d_out_Rp || !d_out_Sp)) return fail(GJ_ERR_ARG, "NULL output column with cap > 0");
    Rel R, S;
    R.keys = d_Rk; R.pays = d_Rp; R.n = nR; R.slot = 0;
    S.keys = d_Sk; S.pays = d_Sp; S.n = nS; S.slot = 1;
    uint64_t m = 0;
    rc = run_join(ctx, R, S, true, d_out_Rp, d_out_Sp, cap, &m, checksum, n_pairs, t);
    if (rc == GJ_OK && n_pairs && *n_pairs != m)
        return fail(GJ_ERR_STATE, "internal: pairs reserved (%llu) != matches (%llu)", (unsigned long long)*n_pairs, (unsigned long long)m);
    return rc;
}

// Late materialisation (reference join_partitioned_varpayload,
// join-primitives.cu:1420-1557, driver outOfGPU_Join_payload_var, hash_join_clustered_probe.cu:542-708)
extern "C" int gj_join_aggregate_late(gj_ctx* ctx, const int32_t* d_Rk, const int32_t* d_Rid, uint64_t nR,
                                      const int32_t* d_Sk, const int32_t* d_Sid, uint64_t nS,
                                      const int32_t* d_Dr, uint32_t cols_r, uint64_t stride_r,
                                      const int32_t* d_Ds, uint32_t cols_s, uint64_t stride_s,
                                      uint64_t* matches, uint64_t* sum, gj_timings* t) {
    int rc = check_caps(ctx, nR, nS);
    if (rc) return rc;
    if ((nR && (!d_Rk || !d_Rid)) || (nS && (!d_Sk || !d_Sid))) return fail(GJ_ERR_ARG, "NULL input column");
    if ((cols_r && !d_Dr) || (cols_s && !d_Ds)) return fail(GJ_ERR_ARG, "NULL side table with a non-zero column count");
    if (cols_r > 64 || cols_s > 64) return fail(GJ_ERR_ARG, "at most 64 side-table columns per relation");
    Rel R, S;
    R.keys = d_Rk; R.pays = d_Rid; R.n = nR; R.slot = 0;
    S.keys = d_Sk; S.pays = d_Sid; S.n = nS; S.slot = 1;
    LateCols late;
    late.cols[0] = d_Dr; late.ncols[0] = cols_r; late.stride[0] = stride_r;
    late.cols[1] = d_Ds; late.ncols[1] = cols_s; late.stride[1] = stride_s;
    return run_join(ctx, R, S, false, nullptr, nullptr, 0, matches, sum, nullptr, t, &late);
}

// Non-partitioned baseline (reference build_ht_chains / chains_probing,
// join-primitives.cu:681-742): memset of the heads + build + probe, no radix pass.
extern "C" int gj_join_aggregate_nopart(gj_ctx* ctx, const int32_t* d_Rk, const int32_t* d_Rp, uint64_t nR,
                                        const int32_t* d_Sk, const int32_t* d_Sp, uint64_t nS,
                                        uint64_t* matches, uint64_t* checksum, gj_timings* t) {
    const auto w0 = std::chrono::steady_clock::now();
    int rc = check_caps(ctx, nR, nS);
    if (rc) return rc;
    if ((nR && (!d_Rk || !d_Rp)) || (nS && (!d_Sk || !d_Sp))) return fail(GJ_ERR_ARG, "NULL input column");
    if (t) memset(t, 0, sizeof(*t));
    if (matches) *matches = 0;
    if (checksum) *checksum = 0;
    CK(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    if (!nR || !nS) return GJ_OK;
    const bool swap = nR > nS;     // build on the smaller relation
    const int32_t* bk = swap ? d_Sk : d_Rk; const int32_t* bp = swap ? d_Sp : d_Rp;
    const int32_t* pk = swap ? d_Rk : d_Sk; const int32_t* pp = swap ? d_Rp : d_Sp;
    const uint64_t nb = swap ? nS : nR, np = swap ? nR : nS;
    uint32_t hb = 4;
    while (hb < 31 && (1ull << hb) < 2 * nb) ++hb;    // load factor <= 1/2
    gj_ctx::NP& q = ctx->np;
    if (q.heads_cap < (1ull << hb)) {
        cudaFree(q.heads); q.heads = nullptr; q.heads_cap = 0;
        if (cudaMalloc(&q.heads, sizeof(uint32_t) << hb) != cudaSuccess) { cudaGetLastError(); return fail(GJ_ERR_NOMEM, "hash-table heads (%.1f MB)", (4ull << hb) * 1e-6); }
        q.heads_cap = 1ull << hb;
    }
    if (q.next_cap < nb) {
        cudaFree(q.next); q.next = nullptr; q.next_cap = 0;
        if (cudaMalloc(&q.next, nb * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); return fail(GJ_ERR_NOMEM, "hash-table links"); }
        q.next_cap = nb;
    }
    cudaStream_t s = ctx->stream;
    CK(cudaMemsetAsync(ctx->zero_common, 0, ctx->zero_common_bytes, s));
    CK(cudaEventRecord(ctx->ev[0], s));
    CK(cudaMemsetAsync(q.heads, 0, sizeof(uint32_t) << hb, s));
    const uint32_t gb = (uint32_t)std::min<uint64_t>((nb + 255) / 256, (uint64_t)ctx->sm_count * 16);
    np_build_kernel<<<gb, 256, 0, s>>>(bk, (uint32_t)nb, hb, q.heads, q.next);
    LAUNCHED();
    CK(cudaEventRecord(ctx->ev[2], s));
    const uint32_t gp = (uint32_t)std::min<uint64_t>((np + 255) / 256, (uint64_t)ctx->sm_count * 16);
    np_probe_kernel<<<gp, 256, 0, s>>>(bk, bp, q.heads, q.next, hb, pk, pp, (uint32_t)np, ctx->result);
    LAUNCHED();
    CK(cudaEventRecord(ctx->ev[3], s));
    CK(cudaMemcpyAsync(ctx->h_result, ctx->result, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (matches) *matches = ctx->h_result[0];
    if (checksum) *checksum = ctx->h_result[1];
    if (t) {
        CK(cudaEventElapsedTime(&t->hist_ms, ctx->ev[0], ctx->ev[2]));   // "hist" slot: table clear + build
        CK(cudaEventElapsedTime(&t->join_ms, ctx->ev[2], ctx->ev[3]));   // probe
        CK(cudaEventElapsedTime(&t->total_ms, ctx->ev[0], ctx->ev[3]));
        t->kernel_launches = ctx->launches;
        t->wall_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - w0).count();
    }
    return GJ_OK;
}

// Perfect-array variant of the non-partitioned baseline (reference build_perfect_array /
// probe_perfect_array, join-primitives.cu:628-668): the build keys must be unique and lie in
// [key_min, key_min + key_range); the key addresses the table.  The build side is the smaller
// relation (R when equal).  The precondition is checked on the device: GJ_ERR_ARG if a build key is
// out of range or occurs twice (the outputs are then not meaningful).
extern "C" int gj_join_aggregate_perfect(gj_ctx* ctx, const int32_t* d_Rk, const int32_t* d_Rp, uint64_t nR,
                                         const int32_t* d_Sk, const int32_t* d_Sp, uint64_t nS,
                                         int32_t key_min, uint64_t key_range,
                                         uint64_t* matches, uint64_t* checksum, gj_timings* t) {
    const auto w0 = std::chrono::steady_clock::now();
    int rc = check_caps(ctx, nR, nS);
    if (rc) return rc;
    if ((nR && (!d_Rk || !d_Rp)) || (nS && (!d_Sk || !d_Sp))) return fail(GJ_ERR_ARG, "NULL input column");
    if (key_range == 0 || key_range > (1ull << 32)) return fail(GJ_ERR_ARG, "key_range must be in [1, 2^32]");
    if (t) memset(t, 0, sizeof(*t));
    if (matches) *matches = 0;
    if (checksum) *checksum = 0;
    CK(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    if (!nR || !nS) return GJ_OK;
    const bool swap = nR > nS;
    const int32_t* bk = swap ? d_Sk : d_Rk; const int32_t* bp = swap ? d_Sp : d_Rp;
    const int32_t* pk = swap ? d_Rk : d_Sk; const int32_t* pp = swap ? d_Rp : d_Sp;
    const uint64_t nb = swap ? nS : nR, np = swap ? nR : nS;
    gj_ctx::NP& q = ctx->np;
    if (q.slots_cap < key_range) {
        cudaFree(q.slots); q.slots = nullptr; q.slots_cap = 0;
        if (cudaMalloc(&q.slots, key_range * sizeof(unsigned long long)) != cudaSuccess) { cudaGetLastError(); return fail(GJ_ERR_NOMEM, "perfect array (%.1f MB)", key_range * 8e-6); }
        q.slots_cap = key_range;
    }
    cudaStream_t s = ctx->stream;
    CK(cudaMemsetAsync(ctx->zero_common, 0, ctx->zero_common_bytes, s));
    uint32_t* status = reinterpret_cast<uint32_t*>(ctx->result + 2);   // result[2] is unused here: two u32 status words
    CK(cudaEventRecord(ctx->ev[0], s));
    CK(cudaMemsetAsync(q.slots, 0, key_range * sizeof(unsigned long long), s));
    const uint32_t range32 = key_range >= (1ull << 32) ? 0xFFFFFFFFu : (uint32_t)key_range;   // 2^32: every key is in range
    const uint32_t gb = (uint32_t)std::min<uint64_t>((nb + 255) / 256, (uint64_t)ctx->sm_count * 16);
    np_build_perfect_kernel<<<gb, 256, 0, s>>>(bk, bp, (uint32_t)nb, key_min, range32, q.slots, status);
    LAUNCHED();
    CK(cudaEventRecord(ctx->ev[2], s));
    const uint32_t gp = (uint32_t)std::min<uint64_t>((np + 255) / 256, (uint64_t)ctx->sm_count * 16);
    np_probe_perfect_kernel<<<gp, 256, 0, s>>>(q.slots, key_min, range32, pk, pp, (uint32_t)np, ctx->result);
    LAUNCHED();
    CK(cudaEventRecord(ctx->ev[3], s));
    CK(cudaMemcpyAsync(ctx->h_result, ctx->result, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const uint32_t out_of_range = (uint32_t)ctx->h_result[2], dups = (uint32_t)(ctx->h_result[2] >> 32);
    if (out_of_range || dups)
        return fail(GJ_ERR_ARG, "perfect array: %u build keys outside [%d, %d + %llu), %u duplicate build keys", out_of_range,
                    key_min, key_min, (unsigned long long)key_range, dups);
    if (matches) *matches = ctx->h_result[0];
    if (checksum) *checksum = ctx->h_result[1];
    if (t) {
        CK(cudaEventElapsedTime(&t->hist_ms, ctx->ev[0], ctx->ev[2]));   // "hist" slot: table clear + build
        CK(cudaEventElapsedTime(&t->join_ms, ctx->ev[2], ctx->ev[3]));   // probe
        CK(cudaEventElapsedTime(&t->total_ms, ctx->ev[0], ctx->ev[3]));
        t->kernel_launches = ctx->launches;
        t->wall_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - w0).count();
    }
    return GJ_OK;
}

// ------------------------------------------------------------------------------------------
// end-to-end host entry: H2D copies chunked on a copy stream, histograms chase the key chunks
// ------------------------------------------------------------------------------------------
static int ensure_host_staging(gj_ctx* ctx) {
    if (ctx->d_in[0] && ctx->d_in[1] && ctx->d_in[2] && ctx->d_in[3]) return GJ_OK;
    const uint64_t caps[4] = {ctx->maxR, ctx->maxR, ctx->maxS, ctx->maxS};
    for (int i = 0; i < 4; ++i)
        if (!ctx->d_in[i] && cudaMalloc(&ctx->d_in[i], caps[i] * sizeof(int32_t)) != cudaSuccess) {
            cudaGetLastError();
            for (int k = 0; k < 4; ++k) { cudaFree(ctx->d_in[k]); ctx->d_in[k] = nullptr; }   // all or nothing
            return fail(GJ_ERR_NOMEM, "cudaMalloc of the host-entry staging columns failed");
        }
    return GJ_OK;
}

extern "C" int gj_join_aggregate_host(gj_ctx* ctx, const int32_t* h_Rk, const int32_t* h_Rp, uint64_t nR,
                                      const int32_t* h_Sk, const int32_t* h_Sp, uint64_t nS,
                                      uint64_t* matches, uint64_t* checksum, gj_timings* t) {
    const auto w0 = std::chrono::steady_clock::now();
    int rc = check_caps(ctx, nR, nS);
    if (rc) return rc;
    if ((nR && (!h_Rk || !h_Rp)) || (nS && (!h_Sk || !h_Sp))) return fail(GJ_ERR_ARG, "NULL input column");
    if (t) memset(t, 0, sizeof(*t));
    if (matches) *matches = 0;
    if (checksum) *checksum = 0;
    CK(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    if (!nR || !nS) return GJ_OK;
    if ((rc = ensure_host_staging(ctx))) return rc;
    cudaStream_t s = ctx->stream, c = ctx->copy_stream;
    const bool swap = nR > nS;
    // role 0 = build, role 1 = probe
    const int32_t* hk[2] = {swap ? h_Sk : h_Rk, swap ? h_Rk : h_Sk};
    const int32_t* hp[2] = {swap ? h_Sp : h_Rp, swap ? h_Rp : h_Sp};
    int32_t* dk[2] = {ctx->d_in[swap ? 2 : 0], ctx->d_in[swap ? 0 : 2]};
    int32_t* dp[2] = {ctx->d_in[swap ? 3 : 1], ctx->d_in[swap ? 1 : 3]};
    const uint64_t nn[2] = {swap ? nS : nR, swap ? nR : nS};
    const int slot[2] = {swap ? 1 : 0, swap ? 0 : 1};
    const Plan pl = choose_plan(ctx, nn[0], 0);
    fill_plan(t, pl);

    CK(cudaMemsetAsync(ctx->zero_block, 0, ctx->zero_bytes, s));
    CK(cudaEventRecord(ctx->ev[0], s));
    CK(cudaStreamWaitEvent(c, ctx->ev[0], 0));   // copies start with the timed window
    int evi = 0;
    const uint64_t chunk = (uint64_t)ctx->opt_h2d_chunk;
    for (int r = 0; r < 2; ++r) {
        for (uint64_t o = 0; o < nn[r]; o += chunk) {
            const uint64_t m = std::min(chunk, nn[r] - o);
            CK(cudaMemcpyAsync(dk[r] + o, hk[r] + o, m * sizeof(int32_t), cudaMemcpyHostToDevice, c));
            cudaEvent_t e = ctx->cev[evi++ % N_EVENTS];
            CK(cudaEventRecord(e, c));
            CK(cudaStreamWaitEvent(s, e, 0));
            if ((rc = enqueue_hist(ctx, s, dk[r] + o, false, m, 0, pl.B, ctx->meta[r].ghist))) return rc;
        }
    }
    cudaEvent_t pay_ev[2];
    for (int r = 0; r < 2; ++r) {
        CK(cudaMemcpyAsync(dp[r], hp[r], nn[r] * sizeof(int32_t), cudaMemcpyHostToDevice, c));
        pay_ev[r] = ctx->cev[evi++ % N_EVENTS];
        CK(cudaEventRecord(pay_ev[r], c));
    }
    CK(cudaEventRecord(ctx->ev[4], c));   // end of all H2D traffic
    if ((rc = enqueue_scan(ctx, s, 0, 2, 1u << pl.B, true))) return rc;
    if ((rc = enqueue_plan(ctx, s, 0, 2, pl, true))) return rc;
    CK(cudaEventRecord(ctx->ev[1], s));
    for (int r = 0; r < 2; ++r) {
        Rel rel;
        rel.keys = dk[r]; rel.pays = dp[r]; rel.n = nn[r]; rel.slot = slot[r];
        CK(cudaStreamWaitEvent(s, pay_ev[r], 0));
        if ((rc = enqueue_scatter(ctx, s, rel, r, pl, ctx->out[slot[r]]))) return rc;
    }
    CK(cudaEventRecord(ctx->ev[2], s));
    if ((rc = enqueue_join(ctx, s, ctx->out[slot[0]], ctx->out[slot[1]], pl, nn[0], nn[1], false, nullptr, nullptr, 0))) return rc;
    CK(cudaEventRecord(ctx->ev[3], s));
    CK(cudaMemcpyAsync(ctx->h_result, ctx->result, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaStreamSynchronize(c));
    if (matches) *matches = ctx->h_result[0];
    if (checksum) *checksum = ctx->h_result[1];
    if (t) {
        CK(cudaEventElapsedTime(&t->hist_ms, ctx->ev[0], ctx->ev[1]));
        CK(cudaEventElapsedTime(&t->part_ms, ctx->ev[1], ctx->ev[2]));
        CK(cudaEventElapsedTime(&t->join_ms, ctx->ev[2], ctx->ev[3]));
        CK(cudaEventElapsedTime(&t->total_ms, ctx->ev[0], ctx->ev[3]));
        CK(cudaEventElapsedTime(&t->h2d_ms, ctx->ev[0], ctx->ev[4]));
        if ((rc = fill_pass_times(ctx, t, pl, 2, 0))) return rc;
        t->kernel_launches = ctx->launches;
        t->wall_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - w0).count();
    }
    return GJ_OK;
}

// Out-of-HBM probe side (reference outOfGPU_Join3_payload,
// hash_join_clustered_probe.cu:1684-1984): R (the build side) is copied and partitioned once and stays
// resident; S streams from host memory in chunks through a double buffer -- H2D of chunk i+1 on the copy
// stream while chunk i is partitioned and joined against R's partitions -- into the same accumulators.
extern "C" int gj_join_aggregate_stream_host(gj_ctx* ctx, const int32_t* h_Rk, const int32_t* h_Rp, uint64_t nR,
                                             const int32_t* h_Sk, const int32_t* h_Sp, uint64_t nS,
                                             uint64_t chunk_tuples, uint64_t* matches, uint64_t* checksum,
                                             gj_timings* t) {
    const auto w0 = std::chrono::steady_clock::now();
    if (!ctx) return fail(GJ_ERR_ARG, "ctx is NULL");
    if (nR > ctx->maxR) return fail(GJ_ERR_ARG, "the build side (%llu tuples) exceeds the context capacity (%llu)", (unsigned long long)nR, (unsigned long long)ctx->maxR);
    if (chunk_tuples == 0 || 2 * chunk_tuples > ctx->maxS) return fail(GJ_ERR_ARG, "chunk_tuples must be in [1, max_S / 2] (double buffer)");
    if ((nR && (!h_Rk || !h_Rp)) || (nS && (!h_Sk || !h_Sp))) return fail(GJ_ERR_ARG, "NULL input column");
    if (t) memset(t, 0, sizeof(*t));
    if (matches) *matches = 0;
    if (checksum) *checksum = 0;
    CK(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    if (!nR || !nS) return GJ_OK;
    int rc;
    if ((rc = ensure_host_staging(ctx))) return rc;
    cudaStream_t s = ctx->stream, c = ctx->copy_stream;
    const Plan pl = choose_plan(ctx, nR, 0);
    fill_plan(t, pl);
    const uint32_t nb = 1u << pl.B;
    int evi = 0;
    auto next_ev = [&]() { return ctx->cev[evi++ % N_EVENTS]; };

    // ---- build side: copy, histogram, offsets, radix passes; its histogram / offsets stay for every chunk
    CK(cudaMemsetAsync(ctx->zero_block, 0, ctx->zero_bytes, s));
    CK(cudaEventRecord(ctx->ev[0], s));
    CK(cudaStreamWaitEvent(c, ctx->ev[0], 0));
    CK(cudaMemcpyAsync(ctx->d_in[0], h_Rk, nR * sizeof(int32_t), cudaMemcpyHostToDevice, c));
    cudaEvent_t e_rk = next_ev(); CK(cudaEventRecord(e_rk, c));
    CK(cudaMemcpyAsync(ctx->d_in[1], h_Rp, nR * sizeof(int32_t), cudaMemcpyHostToDevice, c));
    cudaEvent_t e_rp = next_ev(); CK(cudaEventRecord(e_rp, c));
    CK(cudaStreamWaitEvent(s, e_rk, 0));
    if ((rc = enqueue_hist(ctx, s, ctx->d_in[0], false, nR, 0, pl.B, ctx->meta[0].ghist))) return rc;
    if ((rc = enqueue_scan(ctx, s, 0, 1, nb, false))) return rc;
    if ((rc = enqueue_plan(ctx, s, 0, 1, pl, false))) return rc;
    CK(cudaStreamWaitEvent(s, e_rp, 0));
    Rel R;
    R.keys = ctx->d_in[0]; R.pays = ctx->d_in[1]; R.n = nR; R.slot = 0;
    if ((rc = enqueue_scatter(ctx, s, R, 0, pl, ctx->out[0]))) return rc;
    CK(cudaEventRecord(ctx->ev[1], s));

    // ---- probe side, chunk by chunk
    const size_t res_off = (size_t)(reinterpret_cast<unsigned char*>(ctx->result) - ctx->zero_common);
    cudaEvent_t buf_free[2] = {nullptr, nullptr};
    uint64_t ci = 0;
    for (uint64_t o = 0; o < nS; o += chunk_tuples, ++ci) {
        const uint64_t m = std::min(chunk_tuples, nS - o);
        const int b = (int)(ci & 1);
        int32_t* dk = ctx->d_in[2] + (size_t)b * chunk_tuples;
        int32_t* dp = ctx->d_in[3] + (size_t)b * chunk_tuples;
        if (buf_free[b]) CK(cudaStreamWaitEvent(c, buf_free[b], 0));     // chunk ci - 2 has been scattered out of it
        CK(cudaMemcpyAsync(dk, h_Sk + o, m * sizeof(int32_t), cudaMemcpyHostToDevice, c));
        cudaEvent_t e_k = next_ev(); CK(cudaEventRecord(e_k, c));
        CK(cudaMemcpyAsync(dp, h_Sp + o, m * sizeof(int32_t), cudaMemcpyHostToDevice, c));
        cudaEvent_t e_p = next_ev(); CK(cudaEventRecord(e_p, c));
        CK(cudaMemsetAsync(ctx->zero_role[1], 0, ctx->zero_role_bytes, s));
        CK(cudaStreamWaitEvent(s, e_k, 0));
        if ((rc = enqueue_hist(ctx, s, dk, false, m, 0, pl.B, ctx->meta[1].ghist))) return rc;
        if ((rc = enqueue_scan(ctx, s, 1, 1, nb, false))) return rc;
        if ((rc = enqueue_plan(ctx, s, 1, 1, pl, false))) return rc;
        CK(cudaStreamWaitEvent(s, e_p, 0));
        Rel S;
        S.keys = dk; S.pays = dp; S.n = m; S.slot = 1;
        if ((rc = enqueue_scatter(ctx, s, S, 1, pl, ctx->out[1]))) return rc;
        buf_free[b] = next_ev();
        CK(cudaEventRecord(buf_free[b], s));
        // units of this chunk against the resident build partitions; the accumulators keep counting
        CK(cudaMemsetAsync(ctx->zero_common, 0, res_off, s));
        if ((rc = enqueue_scan(ctx, s, 0, 0, nb, true))) return rc;
        if ((rc = enqueue_plan(ctx, s, 0, 0, pl, true))) return rc;
        if ((rc = enqueue_join(ctx, s, ctx->out[0], ctx->out[1], pl, nR, m, false, nullptr, nullptr, 0))) return rc;
    }
    CK(cudaEventRecord(ctx->ev[4], c));   // end of all H2D traffic
    CK(cudaEventRecord(ctx->ev[3], s));
    CK(cudaMemcpyAsync(ctx->h_result, ctx->result, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaStreamSynchronize(c));
    if (matches) *matches = ctx->h_result[0];
    if (checksum) *checksum = ctx->h_result[1];
    if (t) {
        CK(cudaEventElapsedTime(&t->hist_ms, ctx->ev[0], ctx->ev[1]));    // build side: copy + partition
        CK(cudaEventElapsedTime(&t->join_ms, ctx->ev[1], ctx->ev[3]));    // all probe chunks: copy || partition || join
        CK(cudaEventElapsedTime(&t->total_ms, ctx->ev[0], ctx->ev[3]));
        CK(cudaEventElapsedTime(&t->h2d_ms, ctx->ev[0], ctx->ev[4]));
        t->kernel_launches = ctx->launches;
        t->wall_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - w0).count();
    }
    return GJ_OK;
}

// ------------------------------------------------------------------------------------------
// the partitioner on its own
// ------------------------------------------------------------------------------------------
extern "C" int gj_partition(gj_ctx* ctx, int slot, const int32_t* d_keys, const int32_t* d_pays, uint64_t n,
                            uint32_t radix_bits, const void** d_tuples, const uint32_t** d_offsets,
                            uint32_t* radix_bits_used, gj_timings* t) {
    const auto w0 = std::chrono::steady_clock::now();
    if (!ctx) return fail(GJ_ERR_ARG, "ctx is NULL");
    if (slot != 0 && slot != 1) return fail(GJ_ERR_ARG, "slot must be 0 or 1");
    if (n > (slot ? ctx->maxS : ctx->maxR)) return fail(GJ_ERR_ARG, "n exceeds the capacity of slot %d", slot);
    if (n && (!d_keys || !d_pays)) return fail(GJ_ERR_ARG, "NULL input column");
    if (radix_bits > (uint32_t)MAX_RADIX_BITS) return fail(GJ_ERR_ARG, "radix_bits <= %d", MAX_RADIX_BITS);
    if (t) memset(t, 0, sizeof(*t));
    CK(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    cudaStream_t s = ctx->stream;
    const Plan pl = choose_plan(ctx, n, radix_bits);
    fill_plan(t, pl);
    Rel rel;
    rel.keys = d_keys; rel.pays = d_pays; rel.n = n; rel.slot = slot;
    // the standalone partitioner uses role `slot` of the metadata so two calls (R then S) coexist
    CK(cudaMemsetAsync(ctx->zero_block, 0, ctx->zero_bytes, s));
    CK(cudaEventRecord(ctx->ev[0], s));
    int rc;
    if ((rc = enqueue_hist(ctx, s, d_keys, false, n, 0, pl.B, ctx->meta[slot].ghist))) return rc;
    if ((rc = enqueue_scan(ctx, s, slot, 1, 1u << pl.B, false))) return rc;
    if ((rc = enqueue_plan(ctx, s, slot, 1, pl, false))) return rc;
    CK(cudaEventRecord(ctx->ev[1], s));
    if ((rc = enqueue_scatter(ctx, s, rel, slot, pl, ctx->out[slot]))) return rc;
    CK(cudaEventRecord(ctx->ev[2], s));
    CK(cudaStreamSynchronize(s));
    if (d_tuples) *d_tuples = ctx->out[slot];
    if (d_offsets) *d_offsets = ctx->meta[slot].off;
    if (radix_bits_used) *radix_bits_used = pl.B;
    if (t) {
        CK(cudaEventElapsedTime(&t->hist_ms, ctx->ev[0], ctx->ev[1]));
        CK(cudaEventElapsedTime(&t->part_ms, ctx->ev[1], ctx->ev[2]));
        CK(cudaEventElapsedTime(&t->total_ms, ctx->ev[0], ctx->ev[2]));
        if (n && (rc = fill_pass_times(ctx, t, pl, 1, slot))) return rc;
        t->kernel_launches = ctx->launches;
        t->wall_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - w0).count();
    }
    return GJ_OK;
}

// ------------------------------------------------------------------------------------------
// multi-GPU shuffle step
// ------------------------------------------------------------------------------------------
static int shuffle_args_ok(gj_ctx* ctx, uint64_t n, uint32_t n_gpus, uint32_t gpu_shift, uint32_t* bits) {
    if (!ctx) return fail(GJ_ERR_ARG, "ctx is NULL");
    if (n_gpus == 0 || n_gpus > (uint32_t)NB_MAX || (n_gpus & (n_gpus - 1))) return fail(GJ_ERR_ARG, "n_gpus must be a power of two <= %d", NB_MAX);
    if (gpu_shift > 31) return fail(GJ_ERR_ARG, "gpu_shift <= 31");
    if (n > std::max(ctx->maxR, ctx->maxS)) return fail(GJ_ERR_ARG, "n exceeds the context capacity");
    uint32_t b = 0;
    while ((1u << b) < n_gpus) ++b;
    *bits = b;
    return GJ_OK;
}

extern "C" int gj_shuffle_count(gj_ctx* ctx, const int32_t* d_keys, uint64_t n, uint32_t n_gpus,
                                uint32_t gpu_shift, uint64_t* h_counts) {
    uint32_t bits = 0;
    int rc = shuffle_args_ok(ctx, n, n_gpus, gpu_shift, &bits);
    if (rc) return rc;
    if (!h_counts || (n && !d_keys)) return fail(GJ_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    CK(cudaMemsetAsync(ctx->zero_block, 0, ctx->zero_bytes, s));
    if ((rc = enqueue_hist(ctx, s, d_keys, false, n, gpu_shift, bits, ctx->meta[0].ghist))) return rc;
    uint32_t tmp[NB_MAX];
    CK(cudaMemcpyAsync(tmp, ctx->meta[0].ghist, n_gpus * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (uint32_t g = 0; g < n_gpus; ++g) h_counts[g] = tmp[g];
    return GJ_OK;
}

extern "C" int gj_shuffle_split(gj_ctx* ctx, const int32_t* d_keys, const int32_t* d_pays, uint64_t n,
                                uint32_t n_gpus, uint32_t gpu_shift, void* d_out_tuples, uint64_t* h_counts) {
    uint32_t bits = 0;
    int rc = shuffle_args_ok(ctx, n, n_gpus, gpu_shift, &bits);
    if (rc) return rc;
    if (!h_counts || (n && (!d_keys || !d_pays || !d_out_tuples))) return fail(GJ_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    cudaStream_t s = ctx->stream;
    CK(cudaMemsetAsync(ctx->zero_block, 0, ctx->zero_bytes, s));
    if ((rc = enqueue_hist(ctx, s, d_keys, false, n, gpu_shift, bits, ctx->meta[0].ghist))) return rc;
    Plan pl; pl.B = bits; pl.b1 = bits; pl.b2 = 0;
    if ((rc = enqueue_scan(ctx, s, 0, 1, n_gpus, false))) return rc;
    if ((rc = enqueue_plan(ctx, s, 0, 1, pl, false))) return rc;
    if (n) {
        const ScatterCfg& c1 = scatter_cfg1(ctx);
        const uint32_t T1 = (uint32_t)(c1.threads * c1.ipt);
        ScatterArgs a;
        memset(&a, 0, sizeof(a));
        a.in_keys = d_keys; a.in_pays = d_pays; a.n = (uint32_t)n; a.out = (tup_t*)d_out_tuples;
        a.shift = gpu_shift; a.bits = bits; a.cursors = ctx->meta[0].cur2; a.cursor_stride = 1;
        a.ntiles = (uint32_t)((n + T1 - 1) / T1);
        c1.col<<<a.ntiles, c1.threads, scatter_smem(c1), s>>>(a);
        LAUNCHED();
    }
    uint32_t tmp[NB_MAX];
    CK(cudaMemcpyAsync(tmp, ctx->meta[0].ghist, n_gpus * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (uint32_t g = 0; g < n_gpus; ++g) h_counts[g] = tmp[g];
    return GJ_OK;
}

extern "C" int gj_shuffle_scatter_peers(gj_ctx* ctx, const int32_t* d_keys, const int32_t* d_pays, uint64_t n,
                                        uint32_t n_gpus, uint32_t gpu_shift, void* const* d_peer_bases,
                                        const uint64_t* h_peer_offsets) {
    uint32_t bits = 0;
    int rc = shuffle_args_ok(ctx, n, n_gpus, gpu_shift, &bits);
    if (rc) return rc;
    if (!d_peer_bases || !h_peer_offsets || (n && (!d_keys || !d_pays))) return fail(GJ_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    cudaStream_t s = ctx->stream;
    uint32_t cur[NB_MAX];
    for (uint32_t g = 0; g < n_gpus; ++g) {
        if (h_peer_offsets[g] > 0xFFFFFFFFull) return fail(GJ_ERR_ARG, "peer offset exceeds 2^32 tuples");
        cur[g] = (uint32_t)h_peer_offsets[g];
    }
    CK(cudaMemcpyAsync(ctx->meta[0].cur2, cur, n_gpus * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_dst_bases, d_peer_bases, n_gpus * sizeof(void*), cudaMemcpyHostToDevice, s));
    if (n) {
        const ScatterCfg& c1 = scatter_cfg_shuffle(ctx);
        const uint32_t T1 = (uint32_t)(c1.threads * c1.ipt);
        ScatterArgs a;
        memset(&a, 0, sizeof(a));
        a.in_keys = d_keys; a.in_pays = d_pays; a.n = (uint32_t)n; a.out = nullptr;
        a.dst_bases = ctx->d_dst_bases;
        a.shift = gpu_shift; a.bits = bits; a.cursors = ctx->meta[0].cur2; a.cursor_stride = 1;
        a.ntiles = (uint32_t)((n + T1 - 1) / T1);
        // optional persistent grid ("shuffle_grid" CTAs looping over the tiles): leaves SM resources
        // to the local passes of the other relation running concurrently on another stream
        const uint32_t grid = ctx->opt_shuffle_grid ? std::min<uint32_t>((uint32_t)ctx->opt_shuffle_grid, a.ntiles) : a.ntiles;
        CK(cudaEventRecord(ctx->sev[0][0], s));
        (grid < a.ntiles ? c1.col_persist : c1.col)<<<grid, c1.threads, scatter_smem(c1), s>>>(a);
        LAUNCHED();
        CK(cudaEventRecord(ctx->sev[0][1], s));
    }
    CK(cudaStreamSynchronize(s));
    ctx->last_shuffle_ms = 0.f;
    if (n) CK(cudaEventElapsedTime(&ctx->last_shuffle_ms, ctx->sev[0][0], ctx->sev[0][1]));
    return GJ_OK;
}


// ------------------------------------------------------------------------------------------
// asynchronous / staged entry points: the multi-GPU host overlaps one relation's shuffle with the
// other relation's local radix passes.  Nothing here synchronises until gj_stage_finish.
// ------------------------------------------------------------------------------------------
extern "C" int gj_shuffle_scatter_peers_async(gj_ctx* ctx, int which, const int32_t* d_keys, const int32_t* d_pays,
                                              uint64_t n, uint32_t n_gpus, uint32_t gpu_shift,
                                              void* const* d_peer_bases, const uint64_t* h_peer_offsets,
                                              void* cuda_stream) {
    uint32_t bits = 0;
    int rc = shuffle_args_ok(ctx, n, n_gpus, gpu_shift, &bits);
    if (rc) return rc;
    if (which != 0 && which != 1) return fail(GJ_ERR_ARG, "which must be 0 or 1");
    if (!d_peer_bases || !h_peer_offsets || (n && (!d_keys || !d_pays))) return fail(GJ_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    // pinned staging: [which][cursors u32 x NB_MAX | bases ptr x NB_MAX]
    unsigned char* hs = ctx->h_shuf + (size_t)which * NB_MAX * (sizeof(uint32_t) + sizeof(void*));
    uint32_t* hcur = reinterpret_cast<uint32_t*>(hs);
    void** hbase = reinterpret_cast<void**>(hs + NB_MAX * sizeof(uint32_t));
    for (uint32_t g = 0; g < n_gpus; ++g) {
        if (h_peer_offsets[g] > 0xFFFFFFFFull) return fail(GJ_ERR_ARG, "peer offset exceeds 2^32 tuples");
        hcur[g] = (uint32_t)h_peer_offsets[g];
        hbase[g] = d_peer_bases[g];
    }
    CK(cudaMemcpyAsync(ctx->shuf_cur[which], hcur, n_gpus * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->shuf_bases[which], hbase, n_gpus * sizeof(void*), cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(ctx->sev[which][0], s));
    if (n) {
        const ScatterCfg& c1 = scatter_cfg_shuffle(ctx);
        const uint32_t T1 = (uint32_t)(c1.threads * c1.ipt);
        ScatterArgs a;
        memset(&a, 0, sizeof(a));
        a.in_keys = d_keys; a.in_pays = d_pays; a.n = (uint32_t)n; a.out = nullptr;
        a.dst_bases = ctx->shuf_bases[which];
        a.shift = gpu_shift; a.bits = bits; a.cursors = ctx->shuf_cur[which]; a.cursor_stride = 1;
        a.ntiles = (uint32_t)((n + T1 - 1) / T1);
        const uint32_t grid = ctx->opt_shuffle_grid ? std::min<uint32_t>((uint32_t)ctx->opt_shuffle_grid, a.ntiles) : a.ntiles;
        (grid < a.ntiles ? c1.col_persist : c1.col)<<<grid, c1.threads, scatter_smem(c1), s>>>(a);
        LAUNCHED();
    }
    CK(cudaEventRecord(ctx->sev[which][1], s));
    return GJ_OK;
}

extern "C" int gj_memcpy_d2d_async(void* dst, const void* src, uint64_t bytes, void* cuda_stream) {
    if (bytes && (!dst || !src)) return fail(GJ_ERR_ARG, "NULL argument");
    if (bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)cuda_stream));
    return GJ_OK;
}

extern "C" int gj_shuffle_scatter_ms(gj_ctx* ctx, int which, float* ms) {
    if (!ctx || !ms || (which != 0 && which != 1)) return fail(GJ_ERR_ARG, "bad argument");
    CK(cudaEventSynchronize(ctx->sev[which][1]));
    CK(cudaEventElapsedTime(ms, ctx->sev[which][0], ctx->sev[which][1]));
    return GJ_OK;
}

extern "C" int gj_stage_begin(gj_ctx* ctx, uint64_t nR, uint64_t nS, void* cuda_stream) {
    int rc = check_caps(ctx, nR, nS);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    const bool swap = nR > nS;
    ctx->stage.role_of_side[0] = swap ? 1 : 0;
    ctx->stage.role_of_side[1] = swap ? 0 : 1;
    ctx->stage.n_side[0] = nR; ctx->stage.n_side[1] = nS;
    ctx->stage.pl = choose_plan(ctx, std::min(nR, nS), 0);
    ctx->stage.active = true;
    ctx->stage.scratch_used = false;
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    CK(cudaMemsetAsync(ctx->zero_common, 0, ctx->zero_common_bytes, s));
    CK(cudaEventRecord(ctx->stage_ev[3], s));   // "common block ready"
    return GJ_OK;
}

extern "C" int gj_stage_partition(gj_ctx* ctx, int side, const void* d_tuples, void* cuda_stream) {
    if (!ctx || !ctx->stage.active) return fail(GJ_ERR_STATE, "gj_stage_begin first");
    if (side != 0 && side != 1) return fail(GJ_ERR_ARG, "side must be 0 (R) or 1 (S)");
    const uint64_t n = ctx->stage.n_side[side];
    if (n && (!d_tuples || ((size_t)d_tuples & 7u))) return fail(GJ_ERR_ARG, "packed tuples must be non-NULL and 8-byte aligned");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    const int role = ctx->stage.role_of_side[side];
    const Plan& pl = ctx->stage.pl;
    int rc;
    CK(cudaMemsetAsync(ctx->zero_role[role], 0, ctx->zero_role_bytes, s));
    // 512-thread histogram CTAs: small enough to share an SM with the persistent peer-scatter CTAs
    if ((rc = enqueue_hist(ctx, s, d_tuples, true, n, 0, pl.B, ctx->meta[role].ghist, 512))) return rc;
    if ((rc = enqueue_scan(ctx, s, role, 1, 1u << pl.B, false))) return rc;
    if ((rc = enqueue_plan(ctx, s, role, 1, pl, false))) return rc;
    if (ctx->stage.scratch_used) CK(cudaStreamWaitEvent(s, ctx->stage_ev[2], 0));   // first-pass buffer is shared
    Rel rel;
    rel.tup = (const tup_t*)d_tuples; rel.n = n; rel.slot = side;
    if ((rc = enqueue_scatter(ctx, s, rel, role, pl, ctx->out[side]))) return rc;
    CK(cudaEventRecord(ctx->stage_ev[2], s));
    ctx->stage.scratch_used = true;
    CK(cudaEventRecord(ctx->stage_ev[side], s));
    return GJ_OK;
}

extern "C" int gj_stage_join(gj_ctx* ctx, void* cuda_stream) {
    if (!ctx || !ctx->stage.active) return fail(GJ_ERR_STATE, "gj_stage_begin first");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    const Plan& pl = ctx->stage.pl;
    const uint64_t nR = ctx->stage.n_side[0], nS = ctx->stage.n_side[1];
    CK(cudaStreamWaitEvent(s, ctx->stage_ev[0], 0));
    CK(cudaStreamWaitEvent(s, ctx->stage_ev[1], 0));
    CK(cudaStreamWaitEvent(s, ctx->stage_ev[3], 0));
    int rc;
    if (nR && nS) {
        const int bside = ctx->stage.role_of_side[0] == 0 ? 0 : 1;   // which side is the build role
        if ((rc = enqueue_scan(ctx, s, 0, 0, 1u << pl.B, true))) return rc;
        if ((rc = enqueue_plan(ctx, s, 0, 0, pl, true))) return rc;
        if ((rc = enqueue_join(ctx, s, ctx->out[bside], ctx->out[1 - bside], pl, bside == 0 ? nR : nS,
                               bside == 0 ? nS : nR, false, nullptr, nullptr, 0))) return rc;
    }
    CK(cudaMemcpyAsync(ctx->h_result, ctx->result, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(ctx->ev[3], s));
    return GJ_OK;
}

extern "C" int gj_stage_finish(gj_ctx* ctx, uint64_t* matches, uint64_t* checksum) {
    if (!ctx || !ctx->stage.active) return fail(GJ_ERR_STATE, "gj_stage_begin first");
    CK(cudaEventSynchronize(ctx->ev[3]));
    ctx->stage.active = false;
    if (matches) *matches = ctx->h_result[0];
    if (checksum) *checksum = ctx->h_result[1];
    return GJ_OK;
}

extern "C" int gj_stage_pass_ms(gj_ctx* ctx, float pass_ms[4]) {
    if (!ctx || !pass_ms) return fail(GJ_ERR_ARG, "NULL argument");
    if (ctx->stage.active) return fail(GJ_ERR_STATE, "gj_stage_finish first");
    for (int i = 0; i < 4; ++i) pass_ms[i] = 0.f;
    const Plan& pl = ctx->stage.pl;
    for (int side = 0; side < 2; ++side) {   // [R pass 1, R pass 2, S pass 1, S pass 2]
        if (!ctx->stage.n_side[side]) continue;
        const int role = ctx->stage.role_of_side[side];
        CK(cudaEventSynchronize(ctx->pev[role][pl.b2 ? 2 : 1]));
        CK(cudaEventElapsedTime(&pass_ms[2 * side], ctx->pev[role][0], ctx->pev[role][1]));
        if (pl.b2) CK(cudaEventElapsedTime(&pass_ms[2 * side + 1], ctx->pev[role][1], ctx->pev[role][2]));
    }
    return GJ_OK;
}

// ------------------------------------------------------------------------------------------
// Sharded "partition, then push" pipeline (multi-GPU; kernels.cuh section 3c).  Per relation:
//   gj_pp_local : coarse histogram -> scan -> cursors -> pass 1 (local) -> pass-2 tile list ->
//                 fine histogram of this shard (2^(gpu bits + local bits) counters)
//   [caller: all-gather of the fine histograms, e.g. ncclAllGather on the same stream]
//   gj_pp_push  : write cursors from the gathered histograms -> pass 2, runs stored into the
//                 destination GPUs' partition buffers (local or peer-mapped)
//   [caller: a cross-rank "all pushes have landed" point, e.g. a 1-element all-reduce]
//   gj_pp_join  : unit planning + join over what this GPU received (already fully partitioned)
// Nothing synchronises with the host before gj_pp_finish.  ctx->out[] serve as first-pass buffers.
// ------------------------------------------------------------------------------------------
static int ensure_pp(gj_ctx* ctx) {
    gj_ctx::PP& q = ctx->pp;
    if (q.block) return GJ_OK;
    const size_t NQ = (size_t)1 << PP_MAX_BITS, NC = (size_t)1 << PP_MAX_PASS_BITS;
    size_t b = 0, o_cur[2], o_tp[2], o_bs[2], o_z[2];
    for (int r = 0; r < 2; ++r) {
        o_cur[r] = b; b += NQ * sizeof(uint32_t);
        o_tp[r] = b;  b += (NC + 4) * sizeof(uint32_t);
        o_bs[r] = b;  b += NB_MAX * sizeof(tup_t*);
    }
    q.zero_bytes = 64;   // [tile-scan descriptor u64 x 2][ticket u32 .. pad][status u32 x 4]
    for (int r = 0; r < 2; ++r) { o_z[r] = b; b += q.zero_bytes; }
    if (cudaMalloc(&q.block, b) != cudaSuccess) { cudaGetLastError(); return fail(GJ_ERR_NOMEM, "sharded-pipeline metadata (%.1f MB)", b * 1e-6); }
    for (int r = 0; r < 2; ++r) {
        q.cur_fine[r] = reinterpret_cast<uint32_t*>(q.block + o_cur[r]);
        q.tile_prefix[r] = reinterpret_cast<uint32_t*>(q.block + o_tp[r]);
        q.bases[r] = reinterpret_cast<tup_t**>(q.block + o_bs[r]);
        q.zero[r] = q.block + o_z[r];
        q.tdesc[r] = reinterpret_cast<unsigned long long*>(q.zero[r]);
        q.tticket[r] = reinterpret_cast<uint32_t*>(q.zero[r] + 16);
        q.status[r] = reinterpret_cast<uint32_t*>(q.zero[r] + 32);
    }
    CK(cudaHostAlloc(&q.h_pin, 2 * NB_MAX * sizeof(void*) + 64, cudaHostAllocDefault));
    for (auto& r : q.ev) for (auto& e : r) CK(cudaEventCreate(&e));
    for (auto& e : q.jev) CK(cudaEventCreate(&e));
    return GJ_OK;
}

static const PPCfg& pp_first_cfg(uint32_t bits) { return kPPFirst[bits > 8u ? bits - 8u : 0u]; }
static const PPCfg& pp_push_cfg(const gj_ctx* ctx, uint32_t bits) {
    const uint32_t i = bits > 8u ? bits - 8u : 0u;
    return (ctx->pp.big ? kPPPushBig : kPPPush)[ctx->pp.out ? 1 : 0][i];
}

extern "C" int gj_pp_begin(gj_ctx* ctx, uint64_t n_R_global, uint64_t n_S_global, uint32_t n_gpus, uint32_t rank,
                           uint32_t local_bits, void* cuda_stream) {
    if (!ctx) return fail(GJ_ERR_ARG, "ctx is NULL");
    if (n_gpus == 0 || n_gpus > (uint32_t)NB_MAX || (n_gpus & (n_gpus - 1))) return fail(GJ_ERR_ARG, "n_gpus must be a power of two <= %d", NB_MAX);
    if (rank >= n_gpus) return fail(GJ_ERR_ARG, "rank %u out of range", rank);
    uint32_t g = 0;
    while ((1u << g) < n_gpus) ++g;
    if (local_bits < 1 || local_bits > (uint32_t)MAX_RADIX_BITS) return fail(GJ_ERR_ARG, "local_bits must be in [1, %d]", MAX_RADIX_BITS);
    const uint32_t Btot = g + local_bits;
    if (Btot < 2 || Btot > PP_MAX_BITS) return fail(GJ_ERR_ARG, "gpu bits + local bits = %u outside [2, %u]", Btot, PP_MAX_BITS);
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_pp(ctx))) return rc;
    gj_ctx::PP& q = ctx->pp;
    q.G = n_gpus; q.rank = rank; q.g = g; q.B = local_bits; q.Btot = Btot;
    // the first pass must cover the GPU bits (a first-pass partition belongs to ONE destination)
    // default: the pushing pass gets the smaller half -- its runs (tile / 2^b2 tuples) cross NVLink
    uint32_t b1 = ctx->opt_pass1_bits ? (uint32_t)ctx->opt_pass1_bits : (Btot + 1) / 2;
    b1 = std::max(b1, std::max(g, 1u));
    if (Btot - b1 > PP_MAX_PASS_BITS) b1 = Btot - PP_MAX_PASS_BITS;
    b1 = std::min(b1, std::min(PP_MAX_PASS_BITS, Btot - 1));
    if (b1 < g) return fail(GJ_ERR_ARG, "%u GPU bits do not fit the first pass", g);
    q.b1 = b1; q.b2 = Btot - b1; q.out = (uint32_t)ctx->opt_pp_out; q.big = (uint32_t)ctx->opt_pp_tile16k;
    const bool swap = n_R_global > n_S_global;   // build on the smaller relation -- same choice on every rank
    q.role_of_side[0] = swap ? 1 : 0;
    q.role_of_side[1] = swap ? 0 : 1;
    q.n_glob[0] = n_R_global; q.n_glob[1] = n_S_global;
    q.active = true;
    ctx->launches = 0;
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    CK(cudaMemsetAsync(ctx->zero_common, 0, ctx->zero_common_bytes, s));
    CK(cudaEventRecord(ctx->stage_ev[3], s));   // "common block ready"
    return GJ_OK;
}

extern "C" int gj_pp_local(gj_ctx* ctx, int which, const int32_t* d_keys, const int32_t* d_pays, uint64_t n,
                           uint32_t* d_fine_hist, void* cuda_stream) {
    if (!ctx || !ctx->pp.active) return fail(GJ_ERR_STATE, "gj_pp_begin first");
    if (which != 0 && which != 1) return fail(GJ_ERR_ARG, "which must be 0 (R) or 1 (S)");
    if (n > (which ? ctx->maxS : ctx->maxR)) return fail(GJ_ERR_ARG, "n exceeds the context capacity");
    if (!d_fine_hist || (n && (!d_keys || !d_pays))) return fail(GJ_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    gj_ctx::PP& q = ctx->pp;
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    const int role = q.role_of_side[which];
    const RelMeta& m = ctx->meta[role];
    int rc;
    CK(cudaMemsetAsync(ctx->zero_role[role], 0, ctx->zero_role_bytes, s));
    CK(cudaMemsetAsync(q.zero[which], 0, q.zero_bytes, s));
    CK(cudaMemsetAsync(d_fine_hist, 0, sizeof(uint32_t) << q.Btot, s));
    CK(cudaEventRecord(q.ev[which][0], s));
    // coarse (first-pass) histogram -> offsets -> cursors
    const uint32_t n1 = 1u << q.b1;
    if ((rc = enqueue_hist(ctx, s, d_keys, false, n, q.b2, q.b1, m.ghist))) return rc;
    if ((rc = enqueue_scan(ctx, s, role, 1, n1, false))) return rc;
    Plan p1; p1.B = q.b1; p1.b1 = q.b1; p1.b2 = 0;
    if ((rc = enqueue_plan(ctx, s, role, 1, p1, false))) return rc;
    tup_t* first_out = ctx->out[which];
    if (n) {
        const PPCfg& c1 = pp_first_cfg(q.b1);
        const uint32_t T1 = (uint32_t)(c1.threads * c1.ipt);
        ScatterArgs a;
        memset(&a, 0, sizeof(a));
        a.in_keys = d_keys; a.in_pays = d_pays; a.n = (uint32_t)n; a.out = first_out;
        a.shift = q.b2; a.bits = q.b1; a.cursors = m.cur2; a.cursor_stride = 1;
        a.ntiles = (uint32_t)((n + T1 - 1) / T1);
        c1.fn<<<a.ntiles, c1.threads, pp_smem(c1), s>>>(a);
        LAUNCHED();
    }
    // pass-2 tile list (tiles never straddle a first-pass partition) and the fine counts per tile
    const PPCfg& c2 = pp_push_cfg(ctx, q.b2);
    const uint32_t T2 = (uint32_t)(c2.threads * c2.ipt);
    ScanSeq st;
    st.in = m.off; st.in2 = nullptr; st.out = q.tile_prefix[which]; st.desc = q.tdesc[which]; st.ticket = q.tticket[which];
    st.mode = SCAN_TILES; st.param = T2;
    const uint32_t perm = q.g ? ((q.b1 << 8) | q.g) : 0u;   // destination-interleaved partition order
    st.param2 = perm;
    if ((rc = enqueue_scan_one(ctx, s, st, n1))) return rc;
    tiles3_kernel<<<(n1 + 255) / 256, 256, 0, s>>>(m.off, q.tile_prefix[which], n1, T2, q.b2, m.tiles, perm);
    LAUNCHED();
    if (n) {
        const uint32_t grid = (uint32_t)std::min<uint64_t>(n / T2 + n1 + 2, (uint64_t)ctx->sm_count * 8);
        subhist_tiles_kernel<512><<<grid, 512, 0, s>>>(first_out, m.tiles, q.tile_prefix[which] + n1, q.b2, T2, d_fine_hist);
        LAUNCHED();
    }
    CK(cudaEventRecord(q.ev[which][1], s));
    return GJ_OK;
}

extern "C" int gj_pp_push(gj_ctx* ctx, int which, const uint32_t* d_all_hist, void* const* peer_bases,
                          uint64_t cap_tuples, uint64_t n, void* cuda_stream) {
    if (!ctx || !ctx->pp.active) return fail(GJ_ERR_STATE, "gj_pp_begin first");
    if (which != 0 && which != 1) return fail(GJ_ERR_ARG, "which must be 0 (R) or 1 (S)");
    if (!d_all_hist || !peer_bases) return fail(GJ_ERR_ARG, "NULL argument");
    if (cap_tuples > 0xFFFFFFFFull) return fail(GJ_ERR_ARG, "destination capacity exceeds 2^32 tuples");
    CK(cudaSetDevice(ctx->device));
    gj_ctx::PP& q = ctx->pp;
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    const int role = q.role_of_side[which];
    const RelMeta& m = ctx->meta[role];
    void** hb = reinterpret_cast<void**>(q.h_pin) + (size_t)which * NB_MAX;
    for (uint32_t g = 0; g < q.G; ++g) {
        if (!peer_bases[g] || ((size_t)peer_bases[g] & 15u)) return fail(GJ_ERR_ARG, "destination buffer %u must be non-NULL and 16-byte aligned", g);
        hb[g] = peer_bases[g];
    }
    CK(cudaMemcpyAsync(q.bases[which], hb, q.G * sizeof(void*), cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(q.ev[which][2], s));
    // this GPU's own counts / offsets land in the role's histogram + offset arrays (join planning)
    pp_cursor_kernel<<<q.G, PPC_THREADS, 0, s>>>(d_all_hist, q.G, q.rank, q.B, (uint32_t)cap_tuples, q.cur_fine[which],
                                                 m.ghist, m.off, q.status[which]);
    LAUNCHED();
    if (n) {
        const PPCfg& c2 = pp_push_cfg(ctx, q.b2);
        const uint32_t T2 = (uint32_t)(c2.threads * c2.ipt);
        const uint32_t n1 = 1u << q.b1;
        ScatterArgs b;
        memset(&b, 0, sizeof(b));
        b.in_tup = ctx->out[which]; b.out = nullptr; b.n = (uint32_t)n;
        b.shift = 0; b.bits = q.b2;
        b.cursors = q.cur_fine[which]; b.cursor_stride = 1;
        b.tiles = m.tiles; b.num_tiles = q.tile_prefix[which] + n1;
        b.part_bases = q.bases[which]; b.part_shift = q.B; b.n_dest = q.G; b.abort_flag = q.status[which];
        const uint32_t grid = (uint32_t)(n / T2) + n1 + 2;   // upper bound on tiles
        c2.fn<<<grid, c2.threads, pp_smem(c2), s>>>(b);
        LAUNCHED();
    }
    CK(cudaEventRecord(q.ev[which][3], s));
    return GJ_OK;
}

extern "C" int gj_pp_join(gj_ctx* ctx, const void* d_own_R, const void* d_own_S, uint64_t cap_R, uint64_t cap_S,
                          void* cuda_stream) {
    if (!ctx || !ctx->pp.active) return fail(GJ_ERR_STATE, "gj_pp_begin first");
    if (!d_own_R || !d_own_S) return fail(GJ_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    gj_ctx::PP& q = ctx->pp;
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    CK(cudaStreamWaitEvent(s, ctx->stage_ev[3], 0));
    Plan pl; pl.B = q.B; pl.b1 = q.B; pl.b2 = 0;
    const bool r_builds = q.role_of_side[0] == 0;
    const tup_t* bld = (const tup_t*)(r_builds ? d_own_R : d_own_S);
    const tup_t* prb = (const tup_t*)(r_builds ? d_own_S : d_own_R);
    int rc;
    CK(cudaEventRecord(q.jev[0], s));
    if (q.n_glob[0] && q.n_glob[1]) {
        if ((rc = enqueue_scan(ctx, s, 0, 0, 1u << pl.B, true))) return rc;
        if ((rc = enqueue_plan(ctx, s, 0, 0, pl, true))) return rc;
        if ((rc = enqueue_join(ctx, s, bld, prb, pl, r_builds ? cap_R : cap_S, r_builds ? cap_S : cap_R, false,
                               nullptr, nullptr, 0, nullptr, (int)q.g))) return rc;
    }
    CK(cudaEventRecord(q.jev[1], s));
    CK(cudaMemcpyAsync(ctx->h_result, ctx->result, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    uint32_t* hs = reinterpret_cast<uint32_t*>(q.h_pin + 2 * NB_MAX * sizeof(void*));
    for (int w = 0; w < 2; ++w) CK(cudaMemcpyAsync(hs + 4 * w, q.status[w], 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(ctx->ev[3], s));
    return GJ_OK;
}

extern "C" int gj_pp_finish(gj_ctx* ctx, uint64_t* matches, uint64_t* checksum, uint64_t* n_local_R,
                            uint64_t* n_local_S, float* phase_ms) {
    if (!ctx || !ctx->pp.active) return fail(GJ_ERR_STATE, "gj_pp_begin first");
    gj_ctx::PP& q = ctx->pp;
    CK(cudaEventSynchronize(ctx->ev[3]));
    q.active = false;
    const uint32_t* hs = reinterpret_cast<const uint32_t*>(q.h_pin + 2 * NB_MAX * sizeof(void*));
    if (n_local_R) *n_local_R = hs[1];
    if (n_local_S) *n_local_S = hs[5];
    if (hs[0] || hs[4])
        return fail(GJ_ERR_ARG, "sharded join: a destination GPU would receive more tuples than its buffer holds "
                                "(this GPU: %u R, %u S tuples) -- raise the receive-buffer slack", hs[1], hs[5]);
    if (matches) *matches = ctx->h_result[0];
    if (checksum) *checksum = ctx->h_result[1];
    if (phase_ms) {   // [local R, push R, local S, push S, join]; the events of a relation sit on its stream
        for (int w = 0; w < 2; ++w) {
            CK(cudaEventSynchronize(q.ev[w][3]));
            CK(cudaEventElapsedTime(&phase_ms[2 * w], q.ev[w][0], q.ev[w][1]));
            CK(cudaEventElapsedTime(&phase_ms[2 * w + 1], q.ev[w][2], q.ev[w][3]));
        }
        CK(cudaEventElapsedTime(&phase_ms[4], q.jev[0], q.jev[1]));
    }
    return GJ_OK;
}

extern "C" int gj_pp_plan(gj_ctx* ctx, uint32_t* pass1_bits, uint32_t* pass2_bits) {
    if (!ctx || !ctx->pp.block) return fail(GJ_ERR_STATE, "gj_pp_begin first");
    if (pass1_bits) *pass1_bits = ctx->pp.b1;
    if (pass2_bits) *pass2_bits = ctx->pp.b2;
    return GJ_OK;
}

// ------------------------------------------------------------------------------------------
// Sharded "partition, copy, partition" pipeline (multi-GPU; kernels.cuh section 3d).  The relation
// that builds (the globally smaller one) goes first.  Per relation:
//   gj_pcp_hist : this shard's first-pass histogram on [gpu bits | top local bits] (2^b1 counters)
//   [caller: all-gather of those histograms]
//   [or gj_pcp_hist_exchange: the same exchange through the peers' control blocks, no collective]
//   gj_pcp_part : layout (pcp_layout_kernel) + first radix pass: remote chunks into the stage buffer,
//                 this GPU's own chunks straight into its receive buffer (+ their fine counts)
//   gj_pcp_copy : n_stages x (TMA bulk-copy kernel over a group of first-pass partitions, whose histogram
//                 warps count every piece for the receiver's last pass; then those counts and a flag into
//                 every peer's control block): every remote chunk to its slot in the destination's buffer
//   gj_pcp_recv : per stage: wait for every source's flag, sum the delivered counts, scan + plan + LAST
//                 radix pass over the first-pass partitions of the stage; for the probing relation also
//                 unit planning + join of those partitions -- while later stages still cross NVLink
//   gj_pcp_finish
// Buffers: the first relation stages in ctx->scratch and ends in ctx->out[first]; the second stages
// in ctx->out[second] and ends in ctx->scratch (free again once the first relation's copy is done,
// which every rank's flags for the second relation imply).
// ------------------------------------------------------------------------------------------
static int ensure_pcp(gj_ctx* ctx) {
    gj_ctx::PCP& q = ctx->pcp;
    if (q.block) return GJ_OK;
    const size_t per = (size_t)(PCP_MAX_CHUNKS + 4) * sizeof(uint32_t);
    size_t b = 0;
    size_t o[2][7], o_db[2], o_bs[2], o_fl[2], o_fine[2];
    for (int r = 0; r < 2; ++r) {
        for (int k = 0; k < 7; ++k) { o[r][k] = b; b += per; }
        o_db[r] = b; b += PCP_MAX_CHUNKS * sizeof(tup_t*);
        o_bs[r] = b; b += NB_MAX * sizeof(tup_t*);
        o_fl[r] = b; b += NB_MAX * sizeof(uint32_t*);
        o_fine[r] = b; b += sizeof(uint32_t) << PP_MAX_BITS;
    }
    if (cudaMalloc(&q.block, b) != cudaSuccess) { cudaGetLastError(); return fail(GJ_ERR_NOMEM, "pcp metadata"); }
    for (int r = 0; r < 2; ++r) {
        q.tab[r].cur = reinterpret_cast<uint32_t*>(q.block + o[r][0]);
        q.tab[r].src_start = reinterpret_cast<uint32_t*>(q.block + o[r][1]);
        q.tab[r].dst_start = reinterpret_cast<uint32_t*>(q.block + o[r][2]);
        q.tab[r].cnt = reinterpret_cast<uint32_t*>(q.block + o[r][3]);
        q.tab[r].piece_prefix = reinterpret_cast<uint32_t*>(q.block + o[r][4]);
        q.tab[r].status = reinterpret_cast<uint32_t*>(q.block + o[r][5]);
        q.tab[r].recv_off = reinterpret_cast<uint32_t*>(q.block + o[r][6]);
        q.tab[r].dig_base = reinterpret_cast<tup_t**>(q.block + o_db[r]);
        q.bases[r] = reinterpret_cast<tup_t**>(q.block + o_bs[r]);
        q.ctrl_ptrs[r] = reinterpret_cast<unsigned char**>(q.block + o_fl[r]);
        q.fine[r] = reinterpret_cast<uint32_t*>(q.block + o_fine[r]);
    }
    CK(cudaHostAlloc(&q.h_pin, 4 * NB_MAX * sizeof(void*) + 64, cudaHostAllocDefault));
    for (auto& r : q.ev) for (auto& e : r) CK(cudaEventCreate(&e));
    for (auto& e : q.jev) CK(cudaEventCreate(&e));
    return GJ_OK;
}

static const PPCfg& pcp_last_cfg(uint32_t bits) { return kPcpLast[bits <= 7u ? 0u : bits - 7u]; }
static int pcp_last_ctas_per_sm(uint32_t bits) { return bits <= 7u ? 4 : (bits <= 9u ? 2 : 1); }
// SMs left to the kernels that run next to the copy kernel: with "pcp_copy_ctas" = k <= half the SMs (deep
// ring), k SMs are filled by copy CTAs (196 KB of shared memory each) and host nothing else
static bool pcp_deep_ring(const gj_ctx* ctx) { return ctx->opt_pcp_copy_ctas > 0 && ctx->opt_pcp_copy_ctas <= ctx->sm_count / 2; }
static int pcp_free_sms(const gj_ctx* ctx) { return pcp_deep_ring(ctx) ? ctx->sm_count - (int)ctx->opt_pcp_copy_ctas : ctx->sm_count; }
static tup_t* pcp_stage_buf(gj_ctx* ctx, int which) { return which == ctx->pcp.first ? ctx->scratch : ctx->out[which]; }
static tup_t* pcp_final_buf(gj_ctx* ctx, int which) { return which == ctx->pcp.first ? ctx->out[which] : ctx->scratch; }
static uint64_t pcp_stage_cap(const gj_ctx* ctx, int which) {
    return which == ctx->pcp.first ? std::max(ctx->maxR, ctx->maxS) : (which ? ctx->maxS : ctx->maxR);
}
static uint64_t pcp_final_cap(const gj_ctx* ctx, int which) {
    return which == ctx->pcp.first ? (which ? ctx->maxS : ctx->maxR) : std::max(ctx->maxR, ctx->maxS);
}

extern "C" int gj_pcp_begin(gj_ctx* ctx, uint64_t n_R_global, uint64_t n_S_global, uint32_t n_gpus, uint32_t rank,
                            uint32_t local_bits, void* cuda_stream) {
    if (!ctx) return fail(GJ_ERR_ARG, "ctx is NULL");
    if (n_gpus < 2 || n_gpus > (uint32_t)NB_MAX || (n_gpus & (n_gpus - 1))) return fail(GJ_ERR_ARG, "n_gpus must be a power of two in [2, %d]", NB_MAX);
    if (rank >= n_gpus) return fail(GJ_ERR_ARG, "rank %u out of range", rank);
    uint32_t g = 0;
    while ((1u << g) < n_gpus) ++g;
    if (local_bits < 1 || local_bits > (uint32_t)MAX_RADIX_BITS) return fail(GJ_ERR_ARG, "local_bits must be in [1, %d]", MAX_RADIX_BITS);
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_pcp(ctx))) return rc;
    gj_ctx::PCP& q = ctx->pcp;
    // bl = local bits done at the source (with the GPU bits: b1 = g + bl <= 10), b2 = B - bl at the receiver
    const uint32_t B = local_bits;
    uint32_t bl = ctx->opt_pass1_bits ? ((uint32_t)ctx->opt_pass1_bits > g ? (uint32_t)ctx->opt_pass1_bits - g : 0u)
                                      : (B > g ? (B - g + 1) / 2 : 0u);
    bl = std::min(bl, std::min(PP_MAX_PASS_BITS - g, B - 1));
    bl = std::min(bl, (uint32_t)MAX_PASS_BITS);   // the receiver plans its last pass over <= 256 first-pass partitions
    if (B - bl > PP_MAX_PASS_BITS) bl = B - PP_MAX_PASS_BITS;
    if (g + bl > PP_MAX_PASS_BITS || bl > (uint32_t)MAX_PASS_BITS) return fail(GJ_ERR_ARG, "%u GPU bits + %u local bits do not fit two passes of <= %u", g, B, PP_MAX_PASS_BITS);
    q.G = n_gpus; q.rank = rank; q.g = g; q.B = B; q.bl = bl; q.b1 = g + bl; q.b2 = B - bl;
    const bool swap = n_R_global > n_S_global;
    q.role_of_side[0] = swap ? 1 : 0;
    q.role_of_side[1] = swap ? 0 : 1;
    q.first = swap ? 1 : 0;
    q.n_glob[0] = n_R_global; q.n_glob[1] = n_S_global;
    q.recv_done[0] = q.recv_done[1] = false;
    q.parted[0] = q.parted[1] = false;
    q.active = true;
    ++q.epoch;
    ctx->launches = 0;
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    CK(cudaMemsetAsync(ctx->zero_common, 0, ctx->zero_common_bytes, s));
    CK(cudaEventRecord(ctx->stage_ev[3], s));
    return GJ_OK;
}

extern "C" int gj_pcp_plan(gj_ctx* ctx, uint32_t plan_bits[3]) {
    if (!ctx || !ctx->pcp.block || !plan_bits) return fail(GJ_ERR_STATE, "gj_pcp_begin first");
    plan_bits[0] = ctx->pcp.g; plan_bits[1] = ctx->pcp.bl; plan_bits[2] = ctx->pcp.b2;
    return GJ_OK;
}

extern "C" int gj_pcp_hist(gj_ctx* ctx, int which, const int32_t* d_keys, uint64_t n, uint32_t* d_coarse_hist,
                           void* cuda_stream) {
    if (!ctx || !ctx->pcp.active) return fail(GJ_ERR_STATE, "gj_pcp_begin first");
    if (which != 0 && which != 1) return fail(GJ_ERR_ARG, "which must be 0 (R) or 1 (S)");
    if (!d_coarse_hist || (n && !d_keys)) return fail(GJ_ERR_ARG, "NULL argument");
    gj_ctx::PCP& q = ctx->pcp;
    // every staged chunk has one spare slot (16-byte phase matching)
    if (n + (1ull << q.b1) > pcp_stage_cap(ctx, which) + 16) return fail(GJ_ERR_ARG, "n + %u spare slots exceed the context capacity", 1u << q.b1);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    q.n_loc[which] = n;
    CK(cudaMemsetAsync(d_coarse_hist, 0, sizeof(uint32_t) << q.b1, s));
    CK(cudaMemsetAsync(q.tab[which].status, 0, 4 * sizeof(uint32_t), s));
    CK(cudaMemsetAsync(q.fine[which], 0, sizeof(uint32_t) << (q.g + q.B), s));
    return enqueue_hist(ctx, s, d_keys, false, n, q.B - q.bl, q.b1, d_coarse_hist);
}

extern "C" int gj_pcp_part(gj_ctx* ctx, int which, const int32_t* d_keys, const int32_t* d_pays,
                           const uint32_t* d_all_hist, void* d_own, uint64_t cap_tuples, void* cuda_stream) {
    if (!ctx || !ctx->pcp.active) return fail(GJ_ERR_STATE, "gj_pcp_begin first");
    if (which != 0 && which != 1) return fail(GJ_ERR_ARG, "which must be 0 (R) or 1 (S)");
    gj_ctx::PCP& q = ctx->pcp;
    const uint64_t n = q.n_loc[which];
    if (!d_all_hist || (n && (!d_keys || !d_pays))) return fail(GJ_ERR_ARG, "NULL argument");
    if (!d_own || ((size_t)d_own & 15u)) return fail(GJ_ERR_ARG, "receive buffer must be non-NULL and 16-byte aligned");
    if (cap_tuples > 0xFFFFFFFFull) return fail(GJ_ERR_ARG, "destination capacity exceeds 2^32 tuples");
    if (which != q.first && !q.parted[q.first]) return fail(GJ_ERR_STATE, "partition the building (smaller) relation first");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    const uint32_t perm = (q.b1 << 8) | q.g;
    CK(cudaEventRecord(q.ev[which][0], s));
    pcp_layout_kernel<<<1, PCP_MAX_CHUNKS, 0, s>>>(d_all_hist, q.G, q.rank, q.b1, q.bl, (uint32_t)cap_tuples, perm,
                                                  pcp_stage_buf(ctx, which), (tup_t*)d_own, q.tab[which]);
    LAUNCHED();
    if (n) {
        const PPCfg& c1 = pp_first_cfg(q.b1);
        const uint32_t T1 = (uint32_t)(c1.threads * c1.ipt);
        ScatterArgs a;
        memset(&a, 0, sizeof(a));
        a.in_keys = d_keys; a.in_pays = d_pays; a.n = (uint32_t)n; a.out = nullptr;
        a.dst_bases = q.tab[which].dig_base;        // per chunk: the stage buffer, or this GPU's receive buffer
        a.abort_flag = q.tab[which].status;         // a destination would overflow: nothing is written
        a.shift = q.B - q.bl; a.bits = q.b1; a.cursors = q.tab[which].cur; a.cursor_stride = 1;
        a.ntiles = (uint32_t)((n + T1 - 1) / T1);
        c1.fn<<<a.ntiles, c1.threads, pp_smem(c1), s>>>(a);
        LAUNCHED();
        // fine counts of the chunks this GPU keeps (they never pass through the copy kernel)
        const uint32_t slices = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(16, (n >> (q.b1 > 0 ? q.b1 : 0)) / PCP_SELF_SLICE + 1));
        pcp_self_hist_kernel<<<dim3(slices, 1u << q.bl), 256, 0, s>>>((const tup_t*)d_own, q.tab[which], q.rank, q.bl, q.b2, q.fine[which]);
        LAUNCHED();
    }
    CK(cudaEventRecord(q.ev[which][1], s));
    q.parted[which] = true;
    return GJ_OK;
}

extern "C" uint64_t gj_pcp_ctrl_bytes(uint32_t n_gpus) { return pcp_ctrl_bytes(n_gpus); }

// Exchange of the coarse histograms through the peers' control blocks instead of a collective: push this shard's
// histogram (gj_pcp_hist's output) into every GPU's block, wait (bounded) for every source's, compact them into
// d_all_hist ([n_gpus][2^(g + bl)], the input of gj_pcp_part).  Stream-ordered; replaces the caller's all-gather.
extern "C" int gj_pcp_hist_exchange(gj_ctx* ctx, int which, const uint32_t* d_coarse_hist, void* const* peer_ctrl,
                                    const void* d_ctrl, uint32_t* d_all_hist, void* cuda_stream) {
    if (!ctx || !ctx->pcp.active) return fail(GJ_ERR_STATE, "gj_pcp_begin first");
    if (which != 0 && which != 1) return fail(GJ_ERR_ARG, "which must be 0 (R) or 1 (S)");
    if (!d_coarse_hist || !peer_ctrl || !d_ctrl || !d_all_hist) return fail(GJ_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    gj_ctx::PCP& q = ctx->pcp;
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    void** hb = reinterpret_cast<void**>(q.h_pin) + (size_t)which * 2 * NB_MAX;
    for (uint32_t g = 0; g < q.G; ++g) {
        if (!peer_ctrl[g] || ((size_t)peer_ctrl[g] & 15u)) return fail(GJ_ERR_ARG, "control block %u must be non-NULL and 16-byte aligned", g);
        hb[NB_MAX + g] = peer_ctrl[g];
    }
    CK(cudaMemcpyAsync(q.ctrl_ptrs[which], hb + NB_MAX, q.G * sizeof(void*), cudaMemcpyHostToDevice, s));
    const uint32_t n1 = 1u << q.b1;
    pcp_hist_push_kernel<<<q.G, 256, 0, s>>>(d_coarse_hist, q.ctrl_ptrs[which], q.G, q.rank, (uint32_t)which, n1, q.epoch);
    LAUNCHED();
    pcp_hist_gather_kernel<<<q.G, 256, 0, s>>>((const unsigned char*)d_ctrl, q.G, (uint32_t)which, n1, q.epoch,
                                               (unsigned long long)ctx->opt_pcp_timeout_ms * 1000000ull, d_all_hist, q.tab[which].status);
    LAUNCHED();
    return GJ_OK;
}

extern "C" int gj_pcp_copy(gj_ctx* ctx, int which, void* const* peer_bases, void* const* peer_ctrl, uint32_t n_stages,
                           void* cuda_stream) {
    if (!ctx || !ctx->pcp.active) return fail(GJ_ERR_STATE, "gj_pcp_begin first");
    if (which != 0 && which != 1) return fail(GJ_ERR_ARG, "which must be 0 (R) or 1 (S)");
    if (!peer_bases || !peer_ctrl) return fail(GJ_ERR_ARG, "NULL argument");
    if (n_stages < 1 || n_stages > (uint32_t)PCP_MAX_STAGES) return fail(GJ_ERR_ARG, "n_stages must be in [1, %d]", PCP_MAX_STAGES);
    gj_ctx::PCP& q = ctx->pcp;
    if (!q.parted[which]) return fail(GJ_ERR_STATE, "gj_pcp_part first");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    void** hb = reinterpret_cast<void**>(q.h_pin) + (size_t)which * 2 * NB_MAX;
    for (uint32_t g = 0; g < q.G; ++g) {
        if (!peer_bases[g] || ((size_t)peer_bases[g] & 15u)) return fail(GJ_ERR_ARG, "destination buffer %u must be non-NULL and 16-byte aligned", g);
        if (!peer_ctrl[g] || ((size_t)peer_ctrl[g] & 15u)) return fail(GJ_ERR_ARG, "control block %u must be non-NULL and 16-byte aligned", g);
        hb[g] = peer_bases[g];
        hb[NB_MAX + g] = peer_ctrl[g];
    }
    CK(cudaMemcpyAsync(q.bases[which], hb, q.G * sizeof(void*), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(q.ctrl_ptrs[which], hb + NB_MAX, q.G * sizeof(void*), cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(q.ev[which][2], s));
    const uint32_t nj = 1u << q.bl, K = std::min(n_stages, nj);
    q.stages[which] = K;
    PcpCopyArgs a;
    a.stage = pcp_stage_buf(ctx, which); a.peer_bases = q.bases[which]; a.t = q.tab[which];
    a.b1 = q.b1; a.bl = q.bl; a.perm = (q.b1 << 8) | q.g;
    a.b2 = q.b2; a.fine = q.fine[which];
    const uint64_t pieces_max = q.n_loc[which] / PCP_PIECE + (1ull << q.b1) + 1;
    // few CTAs with the deep ring (they fill their SMs: 196 KB of shared memory each), or one shallow-ring CTA per SM
    const bool deep = pcp_deep_ring(ctx);
    uint32_t grid = ctx->opt_pcp_copy_ctas ? (uint32_t)ctx->opt_pcp_copy_ctas : (uint32_t)ctx->sm_count;
    grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(grid, pieces_max));
    for (uint32_t k = 0; k < K; ++k) {
        // copy positions are (first-pass partition j, destination d) with d fastest: stage k = partitions [j_lo, j_hi)
        const uint32_t j_lo = (uint32_t)((uint64_t)k * nj / K), j_hi = (uint32_t)((uint64_t)(k + 1) * nj / K);
        a.pos_lo = j_lo << q.g;
        a.pos_hi = j_hi << q.g;
        if (q.n_loc[which]) {
            if (deep) pcp_copy_kernel<PCP_NS_DEEP><<<grid, PCP_COPY_THREADS, pcp_copy_smem(PCP_NS_DEEP), s>>>(a);
            else pcp_copy_kernel<PCP_NS><<<grid, PCP_COPY_THREADS, pcp_copy_smem(), s>>>(a);
            LAUNCHED();
        }
        // the stage's fine counts + the flag, into every GPU's control block (this one's included)
        pcp_push_kernel<<<q.G, 256, 0, s>>>(q.fine[which], q.ctrl_ptrs[which], q.G, q.rank, (uint32_t)which, k, q.epoch, q.B,
                                            j_lo << q.b2, j_hi << q.b2, q.tab[which].status);
        LAUNCHED();
    }
    CK(cudaEventRecord(q.ev[which][3], s));
    return GJ_OK;
}

extern "C" int gj_pcp_recv(gj_ctx* ctx, int which, const void* d_own, const void* d_ctrl, uint64_t cap_tuples,
                           void* d_result_out, void* cuda_stream) {
    if (!ctx || !ctx->pcp.active) return fail(GJ_ERR_STATE, "gj_pcp_begin first");
    if (which != 0 && which != 1) return fail(GJ_ERR_ARG, "which must be 0 (R) or 1 (S)");
    if (!d_own || ((size_t)d_own & 15u)) return fail(GJ_ERR_ARG, "receive buffer must be non-NULL and 16-byte aligned");
    if (!d_ctrl) return fail(GJ_ERR_ARG, "control block is NULL");
    gj_ctx::PCP& q = ctx->pcp;
    if (cap_tuples > pcp_final_cap(ctx, which)) return fail(GJ_ERR_ARG, "receive capacity exceeds the context capacity");
    if (!q.stages[which] || !q.parted[which]) return fail(GJ_ERR_STATE, "gj_pcp_part and gj_pcp_copy first");
    const bool joins = (which != q.first);          // the probing relation: join every stage as it lands
    if (joins && !q.recv_done[q.first]) return fail(GJ_ERR_STATE, "receive the building (smaller) relation first");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    const int role = q.role_of_side[which];
    const RelMeta& m = ctx->meta[role];
    const PcpTables& tb = q.tab[which];
    const uint32_t nb = 1u << q.B, nj = 1u << q.bl, K = q.stages[which];
    const int free_sms = pcp_free_sms(ctx);
    int rc;
    CK(cudaStreamWaitEvent(s, q.ev[which][1], 0));   // this GPU's own chunks are in place (source pass done)
    CK(cudaEventRecord(q.ev[which][4], s));
    CK(cudaMemsetAsync(ctx->zero_role[role], 0, ctx->zero_role_bytes, s));
    const bool empty = !q.n_glob[which];
    const PPCfg& c2 = pcp_last_cfg(q.b2);
    const uint32_t T2 = (uint32_t)(c2.threads * c2.ipt);
    const size_t res_off = (size_t)(reinterpret_cast<unsigned char*>(ctx->result) - ctx->zero_common);
    const size_t desc_off = (size_t)(reinterpret_cast<unsigned char*>(m.desc) - ctx->zero_role[role]);
    const bool r_builds = q.role_of_side[0] == 0;
    bool waited_build = false;
    for (uint32_t k = 0; k < K && !empty; ++k) {
        const uint32_t j_lo = (uint32_t)((uint64_t)k * nj / K), j_hi = (uint32_t)((uint64_t)(k + 1) * nj / K);
        pcp_wait_kernel<<<1, std::max(32u, q.G), 0, s>>>((const uint32_t*)d_ctrl, q.G, q.rank, (uint32_t)which * PCP_MAX_STAGES + k,
                                                         q.epoch, (unsigned long long)ctx->opt_pcp_timeout_ms * 1000000ull, tb.status);
        LAUNCHED();
        if (k + 1 == K) CK(cudaEventRecord(q.jev[0], s));      // the last byte of this relation has landed
        // fine histogram of what arrived for [j_lo, j_hi) = the sum of the counts the sources delivered with
        // their flags (taken by their copy kernels: the receiver does not read the data to count it).  The
        // counters of earlier stages stay, so the full-range scan below yields the same offsets for them
        // again and the new ones behind them.
        if (k) CK(cudaMemsetAsync(ctx->zero_role[role] + desc_off, 0, ctx->zero_role_bytes - desc_off, s));
        pcp_sum_hist_kernel<<<std::max(1u, std::min(64u, ((j_hi - j_lo) << q.b2) / 256u)), 256, 0, s>>>(
            (const unsigned char*)d_ctrl, q.G, (uint32_t)which, q.B, j_lo << q.b2, j_hi << q.b2, m.ghist);
        LAUNCHED();
        {
            ScanSeq so;
            so.in = m.ghist; so.in2 = nullptr; so.out = m.off; so.desc = m.desc; so.ticket = m.ticket;
            so.mode = SCAN_PLAIN; so.param = 0; so.param2 = 0;
            if ((rc = enqueue_scan_one(ctx, s, so, nb))) return rc;
        }
        {   // cursors + last-pass tile list of the stage; the receive buffer IS the first-pass output (bl bits done)
            PlanArgs a;
            memset(&a, 0, sizeof(a));
            a.rel[0].off = m.off; a.rel[0].cur1 = m.cur1; a.rel[0].cur2 = m.cur2; a.rel[0].tiles = m.tiles; a.rel[0].num_tiles = m.num_tiles;
            a.rel[1] = a.rel[0];
            a.nrel = 1; a.with_units = 0; a.b1 = q.bl; a.b2 = q.b2; a.tile = T2; a.unit = unit_tuples(ctx);
            a.unit_base = ctx->unit_base; a.units = ctx->units;
            a.j_lo = j_lo; a.j_hi = j_hi;
            const int grid = (int)std::max<uint32_t>(1, std::min<uint32_t>(128, (nb / K + PLAN_THREADS - 1) / PLAN_THREADS + 32));
            plan_kernel<<<grid, PLAN_THREADS, 0, s>>>(a);
            LAUNCHED();
        }
        {
            ScatterArgs b;
            memset(&b, 0, sizeof(b));
            b.in_tup = (const tup_t*)d_own; b.out = pcp_final_buf(ctx, which); b.n = (uint32_t)cap_tuples;
            b.shift = 0; b.bits = q.b2;
            b.cursors = m.cur2; b.cursor_stride = 1; b.tiles = m.tiles; b.num_tiles = m.num_tiles;
            const uint64_t bound = cap_tuples / T2 + (j_hi - j_lo) + 2;   // upper bound on the stage's tiles (skew: all of them)
            const uint32_t grid2 = (uint32_t)std::min<uint64_t>(bound, (uint64_t)free_sms * pcp_last_ctas_per_sm(q.b2) * 2);
            c2.fn<<<grid2, c2.threads, pp_smem(c2), s>>>(b);
            LAUNCHED();
        }
        if (joins && q.n_glob[0] && q.n_glob[1]) {
            if (!waited_build) {
                CK(cudaStreamWaitEvent(s, q.ev[q.first][5], 0));    // the build side is received and partitioned
                CK(cudaStreamWaitEvent(s, ctx->stage_ev[3], 0));    // the accumulators are zeroed
                waited_build = true;
            }
            CK(cudaMemsetAsync(ctx->zero_common, 0, res_off, s));   // re-arm the unit scan; the accumulators keep counting
            ScanSeq su;
            su.in = ctx->meta[0].ghist; su.in2 = ctx->meta[1].ghist; su.out = ctx->unit_base; su.desc = ctx->unit_desc; su.ticket = ctx->unit_ticket;
            su.mode = SCAN_UNITS; su.param = unit_tuples(ctx); su.param2 = 0;
            su.lo = j_lo << q.b2; su.hi = j_hi << q.b2;
            if ((rc = enqueue_scan_one(ctx, s, su, nb))) return rc;
            PlanArgs a;
            memset(&a, 0, sizeof(a));
            a.rel[0].off = ctx->meta[0].off; a.rel[1].off = ctx->meta[1].off;
            a.nrel = 0; a.with_units = 1; a.b1 = q.bl; a.b2 = q.b2; a.tile = T2; a.unit = unit_tuples(ctx);
            a.unit_base = ctx->unit_base; a.units = ctx->units;
            a.j_lo = j_lo; a.j_hi = j_hi;
            plan_kernel<<<std::max(1u, std::min(128u, (nb / K + PLAN_THREADS - 1) / PLAN_THREADS)), PLAN_THREADS, 0, s>>>(a);
            LAUNCHED();
            Plan pl; pl.B = q.B; pl.b1 = q.B; pl.b2 = 0;
            const int bw = r_builds ? 0 : 1;
            const int64_t keep_grid = ctx->opt_join_grid;
            if (!keep_grid && free_sms < ctx->sm_count) ctx->opt_join_grid = free_sms;
            rc = enqueue_join(ctx, s, pcp_final_buf(ctx, bw), pcp_final_buf(ctx, 1 - bw), pl, 0, cap_tuples / K + 1, false,
                              nullptr, nullptr, 0, nullptr, (int)q.g);
            ctx->opt_join_grid = keep_grid;
            if (rc) return rc;
        }
    }
    CK(cudaEventRecord(q.ev[which][5], s));
    q.recv_done[which] = true;
    if (joins) {
        CK(cudaEventRecord(q.jev[1], s));
        // the local aggregate {matches, checksum} also into a device buffer of the caller's (input of its all-reduce)
        if (d_result_out) CK(cudaMemcpyAsync(d_result_out, ctx->result, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(ctx->h_result, ctx->result, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        uint32_t* hs = reinterpret_cast<uint32_t*>(q.h_pin + 4 * NB_MAX * sizeof(void*));
        for (int w = 0; w < 2; ++w) CK(cudaMemcpyAsync(hs + 4 * w, q.tab[w].status, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(ctx->ev[3], s));
    }
    return GJ_OK;
}

extern "C" int gj_pcp_finish(gj_ctx* ctx, uint64_t* matches, uint64_t* checksum, uint64_t* n_local_R,
                             uint64_t* n_local_S, float* phase_ms, uint32_t* plan_bits) {
    if (!ctx || !ctx->pcp.active) return fail(GJ_ERR_STATE, "gj_pcp_begin first");
    gj_ctx::PCP& q = ctx->pcp;
    if (!q.recv_done[0] || !q.recv_done[1]) return fail(GJ_ERR_STATE, "gj_pcp_recv of both relations first");
    CK(cudaEventSynchronize(ctx->ev[3]));
    q.active = false;
    q.stages[0] = q.stages[1] = 0;
    const uint32_t* hs = reinterpret_cast<const uint32_t*>(q.h_pin + 4 * NB_MAX * sizeof(void*));
    if (n_local_R) *n_local_R = hs[1];
    if (n_local_S) *n_local_S = hs[5];
    if (plan_bits) { plan_bits[0] = q.g; plan_bits[1] = q.bl; plan_bits[2] = q.b2; }
    if (hs[3] || hs[7])
        return fail(GJ_ERR_STATE, "sharded join: timed out waiting for a peer's data (option \"pcp_timeout_ms\" = %lld)", (long long)ctx->opt_pcp_timeout_ms);
    if (hs[0] || hs[4])
        return fail(GJ_ERR_ARG, "sharded join: a destination GPU would receive more tuples than its buffer holds "
                                "(this GPU: %u R, %u S tuples) -- raise the receive-buffer slack", hs[1], hs[5]);
    if (matches) *matches = ctx->h_result[0];
    if (checksum) *checksum = ctx->h_result[1];
    if (phase_ms) {   // [part R, copy R, recv R, part S, copy S, recv S, tail = last byte of the probe side landed -> last join done]
        for (int w = 0; w < 2; ++w) {
            CK(cudaEventSynchronize(q.ev[w][5]));
            CK(cudaEventSynchronize(q.ev[w][3]));
            for (int k = 0; k < 3; ++k) CK(cudaEventElapsedTime(&phase_ms[3 * w + k], q.ev[w][2 * k], q.ev[w][2 * k + 1]));
        }
        phase_ms[6] = 0.f;
        if (q.n_glob[1 - q.first]) CK(cudaEventElapsedTime(&phase_ms[6], q.jev[0], q.jev[1]));
    }
    return GJ_OK;
}

extern "C" int gj_ipc_export(void* d_ptr, char handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    if (!d_ptr || !handle) return fail(GJ_ERR_ARG, "NULL argument");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle, &h, 64);
    return GJ_OK;
}
extern "C" int gj_ipc_open(const char handle[64], void** d_ptr) {
    if (!d_ptr || !handle) return fail(GJ_ERR_ARG, "NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GJ_OK;
}
extern "C" int gj_ipc_close(void* d_ptr) { CK(cudaIpcCloseMemHandle(d_ptr)); return GJ_OK; }

// ------------------------------------------------------------------------------------------
// synthetic data, memory helpers
// ------------------------------------------------------------------------------------------
extern "C" int gj_generate_unique(gj_ctx* ctx, int32_t* d_keys, int32_t* d_pays, uint64_t row_begin,
                                  uint64_t n_rows, uint64_t n_total, uint32_t seed, uint32_t pay_seed) {
    if (!ctx) return fail(GJ_ERR_ARG, "ctx is NULL");
    if (n_total == 0 || n_total > (1ull << 32) || row_begin + n_rows > n_total) return fail(GJ_ERR_ARG, "rows out of range");
    if (n_rows && (!d_keys || !d_pays)) return fail(GJ_ERR_ARG, "NULL output column");
    if (!n_rows) return GJ_OK;
    CK(cudaSetDevice(ctx->device));
    const int grid = (int)std::min<uint64_t>((n_rows + 255) / 256, (uint64_t)ctx->sm_count * 16);
    generate_unique_kernel<<<grid, 256, 0, ctx->stream>>>(d_keys, d_pays, row_begin, n_rows, n_total, seed, pay_seed);
    LAUNCHED();
    CK(cudaStreamSynchronize(ctx->stream));
    return GJ_OK;
}

extern "C" uint32_t gj_bijection(uint64_t row, uint64_t n_total, uint32_t seed) { return bijection(row, n_total, seed); }
extern "C" int32_t gj_payload_of_key(uint32_t key, uint32_t pay_seed) { return payload_of_key(key, pay_seed); }

// One process driving several GPUs (tests, profiling with ncu): let kernels on `device` dereference `peer`'s memory.
extern "C" int gj_enable_peer_access(int device, int peer) {
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, device, peer));
    if (!can) return fail(GJ_ERR_CUDA, "device %d cannot access device %d", device, peer);
    CK(cudaSetDevice(device));
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(GJ_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", device, peer, cudaGetErrorString(e));
    cudaGetLastError();
    return GJ_OK;
}

extern "C" int gj_device_count(int* n) {
    if (!n) return fail(GJ_ERR_ARG, "NULL argument");
    CK(cudaGetDeviceCount(n));
    return GJ_OK;
}
extern "C" int gj_malloc_device(void** p, uint64_t bytes) {
    if (!p) return fail(GJ_ERR_ARG, "NULL argument");
    if (cudaMalloc(p, std::max<uint64_t>(bytes, 1)) != cudaSuccess) { cudaGetLastError(); return fail(GJ_ERR_NOMEM, "cudaMalloc(%llu) failed", (unsigned long long)bytes); }
    return GJ_OK;
}
extern "C" int gj_free_device(void* p) { CK(cudaFree(p)); return GJ_OK; }
extern "C" int gj_malloc_pinned(void** p, uint64_t bytes) {
    if (!p) return fail(GJ_ERR_ARG, "NULL argument");
    if (cudaHostAlloc(p, std::max<uint64_t>(bytes, 1), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return fail(GJ_ERR_NOMEM, "cudaHostAlloc(%llu) failed", (unsigned long long)bytes); }
    return GJ_OK;
}
extern "C" int gj_free_pinned(void* p) { CK(cudaFreeHost(p)); return GJ_OK; }
extern "C" int gj_memcpy_h2d(void* d, const void* h, uint64_t bytes) { CK(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice)); return GJ_OK; }
extern "C" int gj_memcpy_d2h(void* h, const void* d, uint64_t bytes) { CK(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost)); return GJ_OK; }
extern "C" int gj_memset_device(void* d, int value, uint64_t bytes) { CK(cudaMemset(d, value, bytes)); return GJ_OK; }
extern "C" int gj_device_synchronize(void) { CK(cudaDeviceSynchronize()); return GJ_OK; }

extern "C" int gj_flush_l2(gj_ctx* ctx) {
    if (!ctx) return fail(GJ_ERR_ARG, "ctx is NULL");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->flush_buf) {
        ctx->flush_bytes = 256ull << 20;   // 2x the 126 MB L2
        if (cudaMalloc(&ctx->flush_buf, ctx->flush_bytes) != cudaSuccess) { cudaGetLastError(); return fail(GJ_ERR_NOMEM, "flush buffer"); }
    }
    static uint32_t v = 0;
    flush_kernel<<<ctx->sm_count * 4, 512, 0, ctx->stream>>>((uint4*)ctx->flush_buf, ctx->flush_bytes / 16, ++v);
    CK(cudaGetLastError());
    return GJ_OK;
}
