/*
 * oracle/oracle_join.c -- TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * CPU restatement of what the reference (psiul/ICDE2019-GPU-Join) computes on its
 * radix hash-join hot path, used ONLY as the parity checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * The product (libgpujoin.so) never links, loads or calls anything in here.
 *
 * What is restated, and where it comes from (all file:line into /root/reference/src):
 *   - ETHZ-style generators ............ generator_ETHZ.cu:115-122 (random_gen), :127-149
 *     (random_unique_gen), :162-187 (create_relation_fk_from_pk), :194-212 (knuth_shuffle,
 *     knuth_shuffle48), :236-258 (gen_alphabet), :265-294 (gen_zipf_lut), :299-348 (gen_zipf).
 *     The reference seeds random_unique_gen from time(NULL) (:133-135); here the seed is an
 *     explicit argument, everything else (libc rand()/nrand48() streams, order of draws)
 *     is identical, so for equal seeds the byte streams are identical.
 *   - join semantics ................... join-primitives.cu:885-1095 (join_partitioned_aggregate):
 *     inner equi-join on the full 32-bit key (13 partition bits + 10 hash bits + remnant
 *     compare cover all 32 bits, :1030-1032,:1066-1072), all pairs (N:M), aggregate
 *     SUM(Pr*Ps).  The reference accumulates in int32 (:914,:1092) and prints "%d results"
 *     (hash_join_clustered_probe.cu:984-986); here the sum is kept mod 2^64 -- its low 32
 *     bits ARE the reference's printed value (ring homomorphism Z/2^64 -> Z/2^32).
 *   - radix partition semantics ........ join-primitives.cu:58-283,:338-535: a tuple's
 *     partition id is (hasht(key) >> first_bit) & (parts-1) with hasht = identity
 *     (common.h:45-47).  Only the partition *contents* (multiset per partition id) are part
 *     of the contract; the reference's bucket-chain layout is not.
 *   - host radix partition structure ... partition-primitives.cu:40-125 (histogram, prefix,
 *     scatter per thread) is the model for the multithreaded checker below.
 *
 * PARITY PIN: the reference ships no tests or golden vectors.  This oracle is
 * pinned by (1) bit-exact comparison of every generator against the reference's own
 * generator object code (oracle/_ref/libref_generator.so, built from the sources where
 * they lie by oracle/Makefile; tests/test_oracle_vs_ref.py + tests/golden/), (2) the
 * closed-form known answers the generators imply, (3) a brute-force
 * nested-loop join on small inputs, and (4) on the GPU box, the reference's CUDA kernels
 * rebuilt for sm_100a (oracle/_ref/bench_ref) run on identical .bin inputs.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------ */
/* 1. Generators                                                                         */
/* ------------------------------------------------------------------------------------ */

/* generator_ETHZ.cu:16  RAND_RANGE(N)   = rand()    / (RAND_MAX+1) * N   (double math) */
static inline double unit_rand(void) { return (double)rand() / ((double)RAND_MAX + 1.0); }
/* generator_ETHZ.cu:17  RAND_RANGE48(N) = nrand48() / (RAND_MAX+1) * N                 */
static inline double unit_rand48(unsigned short st[3]) {
    return (double)nrand48(st) / ((double)RAND_MAX + 1.0);
}

static inline void swap32(int32_t *a, int32_t *b) { int32_t t = *a; *a = *b; *b = t; }

/* generator_ETHZ.cu:22-26 seed_generator */
void orc_seed_generator(unsigned int seed) { srand(seed); }

/* generator_ETHZ.cu:204-212.  j is drawn from [0,i) -- never i -- i.e. Sattolo's variant. */
void orc_knuth_shuffle48(int32_t *rel, uint64_t n, unsigned short st[3]) {
    for (int64_t i = (int64_t)n - 1; i > 0; --i) {
        int64_t j = (int64_t)(unit_rand48(st) * (double)i);
        swap32(&rel[i], &rel[j]);
    }
}

/* generator_ETHZ.cu:194-202, same walk driven by libc rand(). */
void orc_knuth_shuffle(int32_t *rel, uint64_t n) {
    for (int64_t i = (int64_t)n - 1; i > 0; --i) {
        int64_t j = (int64_t)(unit_rand() * (double)i);
        swap32(&rel[i], &rel[j]);
    }
}

/* generator_ETHZ.cu:115-122 */
void orc_random_gen(int32_t *rel, uint64_t n, int64_t maxid) {
    for (uint64_t i = 0; i < n; ++i) rel[i] = (int32_t)(unit_rand() * (double)maxid);
}

/* The un-shuffled key sequence of random_unique_gen (generator_ETHZ.cu:137-144):
 * 0,1,...,maxid,1,2,...,maxid,1,...  (the wrap happens AFTER maxid was emitted and restarts at 1). */
void orc_unique_sequence(int32_t *rel, uint64_t n, int64_t maxid) {
    uint64_t next = 0;
    for (uint64_t i = 0; i < n; ++i) {
        rel[i] = (int32_t)next;
        if ((int64_t)next == maxid) next = 0;
        ++next;
    }
}

/* generator_ETHZ.cu:127-149 with the time(NULL) seed made explicit.
 * State layout as in the reference: the 4 seed bytes are memcpy'd over a zeroed short[3]. */
void orc_random_unique_gen(int32_t *rel, uint64_t n, int64_t maxid, unsigned int seed) {
    unsigned short st[3] = {0, 0, 0};
    memcpy(st, &seed, sizeof(seed));
    orc_unique_sequence(rel, n, maxid);
    orc_knuth_shuffle48(rel, n, st);
}

/* generator_ETHZ.cu:162-187 (the generation branch; the file cache is not part of the math). */
void orc_fk_from_pk(int32_t *fk, uint64_t nfk, const int32_t *pk, uint64_t npk) {
    uint64_t whole = nfk / npk, done = 0;
    for (uint64_t c = 0; c < whole; ++c, done += npk) memcpy(fk + done, pk, npk * sizeof(int32_t));
    if (nfk > done) memcpy(fk + done, pk, (nfk - done) * sizeof(int32_t));
    orc_knuth_shuffle(fk, nfk);
}

/* generator_ETHZ.cu:299-348 incl. gen_alphabet :236-258 and gen_zipf_lut :265-294.
 * Draw order kept exactly: alphabet permutation, then 64 discarded rand() calls (:308-311,
 * the unused seeds[] array), then one rand() per output tuple. */
void orc_gen_zipf(uint64_t n, unsigned int alphabet_size, double z, int32_t *out) {
    uint32_t *alpha = (uint32_t *)malloc((size_t)alphabet_size * sizeof(uint32_t));
    double *cdf = (double *)malloc((size_t)alphabet_size * sizeof(double));
    for (unsigned int i = 0; i < alphabet_size; ++i) alpha[i] = i + 1; /* 0 is never a symbol */
    for (unsigned int i = alphabet_size - 1; i > 0; --i) {
        unsigned int k = (unsigned int)((unsigned long)i * (unsigned long)rand() / RAND_MAX);
        uint32_t t = alpha[i]; alpha[i] = alpha[k]; alpha[k] = t;
    }
    double norm = 0.0;
    for (unsigned int i = 1; i <= alphabet_size; ++i) norm += 1.0 / pow((double)i, z);
    double run = 0.0;
    for (unsigned int i = 1; i <= alphabet_size; ++i) {
        run += 1.0 / pow((double)i, z);
        cdf[i - 1] = run / norm;
    }
    for (int i = 0; i < 64; ++i) (void)rand();
    for (uint64_t t = 0; t < n; ++t) {
        double r = (double)rand() / RAND_MAX;
        unsigned int pos;
        if (cdf[0] >= r) {
            pos = 0;
        } else {
            unsigned int lo = 0, hi = alphabet_size - 1;
            while (hi - lo > 1) {
                unsigned int mid = (lo + hi) / 2;
                if (cdf[mid] < r) lo = mid; else hi = mid;
            }
            pos = hi;
        }
        out[t] = (int32_t)alpha[pos];
    }
    free(cdf);
    free(alpha);
}

/* ------------------------------------------------------------------------------------ */
/* 2. Result arithmetic shared by every checker                                          */
/* ------------------------------------------------------------------------------------ */

/* Order-independent fingerprint of the multiset of materialised (Pr,Ps) pairs:
 * SUM over pairs of splitmix64(Pr<<32 | Ps) mod 2^64.  (New definition -- the reference's
 * materialised ring is not decodable, join-primitives.cu:1097-1099,1403-1413.) */
static inline uint64_t pair_mix(int32_t pr, int32_t ps) {
    uint64_t x = ((uint64_t)(uint32_t)pr << 32) | (uint64_t)(uint32_t)ps;
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
uint64_t orc_pair_mix(int32_t pr, int32_t ps) { return pair_mix(pr, ps); }

/* SUM(Pr*Ps): join-primitives.cu:987,:1073 `count += pval*payload[pos]`, widened to 64 bit. */
static inline uint64_t pay_prod(int32_t pr, int32_t ps) {
    return (uint64_t)((int64_t)pr * (int64_t)ps);
}

typedef struct { uint64_t matches, checksum, pairhash; } orc_result;

/* Brute force: every (r,s) pair compared.  Pins the radix checker on small inputs. */
void orc_join_naive(const int32_t *Rk, const int32_t *Rp, uint64_t nR, const int32_t *Sk,
                    const int32_t *Sp, uint64_t nS, uint64_t out[3]) {
    orc_result a = {0, 0, 0};
    for (uint64_t i = 0; i < nR; ++i)
        for (uint64_t j = 0; j < nS; ++j)
            if (Rk[i] == Sk[j]) {
                a.matches++;
                a.checksum += pay_prod(Rp[i], Sp[j]);
                a.pairhash += pair_mix(Rp[i], Sp[j]);
            }
    out[0] = a.matches; out[1] = a.checksum; out[2] = a.pairhash;
}

/* ------------------------------------------------------------------------------------ */
/* 3. Radix partition oracle (stable counting sort on one digit)                         */
/* ------------------------------------------------------------------------------------ */

/* digit = (hasht(key) >> shift) & (2^bits-1), hasht = identity (common.h:45-47);
 * join-primitives.cu:126 (pass one), :395 (pass two). Keys are treated as unsigned bits. */
static inline uint32_t digit_of(int32_t key, uint32_t shift, uint32_t bits) {
    return ((uint32_t)key >> shift) & ((1u << bits) - 1u);
}

/* offsets has 2^bits+1 entries. Output is stable: inside a partition tuples keep input order. */
void orc_partition(const int32_t *keys, const int32_t *pays, uint64_t n, uint32_t shift,
                   uint32_t bits, uint64_t *offsets, int32_t *keys_out, int32_t *pays_out) {
    uint64_t parts = 1ull << bits;
    memset(offsets, 0, (parts + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; ++i) offsets[digit_of(keys[i], shift, bits) + 1]++;
    for (uint64_t p = 0; p < parts; ++p) offsets[p + 1] += offsets[p];
    uint64_t *cur = (uint64_t *)malloc(parts * sizeof(uint64_t));
    memcpy(cur, offsets, parts * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t at = cur[digit_of(keys[i], shift, bits)]++;
        keys_out[at] = keys[i];
        if (pays_out) pays_out[at] = pays ? pays[i] : 0;
    }
    free(cur);
}

/* ------------------------------------------------------------------------------------ */
/* 4. Multithreaded host radix join (the checker and the reported CPU baseline)          */
/* ------------------------------------------------------------------------------------ */

typedef struct { int32_t k, p; } tup;

static int pick_threads(int want) {
#ifdef _OPENMP
    /* want <= 0: OpenMP's default (OMP_NUM_THREADS / all cores).  An explicit count may exceed
     * that default -- torch.distributed.run exports OMP_NUM_THREADS=1 -- up to the processors
     * this process may run on. */
    int mx = want > 0 ? omp_get_num_procs() : omp_get_max_threads();
    if (want <= 0 || want > mx) want = mx;
    return want < 1 ? 1 : want;
#else
    (void)want; return 1;
#endif
}
int orc_max_threads(void) { return pick_threads(0); }

/* One parallel radix pass over [0,n): per-thread histograms, prefix, scatter through
 * software write-combining buffers -- one cache line (8 tuples) per partition and thread,
 * written out when full (structure of partition-primitives.cu:60-125; plain 64-byte copies
 * where the reference streams with AVX2 non-temporal stores). */
#define SWWC_TUPLES 8
#define SWWC_MAX_PARTS 2048
static void pass_parallel(const int32_t *k, const int32_t *p, const tup *in, uint64_t n,
                          uint32_t shift, uint32_t bits, tup *out, uint64_t *offsets, int T) {
    uint64_t parts = 1ull << bits;
    uint64_t *hist = (uint64_t *)calloc((size_t)T * parts, sizeof(uint64_t));
    const int swwc = parts <= SWWC_MAX_PARTS;
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        uint64_t lo = n * (uint64_t)t / T, hi = n * (uint64_t)(t + 1) / T;
        uint64_t *h = hist + (size_t)t * parts;
        if (in) for (uint64_t i = lo; i < hi; ++i) h[digit_of(in[i].k, shift, bits)]++;
        else    for (uint64_t i = lo; i < hi; ++i) h[digit_of(k[i], shift, bits)]++;
#pragma omp barrier
#pragma omp single
        {
            uint64_t run = 0;
            for (uint64_t d = 0; d < parts; ++d) {
                offsets[d] = run;
                for (int tt = 0; tt < T; ++tt) {
                    uint64_t c = hist[(size_t)tt * parts + d];
                    hist[(size_t)tt * parts + d] = run;
                    run += c;
                }
            }
            offsets[parts] = run;
        }
        if (swwc) {
            tup *buf = (tup *)aligned_alloc(64, parts * SWWC_TUPLES * sizeof(tup));
            uint8_t *fill = (uint8_t *)calloc(parts, 1);
            for (uint64_t i = lo; i < hi; ++i) {
                tup x;
                if (in) x = in[i]; else { x.k = k[i]; x.p = p[i]; }
                uint64_t d = digit_of(x.k, shift, bits);
                tup *line = buf + d * SWWC_TUPLES;
                line[fill[d]] = x;
                if (++fill[d] == SWWC_TUPLES) {
                    memcpy(out + h[d], line, SWWC_TUPLES * sizeof(tup));
                    h[d] += SWWC_TUPLES;
                    fill[d] = 0;
                }
            }
            for (uint64_t d = 0; d < parts; ++d)
                for (uint8_t j = 0; j < fill[d]; ++j) out[h[d]++] = buf[d * SWWC_TUPLES + j];
            free(fill);
            free(buf);
        } else if (in) {
            for (uint64_t i = lo; i < hi; ++i) out[h[digit_of(in[i].k, shift, bits)]++] = in[i];
        } else {
            for (uint64_t i = lo; i < hi; ++i) {
                tup x = {k[i], p[i]};
                out[h[digit_of(k[i], shift, bits)]++] = x;
            }
        }
    }
    free(hist);
}

/* Second pass: every first-pass partition is sub-partitioned by one thread. */
static void pass_per_partition(const tup *in, const uint64_t *off1, uint32_t bits1,
                               uint32_t bits2, tup *out, uint64_t *off2, int T) {
    uint64_t p1n = 1ull << bits1, p2n = 1ull << bits2;
#pragma omp parallel num_threads(T)
    {
        uint64_t *h = (uint64_t *)malloc(p2n * sizeof(uint64_t));
#pragma omp for schedule(dynamic, 1)
        for (uint64_t a = 0; a < p1n; ++a) {
            uint64_t lo = off1[a], hi = off1[a + 1];
            memset(h, 0, p2n * sizeof(uint64_t));
            for (uint64_t i = lo; i < hi; ++i) h[digit_of(in[i].k, 0, bits2)]++;
            uint64_t run = lo;
            for (uint64_t d = 0; d < p2n; ++d) {
                uint64_t c = h[d];
                h[d] = run;
                off2[a * p2n + d] = run;
                run += c;
            }
            for (uint64_t i = lo; i < hi; ++i) out[h[digit_of(in[i].k, 0, bits2)]++] = in[i];
        }
        free(h);
    }
    off2[p1n * p2n] = off1[p1n];
}

/* Partition one relation on its low `bits` key bits into 2^bits contiguous partitions
 * (two passes when bits > 11, high digit first, like join-primitives.cu:1598-1609). */
static tup *radix_relation(const int32_t *k, const int32_t *p, uint64_t n, uint32_t bits,
                           uint64_t *off /* 2^bits+1 */, int T) {
    tup *a = (tup *)malloc((n ? n : 1) * sizeof(tup));
    if (bits <= 11) {
        pass_parallel(k, p, NULL, n, 0, bits, a, off, T);
        return a;
    }
    uint32_t b2 = bits / 2, b1 = bits - b2;
    uint64_t *off1 = (uint64_t *)malloc(((1ull << b1) + 1) * sizeof(uint64_t));
    pass_parallel(k, p, NULL, n, b2, b1, a, off1, T);
    tup *b = (tup *)malloc((n ? n : 1) * sizeof(tup));
    pass_per_partition(a, off1, b1, b2, b, off, T);
    free(a);
    free(off1);
    return b;
}

static uint32_t pick_bits(uint64_t nR) {
    uint32_t bits = 0;
    while (bits < 22 && (nR >> bits) > 2048) ++bits; /* ~<=2K build tuples per partition */
    return bits;
}

typedef struct { int32_t *rp, *sp; uint64_t cap; uint64_t n; } pair_sink;

/* Chained hash join of one partition pair (structure of join-primitives.cu:1005-1085: heads
 * + next links, full-key compare).  Build side is R, probe side is S. */
static void join_partition(const tup *R, uint64_t nR, const tup *S, uint64_t nS, uint32_t skip,
                           int32_t *heads, uint64_t nheads, int32_t *next, orc_result *acc,
                           pair_sink *sink) {
    if (!nR || !nS) return;
    uint64_t mask = nheads - 1;
    memset(heads, 0xFF, nheads * sizeof(int32_t));
    for (uint64_t i = 0; i < nR; ++i) {
        uint64_t h = ((uint32_t)R[i].k >> skip) & mask;
        next[i] = heads[h];
        heads[h] = (int32_t)i;
    }
    for (uint64_t j = 0; j < nS; ++j) {
        uint64_t h = ((uint32_t)S[j].k >> skip) & mask;
        for (int32_t c = heads[h]; c >= 0; c = next[c]) {
            if (R[c].k != S[j].k) continue;
            acc->matches++;
            acc->checksum += pay_prod(R[c].p, S[j].p);
            acc->pairhash += pair_mix(R[c].p, S[j].p);
            if (sink) {
                uint64_t at;
#pragma omp atomic capture
                at = sink->n++;
                if (at < sink->cap) { sink->rp[at] = R[c].p; sink->sp[at] = S[j].p; }
            }
        }
    }
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static double join_impl(const int32_t *Rk, const int32_t *Rp, uint64_t nR, const int32_t *Sk,
                        const int32_t *Sp, uint64_t nS, int threads, uint64_t out[3],
                        pair_sink *sink) {
    int T = pick_threads(threads);
    double t0 = now_s();
    uint32_t bits = pick_bits(nR);
    uint64_t parts = 1ull << bits;
    uint64_t *offR = (uint64_t *)malloc((parts + 1) * sizeof(uint64_t));
    uint64_t *offS = (uint64_t *)malloc((parts + 1) * sizeof(uint64_t));
    tup *R = radix_relation(Rk, Rp, nR, bits, offR, T);
    tup *S = radix_relation(Sk, Sp, nS, bits, offS, T);
    uint64_t maxR = 0;
    for (uint64_t p = 0; p < parts; ++p)
        if (offR[p + 1] - offR[p] > maxR) maxR = offR[p + 1] - offR[p];
    uint64_t nheads = 16;
    while (nheads < maxR && nheads < (1ull << 30)) nheads <<= 1;
    orc_result total = {0, 0, 0};
#pragma omp parallel num_threads(T)
    {
        int32_t *heads = (int32_t *)malloc(nheads * sizeof(int32_t));
        int32_t *next = (int32_t *)malloc((maxR ? maxR : 1) * sizeof(int32_t));
        orc_result acc = {0, 0, 0};
#pragma omp for schedule(dynamic, 8)
        for (uint64_t p = 0; p < parts; ++p) {
            uint64_t r = offR[p + 1] - offR[p];
            uint64_t nh = 16;
            while (nh < r) nh <<= 1;
            join_partition(R + offR[p], r, S + offS[p], offS[p + 1] - offS[p], bits, heads, nh,
                           next, &acc, sink);
        }
#pragma omp critical
        {
            total.matches += acc.matches;
            total.checksum += acc.checksum;
            total.pairhash += acc.pairhash;
        }
        free(next);
        free(heads);
    }
    double t1 = now_s();
    free(R); free(S); free(offR); free(offS);
    out[0] = total.matches; out[1] = total.checksum; out[2] = total.pairhash;
    return t1 - t0;
}

/* Returns elapsed seconds (partition both sides + build/probe, inputs already in RAM).
 * out = {matches, checksum64, pairhash64}. threads<=0 -> all cores. */
double orc_join_check(const int32_t *Rk, const int32_t *Rp, uint64_t nR, const int32_t *Sk,
                      const int32_t *Sp, uint64_t nS, int threads, uint64_t out[3]) {
    return join_impl(Rk, Rp, nR, Sk, Sp, nS, threads, out, NULL);
}

/* Same, also writing up to `cap` (Pr,Ps) pairs in unspecified order; returns the exact
 * number of result pairs (which may exceed cap). */
uint64_t orc_join_materialize(const int32_t *Rk, const int32_t *Rp, uint64_t nR,
                              const int32_t *Sk, const int32_t *Sp, uint64_t nS, int threads,
                              int32_t *out_rp, int32_t *out_sp, uint64_t cap, uint64_t out[3]) {
    pair_sink sink = {out_rp, out_sp, cap, 0};
    join_impl(Rk, Rp, nR, Sk, Sp, nS, threads, out, &sink);
    return sink.n;
}

/* Fingerprint of an explicit pair list (used on device-materialised output). */
/* Late materialisation (reference join_partitioned_varpayload, join-primitives.cu:1420-1557): the join's
 * payloads are row ids; for every result pair the reference adds Dr[pval + z*rel_size] for z < col_num1
 * and Ds[bval + z*rel_size] for z < col_num2 (:1531-1536) into an int32 (:1460, 1552).  Here: the same sum
 * over already materialised (R row id, S row id) pairs, widened to int64 mod 2^64 (its low 32 bits are the
 * reference's value). */
uint64_t orc_late_sum(const int32_t *rid, const int32_t *sid, uint64_t npairs, const int32_t *Dr,
                      uint32_t cols_r, uint64_t stride_r, const int32_t *Ds, uint32_t cols_s,
                      uint64_t stride_s) {
    uint64_t sum = 0;
    for (uint64_t i = 0; i < npairs; ++i) {
        for (uint32_t z = 0; z < cols_r; ++z) sum += (uint64_t)(int64_t)Dr[(uint64_t)z * stride_r + (uint64_t)(uint32_t)rid[i]];
        for (uint32_t z = 0; z < cols_s; ++z) sum += (uint64_t)(int64_t)Ds[(uint64_t)z * stride_s + (uint64_t)(uint32_t)sid[i]];
    }
    return sum;
}

uint64_t orc_pairs_hash(const int32_t *rp, const int32_t *sp, uint64_t n) {
    uint64_t h = 0;
#pragma omp parallel for reduction(+ : h)
    for (uint64_t i = 0; i < n; ++i) h += pair_mix(rp[i], sp[i]);
    return h;
}

/* Per-partition multiset fingerprints of a partitioned relation: for every partition id
 * the tuple count, SUM mix(key,pay) -- lets tests compare partition CONTENTS irrespective of
 * intra-partition order. counts/hashes have 2^bits entries. */
void orc_partition_fingerprint(const int32_t *keys, const int32_t *pays, uint64_t n,
                               uint32_t shift, uint32_t bits, uint64_t *counts,
                               uint64_t *hashes) {
    uint64_t parts = 1ull << bits;
    memset(counts, 0, parts * sizeof(uint64_t));
    memset(hashes, 0, parts * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t d = digit_of(keys[i], shift, bits);
        counts[d]++;
        hashes[d] += pair_mix(keys[i], pays ? pays[i] : 0);
    }
}

/* ------------------------------------------------------------------------------------ */
/* 5. Synthetic unique relations of BASELINE config 5 (no reference counterpart)         */
/* ------------------------------------------------------------------------------------ */
/* Independent restatement of the engine's device generator contract (include/gpujoin.h,
 * gj_generate_unique): key = seeded bijection of the row id on [0,n), payload = f(key, seed).
 * With both relations holding every key of [0,n) exactly once the join has n matches and
 * checksum = SUM_k f(k,seedR)*f(k,seedS) mod 2^64 -- a closed form over the key set. */
static inline uint32_t orc_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
static inline int32_t orc_payload(uint32_t key, uint32_t pay_seed) {
    return (int32_t)orc_mix32(key ^ (pay_seed * 0xC2B2AE35u + 0x27D4EB2Fu));
}
void orc_payload_of_keys(const int32_t *keys, uint64_t n, uint32_t pay_seed, int32_t *out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = orc_payload((uint32_t)keys[i], pay_seed);
}
uint32_t orc_bijection(uint64_t row, uint64_t n_total, uint32_t seed) {
    uint32_t bits = 2;
    while (bits < 32 && (1ull << bits) < n_total) ++bits;
    uint32_t half = (bits + 1) >> 1, hm = (1u << half) - 1u;
    uint64_t x = row;
    do {
        uint32_t L = (uint32_t)(x >> half) & hm, R = (uint32_t)x & hm;
        for (uint32_t r = 0; r < 4; ++r) {
            uint32_t f = orc_mix32(R + seed * 0x9E3779B9u + r * 0x85EBCA6Bu) & hm;
            uint32_t nl = R; R = L ^ f; L = nl;
        }
        x = ((uint64_t)L << half) | R;
    } while (x >= n_total);
    return (uint32_t)x;
}
/* SUM over keys k in [k_begin,k_end) of f(k,a)*f(k,b) mod 2^64 */
uint64_t orc_unique_join_checksum(uint64_t k_begin, uint64_t k_end, uint32_t seed_a, uint32_t seed_b) {
    uint64_t s = 0;
#pragma omp parallel for reduction(+ : s)
    for (uint64_t k = k_begin; k < k_end; ++k)
        s += pay_prod(orc_payload((uint32_t)k, seed_a), orc_payload((uint32_t)k, seed_b));
    return s;
}
