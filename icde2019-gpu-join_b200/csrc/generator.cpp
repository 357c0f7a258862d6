// generator.cpp -- ETHZ-style relation generator (host).  Interface: include/gpujoin_generator.h.
// Follows the algorithms of the reference's generator_ETHZ.cu (file:line cited per function);
// written from their description in SURVEY.md Appendix C, with explicit seeds.
#include "../../include/gpujoin_generator.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// generator_ETHZ.cu:16-17: rand()/nrand48() scaled to [0, N) in double arithmetic
inline uint64_t scaled(long draw, uint64_t n) {
    return (uint64_t)(((double)draw / ((double)RAND_MAX + 1.0)) * (double)n);
}

template <class Draw>
void sattolo_walk(int32_t* rel, uint64_t n, Draw&& draw) {
    // generator_ETHZ.cu:194-212: i runs n-1..1, partner index is drawn from [0, i)
    if (n < 2) return;
    for (uint64_t i = n - 1; i > 0; --i) std::swap(rel[i], rel[scaled(draw(), i)]);
}

// generator_ETHZ.cu:137-144: 0,1,...,maxid,1,2,...,maxid,1,...
inline int32_t unique_sequence_at(uint64_t i, uint64_t maxid) {
    if (i <= maxid || maxid == 0) return (int32_t)i;
    return (int32_t)((i - maxid - 1) % maxid + 1);
}

inline uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

inline uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// seeded bijection on [0, n): cycle-walking 4-round Feistel network
inline uint64_t permute_index(uint64_t i, uint64_t n, uint32_t seed) {
    uint32_t bits = 2;
    while (bits < 40 && (1ull << bits) < n) ++bits;
    const uint32_t half = (bits + 1) >> 1;
    const uint64_t hm = (1ull << half) - 1;
    uint64_t x = i;
    do {
        uint64_t L = (x >> half) & hm, R = x & hm;
        for (uint32_t r = 0; r < 4; ++r) {
            const uint64_t f = mix32((uint32_t)R * 0x9E3779B1u + seed + r * 0x85EBCA6Bu + (uint32_t)(R >> 20)) & hm;
            const uint64_t nl = R;
            R = L ^ f;
            L = nl;
        }
        x = (L << half) | R;
    } while (x >= n);
    return x;
}

int pick_threads(int want) {
#ifdef _OPENMP
    const int mx = omp_get_max_threads();
    return (want <= 0 || want > mx) ? mx : want;
#else
    (void)want;
    return 1;
#endif
}

}  // namespace

extern "C" void gj_seed_generator(unsigned int seed) { srand(seed); }

extern "C" int gj_read_relation(const char* filename, int32_t* relation, uint64_t n) {
    if (!filename) return 1;
    FILE* fp = fopen(filename, "rb");
    if (!fp) return 1;
    printf("Reading file %s ", filename);   // generator_ETHZ.cu:45
    fflush(stdout);
    const size_t got = fread(relation, sizeof(int32_t), n, fp);
    fclose(fp);
    return got == n ? 0 : 2;
}

extern "C" int gj_write_relation(const char* filename, const int32_t* relation, uint64_t n) {
    if (!filename) return 0;
    FILE* fp = fopen(filename, "wb");
    if (!fp) return 1;
    const size_t put = fwrite(relation, sizeof(int32_t), n, fp);
    return (fclose(fp) == 0 && put == n) ? 0 : 1;
}

extern "C" void gj_random_gen(int32_t* rel, uint64_t n, int64_t maxid) {
    for (uint64_t i = 0; i < n; ++i) rel[i] = (int32_t)scaled(rand(), (uint64_t)maxid);
}

extern "C" void gj_knuth_shuffle(int32_t* rel, uint64_t n) {
    sattolo_walk(rel, n, [] { return (long)rand(); });
}

extern "C" void gj_knuth_shuffle48(int32_t* rel, uint64_t n, unsigned short state[3]) {
    sattolo_walk(rel, n, [state] { return nrand48(state); });
}

extern "C" void gj_random_unique_gen(int32_t* rel, uint64_t n, int64_t maxid, unsigned int seed) {
    for (uint64_t i = 0; i < n; ++i) rel[i] = unique_sequence_at(i, (uint64_t)maxid);
    unsigned short state[3] = {(unsigned short)(seed & 0xFFFFu), (unsigned short)(seed >> 16), 0};
    gj_knuth_shuffle48(rel, n, state);
}

// generator_ETHZ.cu:236-343
extern "C" void gj_gen_zipf(uint64_t n, unsigned int alphabet_size, double z, int32_t* out) {
    std::vector<uint32_t> alphabet(alphabet_size);
    for (unsigned int i = 0; i < alphabet_size; ++i) alphabet[i] = i + 1;
    for (unsigned int i = alphabet_size; i-- > 1;) {
        const unsigned int k = (unsigned int)((unsigned long)i * (unsigned long)rand() / RAND_MAX);
        std::swap(alphabet[i], alphabet[k]);
    }
    std::vector<double> lut(alphabet_size);
    double scaling = 0.0;
    for (unsigned int i = 1; i <= alphabet_size; ++i) scaling += 1.0 / pow((double)i, z);
    double acc = 0.0;
    for (unsigned int i = 1; i <= alphabet_size; ++i) {
        acc += 1.0 / pow((double)i, z);
        lut[i - 1] = acc / scaling;
    }
    for (int i = 0; i < 64; ++i) (void)rand();   // the reference draws 64 unused seeds (:308-311)
    for (uint64_t t = 0; t < n; ++t) {
        const double r = (double)rand() / RAND_MAX;
        unsigned int left = 0, right = alphabet_size - 1, pos;
        if (lut[0] >= r) pos = 0;
        else {
            while (right - left > 1) {
                const unsigned int m = (left + right) / 2;
                if (lut[m] < r) left = m; else right = m;
            }
            pos = right;
        }
        out[t] = (int32_t)alphabet[pos];
    }
}

extern "C" int gj_create_relation_unique(const char* filename, int32_t* rel, uint64_t n, int64_t maxid, unsigned int seed) {
    if (gj_read_relation(filename, rel, n) == 0) return 0;
    gj_random_unique_gen(rel, n, maxid, seed);
    return gj_write_relation(filename, rel, n);
}

extern "C" int gj_create_relation_nonunique(const char* filename, int32_t* rel, uint64_t n, int64_t maxid) {
    if (gj_read_relation(filename, rel, n) == 0) return 0;
    gj_random_gen(rel, n, maxid);
    return gj_write_relation(filename, rel, n);
}

// generator_ETHZ.cu:162-187: whole copies of the primary keys, then a prefix, then one shuffle
extern "C" int gj_create_relation_fk_from_pk(const char* filename, int32_t* fk, uint64_t nfk, const int32_t* pk, uint64_t npk) {
    if (gj_read_relation(filename, fk, nfk) == 0) return 0;
    for (uint64_t at = 0; at < nfk; at += npk) memcpy(fk + at, pk, std::min(npk, nfk - at) * sizeof(int32_t));
    gj_knuth_shuffle(fk, nfk);
    return gj_write_relation(filename, fk, nfk);
}

extern "C" int gj_create_relation_zipf(const char* filename, int32_t* rel, uint64_t n, int64_t maxid, double z) {
    if (gj_read_relation(filename, rel, n) == 0) return 0;
    gj_gen_zipf(n, (unsigned int)maxid, z, rel);
    return gj_write_relation(filename, rel, n);
}

extern "C" int gj_create_relation_n(const int32_t* in, int32_t* out, uint64_t n, uint64_t copies) {
    for (uint64_t c = 0; c < copies; ++c) memcpy(out + c * n, in, n * sizeof(int32_t));
    return 0;
}

extern "C" int gj_create_relation_unique_parallel(int32_t* rel, uint64_t n, int64_t maxid, unsigned int seed, int threads) {
    if (maxid <= 0 && n > 1) return 1;
    const int T = pick_threads(threads);
    (void)T;
#pragma omp parallel for num_threads(T) schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i)
        rel[permute_index((uint64_t)i, n, seed)] = unique_sequence_at((uint64_t)i, (uint64_t)maxid);
    return 0;
}

extern "C" int gj_create_relation_zipf_parallel(int32_t* rel, uint64_t n, unsigned int alphabet_size, double z, unsigned int seed, int threads) {
    if (!alphabet_size) return 1;
    const int T = pick_threads(threads);
    std::vector<double> lut(alphabet_size);
    // CDF: per-thread partial sums of k^-z, then a prefix over the thread blocks
    std::vector<double> block(T + 1, 0.0);
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
        const uint64_t lo = (uint64_t)alphabet_size * t / T, hi = (uint64_t)alphabet_size * (t + 1) / T;
        double s = 0.0;
        for (uint64_t i = lo; i < hi; ++i) { s += 1.0 / pow((double)(i + 1), z); lut[i] = s; }
        block[t + 1] = s;
#pragma omp barrier
#pragma omp single
        for (int k = 0; k < T; ++k) block[k + 1] += block[k];
        const double total = block[T], base = block[t];
        for (uint64_t i = lo; i < hi; ++i) lut[i] = (lut[i] + base) / total;
    }
    lut[alphabet_size - 1] = 1.0;
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
        uint64_t st = ((uint64_t)seed << 32) ^ (0xD1B54A32D192ED03ull * (uint64_t)(t + 1));
        const uint64_t lo = n * (uint64_t)t / T, hi = n * (uint64_t)(t + 1) / T;
        for (uint64_t i = lo; i < hi; ++i) {
            const double r = (double)(splitmix(st) >> 11) * (1.0 / 9007199254740992.0);
            const uint64_t pos = (uint64_t)(std::lower_bound(lut.begin(), lut.end(), r) - lut.begin());
            // alphabet = seeded permutation of 1..alphabet_size (symbol 0 never occurs)
            rel[i] = (int32_t)(permute_index(std::min<uint64_t>(pos, alphabet_size - 1), alphabet_size, seed ^ 0xA5A5A5A5u) + 1);
        }
    }
    return 0;
}
