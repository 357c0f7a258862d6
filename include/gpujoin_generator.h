/*
 * gpujoin_generator.h -- ETHZ-style workload generator of the B200 radix-join engine (host side).
 *
 * Same functions, argument meaning and raw-int32 file format as the reference's
 * generator_ETHZ.cuh:11-23 (psiul/ICDE2019-GPU-Join), with two differences:
 *   - seeds are explicit arguments (the reference takes time(NULL), generator_ETHZ.cu:32,134), so
 *     inputs are reproducible; for equal seeds the libc rand()/nrand48() draw order is the
 *     reference's, hence the key streams are byte-identical to the reference's object code;
 *   - *_parallel variants generate the >= 128 M-tuple relations of the benchmark configurations
 *     with all host cores (same distribution definitions, different random streams).
 * File format (generator_ETHZ.cu:38-72): raw little-endian int32[n], no header.
 * All create_* functions first try to read `filename` (when not NULL) and only generate + write
 * it on failure, like the reference's file cache (generator_ETHZ.cu:74-94,214-225).
 * Return 0 on success, non-zero on I/O failure.
 */
#ifndef GPUJOIN_GENERATOR_H
#define GPUJOIN_GENERATOR_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

void gj_seed_generator(unsigned int seed);                                   /* generator_ETHZ.cu:23 */
int gj_read_relation(const char* filename, int32_t* relation, uint64_t n);    /* readFromFile :38 */
int gj_write_relation(const char* filename, const int32_t* relation, uint64_t n); /* writeToFile :61 */

void gj_random_gen(int32_t* rel, uint64_t n, int64_t maxid);                  /* :115 */
void gj_random_unique_gen(int32_t* rel, uint64_t n, int64_t maxid, unsigned int seed); /* :127 */
void gj_knuth_shuffle(int32_t* rel, uint64_t n);                              /* :194 */
void gj_knuth_shuffle48(int32_t* rel, uint64_t n, unsigned short state[3]);   /* :204 */
void gj_gen_zipf(uint64_t n, unsigned int alphabet_size, double z, int32_t* out); /* :299 */

int gj_create_relation_unique(const char* filename, int32_t* rel, uint64_t n, int64_t maxid,
                              unsigned int seed);                             /* :86 */
int gj_create_relation_nonunique(const char* filename, int32_t* rel, uint64_t n, int64_t maxid); /* :74 */
int gj_create_relation_fk_from_pk(const char* filename, int32_t* fk, uint64_t nfk,
                                  const int32_t* pk, uint64_t npk);           /* :162 */
int gj_create_relation_zipf(const char* filename, int32_t* rel, uint64_t n, int64_t maxid,
                            double z);                                        /* :214 */
int gj_create_relation_n(const int32_t* in, int32_t* out, uint64_t n, uint64_t copies); /* :97 */

/* Multithreaded variants for the large benchmark relations (threads <= 0: all cores).
 * unique: the reference's key multiset 0,1..maxid,1..maxid,... (generator_ETHZ.cu:137-144)
 * scattered by a seeded bijection of the row index instead of a sequential Knuth shuffle.
 * zipf: the reference's definition (alphabet = random permutation of 1..alphabet_size, CDF
 * inversion by binary search, generator_ETHZ.cu:236-343) with per-thread counter-based
 * random streams. */
int gj_create_relation_unique_parallel(int32_t* rel, uint64_t n, int64_t maxid, unsigned int seed,
                                       int threads);
int gj_create_relation_zipf_parallel(int32_t* rel, uint64_t n, unsigned int alphabet_size,
                                     double z, unsigned int seed, int threads);

#ifdef __cplusplus
}
#endif
#endif
