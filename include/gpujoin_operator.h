/*
 * gpujoin_operator.h -- the reference's operator-level interface, served by libgpujoin.so.
 *
 * C++ declarations with the reference's own names, argument order and meaning, so that the
 * reference driver (src/main.cu) links against libgpujoin.so unchanged:
 *   hashJoinClusteredProbe   hash_join_clustered_probe.cu:2062  (the algs[] entry, main.cu:64)
 *   hj_ClusteredProbe        hash_join_clustered_probe.cu:1990
 *   outOfGPU_Join1_payload   hash_join_clustered_probe.cu:802
 * `args` mirrors common-host.h:39-52 field for field; `timingInfo` (common.h:101-119) is only
 * passed through as a pointer and never dereferenced here.
 *
 * Behavioural differences, all deliberate:
 *   - no size dispatch to the PCIe streaming / CPU co-processing modes
 *     (hash_join_clustered_probe.cu:2001-2009): every size that fits in HBM runs in-GPU;
 *   - CUDA failures print "GPU Error: ..." and make the call return ~0u instead of exit()
 *     (common.h:132-141);
 *   - log_parts1/log_parts2 are taken as a fan-out REQUEST (0 = engine chooses); first_bit is
 *     ignored exactly like the reference does (hash_join_clustered_probe.cu:813).
 * Stdout keeps the reference's lines (hash_join_clustered_probe.cu:937-940, 986-991).
 */
#ifndef GPUJOIN_OPERATOR_H
#define GPUJOIN_OPERATOR_H
#include <stddef.h>
#include <stdint.h>

#ifndef COMMON_HOST_H_   /* the reference's own header defines `args` when compiled with it */
typedef struct args {
    int* S;
    size_t S_els;
    char S_filename[50];
    int* R;
    size_t R_els;
    char R_filename[50];
    int threadsNum;
    unsigned int sharedMem;
    unsigned int pivotsNum;
} args;
#endif
struct timingInfo;

unsigned int hashJoinClusteredProbe(args* inputAttrs, timingInfo* time);
unsigned int hj_ClusteredProbe(int* R, size_t RelsNum, int* S, size_t SelsNum, timingInfo* time);
unsigned int outOfGPU_Join1_payload(int* R, int* Pr, size_t RelsNum, int* S, int* Ps, size_t SelsNum,
                                    timingInfo* time, unsigned int log_parts1, unsigned int log_parts2,
                                    unsigned int first_bit);

/* Results of the last outOfGPU_Join1_payload call on this thread (the reference reports them on
 * stdout only). */
extern "C" {
typedef struct gj_operator_result {
    uint64_t matches;          /* exact result size */
    uint64_t checksum;         /* SUM Pr*Ps mod 2^64 */
    uint64_t pairs_materialized;
    int32_t ref_results;       /* what the reference prints as "%d results" */
    double partition_mbps[2];  /* [0] materialising run, [1] aggregate run */
    double join_mbps[2];
    double total_mbps[2];
    int status;                /* gj_status of the last call */
} gj_operator_result;
const gj_operator_result* gj_operator_last_result(void);
void gj_operator_release(void); /* frees the cached engine context */
}
#endif
