/* oracle/refbuild/numa.h -- stand-in for libnuma (absent in this image) so the reference
 * sources compile unmodified. Only the reference's out-of-scope streaming / co-processing
 * modes call these (hash_join_clustered_probe.cu:1801-1823, partition-primitives.cu). */
#ifndef GJ_REFBUILD_NUMA_STUB_H
#define GJ_REFBUILD_NUMA_STUB_H
#include <stdlib.h>
static inline void *numa_alloc_onnode(size_t bytes, int node) { (void)node; return malloc(bytes); }
static inline void numa_free(void *p, size_t bytes) { (void)bytes; free(p); }
static inline int numa_node_of_cpu(int cpu) { (void)cpu; return 0; }
static inline int numa_available(void) { return -1; }
#endif
