/* oracle/refbuild/shim.h -- force-included when rebuilding the reference for sm_100a.
 * The reference uses pre-Volta warp intrinsics (join-primitives.cu:150,410,815-816,1212,
 * 1228,1248,1344,1358,1376,1408) that no longer exist. Mask choice:
 * __any in the partition kernels (lines < 1000) runs under divergence -> __activemask();
 * the materialising join's collectives (lines >= 1000) are whole-warp -> full mask. */
#ifndef GJ_REFBUILD_SHIM_H
#define GJ_REFBUILD_SHIM_H
#ifdef __CUDACC__
#define __shfl(v, l) __shfl_sync(0xffffffffu, (v), (l))
#define __ballot(p) __ballot_sync(0xffffffffu, (p))
#define __any(p) __any_sync(((__LINE__) >= 1000) ? 0xffffffffu : __activemask(), (p))
#endif
#endif
