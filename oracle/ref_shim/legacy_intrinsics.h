/* Force-included (-include) when the reference is compiled for sm_100a: the pre-Volta
 * __shfl_up the reference calls (bhsparse_cuda.h:1031,1081,1133) no longer exists; all three
 * call sites are executed by full warps, so the _sync form with a full mask is equivalent.
 * TEST INFRASTRUCTURE (oracle/_ref build only). */
#ifndef BHB200_REF_SHIM_LEGACY_INTRINSICS_H
#define BHB200_REF_SHIM_LEGACY_INTRINSICS_H
#define __shfl_up(var, delta) __shfl_up_sync(0xffffffffu, (var), (delta))
#endif
