/* Stand-in for the CUDA-samples header the reference includes (SpGEMM_cuda/common.h:22).
 * TEST INFRASTRUCTURE (oracle/_ref build only).  The reference aborts the process on a CUDA
 * error (samples' checkCudaErrors); here the error is recorded and the call returns, so the
 * test harness can report it instead of dying. */
#ifndef BHB200_REF_SHIM_HELPER_CUDA_H
#define BHB200_REF_SHIM_HELPER_CUDA_H
#include <cuda_runtime.h>
#include <stdio.h>

extern "C" int bhref_cuda_error_count;
inline void bhref_check(cudaError_t e, const char *what, const char *file, int line)
{
    if (e != cudaSuccess) {
        ++bhref_cuda_error_count;
        fprintf(stderr, "[oracle/_ref] CUDA error %s at %s:%d: %s\n", cudaGetErrorString(e), file, line, what);
    }
}
#define checkCudaErrors(call) bhref_check((call), #call, __FILE__, __LINE__)
#define getLastCudaError(msg) bhref_check(cudaGetLastError(), (msg), __FILE__, __LINE__)
#endif
