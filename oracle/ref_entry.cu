// ref_entry.cu -- C-ABI harness around the UNMODIFIED reference (bhSPARSE, SpGEMM_cuda/).
//
// TEST INFRASTRUCTURE ONLY (oracle/_ref): compiled by oracle/build_ref.py from the reference's
// own headers WHERE THEY LIE under /root/reference/SpGEMM_cuda (-I, nothing is copied) into
// oracle/_ref/libbhsparse_ref_{f64,f32}.so.  Only tests/, __graft_entry__.smoke() and bench.py's
// reference legs may load it; the product library never does.
//
// What is replaced, and why the result is still "the reference":
//   * main.cu / ref_spgemm.h (the CUSP-based driver and check) are not compiled: CUSP v0.4.0
//     (README.md:91) is not vendored.  This file plays main.cu:104-135's role: the same call
//     protocol initPlatform -> initData -> spgemm -> get_nnzC -> get_C -> free_mem -> freePlatform
//     on caller-supplied CSR arrays.
//   * helper_functions.h / helper_cuda.h (CUDA samples, common.h:21-22): stand-ins in
//     oracle/ref_shim/ (a host stopwatch and an error check that records instead of exit()).
//   * __shfl_up (bhsparse_cuda.h:1031,1081,1133): mapped to __shfl_up_sync(full mask) by the
//     force-included oracle/ref_shim/legacy_intrinsics.h.
//   * value_type (common.h:31) is `double`; the f32 build re-types it the way README.md:84-86
//     tells users to (edit the typedef) -- done here with a macro around the first inclusion of
//     common.h instead of editing the file.
// Every kernel, the host binning (bhsparse.h:365-481) and the re-allocation loop
// (bhsparse_cuda.h:2527-2780) are the reference's, compiled for sm_100a.
#ifdef BHREF_F32
#define value_type bhref_unused_value_type
#include "common.h"      // SpGEMM_cuda/common.h (include guard COMMON_H): typedef double bhref_unused_value_type
#undef value_type
typedef float value_type;
#endif
#include "bhsparse.h"    // SpGEMM_cuda/bhsparse.h -> bhsparse_cuda.h -> common.h

#include <chrono>
#include <sstream>
#include <string>

extern "C" {
int bhref_cuda_error_count = 0;
}

namespace {
bhsparse *g_bh = nullptr;
bool g_platforms[NUM_PLATFORMS];
std::string g_log;

// the reference prints its stage times to std::cout; keep the caller's stdout clean
struct CoutCapture {
    std::ostringstream os;
    std::streambuf *old;
    CoutCapture() : old(std::cout.rdbuf(os.rdbuf())) {}
    ~CoutCapture()
    {
        std::cout.rdbuf(old);
        g_log += os.str();
    }
};
}  // namespace

#define BHREF_API extern "C" __attribute__((visibility("default")))

BHREF_API int bhref_value_size(void)
{
    return (int)sizeof(value_type);
}

BHREF_API const char *bhref_log(void)
{
    return g_log.c_str();
}

BHREF_API int bhref_cuda_errors(void)
{
    return bhref_cuda_error_count;
}

// main.cu:104-122: initPlatform, initData, [warmup x n], spgemm, get_nnzC.  rowptrC (m+1 ints) is
// caller-allocated and retained until bhref_get_C (bhsparse.h:213).  *ms = wall time of spgemm()
// (what the reference prints as "SpGEMM time", bhsparse.h:268-289).  Returns the reference's
// error code, or -1000 - (number of CUDA errors its calls raised).
BHREF_API int bhref_spgemm(int m, int k, int n, int nnzA, const void *valA, const int *rowptrA, const int *colA, int nnzB,
                           const void *valB, const int *rowptrB, const int *colB, int *rowptrC, int warmups, int *nnzC,
                           double *ms)
{
    g_log.clear();
    bhref_cuda_error_count = 0;
    CoutCapture cap;
    if (g_bh) return -2;   // previous product not collected
    g_bh = new bhsparse();
    memset(g_platforms, 0, sizeof(g_platforms));
    g_platforms[BHSPARSE_CUDA] = true;
    int err = g_bh->initPlatform(g_platforms);
    if (err) return err;
    err = g_bh->initData(m, k, n, nnzA, (value_type *)valA, (index_type *)rowptrA, (index_type *)colA, nnzB,
                         (value_type *)valB, (index_type *)rowptrB, (index_type *)colB, rowptrC);
    if (err) return err;
    for (int i = 0; i < warmups; ++i) err |= g_bh->warmup();
    cudaDeviceSynchronize();
    const auto t0 = std::chrono::steady_clock::now();
    err |= g_bh->spgemm();
    cudaDeviceSynchronize();
    if (ms) *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (nnzC) *nnzC = g_bh->get_nnzC();
    if (bhref_cuda_error_count) return -1000 - bhref_cuda_error_count;
    return err;
}

// main.cu:125-135: get_C into caller arrays of get_nnzC() entries, then free_mem, freePlatform.
// colC/valC may be null (result discarded).
BHREF_API int bhref_get_C(int *colC, void *valC)
{
    CoutCapture cap;
    if (!g_bh) return -2;
    int err = 0;
    if (colC && valC) err = g_bh->get_C(colC, (value_type *)valC);
    err |= g_bh->free_mem();
    err |= g_bh->freePlatform();
    delete g_bh;
    g_bh = nullptr;
    if (bhref_cuda_error_count) return -1000 - bhref_cuda_error_count;
    return err;
}
