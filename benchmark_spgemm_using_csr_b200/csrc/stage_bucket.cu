// stage_bucket.cu -- the column CDF of the intermediate products (k_colcount, k_cdf_hist, k_cdf_scan)
// and the launchers of the bucket-sort kernel (stage_bucket.cuh).
#include "stage_bucket.cuh"

namespace bhb {

// entries of A per column = how often each row of B is referenced
__global__ void __launch_bounds__(256) k_colcount(const int nnzA, const int k, const int *__restrict__ colA, int *__restrict__ cnt)
{
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nnzA; j += (long long)gridDim.x * blockDim.x) {
        const int c = colA[j];
        if ((unsigned)c < (unsigned)k) atomicAdd(&cnt[c], 1);
    }
}

// histogram of the product columns over CDF_KNOTS equal slices of [0, n): entry (r, c) of B counts colcount[r] times
__global__ void __launch_bounds__(256) k_cdf_hist(const int k, const int *__restrict__ rowptrB, const int *__restrict__ colB,
                                                  const int *__restrict__ colcount, const int shift,
                                                  unsigned long long *__restrict__ hist)
{
    __shared__ unsigned long long s_hist[CDF_KNOTS];
    for (int i = threadIdx.x; i < CDF_KNOTS; i += blockDim.x) s_hist[i] = 0ull;
    __syncthreads();
    const int gl = threadIdx.x & 7;
    const long long stride = (long long)gridDim.x * (blockDim.x >> 3);
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < k; r += stride) {
        const unsigned long long w = (unsigned long long)colcount[r];
        if (w == 0ull) continue;
        const int s = rowptrB[r], e = rowptrB[r + 1];
        for (int p = s + gl; p < e; p += 8) atomicAdd(&s_hist[min(colB[p] >> shift, CDF_KNOTS - 1)], w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CDF_KNOTS; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

// cdf[i] = F(i << shift) * 2^32, cdf[0] = 0, cdf[CDF_KNOTS] = 2^32 - 1 (one block of 1024 threads, 4 knots each)
__global__ void __launch_bounds__(1024) k_cdf_scan(const unsigned long long *__restrict__ hist, unsigned *__restrict__ cdf)
{
    __shared__ double s_warp[33];
    const int t = threadIdx.x;
    double v[4], mine = 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        v[u] = (double)hist[t * 4 + u];
        mine += v[u];
    }
    double incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double y = __shfl_up_sync(FULL, incl, d);
        if ((t & 31) >= d) incl += y;
    }
    if ((t & 31) == 31) s_warp[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        double w = s_warp[t];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double y = __shfl_up_sync(FULL, w, d);
            if (t >= d) w += y;
        }
        s_warp[t] = w;
    }
    __syncthreads();
    const double total = s_warp[31];
    double run = ((t >> 5) ? s_warp[(t >> 5) - 1] : 0.0) + incl - mine;   // products below knot 4t
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = t * 4 + u;
        // (all counts are integers below 2^53: the sums are exact, so the table is non-decreasing)
        const double f = total > 0.0 ? run / total : (double)i / (double)CDF_KNOTS;
        cdf[i] = (unsigned)fmin(f * 4294967296.0, 4294967295.0);
        run += v[u];
    }
    if (t == 0) cdf[CDF_KNOTS] = 0xffffffffu;
}

cudaError_t launch_build_cdf(const LaunchCtx &lc, int m, int k, int n, int nnzA, Csr A, Csr B, int *colcountA,
                             unsigned long long *hist, unsigned *cdf, int *shift_out)
{
    (void)m;
    int bits = 0;
    while (bits < 31 && (1ll << bits) < (long long)n) ++bits;
    const int shift = bits > CDF_BITS ? bits - CDF_BITS : 0;
    *shift_out = shift;
    cudaError_t e = cudaMemsetAsync(colcountA, 0, ((size_t)k + 1) * 4, lc.stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(hist, 0, (size_t)CDF_KNOTS * 8, lc.stream);
    if (e != cudaSuccess) return e;
    if (nnzA > 0) {
        long long blocks = ((long long)nnzA + 255) / 256;
        if (blocks > (long long)lc.sm_count * 16) blocks = (long long)lc.sm_count * 16;
        ++*lc.launches;
        k_colcount<<<(int)blocks, 256, 0, lc.stream>>>(nnzA, k, A.col, colcountA);
    }
    if (k > 0) {
        long long blocks = ((long long)k * 8 + 255) / 256;
        if (blocks > (long long)lc.sm_count * 4) blocks = (long long)lc.sm_count * 4;
        ++*lc.launches;
        k_cdf_hist<<<(int)blocks, 256, 0, lc.stream>>>(k, B.rowptr, B.col, colcountA, shift, hist);
    }
    ++*lc.launches;
    k_cdf_scan<<<1, 1024, 0, lc.stream>>>(hist, cdf);
    return cudaGetLastError();
}

}  // namespace bhb
