// stage_symbolic.cu -- exact nnz(C_i) per row, binned by the upper bound p.
//
// k_sym_group<G,LOG2T> : a group of G lanes (8 or 32) owns a row and a T-slot column
//                        table in shared memory; B rows are streamed one per step with
//                        coalesced loads of their column segments.
//                        (tables up to 1024 slots; larger ones only for sampling and retries).
// k_sym_block<LOG2T>   : one CTA per row, 2048 to 32768 slots (128 KB) -- the sizes for which
//                        the reference needs its iterative merge with host re-allocation
//                        (EM_mergepath, bhsparse_cuda.h:1902-2157, host loop :2527-2780).
// k_sym_large          : rows beyond that: column bitmap in global memory (L2 resident).
// None of these has a reference counterpart as a separate pass: the reference finds
// nnz(C_i) as a by-product of its numeric kernels and compacts afterwards
// (copyCt2C, :2813-2911); counting first lets C be allocated exactly and written once.
#include "common.cuh"
#include <cstdlib>

namespace bhb {

template <int G, int LOG2T>
__global__ void __launch_bounds__(256) k_sym_group(const int *__restrict__ queue, const int count_,
                                                   const int *__restrict__ rowptrA, const int *__restrict__ colA,
                                                   const int *__restrict__ rowptrB, const int *__restrict__ colB,
                                                   int *__restrict__ rc, const int qstride,
                                                   int *__restrict__ bin_max, const int *__restrict__ dcount,
                                                   unsigned long long *__restrict__ bin_sum)
{
    constexpr int T = 1 << LOG2T;
    extern __shared__ int smem_i[];
    const int lane = threadIdx.x & 31;
    const int gl = threadIdx.x & (G - 1);
    const int count = dcount ? min(*dcount, count_) : count_;
    const int gib = threadIdx.x / G;
    const int groups_per_block = blockDim.x / G;
    int *keys = smem_i + gib * T;
    (void)lane;

    // All loops are WARP-uniform (trip counts are maxima over the 32/G groups of the warp,
    // groups with less work are predicated off): sub-warp groups with their own loop
    // counters never reconverge on sm_100a and then run at G/32 efficiency.
    for (int q0 = blockIdx.x * groups_per_block + (gib & ~(32 / G - 1)); q0 < count;
         q0 += gridDim.x * groups_per_block) {
        const int q = q0 + (gib & (32 / G - 1));
        const bool active = q < count;
        const int row = active ? queue[(long long)q * qstride] : 0;
#pragma unroll 4
        for (int s = gl; s < T; s += G) keys[s] = EMPTY_KEY;
        __syncwarp();
        const int a0 = active ? rowptrA[row] : 0;
        const int na = active ? rowptrA[row + 1] - a0 : 0;
        const int max_na = (G == 32) ? na : __reduce_max_sync(FULL, na);   // uniform already for G == 32
        int newcnt = 0;
        for (int base = 0; base < max_na; base += G) {
            const int j = base + gl;
            int bs = 0, len = 0;
            if (j < na) {
                const int k = colA[a0 + j];
                bs = rowptrB[k];
                len = rowptrB[k + 1] - bs;
            }
            const int cnt = min(G, max_na - base);
            for (int t = 0; t < cnt; ++t) {
                const int s_bs = __shfl_sync(FULL, bs, t, G);
                const int s_len = __shfl_sync(FULL, len, t, G);
                const int max_len = (G == 32) ? s_len : __reduce_max_sync(FULL, s_len);   // uniform already for G == 32
                for (int off0 = 0; off0 < max_len; off0 += G) {
                    const int off = off0 + gl;
                    if (off < s_len) {
                        bool is_new;
                        table_insert<LOG2T>(keys, colB[s_bs + off], is_new);
                        newcnt += is_new;
                    }
                }
            }
        }
#pragma unroll
        for (int d = G >> 1; d > 0; d >>= 1) newcnt += __shfl_xor_sync(FULL, newcnt, d, G);
        if (gl == 0 && active) {
            rc[row] = newcnt;
            if (bin_max) atomicMax(bin_max, newcnt);
            if (bin_sum) atomicAdd(bin_sum, (unsigned long long)newcnt);
        }
        __syncwarp();
    }
}

template <int LOG2T>
__global__ void __launch_bounds__(1024) k_sym_block(const int *__restrict__ queue, const int count,
                                                   const int *__restrict__ rowptrA, const int *__restrict__ colA,
                                                   const int *__restrict__ rowptrB, const int *__restrict__ colB,
                                                   int *__restrict__ rc)
{
    constexpr int T = 1 << LOG2T;
    extern __shared__ int smem_i[];
    __shared__ int s_red[33];
    __shared__ int s_next;
    int *keys = smem_i;
    const int lane = threadIdx.x & 31;
    for (int q = blockIdx.x; q < count; q += gridDim.x) {
        const int row = queue[q];
        for (int s = threadIdx.x; s < T; s += blockDim.x) keys[s] = EMPTY_KEY;
        if (threadIdx.x == 0) s_next = 0;
        __syncthreads();
        const int a0 = rowptrA[row], a1 = rowptrA[row + 1];
        int newcnt = 0;
        for (int j = a0 + take_next(&s_next, lane); j < a1; j = a0 + take_next(&s_next, lane)) {
            const int k = colA[j];
            const int bs = rowptrB[k], be = rowptrB[k + 1];
            for (int p = bs + lane; p < be; p += 32) {
                bool is_new;
                table_insert<LOG2T>(keys, colB[p], is_new);
                newcnt += is_new;
            }
        }
        const int tot = block_sum(newcnt, s_red);
        if (threadIdx.x == 0) rc[row] = tot;
        __syncthreads();
    }
}

// Rows whose upper bound exceeds the largest shared-memory table.  Each resident CTA
// owns an all-zero bitmap of n bits in global memory (it stays in the 126 MB L2); set
// bits with atomicOr, count and clear the touched word range.
__global__ void __launch_bounds__(1024) k_sym_large(const int *__restrict__ queue, const int count,
                                                    const int *__restrict__ rowptrA, const int *__restrict__ colA,
                                                    const int *__restrict__ rowptrB, const int *__restrict__ colB,
                                                    int *__restrict__ rc, unsigned *__restrict__ bitmap_all,
                                                    const int nwords)
{
    __shared__ int s_red[33];
    __shared__ int s_lo, s_hi, s_next, s_nlong, s_long[LONG_CAP];
    unsigned *bm = bitmap_all + (size_t)blockIdx.x * nwords;
    const int lane = threadIdx.x & 31;
    for (int q = blockIdx.x; q < count; q += gridDim.x) {
        const int row = queue[q];
        if (threadIdx.x == 0) {
            s_lo = 0x7fffffff;
            s_hi = -1;
            s_next = 0;
            s_nlong = 0;
        }
        __syncthreads();
        const int a0 = rowptrA[row], a1 = rowptrA[row + 1];
        int wlo = 0x7fffffff, whi = -1;
        cta_for_each_b_row(a0, a1, colA, rowptrB, &s_next, &s_nlong, s_long, [&](int, int p0, int pe, int stride) {
            for (int p = p0; p < pe; p += stride) {
                const int c = colB[p];
                const int w = c >> 5;
                atomicOr(&bm[w], 1u << (c & 31));
                wlo = min(wlo, w);
                whi = max(whi, w);
            }
        });
        if (whi >= 0) {
            atomicMin(&s_lo, wlo);
            atomicMax(&s_hi, whi);
        }
        __threadfence();
        __syncthreads();
        const int lo = s_lo, hi = s_hi;
        int cnt = 0;
        for (int w = lo + (int)threadIdx.x; w <= hi; w += blockDim.x) {
            const unsigned bits = __ldcg(bm + w);
            if (bits) {
                cnt += __popc(bits);
                bm[w] = 0u;
            }
        }
        const int tot = block_sum(cnt, s_red);
        if (threadIdx.x == 0) rc[row] = tot;
        __threadfence();
        __syncthreads();
    }
}

// ---- launchers -------------------------------------------------------------
template <int G, int LOG2T>
static cudaError_t launch_sym_group_t(const LaunchCtx &lc, const int *queue, int count, Csr A, Csr B, int *rc,
                                      int qstride, int *bin_max, const int *dcount, unsigned long long *bin_sum)
{
    constexpr int T = 1 << LOG2T;
    constexpr size_t per_group = (size_t)T * 4;
    // groups per block: as many as fit ~64 KB, at most 256 threads, at least one warp
    int groups = (int)((64 * 1024) / per_group);
    const int max_groups = 256 / G;
    if (groups > max_groups) groups = max_groups;
    const int min_groups = 32 / G;
    groups -= groups % min_groups;   // whole warps only
    if (groups < min_groups) groups = min_groups;
    const int threads = groups * G;
    const size_t smem = per_group * groups;
    if (smem > 48 * 1024) {   // per device, so not cached in a static
        cudaError_t e = cudaFuncSetAttribute(k_sym_group<G, LOG2T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    long long blocks = ((long long)count + groups - 1) / groups;
    const int per_sm = resident_blocks(k_sym_group<G, LOG2T>, threads, smem);
    const long long cap = (long long)lc.sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    ++*lc.launches;
    k_sym_group<G, LOG2T><<<(int)blocks, threads, smem, lc.stream>>>(queue, count, A.rowptr, A.col, B.rowptr, B.col, rc,
                                                                      qstride, bin_max, dcount, bin_sum);
    return cudaGetLastError();
}

template <int LOG2T>
static cudaError_t launch_sym_block_t(const LaunchCtx &lc, const int *queue, int count, Csr A, Csr B, int *rc,
                                      int threads = 512)
{
    constexpr int T = 1 << LOG2T;
    const size_t smem = (size_t)T * 4;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_sym_block<LOG2T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int per_sm = resident_blocks(k_sym_block<LOG2T>, threads, smem);
    long long blocks = count;
    const long long cap = (long long)lc.sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    ++*lc.launches;
    k_sym_block<LOG2T><<<(int)blocks, threads, smem, lc.stream>>>(queue, count, A.rowptr, A.col, B.rowptr, B.col, rc);
    return cudaGetLastError();
}

cudaError_t launch_sym_hash(const LaunchCtx &lc, int bin, int G, const int *queue, int count, Csr A, Csr B, int *rc,
                            int qstride, int *bin_max, const int *dcount, unsigned long long *bin_sum)
{
    if (count <= 0) return cudaSuccess;
    // 2048/4096-slot tables: a warp per table leaves 24/12 warps per SM; a 128-thread CTA per
    // row keeps the SM full. (Sampling and retry launches keep the group kernel.)
    if (qstride == 1 && !bin_max && !dcount) {
        if (bin == SB_G2048) return launch_sym_block_t<11>(lc, queue, count, A, B, rc, 128);
        if (bin == SB_G4096) return launch_sym_block_t<12>(lc, queue, count, A, B, rc, 256);
    }
#define SYM_GROUP_CASE(BIN, L2T)                                                                   \
    case BIN:                                                                                      \
        return (G == 8 && L2T <= 10) ? launch_sym_group_t<8, L2T>(lc, queue, count, A, B, rc, qstride, bin_max, dcount, bin_sum)  \
                                     : launch_sym_group_t<32, L2T>(lc, queue, count, A, B, rc, qstride, bin_max, dcount, bin_sum);
    switch (bin) {
        SYM_GROUP_CASE(SB_G128, 7)
        SYM_GROUP_CASE(SB_G256, 8)
        SYM_GROUP_CASE(SB_G512, 9)
        SYM_GROUP_CASE(SB_G1024, 10)
        SYM_GROUP_CASE(SB_G2048, 11)
        SYM_GROUP_CASE(SB_G4096, 12)
    case SB_B8192: return launch_sym_block_t<13>(lc, queue, count, A, B, rc);
    case SB_B16384: return launch_sym_block_t<14>(lc, queue, count, A, B, rc);
    case SB_B32768: return launch_sym_block_t<15>(lc, queue, count, A, B, rc, 1024);
    default: return cudaErrorInvalidValue;
    }
#undef SYM_GROUP_CASE
}

int large_scratch_blocks(int sm_count, int n)
{
    // Two CTAs per SM whatever n is. (At n = 2 M columns the 296 bitmap + prefix sets are 150 MB and
    // spill out of the 126 MB L2, but fewer sets are slower still: 148 / 96 / 74 CTAs ran the large
    // bin of R-MAT scale 21 in 4.2 / 5.3 / 6.9 ms against 4.0 ms.)
    (void)n;
    return sm_count * 2;
}

cudaError_t launch_sym_large(const LaunchCtx &lc, const int *queue, int count, int n, Csr A, Csr B, int *rc,
                             unsigned *bitmap_scratch, int scratch_blocks)
{
    if (count <= 0) return cudaSuccess;
    const int nwords = large_nwords(n);
    int blocks = count < scratch_blocks ? count : scratch_blocks;
    ++*lc.launches;
    k_sym_large<<<blocks, 1024, 0, lc.stream>>>(queue, count, A.rowptr, A.col, B.rowptr, B.col, rc, bitmap_scratch, nwords);
    return cudaGetLastError();
}

}  // namespace bhb
