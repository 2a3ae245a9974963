// stage_pattern.cu -- kernels of the diagonal-pattern mode (see stage_pattern.cuh).
#include "stage_pattern.cuh"

namespace bhb {

// ---------------------------------------------------------------------------------------------
// k_offset_set: the set of (column - row) over all entries of a CSR matrix, in a small global
// hash set.  A hit is a plain load (the few distinct keys sit in L1), only a new offset costs an
// atomic.  More than PAT_MAX_OFFS distinct offsets raise `overflow`; every row checks the flag
// first, so an unstructured matrix costs a few microseconds.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pat_set_insert(PatSet *set, const int d)
{
    if (d == PAT_EMPTY) {   // (an offset no int32-indexed matrix of sane size has)
        set->overflow = 1;
        return;
    }
    unsigned h = ((unsigned)d * 2654435761u) >> (32 - 9);
    static_assert(PAT_SET_SLOTS == 512, "hash shift");
    volatile int *vs = set->slot;
    for (int probe = 0; probe < PAT_SET_SLOTS; ++probe) {
        int v = vs[h];
        if (v == d) return;
        if (v == PAT_EMPTY) {
            const int old = atomicCAS(&set->slot[h], PAT_EMPTY, d);
            if (old == PAT_EMPTY) {
                if (atomicAdd(&set->count, 1) + 1 > PAT_MAX_OFFS) set->overflow = 1;
                return;
            }
            if (old == d) return;
        }
        h = (h + 1) & (PAT_SET_SLOTS - 1);
    }
    set->overflow = 1;
}

__global__ void __launch_bounds__(256) k_offset_set(const int rows, const int *__restrict__ rowptr,
                                                    const int *__restrict__ col, PatSet *__restrict__ set)
{
    const int gl = threadIdx.x & 7;
    const long long stride = (long long)gridDim.x * (blockDim.x >> 3);
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < rows; r += stride) {
        if (*(volatile int *)&set->overflow) return;
        const int s = rowptr[r], e = rowptr[r + 1];
        for (int p = s + gl; p < e; p += 8) pat_set_insert(set, col[p] - (int)r);
    }
}

cudaError_t launch_offset_set(const LaunchCtx &lc, int rows, const int *rowptr, const int *col, PatSet *set)
{
    if (rows <= 0) return cudaSuccess;
    long long blocks = ((long long)rows * 8 + 255) / 256;
    const long long cap = (long long)lc.sm_count * 8;
    if (blocks > cap) blocks = cap;
    ++*lc.launches;
    k_offset_set<<<(int)blocks, 256, 0, lc.stream>>>(rows, rowptr, col, set);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// k_pat_codes: code[p] = index of (col[p] - row) in the sorted offset list (binary search in
// shared memory); rowmask[r] = OR of 1 << code over the row (B only).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pat_codes(const int rows, const int *__restrict__ rowptr,
                                                   const int *__restrict__ col, const int *__restrict__ offs,
                                                   const int noffs, unsigned char *__restrict__ code,
                                                   unsigned long long *__restrict__ rowmask)
{
    __shared__ int s_offs[PAT_MAX_OFFS];
    if (threadIdx.x < PAT_MAX_OFFS) s_offs[threadIdx.x] = threadIdx.x < noffs ? offs[threadIdx.x] : 0x7fffffff;
    __syncthreads();
    const int gl = threadIdx.x & 7;
    const long long stride = (long long)gridDim.x * (blockDim.x >> 3);
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < rows; r += stride) {
        const int s = rowptr[r], e = rowptr[r + 1];
        unsigned long long mask = 0ull;
        for (int p = s + gl; p < e; p += 8) {
            const int d = col[p] - (int)r;
            int lo = 0;   // first index with s_offs[idx] >= d (d is in the list)
#pragma unroll
            for (int step = PAT_MAX_OFFS / 2; step > 0; step >>= 1)
                if (s_offs[lo + step - 1] < d) lo += step;
            code[p] = (unsigned char)lo;
            mask |= 1ull << lo;
        }
        if (rowmask) {
            mask |= __shfl_xor_sync(group_mask<8>(threadIdx.x & 31), mask, 1, 8);
            mask |= __shfl_xor_sync(group_mask<8>(threadIdx.x & 31), mask, 2, 8);
            mask |= __shfl_xor_sync(group_mask<8>(threadIdx.x & 31), mask, 4, 8);
            if (gl == 0) rowmask[r] = mask;
        }
    }
}

cudaError_t launch_pat_codes(const LaunchCtx &lc, int rows, const int *rowptr, const int *col, const int *offs, int noffs,
                             unsigned char *code, unsigned long long *rowmask)
{
    if (rows <= 0) return cudaSuccess;
    long long blocks = ((long long)rows * 8 + 255) / 256;
    const long long cap = (long long)lc.sm_count * 16;
    if (blocks > cap) blocks = cap;
    ++*lc.launches;
    k_pat_codes<<<(int)blocks, 256, 0, lc.stream>>>(rows, rowptr, col, offs, noffs, code, rowmask);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// k_pat_symbolic: 8 lanes per row of A.  The output-offset mask of row i is the OR over its
// entries (ja, k) of the image of B row k's offset mask under jb -> M[ja][jb]; for a B row that
// holds every offset of DB (all interior rows of a stencil) that image is the precomputed
// P[ja].  nnz(C_i) = popcount.  The mask is kept for the numeric kernel.
// ---------------------------------------------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(256) k_pat_symbolic(const int m, const int *__restrict__ rowptrA,
                                                      const int *__restrict__ colA,
                                                      const unsigned char *__restrict__ ta,
                                                      const unsigned long long *__restrict__ maskB, const PatTables t,
                                                      unsigned *__restrict__ outmask, int *__restrict__ rc)
{
    const int gl = threadIdx.x & 7;
    const unsigned gmask = group_mask<8>(threadIdx.x & 31);
    const long long stride = (long long)gridDim.x * (blockDim.x >> 3);
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < m; r += stride) {
        const int a0 = rowptrA[r], a1 = rowptrA[r + 1];
        unsigned mk[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) mk[w] = 0u;
        for (int j = a0 + gl; j < a1; j += 8) {
            const int ja = ta[j];
            unsigned long long mb = maskB[colA[j]];
            if (mb == t.fullB) {
#pragma unroll
                for (int w = 0; w < NW; ++w) mk[w] |= __ldg(t.pfull + ja * NW + w);
            } else {
                while (mb) {
                    const int jb = __ffsll((long long)mb) - 1;
                    mb &= mb - 1;
                    const int o = t.mlog[ja * t.nDB + jb];
#pragma unroll
                    for (int w = 0; w < NW; ++w) mk[w] |= ((o >> 5) == w) ? (1u << (o & 31)) : 0u;
                }
            }
        }
        int cnt = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            mk[w] |= __shfl_xor_sync(gmask, mk[w], 1, 8);
            mk[w] |= __shfl_xor_sync(gmask, mk[w], 2, 8);
            mk[w] |= __shfl_xor_sync(gmask, mk[w], 4, 8);
            cnt += __popc(mk[w]);
            if (gl == (w & 7)) outmask[r * NW + w] = mk[w];
        }
        if (gl == 0) rc[r] = cnt;
    }
}

cudaError_t launch_pat_symbolic(const LaunchCtx &lc, int m, Csr A, const unsigned char *ta, const unsigned long long *maskB,
                                PatTables t, unsigned *outmask, int *rc)
{
    if (m <= 0) return cudaSuccess;
    long long blocks = ((long long)m * 8 + 255) / 256;
    const long long cap = (long long)lc.sm_count * 16;
    if (blocks > cap) blocks = cap;
    ++*lc.launches;
    switch (t.nw) {
    case 1: k_pat_symbolic<1><<<(int)blocks, 256, 0, lc.stream>>>(m, A.rowptr, A.col, ta, maskB, t, outmask, rc); break;
    case 2: k_pat_symbolic<2><<<(int)blocks, 256, 0, lc.stream>>>(m, A.rowptr, A.col, ta, maskB, t, outmask, rc); break;
    case 4: k_pat_symbolic<4><<<(int)blocks, 256, 0, lc.stream>>>(m, A.rowptr, A.col, ta, maskB, t, outmask, rc); break;
    case 8: k_pat_symbolic<8><<<(int)blocks, 256, 0, lc.stream>>>(m, A.rowptr, A.col, ta, maskB, t, outmask, rc); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// k_pat_numeric: a group of G lanes owns a row of C (G = 32 unless B rows have <= 8 / <= 16
// entries).  The row's A entries are staged as 16-byte records {B row start, length | ja << 16,
// value} in shared memory and read back as one broadcast LDS.128 per B row (four shuffles in the
// hash kernels).  Per product: one byte of B's offset codes, one value of B, one byte of the
// position table, LDS/FMA/STS on the dense accumulator.  The accumulator layout (PatTables::pos)
// is chosen on the host so that the products of one B row fall into distinct banks.
// Rows are taken in natural order: neighbouring rows read the same rows of B.
// All loops are warp-uniform (see k_num_group).
// ---------------------------------------------------------------------------------------------
template <typename VT>
struct PatRec;
template <>
struct PatRec<double> {
    static __device__ __forceinline__ int4 pack(int bs, int lj, double v)
    {
        return make_int4(bs, lj, __double2loint(v), __double2hiint(v));
    }
    static __device__ __forceinline__ double val(const int4 &r) { return __hiloint2double(r.w, r.z); }
};
template <>
struct PatRec<float> {
    static __device__ __forceinline__ int4 pack(int bs, int lj, float v) { return make_int4(bs, lj, __float_as_int(v), 0); }
    static __device__ __forceinline__ float val(const int4 &r) { return __int_as_float(r.z); }
};

template <typename VT, int G>
__global__ void __launch_bounds__(256, 8)
k_pat_numeric(const int m, const int *__restrict__ rowptrA, const int *__restrict__ colA, const VT *__restrict__ valA,
              const unsigned char *__restrict__ ta, const int *__restrict__ rowptrB, const unsigned char *__restrict__ tb,
              const VT *__restrict__ valB, const PatTables t, const unsigned *__restrict__ outmask,
              const int64_t *__restrict__ rowoff, int *__restrict__ colC, VT *__restrict__ valC)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // CTA-wide tables: dcol int[nD] | mphys u8[nDA*nDB] | pos u8[nD]   (padded to 16 bytes)
    int *s_dcol = reinterpret_cast<int *>(smem_raw);
    unsigned char *s_mphys = smem_raw + (size_t)t.nD * 4;
    unsigned char *s_pos = s_mphys + (size_t)t.nDA * t.nDB;
    const size_t tab_bytes = ((size_t)t.nD * 4 + (size_t)t.nDA * t.nDB + (size_t)t.nD + 15) & ~(size_t)15;
    for (int i = threadIdx.x; i < t.nD; i += blockDim.x) {
        s_dcol[i] = t.dcol[i];
        s_pos[i] = t.pos[i];
    }
    for (int i = threadIdx.x; i < t.nDA * t.nDB; i += blockDim.x) s_mphys[i] = t.mphys[i];
    __syncthreads();

    constexpr int GPW = 32 / G;   // groups per warp
    const int lane = threadIdx.x & 31;
    const int gl = threadIdx.x & (G - 1);
    const int gib = threadIdx.x / G;
    const int groups_per_block = blockDim.x / G;
    const int gshift = lane & ~(G - 1);
    const unsigned gbits = (G == 32) ? FULL : ((1u << (G & 31)) - 1u);
    const int acc_len = t.acc_len;
    const size_t per_group = (size_t)acc_len * sizeof(VT) + (size_t)G * 16;
    unsigned char *mine = smem_raw + tab_bytes + (size_t)gib * per_group;
    VT *acc = reinterpret_cast<VT *>(mine);
    int4 *rec = reinterpret_cast<int4 *>(mine + (size_t)acc_len * sizeof(VT));
    const int nDB = t.nDB, nD = t.nD, nw = t.nw;
    (void)gshift;
    (void)gbits;

    for (long long q0 = (long long)blockIdx.x * groups_per_block + (gib & ~(GPW - 1)); q0 < m;
         q0 += (long long)gridDim.x * groups_per_block) {
        const long long q = q0 + (gib & (GPW - 1));
        const bool active = q < m;
        const int row = active ? (int)q : 0;
        for (int s = gl; s < acc_len; s += G) acc[s] = VT(0);
        const int a0 = active ? rowptrA[row] : 0;
        const int na = active ? rowptrA[row + 1] - a0 : 0;
        const int max_na = (G == 32) ? na : __reduce_max_sync(FULL, na);
        for (int base = 0; base < max_na; base += G) {
            const int j = base + gl;
            if (j < na) {
                const int k = colA[a0 + j];
                const int bs = rowptrB[k];
                const int len = rowptrB[k + 1] - bs;
                rec[gl] = PatRec<VT>::pack(bs, len | ((int)ta[a0 + j] << 16), valA[a0 + j]);
            }
            __syncwarp();
            const int cnt = min(G, max_na - base);   // warp-uniform
            const int mycnt = na - base;              // this group's entries in the chunk (may be <= 0)
            for (int tt = 0; tt < cnt; ++tt) {
                const int4 r = rec[tt];
                const int len = (tt < mycnt) ? (r.y & 0xffff) : 0;
                const VT av = PatRec<VT>::val(r);
                const unsigned char *mrow = s_mphys + (r.y >> 16) * nDB;
                const int max_len = (G == 32) ? len : __reduce_max_sync(FULL, len);
                for (int off0 = 0; off0 < max_len; off0 += G) {
                    const int off = off0 + gl;
                    if (off < len) {
                        const int p = mrow[tb[r.x + off]];
                        acc[p] = fma(av, valB[r.x + off], acc[p]);
                    }
                }
                __syncwarp();   // the next B row may hit the same accumulators from other lanes
            }
        }
        // ---- emit: output index o (ascending offset = ascending column) -> rank by popcount ----
        const int64_t o0 = active ? rowoff[row] : 0;
        const unsigned *mrow = outmask + (size_t)row * nw;
        int rank0 = 0;
        for (int ob = 0; ob < nD; ob += G) {
            const unsigned word = active ? __ldg(mrow + (ob >> 5)) : 0u;
            const unsigned bits = (G == 32) ? word : ((word >> (ob & 31)) & gbits);
            const int o = ob + gl;
            if ((bits >> gl) & 1u) {
                const int rnk = rank0 + __popc(bits & ((1u << gl) - 1u));
                colC[o0 + rnk] = row + s_dcol[o];
                valC[o0 + rnk] = acc[s_pos[o]];
            }
            rank0 += __popc(bits);
        }
        __syncwarp();
    }
}

template <typename VT, int G>
static cudaError_t launch_pat_numeric_t(const LaunchCtx &lc, int m, Csr A, Csr B, const unsigned char *ta,
                                        const unsigned char *tb, const PatTables &t, const unsigned *outmask,
                                        const int64_t *rowoff, int *colC, VT *valC)
{
    const int threads = 256;
    const int groups = threads / G;
    const size_t tab_bytes = ((size_t)t.nD * 4 + (size_t)t.nDA * t.nDB + (size_t)t.nD + 15) & ~(size_t)15;
    const size_t smem = tab_bytes + (size_t)groups * ((size_t)t.acc_len * sizeof(VT) + (size_t)G * 16);
    auto kern = k_pat_numeric<VT, G>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    long long blocks = ((long long)m + groups - 1) / groups;
    const long long cap = (long long)lc.sm_count * resident_blocks(kern, threads, smem);
    if (blocks > cap) blocks = cap;
    ++*lc.launches;
    kern<<<(int)blocks, threads, smem, lc.stream>>>(m, A.rowptr, A.col, (const VT *)A.val, ta, B.rowptr, tb,
                                                    (const VT *)B.val, t, outmask, rowoff, colC, valC);
    return cudaGetLastError();
}

cudaError_t launch_pat_numeric(const LaunchCtx &lc, int dtype, int m, Csr A, Csr B, const unsigned char *ta,
                               const unsigned char *tb, PatTables t, const unsigned *outmask, const int64_t *rowoff,
                               int *colC, void *valC)
{
    if (m <= 0) return cudaSuccess;
    const int G = t.nDB <= 8 ? 8 : t.nDB <= 16 ? 16 : 32;
    if (dtype == 1) {   // BHB200_DTYPE_F64
        double *v = (double *)valC;
        if (G == 8) return launch_pat_numeric_t<double, 8>(lc, m, A, B, ta, tb, t, outmask, rowoff, colC, v);
        if (G == 16) return launch_pat_numeric_t<double, 16>(lc, m, A, B, ta, tb, t, outmask, rowoff, colC, v);
        return launch_pat_numeric_t<double, 32>(lc, m, A, B, ta, tb, t, outmask, rowoff, colC, v);
    }
    float *v = (float *)valC;
    if (G == 8) return launch_pat_numeric_t<float, 8>(lc, m, A, B, ta, tb, t, outmask, rowoff, colC, v);
    if (G == 16) return launch_pat_numeric_t<float, 16>(lc, m, A, B, ta, tb, t, outmask, rowoff, colC, v);
    return launch_pat_numeric_t<float, 32>(lc, m, A, B, ta, tb, t, outmask, rowoff, colC, v);
}

}  // namespace bhb
