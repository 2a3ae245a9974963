// stage_range.cu -- symbolic range kernel (see stage_range.cuh).
#include "stage_range_vec.cuh"

namespace bhb {

__global__ void __launch_bounds__(512)
k_sym_range(const int *__restrict__ queue, const int count, const int *__restrict__ rowptrA,
            const int *__restrict__ colA, const int *__restrict__ rowptrB, const int *__restrict__ colB,
            const int *__restrict__ rlo, const int nsum, const int vec, int *__restrict__ rc,
            unsigned long long *__restrict__ pool_cursor, const WordLists wl)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nwords = nsum * 32;
    unsigned char *mine = smem_raw + (size_t)warp * sym_range_warp_bytes(nsum);
    unsigned long long *bm64 = reinterpret_cast<unsigned long long *>(mine);   // [nwords] + 2 sink words
    unsigned *bm32 = reinterpret_cast<unsigned *>(mine);
    const int *colB_lane = colB + lane;
    for (int i = lane; i < nwords + 2; i += 32) bm64[i] = 0ull;
    __syncwarp();
    // this lane's words in the sweeps: [lane*nsum, (lane+1)*nsum), read as nsum/2 16-byte vectors
    // (nsum is even; a lane stride of nsum*8 bytes keeps the 128-bit loads bank-conflict free)
    uint4 *myvec = reinterpret_cast<uint4 *>(bm64 + (size_t)lane * nsum);
    const int nvec = nsum >> 1;

    for (int q = blockIdx.x * nwarps + warp; q < count; q += gridDim.x * nwarps) {
        const int row = queue[q];
        const int base = rlo[row] & ~63;
        if (vec)
            range_mark_row_vec(rowptrA[row], rowptrA[row + 1], base, base + ((nwords + 1) << 6), lane, colA, rowptrB, colB,
                               bm32);
        else
            range_mark_row(rowptrA[row], rowptrA[row + 1], base, base + ((nwords + 1) << 6), lane, colA, rowptrB,
                           colB_lane, bm32);
        __syncwarp();
        // ---- sweep 1: set bits and non-empty words of this lane's chunk ----
        int nbits = 0, nwz = 0;
        for (int j = 0; j < nvec; ++j) {
            const uint4 v = myvec[j];
            if (v.x | v.y | v.z | v.w) {
                nbits += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
                nwz += ((v.x | v.y) != 0u) + ((v.z | v.w) != 0u);
            }
        }
        const int incl = warp_incl_scan(nwz, lane);
        const int total_w = __shfl_sync(FULL, incl, 31);
        nbits = warp_sum(nbits);
        // ---- hand the words to the numeric pass (device pool, one atomic per row) ----
        long long off = 0;
        if (lane == 0) off = (long long)atomicAdd(pool_cursor, (unsigned long long)total_w);
        off = __shfl_sync(FULL, off, 0);
        const bool fits = (off + total_w <= wl.cap);
        // ---- sweep 2: write (index, bits) in ascending word order, clear the bitmap ----
        if (nwz) {
            long long pos = off + incl - nwz;
            const int w0 = lane * nsum;
            for (int j = 0; j < nvec; ++j) {
                const uint4 v = myvec[j];
                if (v.x | v.y | v.z | v.w) {
                    if (fits) {
                        if (v.x | v.y) {
                            wl.idx[pos] = (unsigned)(w0 + 2 * j);
                            wl.bits[pos] = ((unsigned long long)v.y << 32) | v.x;
                            ++pos;
                        }
                        if (v.z | v.w) {
                            wl.idx[pos] = (unsigned)(w0 + 2 * j + 1);
                            wl.bits[pos] = ((unsigned long long)v.w << 32) | v.z;
                            ++pos;
                        }
                    }
                    myvec[j] = make_uint4(0u, 0u, 0u, 0u);
                }
            }
        }
        if (lane == 0) {
            rc[row] = nbits;
            wl.off[row] = off;
            wl.cnt[row] = fits ? total_w : -1;
        }
        __syncwarp();
    }
}

cudaError_t launch_sym_range(const LaunchCtx &lc, int nsum, const int *queue, int count, Csr A, Csr B, const int *rlo,
                             int *rc, Counters *ctr, WordLists wl)
{
    if (count <= 0) return cudaSuccess;
    const size_t wb = sym_range_warp_bytes(nsum);
    int wpb, bps;
    range_launch_shape(wb, wpb, bps);
    const size_t smem = wb * wpb;
    cudaError_t e = cudaFuncSetAttribute(k_sym_range, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    bps = resident_blocks(k_sym_range, wpb * 32, smem);
    long long blocks = ((long long)count + wpb - 1) / wpb;
    const long long cap = (long long)lc.sm_count * bps;
    if (blocks > cap) blocks = cap;
    ++*lc.launches;
    k_sym_range<<<(int)blocks, wpb * 32, smem, lc.stream>>>(queue, count, A.rowptr, A.col, B.rowptr, B.col, rlo, nsum,
                                                            range_vec_aligned(B) ? 1 : 0, rc, &ctr->pool_cursor, wl);
    return cudaGetLastError();
}

}  // namespace bhb
