// stage_count.cu -- stage 1 (per-row upper bound + binning) and the row-pointer
// scan.  Replaces compute_nnzCt_cudakernel (SpGEMM_cuda/bhsparse_cuda.h:210-237),
// the HOST passes of bhsparse::statistics (bhsparse.h:365-481) and the HOST scan
// of bhsparse_cuda::create_C (bhsparse_cuda.h:2783-2811) with device kernels.
#include "common.cuh"

namespace bhb {

// ---------------------------------------------------------------------------
// k_row_products: G lanes per row of A.  prod[i] = sum_{k in A_i} len(B_k).
// The reference uses one thread per row (uncoalesced colA reads); here a group
// of G lanes reads G consecutive column indices per step and the rowptrB pair
// gathers of a warp are issued together.  Also builds the symbolic-bin
// histogram, the product total (int64) and settles rows with p <= 1.
// ---------------------------------------------------------------------------
// k_b_row_ranges: one 16-byte record per row of B: {start, length, first column, last column}
// ({.., 0, INT_MAX, -1} for empty rows).  k_row_products then needs ONE aligned 16-byte gather per
// entry of A for both the product count and the column span of the row of C (instead of
// rowptrB[k], rowptrB[k+1] and two columns).
// The kernel reads EVERY column of B (8 lanes per row): min / max are taken over the whole row and
// the precondition of the whole pipeline -- columns of a B row strictly ascending (sorted, no
// duplicates; the reference's merge kernels need the same, bhsparse_cuda.h:1730,1762) and inside
// [0, n) -- is checked on the way.  A violation raises ctr->bad_B and bhb200_spgemm returns
// BHB200_ERR_INVALID instead of corrupting shared memory (range kernels) or losing duplicate
// contributions (the hash kernels' non-atomic accumulate).
__global__ void __launch_bounds__(256) k_b_row_ranges(const int k, const int n, const int *__restrict__ rowptrB,
                                                      const int *__restrict__ colB, int4 *__restrict__ brange,
                                                      Counters *__restrict__ ctr)
{
    const int gl = threadIdx.x & 7;
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (r >= k) return;   // (whole 8-lane groups leave together)
    const unsigned gmask = group_mask<8>(threadIdx.x & 31);
    const int s = rowptrB[r], e = rowptrB[r + 1];
    int lo = 0x7fffffff, hi = -1;
    bool bad = false;
    for (int p = s + gl; p < e; p += 8) {
        const int c = colB[p];
        lo = min(lo, c);
        hi = max(hi, c);
        bad |= (p > s && colB[p - 1] >= c);
    }
#pragma unroll
    for (int d = 4; d > 0; d >>= 1) {
        lo = min(lo, __shfl_xor_sync(gmask, lo, d, 8));
        hi = max(hi, __shfl_xor_sync(gmask, hi, d, 8));
    }
    bad |= e < s || (e > s && (lo < 0 || hi >= n));
    if (bad) ctr->bad_B = 1;
    if (gl == 0) brange[r] = (e > s) ? make_int4(s, e - s, lo, hi) : make_int4(s, 0, 0x7fffffff, -1);
}

cudaError_t launch_b_row_ranges(const LaunchCtx &lc, int k, int n, Csr B, int4 *brange, Counters *ctr)
{
    if (k <= 0) return cudaSuccess;
    ++*lc.launches;
    k_b_row_ranges<<<(int)(((long long)k * 8 + 255) / 256), 256, 0, lc.stream>>>(k, n, B.rowptr, B.col, brange, ctr);
    return cudaGetLastError();
}

template <int G>
__global__ void __launch_bounds__(256) k_row_products(const int m, const int *__restrict__ rowptrA,
                                                      const int *__restrict__ colA,
                                                      const int *__restrict__ rowptrB,
                                                      const int4 *__restrict__ brange, int *__restrict__ prod,
                                                      int *__restrict__ rc, int *__restrict__ rlo,
                                                      int *__restrict__ rspan, const int max_span, const int k,
                                                      Counters *__restrict__ ctr)
{
    __shared__ int s_hist[MAX_BINS];
    __shared__ unsigned long long s_hprod[MAX_BINS];
    __shared__ unsigned long long s_total;
    __shared__ int s_max, s_ovf;
    if (threadIdx.x < MAX_BINS) {
        s_hist[threadIdx.x] = 0;
        s_hprod[threadIdx.x] = 0ull;
    }
    if (threadIdx.x == 0) {
        s_total = 0ull;
        s_max = 0;
        s_ovf = 0;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int gl = threadIdx.x & (G - 1);
    const unsigned gmask = group_mask<G>(lane);
    const int groups_per_block = blockDim.x / G;
    const long long stride = (long long)gridDim.x * groups_per_block;
    unsigned long long my_total = 0ull;
    int my_max = 0;
    int h_bin = -1, h_cnt = 0;   // run-length aggregation of the histogram updates (rows of a thread mostly share a bin)
    unsigned long long h_prod = 0ull;

    // warp-uniform loops (maxima over the groups of the warp): sub-warp groups with their
    // own trip counts would not reconverge
    (void)gmask;
    const int gib = threadIdx.x / G;
    for (long long r0 = (long long)blockIdx.x * groups_per_block + (gib & ~(32 / G - 1)); r0 < m; r0 += stride) {
        const long long r = r0 + (gib & (32 / G - 1));
        const bool active = r < m;
        const int row = active ? (int)r : 0;
        const int a0 = active ? rowptrA[row] : 0;
        const int na = active ? rowptrA[row + 1] - a0 : 0;
        const int max_na = __reduce_max_sync(FULL, na);
        long long s = 0;
        int lo = 0x7fffffff, hi = -1;
        for (int j0 = 0; j0 < max_na; j0 += G) {
            const int j = j0 + gl;
            if (j < na) {
                const int ck = colA[a0 + j];
                if ((unsigned)ck >= (unsigned)k) {   // column of A outside B's rows: reported, not followed
                    ctr->bad_A = 1;
                    continue;
                }
                const int4 br = __ldg(brange + ck);
                s += (long long)br.y;
                lo = min(lo, br.z);
                hi = max(hi, br.w);
            }
        }
#pragma unroll
        for (int d = G >> 1; d > 0; d >>= 1) {
            s += __shfl_xor_sync(FULL, s, d, G);
            lo = min(lo, __shfl_xor_sync(FULL, lo, d, G));
            hi = max(hi, __shfl_xor_sync(FULL, hi, d, G));
        }
        if (gl == 0 && active) {
            int span = (hi >= lo) ? hi - lo + 1 : 0;
            if (span > max_span) span = 0x7fffffff;   // hash path
            rlo[row] = lo;
            rspan[row] = span;
            int p;
            if (s > 0x7fffffffLL) {
                p = 0x7fffffff;
                s_ovf = 1;
            } else {
                p = (int)s;
            }
            prod[row] = p;
            if (p <= 1) rc[row] = p;
            const int sbin = sym_bin_of(p, span);
            if (sbin != h_bin) {
                if (h_cnt) {
                    atomicAdd(&s_hist[h_bin], h_cnt);
                    atomicAdd(&s_hprod[h_bin], h_prod);
                }
                h_bin = sbin;
                h_cnt = 0;
                h_prod = 0ull;
            }
            ++h_cnt;
            h_prod += (unsigned long long)p;
            my_total += (unsigned long long)s;
            my_max = max(my_max, p);
        }
    }
    if (h_cnt) {
        atomicAdd(&s_hist[h_bin], h_cnt);
        atomicAdd(&s_hprod[h_bin], h_prod);
    }
    // block reduction of totals
    my_total = warp_sum(my_total);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) my_max = max(my_max, __shfl_xor_sync(FULL, my_max, d));
    if (lane == 0) {
        atomicAdd(&s_total, my_total);
        atomicMax(&s_max, my_max);
    }
    __syncthreads();
    if (threadIdx.x < MAX_BINS && s_hist[threadIdx.x]) {
        atomicAdd(&ctr->sym_bin[threadIdx.x], s_hist[threadIdx.x]);
        atomicAdd(&ctr->sym_bin_products[threadIdx.x], s_hprod[threadIdx.x]);
    }
    if (threadIdx.x == 0) {
        if (s_total) atomicAdd(&ctr->products, s_total);
        atomicMax(&ctr->max_row_products, s_max);
        if (s_ovf) ctr->row_overflow = 1;
    }
}

cudaError_t launch_row_products(const LaunchCtx &lc, int m, int k, int nnzA, Csr A, Csr B, const int4 *brange, int *prod,
                                int *rc, int *rlo, int *rspan, Counters *ctr)
{
    if (m <= 0) return cudaSuccess;
    const double avg = (double)nnzA / (double)m;
    const int threads = 256;
    // narrow groups on purpose: the kernel is a chain of dependent gathers (rowptrA -> colA -> B row
    // record), so several rows per warp in flight matter more than lane utilisation
    int G = avg <= 4.0 ? 2 : avg <= 12.0 ? 4 : avg <= 48.0 ? 8 : avg <= 160.0 ? 16 : 32;
    const long long rows_per_block = threads / G;
    long long blocks = (m + rows_per_block - 1) / rows_per_block;
    const long long cap = (long long)lc.sm_count * 64;
    if (blocks > cap) blocks = cap;
    ++*lc.launches;
    switch (G) {
    case 2: k_row_products<2><<<(int)blocks, threads, 0, lc.stream>>>(m, A.rowptr, A.col, B.rowptr, brange, prod, rc, rlo, rspan, lc.max_span, k, ctr); break;
    case 4: k_row_products<4><<<(int)blocks, threads, 0, lc.stream>>>(m, A.rowptr, A.col, B.rowptr, brange, prod, rc, rlo, rspan, lc.max_span, k, ctr); break;
    case 8: k_row_products<8><<<(int)blocks, threads, 0, lc.stream>>>(m, A.rowptr, A.col, B.rowptr, brange, prod, rc, rlo, rspan, lc.max_span, k, ctr); break;
    case 16: k_row_products<16><<<(int)blocks, threads, 0, lc.stream>>>(m, A.rowptr, A.col, B.rowptr, brange, prod, rc, rlo, rspan, lc.max_span, k, ctr); break;
    default: k_row_products<32><<<(int)blocks, threads, 0, lc.stream>>>(m, A.rowptr, A.col, B.rowptr, brange, prod, rc, rlo, rspan, lc.max_span, k, ctr); break;
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// k_bin_scatter: rows -> queue, bins contiguous (offsets come from the bin
// histogram; the host only turns 13 counts into 13 offsets).  One block handles
// a contiguous chunk of rows and claims one range per bin with a single global
// atomic, so rows stay in ascending order per bin up to block granularity --
// neighbouring rows (which share rows of B) are processed by neighbouring warps.
// Replaces the second host pass of bhsparse::statistics (bhsparse.h:434-478).
// ---------------------------------------------------------------------------
constexpr int SCATTER_ITEMS = 4;
template <bool NUMERIC>
__global__ void __launch_bounds__(256) k_bin_scatter(const int m, const int *__restrict__ prod,
                                                     const int *__restrict__ rc,
                                                     const int *__restrict__ rspan, const unsigned spec_mask,
                                                     const long long *__restrict__ ct_off, const BinOffsets offs,
                                                     int *__restrict__ cursor, int *__restrict__ queue)
{
    __shared__ int s_cnt[MAX_BINS];
    __shared__ int s_base[MAX_BINS];
    if (threadIdx.x < MAX_BINS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const long long chunk0 = (long long)blockIdx.x * blockDim.x * SCATTER_ITEMS;
    int bin[SCATTER_ITEMS], rank[SCATTER_ITEMS];
#pragma unroll
    for (int it = 0; it < SCATTER_ITEMS; ++it) {
        const long long row = chunk0 + (long long)it * blockDim.x + threadIdx.x;
        bin[it] = -1;
        if (row < m) {
            const int p = prod[row];
            const int span = rspan[row];
            const int sb = sym_bin_of(p, span);
            if (!NUMERIC)
                bin[it] = sb;
            else if (((spec_mask >> sb) & 1u) && ct_off[row] >= 0)
                bin[it] = NB_COPY;        // computed by the direct mode already
            else
                bin[it] = num_bin_of(p, rc[row], span);
            rank[it] = atomicAdd(&s_cnt[bin[it]], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x < MAX_BINS) {
        const int c = s_cnt[threadIdx.x];
        s_base[threadIdx.x] = c ? offs.off[threadIdx.x] + atomicAdd(&cursor[threadIdx.x], c) : 0;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < SCATTER_ITEMS; ++it) {
        if (bin[it] >= 0) {
            const long long row = chunk0 + (long long)it * blockDim.x + threadIdx.x;
            queue[s_base[bin[it]] + rank[it]] = (int)row;
        }
    }
}

cudaError_t launch_bin_scatter(const LaunchCtx &lc, bool numeric, int m, const int *prod, const int *rc,
                               const int *rspan, unsigned spec_mask, const long long *ct_off, const BinOffsets &offs,
                               Counters *ctr, int *queue)
{
    if (m <= 0) return cudaSuccess;
    const int threads = 256;
    const long long per_block = (long long)threads * SCATTER_ITEMS;
    const int blocks = (int)((m + per_block - 1) / per_block);
    ++*lc.launches;
    if (numeric)
        k_bin_scatter<true><<<blocks, threads, 0, lc.stream>>>(m, prod, rc, rspan, spec_mask, ct_off, offs, ctr->num_cursor, queue);
    else
        k_bin_scatter<false><<<blocks, threads, 0, lc.stream>>>(m, prod, rc, rspan, 0u, nullptr, offs, ctr->sym_cursor, queue);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Row-pointer scan: rc[i] = nnz(C_i)  ->  rowoff64[0..m], rowptr32[0..m]
// (three small kernels: chunk sums + numeric-bin histogram, scan of the chunk
// sums, chunk-local scan + offset).  Replaces the D2H / host loop / H2D of
// create_C (bhsparse_cuda.h:2787-2808).  Totals are int64; rowptr32 saturates.
// ---------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

size_t scan_blocksum_count(int m)
{
    return (size_t)((m + SCAN_CHUNK - 1) / SCAN_CHUNK) + 1;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const int m, const int *__restrict__ rowptrA,
                                                              const int *__restrict__ prod,
                                                              const int *__restrict__ rc,
                                                              const int *__restrict__ rspan,
                                                              const unsigned spec_mask,
                                                              const long long *__restrict__ ct_off,
                                                              long long *__restrict__ blocksums,
                                                              Counters *__restrict__ ctr)
{
    __shared__ int s_hist[MAX_BINS];
    __shared__ unsigned long long s_work[3][MAX_BINS];   // products, nnz(C), nnz(A) per numeric bin
    __shared__ long long s_red[33];
    if (threadIdx.x < MAX_BINS) {
        s_hist[threadIdx.x] = 0;
        s_work[0][threadIdx.x] = 0ull;
        s_work[1][threadIdx.x] = 0ull;
        s_work[2][threadIdx.x] = 0ull;
    }
    __syncthreads();
    const long long base = (long long)blockIdx.x * SCAN_CHUNK;
    long long s = 0;
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; ++it) {
        const long long i = base + (long long)it * SCAN_THREADS + threadIdx.x;
        int b = -1;
        unsigned long long wp = 0ull, wc = 0ull, wa = 0ull;
        if (i < m) s += rc[i];
        if (i < m && rspan) {   // (rspan == nullptr: row pointers only, no numeric bins -- pattern mode)
            const int c = rc[i];
            const int p = prod[i];
            const int span = rspan[i];
            const int sb = sym_bin_of(p, span);
            b = (((spec_mask >> sb) & 1u) && ct_off[i] >= 0) ? (int)NB_COPY : num_bin_of(p, c, span);
            wp = (unsigned long long)p;
            wc = (unsigned long long)c;
            wa = (unsigned long long)(rowptrA[i + 1] - rowptrA[i]);
        }
        // neighbouring rows nearly always share a bin: one set of shared-memory atomics per warp
        // instead of per row (the per-row version spent 0.3 ms on same-address atomics for 2 M rows)
        int uniform;
        __match_all_sync(FULL, b, &uniform);
        if (uniform) {
            if (b >= 0) {
                wp = warp_sum(wp);
                wc = warp_sum(wc);
                wa = warp_sum(wa);
                if ((threadIdx.x & 31) == 0) {
                    atomicAdd(&s_hist[b], 32);
                    atomicAdd(&s_work[0][b], wp);
                    atomicAdd(&s_work[1][b], wc);
                    atomicAdd(&s_work[2][b], wa);
                }
            }
        } else if (b >= 0) {
            atomicAdd(&s_hist[b], 1);
            atomicAdd(&s_work[0][b], wp);
            atomicAdd(&s_work[1][b], wc);
            atomicAdd(&s_work[2][b], wa);
        }
    }
    const long long tot = block_sum(s, s_red);
    if (threadIdx.x == 0) blocksums[blockIdx.x] = tot;
    if (threadIdx.x < MAX_BINS && s_hist[threadIdx.x]) {
        atomicAdd(&ctr->num_bin[threadIdx.x], s_hist[threadIdx.x]);
        atomicAdd(&ctr->num_bin_products[threadIdx.x], s_work[0][threadIdx.x]);
        atomicAdd(&ctr->num_bin_nnzc[threadIdx.x], s_work[1][threadIdx.x]);
        atomicAdd(&ctr->num_bin_nnza[threadIdx.x], s_work[2][threadIdx.x]);
    }
}

// one block: exclusive scan of blocksums[0..nb) in place; total -> ctr->nnzC and blocksums[nb]
__global__ void __launch_bounds__(1024) k_scan_blocks(const int nb, long long *__restrict__ blocksums,
                                                      Counters *__restrict__ ctr)
{
    __shared__ long long s_warp[32];
    __shared__ long long s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const long long v = (i < nb) ? blocksums[i] : 0;
        long long x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long y = __shfl_up_sync(FULL, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            long long w = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long y = __shfl_up_sync(FULL, w, d);
                if (lane >= d) w += y;
            }
            s_warp[lane] = w;   // inclusive over warps
        }
        __syncthreads();
        const long long carry = s_carry;
        const long long warp_off = warp ? s_warp[warp - 1] : 0;
        if (i < nb) blocksums[i] = carry + warp_off + x - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        blocksums[nb] = s_carry;
        ctr->nnzC = (unsigned long long)s_carry;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_write(const int m, const int *__restrict__ rc,
                                                             const long long *__restrict__ blocksums,
                                                             const int nb, int64_t *__restrict__ rowoff64,
                                                             int *__restrict__ rowptr32)
{
    __shared__ long long s_warp[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // each thread owns SCAN_ITEMS consecutive rows
    const long long first = (long long)blockIdx.x * SCAN_CHUNK + (long long)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    long long s = 0;
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; ++it) {
        const long long i = first + it;
        v[it] = (i < m) ? rc[i] : 0;
        s += v[it];
    }
    long long x = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long y = __shfl_up_sync(FULL, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    long long off = blocksums[blockIdx.x] + (x - s);
    for (int w = 0; w < warp; ++w) off += s_warp[w];
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; ++it) {
        const long long i = first + it;
        if (i < m) {
            rowoff64[i] = off;
            rowptr32[i] = off > 0x7fffffffLL ? 0x7fffffff : (int)off;
        }
        off += v[it];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const long long tot = blocksums[nb];
        rowoff64[m] = tot;
        rowptr32[m] = tot > 0x7fffffffLL ? 0x7fffffff : (int)tot;
    }
}

cudaError_t launch_scan(const LaunchCtx &lc, int m, const int *rowptrA, const int *prod, const int *rc,
                        const int *rspan, unsigned spec_mask, const long long *ct_off, int64_t *rowoff64,
                        int *rowptr32, long long *blocksums, Counters *ctr)
{
    const int nb = (m + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (nb > 0) {
        ++*lc.launches;
        k_scan_reduce<<<nb, SCAN_THREADS, 0, lc.stream>>>(m, rowptrA, prod, rc, rspan, spec_mask, ct_off, blocksums, ctr);
    }
    ++*lc.launches;
    k_scan_blocks<<<1, 1024, 0, lc.stream>>>(nb, blocksums, ctr);
    ++*lc.launches;
    k_scan_write<<<nb > 0 ? nb : 1, SCAN_THREADS, 0, lc.stream>>>(m, rc, blocksums, nb, rowoff64, rowptr32);
    return cudaGetLastError();
}

}  // namespace bhb
