/*
 * spgemm_oracle.c -- CPU restatement of the reference's result definition for
 * C = A * B (CSR x CSR -> CSR).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke test
 * in __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product path (benchmark_spgemm_using_csr_b200) never calls it.
 *
 * PARITY STATUS: PINNED by outputs of the reference itself.  The reference's own check
 * (SpGEMM_cuda/ref_spgemm.h:65-127) calls cusp::multiply (CUSP v0.4.0, README.md:91), an
 * un-vendored dependency, and stores no expected outputs -- but its GPU path compiles:
 * oracle/build_ref.py builds bhsparse.h + bhsparse_cuda.h for sm_100a (oracle/_ref), and
 * tests/golden/ref_outputs.npz holds what it produced on a B200 for the 18 x 2 cases of
 * tests/ref_cases.py (tests/golden/make_ref_golden.py).  tests/test_oracle_pinned.py checks
 * this oracle against those vectors on the CPU; tests/test_reference_gpu.py re-runs the
 * reference live beside the library.  Also pinned by the hand-derived result of
 * test_small_spgemm (main.cu:149-246), cage4.mtx squared and scipy.sparse (tests/test_oracle.py).
 *
 * What it restates (all citations relative to /root/reference/SpGEMM_cuda):
 *  - the result contract checked by ref_spgemm::compData (ref_spgemm.h:79-126):
 *    nnzC, rowptrC exact, column indices exact after a per-row ascending sort
 *    (csr_sort_indices, ref_spgemm.h:37-62), values within a tolerance;
 *  - the structural semantics of the GPU path: one intermediate product for
 *    every (i,k) in A, (k,j) in B, duplicates summed, NO dropping of entries
 *    whose sum is numerically zero (there is no value test anywhere in
 *    bhsparse_cuda.h, e.g. :699-706, :1457-1473, :2043-2053);
 *  - the per-row upper bound nnzCt[i] = sum_{k in A_i} len(B_k)
 *    (compute_nnzCt_cudakernel, bhsparse_cuda.h:210-237);
 *  - the reference's bin assignment (bhsparse::statistics, bhsparse.h:373-407),
 *    exposed for tests of the device binning.
 *
 * Algorithm: row-wise Gustavson (SMMP) with a dense accumulator and a
 * touched-column list per thread, sum in A-row order then B-row order, per-row
 * ascending sort of the touched list.  All counters are int64.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

ORACLE_API int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORACLE_API void oracle_set_num_threads(int t)
{
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}

/* bhsparse_cuda.h:210-237 -- per-row count of intermediate products.
 * Returns the total (the reference's _nnzCt_full, bhsparse.h:381-405). */
ORACLE_API int64_t oracle_row_products(int m, const int32_t *rowptrA, const int32_t *colA,
                                       const int32_t *rowptrB, int64_t *row_products)
{
    int64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int i = 0; i < m; i++) {
        int64_t c = 0;
        for (int32_t p = rowptrA[i]; p < rowptrA[i + 1]; p++) {
            int32_t k = colA[p];
            c += (int64_t)(rowptrB[k + 1] - rowptrB[k]);
        }
        if (row_products) row_products[i] = c;
        total += c;
    }
    return total;
}

/* bhsparse.h:373-407 -- the reference's segment ("bin") id for a row with
 * `count` intermediate products. */
ORACLE_API int oracle_reference_bin(int64_t count)
{
    if (count <= 121) return (int)count;
    if (count <= 128) return 122;
    if (count <= 256) return 123;
    if (count <= 512) return 124;
    return 127;
}

/* ascending int32 sort: quicksort (median of 3) down to runs of 24, then one
 * insertion pass -- the touched-column lists are short and often nearly sorted */
static void sort_i32(int32_t *a, int64_t n)
{
    int64_t stack[128];
    int sp = 0;
    int64_t lo = 0, hi = n - 1;
    for (;;) {
        while (hi - lo > 24) {
            int64_t mid = lo + (hi - lo) / 2;
            int32_t t;
            if (a[mid] < a[lo]) { t = a[mid]; a[mid] = a[lo]; a[lo] = t; }
            if (a[hi] < a[lo]) { t = a[hi]; a[hi] = a[lo]; a[lo] = t; }
            if (a[hi] < a[mid]) { t = a[hi]; a[hi] = a[mid]; a[mid] = t; }
            int32_t pivot = a[mid];
            int64_t i = lo, j = hi;
            while (i <= j) {
                while (a[i] < pivot) i++;
                while (a[j] > pivot) j--;
                if (i <= j) { t = a[i]; a[i] = a[j]; a[j] = t; i++; j--; }
            }
            /* recurse on the smaller part first to bound the stack */
            if (j - lo < hi - i) {
                if (i < hi) { stack[sp++] = i; stack[sp++] = hi; }
                hi = j;
            } else {
                if (lo < j) { stack[sp++] = lo; stack[sp++] = j; }
                lo = i;
            }
        }
        if (sp == 0) break;
        hi = stack[--sp];
        lo = stack[--sp];
    }
    for (int64_t i = 1; i < n; i++) {
        int32_t v = a[i];
        int64_t j = i - 1;
        while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; j--; }
        a[j + 1] = v;
    }
}

/* Symbolic phase: rowptrC64[m+1] (exclusive scan of per-row structural counts).
 * Returns nnzC, or -1 on allocation failure. */
ORACLE_API int64_t oracle_spgemm_symbolic(int m, int n, const int32_t *rowptrA, const int32_t *colA,
                                          const int32_t *rowptrB, const int32_t *colB,
                                          int64_t *rowptrC64)
{
    int fail = 0;
    rowptrC64[0] = 0;
#pragma omp parallel
    {
        int32_t *mark = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
        if (!mark) {
#pragma omp atomic write
            fail = 1;
        } else {
            for (int j = 0; j < n; j++) mark[j] = -1;
#pragma omp for schedule(dynamic, 256)
            for (int i = 0; i < m; i++) {
                int64_t cnt = 0;
                for (int32_t p = rowptrA[i]; p < rowptrA[i + 1]; p++) {
                    int32_t k = colA[p];
                    for (int32_t q = rowptrB[k]; q < rowptrB[k + 1]; q++) {
                        int32_t j = colB[q];
                        if (mark[j] != i) {
                            mark[j] = i;
                            cnt++;
                        }
                    }
                }
                rowptrC64[i + 1] = cnt;
            }
            free(mark);
        }
    }
    if (fail) return -1;
    for (int i = 0; i < m; i++) rowptrC64[i + 1] += rowptrC64[i];
    return rowptrC64[m];
}

#define DEFINE_NUMERIC(NAME, VT)                                                                  \
    ORACLE_API int NAME(int m, int n, const int32_t *rowptrA, const int32_t *colA, const VT *valA, \
                        const int32_t *rowptrB, const int32_t *colB, const VT *valB,               \
                        const int64_t *rowptrC64, int32_t *colC, VT *valC)                         \
    {                                                                                              \
        int fail = 0;                                                                              \
        _Pragma("omp parallel")                                                                    \
        {                                                                                          \
            size_t nn = (size_t)(n > 0 ? n : 1);                                                   \
            int32_t *mark = (int32_t *)malloc(sizeof(int32_t) * nn);                               \
            VT *acc = (VT *)malloc(sizeof(VT) * nn);                                               \
            if (!mark || !acc) {                                                                   \
                _Pragma("omp atomic write") fail = 1;                                              \
            } else {                                                                               \
                for (int j = 0; j < n; j++) mark[j] = -1;                                          \
                _Pragma("omp for schedule(dynamic, 256)") for (int i = 0; i < m; i++)              \
                {                                                                                  \
                    int64_t base = rowptrC64[i];                                                   \
                    int64_t cnt = 0;                                                               \
                    for (int32_t p = rowptrA[i]; p < rowptrA[i + 1]; p++) {                        \
                        int32_t k = colA[p];                                                       \
                        VT a = valA[p];                                                            \
                        for (int32_t q = rowptrB[k]; q < rowptrB[k + 1]; q++) {                    \
                            int32_t j = colB[q];                                                   \
                            VT prod = a * valB[q];                                                 \
                            if (mark[j] != i) {                                                    \
                                mark[j] = i;                                                       \
                                acc[j] = prod;                                                     \
                                colC[base + cnt] = j;                                              \
                                cnt++;                                                             \
                            } else {                                                               \
                                acc[j] += prod;                                                    \
                            }                                                                      \
                        }                                                                          \
                    }                                                                              \
                    /* ref_spgemm.h:37-62: per-row ascending sort by column */                     \
                    sort_i32(colC + base, cnt);                     \
                    for (int64_t t = 0; t < cnt; t++) valC[base + t] = acc[colC[base + t]];        \
                }                                                                                  \
            }                                                                                      \
            free(mark);                                                                            \
            free(acc);                                                                             \
        }                                                                                          \
        return fail ? -1 : 0;                                                                      \
    }

DEFINE_NUMERIC(oracle_spgemm_numeric_f64, double)
DEFINE_NUMERIC(oracle_spgemm_numeric_f32, float)

/* ref_spgemm.h:37-62 -- csr_sort_indices: sort each row's (col,val) pairs by
 * column (used by the reference on .mtx inputs, main.cu:62-64). */
#define DEFINE_SORT(NAME, VT)                                                              \
    typedef struct {                                                                       \
        int32_t c;                                                                         \
        VT v;                                                                              \
    } pair_##VT;                                                                           \
    static int cmp_pair_##VT(const void *a, const void *b)                                 \
    {                                                                                      \
        int32_t x = ((const pair_##VT *)a)->c, y = ((const pair_##VT *)b)->c;              \
        return (x > y) - (x < y);                                                          \
    }                                                                                      \
    ORACLE_API int NAME(int rows, const int32_t *rowptr, int32_t *col, VT *val)            \
    {                                                                                      \
        int fail = 0;                                                                      \
        _Pragma("omp parallel for schedule(dynamic, 256)") for (int i = 0; i < rows; i++)  \
        {                                                                                  \
            int32_t s = rowptr[i], e = rowptr[i + 1];                                      \
            if (e - s < 2) continue;                                                       \
            pair_##VT *tmp = (pair_##VT *)malloc(sizeof(pair_##VT) * (size_t)(e - s));     \
            if (!tmp) {                                                                    \
                fail = 1;                                                                  \
                continue;                                                                  \
            }                                                                              \
            for (int32_t p = s; p < e; p++) {                                              \
                tmp[p - s].c = col[p];                                                     \
                tmp[p - s].v = val[p];                                                     \
            }                                                                              \
            qsort(tmp, (size_t)(e - s), sizeof(pair_##VT), cmp_pair_##VT);                 \
            for (int32_t p = s; p < e; p++) {                                              \
                col[p] = tmp[p - s].c;                                                     \
                val[p] = tmp[p - s].v;                                                     \
            }                                                                              \
            free(tmp);                                                                     \
        }                                                                                  \
        return fail ? -1 : 0;                                                              \
    }

DEFINE_SORT(oracle_csr_sort_indices_f64, double)
DEFINE_SORT(oracle_csr_sort_indices_f32, float)
