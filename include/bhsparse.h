/*
 * bhsparse.h -- drop-in replacement for the reference's `bhsparse` class
 * (weifengliu-ssslab/Benchmark_SpGEMM_using_CSR, SpGEMM_cuda/bhsparse.h:17-34):
 * same public methods, argument order (val, rowptr, colidx) and int error
 * convention, forwarding to the C-ABI of libbhsparse_b200 (bhsparse_b200.h).
 * Header-only and all-inline (the reference defines non-inline members in its
 * header, an ODR hazard: bhsparse.h:91).
 *
 *     #include "bhsparse.h"          // instead of the reference's header
 *     bhsparse *bh = new bhsparse();
 *     bh->initPlatform(platforms);   // platforms[BHSPARSE_CUDA] = true
 *     bh->initData(m, k, n, nnzA, valA, rowptrA, colA, nnzB, valB, rowptrB, colB, rowptrC);
 *     bh->warmup(); bh->spgemm();
 *     int nnzC = bh->get_nnzC();  bh->get_C(colC, valC);
 *     bh->free_mem(); bh->freePlatform();
 *
 * value_type follows the reference's typedef (common.h:31): double unless
 * BHSPARSE_VALUE_FLOAT is defined before including this header.
 */
#ifndef BHSPARSE_H
#define BHSPARSE_H

#include <cstdio>

#include "bhsparse_b200.h"

#define BHSPARSE_SUCCESS 0   /* common.h:26 */
#define NUM_PLATFORMS 9      /* common.h:33 */
#define NAIVE 0
#define BHSPARSE_CUDA 1      /* common.h:36 */
#define BHSPARSE_OPENCL 2

typedef int index_type;      /* common.h:30 */
#ifdef BHSPARSE_VALUE_FLOAT
typedef float value_type;
#else
typedef double value_type;   /* common.h:31 */
#endif

class bhsparse {
public:
    bhsparse() : _ctx(nullptr), _rowptrC(nullptr), _verbose(true) {}
    ~bhsparse()
    {
        if (_ctx) bhb200_destroy(_ctx);
    }
    bhsparse(const bhsparse &) = delete;
    bhsparse &operator=(const bhsparse &) = delete;

    /* The reference prints stage times and GFLOPS from spgemm() (bhsparse.h:286-336);
     * keep that for CLI users, switch it off for library use. */
    void set_verbose(bool v) { _verbose = v; }

    /* bhsparse.h:96-124.  Only the CUDA platform exists. */
    int initPlatform(bool *spgemm_platform)
    {
        if (!spgemm_platform || !spgemm_platform[BHSPARSE_CUDA]) return BHB200_ERR_INVALID;
        if (_ctx) return BHSPARSE_SUCCESS;
        int err = bhb200_create(&_ctx, 0);   /* device 0, like bhsparse_cuda.h:100-101 */
        if (err == BHSPARSE_SUCCESS && _verbose)
            printf("Device [0] %s. %d SMs.\n", bhb200_device_name(_ctx), bhb200_sm_count(_ctx));
        return err;
    }

    /* bhsparse.h:180-258.  csrRowPtrC (m+1 ints) is caller-owned and written by get_C. */
    int initData(int m, int k, int n, int nnzA, value_type *csrValA, index_type *csrRowPtrA, index_type *csrColIndA,
                 int nnzB, value_type *csrValB, index_type *csrRowPtrB, index_type *csrColIndB,
                 index_type *csrRowPtrC)
    {
        if (!_ctx) return BHB200_ERR_INVALID;
        _rowptrC = csrRowPtrC;
#ifdef BHSPARSE_VALUE_FLOAT
        return bhb200_init_data_f32(_ctx, m, k, n, nnzA, csrValA, csrRowPtrA, csrColIndA, nnzB, csrValB, csrRowPtrB,
                                    csrColIndB);
#else
        return bhb200_init_data_f64(_ctx, m, k, n, nnzA, csrValA, csrRowPtrA, csrColIndA, nnzB, csrValB, csrRowPtrB,
                                    csrColIndB);
#endif
    }

    int warmup() { return _ctx ? bhb200_warmup(_ctx) : BHB200_ERR_INVALID; }   /* bhsparse.h:341-363 */

    /* bhsparse.h:260-339 */
    int spgemm()
    {
        if (!_ctx) return BHB200_ERR_INVALID;
        int err = bhb200_spgemm(_ctx);
        if (err != BHSPARSE_SUCCESS) {
            printf("spgemm error = %d (%s)\n", err, bhb200_last_error(_ctx));
            return err;
        }
        /* create_C leaves the row pointers of C in the caller's array at the end of spgemm()
         * (bhsparse_cuda.h:2787-2808); get_C writes them again (:3016).  Results beyond
         * INT32_MAX entries have no int32 row pointers (get_nnzC() == -1). */
        if (_rowptrC && bhb200_get_nnzC(_ctx) <= 0x7fffffffLL) {
#ifdef BHSPARSE_VALUE_FLOAT
            err = bhb200_get_C_f32(_ctx, _rowptrC, nullptr, nullptr);
#else
            err = bhb200_get_C_f64(_ctx, _rowptrC, nullptr, nullptr);
#endif
            if (err != BHSPARSE_SUCCESS) return err;
        }
        if (_verbose) {
            bhb200_stats st;
            bhb200_get_stats(_ctx, &st);
            printf("STAGE 1 time: %g ms.\nSTAGE 2 time: %g ms.\nSTAGE 3 time: %g ms.\nSTAGE 4 time: %g ms.\n",
                   st.ms_count, st.ms_symbolic, st.ms_scan, st.ms_numeric);
            printf("[ CUDA ] SpGEMM time: %g ms. Gflops = %g\n", st.ms_total,
                   2.0 * (double)st.products / (st.ms_total * 1.0e+6));
        }
        return err;
    }

    int get_nnzC()   /* bhsparse_cuda.h:3006-3009; -1 if it does not fit an int */
    {
        long long c = _ctx ? bhb200_get_nnzC(_ctx) : -1;
        return (c < 0 || c > 0x7fffffffLL) ? -1 : (int)c;
    }

    /* Additions (no reference counterpart): new values, same patterns -> values of C only. */
    int update_values(const value_type *csrValA, const value_type *csrValB)
    {
        if (!_ctx) return BHB200_ERR_INVALID;
#ifdef BHSPARSE_VALUE_FLOAT
        return bhb200_update_values_f32(_ctx, csrValA, csrValB);
#else
        return bhb200_update_values_f64(_ctx, csrValA, csrValB);
#endif
    }
    int spgemm_numeric() { return _ctx ? bhb200_spgemm_numeric(_ctx) : BHB200_ERR_INVALID; }

    int get_C(index_type *csrColIndC, value_type *csrValC)   /* bhsparse_cuda.h:3011-3020 */
    {
        if (!_ctx) return BHB200_ERR_INVALID;
#ifdef BHSPARSE_VALUE_FLOAT
        return bhb200_get_C_f32(_ctx, _rowptrC, csrColIndC, csrValC);
#else
        return bhb200_get_C_f64(_ctx, _rowptrC, csrColIndC, csrValC);
#endif
    }

    int free_mem() { return _ctx ? bhb200_free_mem(_ctx) : BHSPARSE_SUCCESS; }   /* bhsparse.h:150-177 */

    int freePlatform()   /* bhsparse.h:126-148 */
    {
        int err = _ctx ? bhb200_destroy(_ctx) : BHSPARSE_SUCCESS;
        _ctx = nullptr;
        return err;
    }

    bhb200_ctx *native_handle() { return _ctx; }

private:
    bhb200_ctx *_ctx;
    index_type *_rowptrC;
    bool _verbose;
};

#endif /* BHSPARSE_H */
