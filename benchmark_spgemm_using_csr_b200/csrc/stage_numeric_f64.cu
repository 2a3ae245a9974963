// stage_numeric_f64.cu -- double instantiation of the numeric kernels (stage_numeric.cuh).
#include "stage_numeric.cuh"

namespace bhb {

cudaError_t launch_num_hash_f64(const LaunchCtx &lc, int bin, int G, const int *queue, int count, Csr A, Csr B,
                                const int64_t *rowoff, int *colC, double *valC)
{
    return launch_num_hash_t<double>(lc, bin, G, queue, count, A, B, rowoff, colC, valC);
}

cudaError_t launch_num_large_f64(const LaunchCtx &lc, const int *queue, int count, int n, Csr A, Csr B,
                                 const int64_t *rowoff, int *colC, double *valC, unsigned *bitmap_scratch,
                                 int *prefix_scratch, int scratch_blocks)
{
    return launch_num_large_t<double>(lc, queue, count, n, A, B, rowoff, colC, valC, bitmap_scratch, prefix_scratch,
                                   scratch_blocks);
}

}  // namespace bhb
