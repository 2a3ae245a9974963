// smem_ubench.cu -- shared-memory op throughput on B200 (design evidence for the
// accumulator choice: plain LDS/STS vs ATOMS.OR / ATOMS.CAS, by address pattern).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_ubench smem_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

enum Op { LDS = 0, STS = 1, ATOM_OR = 2, ATOM_OR_RET = 3, ATOM_CAS = 4, ATOM_ADD = 5, LDS64_FMA_STS64 = 6, MATCH_ANY = 7, REDUX_OR = 8 };
enum Pat { SPREAD = 0, RANDOM = 1, CLUSTER3 = 2, SAME = 3 };

template <int OP, int PAT>
__global__ void __launch_bounds__(256) k(int iters, unsigned *out)
{
    __shared__ unsigned sm[4096];
    __shared__ double smd[2048];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 0;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) smd[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wbase = (threadIdx.x >> 5) * 512;
    unsigned acc = 0;
    unsigned x = threadIdx.x * 2654435761u + 12345u;
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        int a;
        if (PAT == SPREAD) a = (lane + it) & 511;
        else if (PAT == RANDOM) a = (x >> 12) & 511;
        else if (PAT == CLUSTER3) a = ((lane / 3) * 37 + it) & 511;
        else a = it & 511;
        unsigned bit = 1u << (x & 31);
        if (OP == LDS) acc += sm[wbase + a];
        else if (OP == STS) sm[wbase + a] = bit;
        else if (OP == ATOM_OR) atomicOr(&sm[wbase + a], bit);
        else if (OP == ATOM_OR_RET) acc += atomicOr(&sm[wbase + a], bit);
        else if (OP == ATOM_CAS) acc += atomicCAS(&sm[wbase + a], 0u, bit);
        else if (OP == ATOM_ADD) atomicAdd(&sm[wbase + a], 1u);
        else if (OP == LDS64_FMA_STS64) { double *p = &smd[(wbase >> 1) + (a & 255)]; *p = fma((double)bit, 1.5, *p); }
        else if (OP == MATCH_ANY) acc += __match_any_sync(0xffffffffu, a);
        else if (OP == REDUX_OR) acc += __reduce_or_sync(0xffffffffu, bit);
    }
    __syncthreads();
    if (acc == 0x12345678u || threadIdx.x == 0) out[blockIdx.x] = acc + sm[threadIdx.x] + (unsigned)smd[threadIdx.x];
}

template <int OP, int PAT>
void run(const char *name, int sms, double ghz)
{
    unsigned *out;
    cudaMalloc(&out, 1 << 20);
    const int iters = 20000, blocks = sms * 4, threads = 256;
    k<OP, PAT><<<blocks, threads>>>(100, out);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP, PAT><<<blocks, threads>>>(iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr_per_sm = (double)iters * 4 * (threads / 32);
    double cycles = ms * 1e-3 * ghz * 1e9;
    printf("%-34s %8.3f ms  %7.2f cycles per warp-op per SM\n", name, ms, cycles / warp_instr_per_sm);
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double ghz = 1.965;
    printf("%s, %d SMs; 32 warps/SM resident; cycles assume %.3f GHz (includes ~6 ALU ops of address math per op)\n", p.name, sms, ghz);
#define R(OP, PAT) run<OP, PAT>(#OP " " #PAT, sms, ghz)
    R(LDS, SPREAD); R(LDS, RANDOM); R(LDS, CLUSTER3);
    R(STS, SPREAD); R(STS, RANDOM);
    R(ATOM_OR, SPREAD); R(ATOM_OR, RANDOM); R(ATOM_OR, CLUSTER3); R(ATOM_OR, SAME);
    R(ATOM_OR_RET, SPREAD); R(ATOM_OR_RET, RANDOM); R(ATOM_OR_RET, CLUSTER3);
    R(ATOM_CAS, SPREAD); R(ATOM_CAS, RANDOM);
    R(ATOM_ADD, SPREAD); R(ATOM_ADD, RANDOM); R(ATOM_ADD, CLUSTER3);
    R(LDS64_FMA_STS64, SPREAD); R(LDS64_FMA_STS64, RANDOM);
    R(MATCH_ANY, CLUSTER3); R(REDUX_OR, SPREAD);
    return 0;
}
