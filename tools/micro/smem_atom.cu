// Micro-benchmark: shared-memory atomic / load / store throughput per SM on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atom smem_atom.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE, int OP>
__global__ void __launch_bounds__(512) k(const uint32_t* __restrict__ idx, uint32_t* out, int iters)
{
    __shared__ uint32_t sh[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    uint32_t a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        uint32_t r = idx[(blockIdx.x * blockDim.x + threadIdx.x) * 8 + j];
        if (MODE == 0) a[j] = lane + 32 * j;               // conflict-free, distinct addresses
        else if (MODE == 1) a[j] = 7;                      // all lanes same address
        else if (MODE == 2) a[j] = r & 255;                // random over 256 bins
        else if (MODE == 3) a[j] = (r & 255) * 4 + (lane & 3);  // 4 interleaved copies
        else if (MODE == 4) a[j] = (r & 1023) ;            // random over 1024 words
        else a[j] = (lane >> 1) + 32 * j;                  // pairs of lanes share an address
    }
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (OP == 0) atomicAdd(&sh[a[j]], 1u);
            else if (OP == 1) atomicOr(&sh[a[j]], acc + it);
            else if (OP == 2) acc += sh[a[j]];
            else if (OP == 3) sh[a[j]] = acc + it;
            else acc += atomicAdd(&sh[a[j]], 1u);
            a[j] = (a[j] + (OP == 2 ? acc & 0 : 0)) & 4095;
        }
    }
    __syncthreads();
    if (acc == 0xdeadbeef || threadIdx.x == 0) out[blockIdx.x] = sh[threadIdx.x] + acc;
}

template <int MODE, int OP>
void run(const char* name, const uint32_t* idx, uint32_t* out)
{
    const int iters = 2000, grid = 148 * 2, block = 512;
    k<MODE, OP><<<grid, block>>>(idx, out, 10);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE, OP><<<grid, block>>>(idx, out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // warp-instructions per SM: grid/148 CTAs * 16 warps * iters * 8
    const double winst = (double)grid / 148 * 16 * iters * 8;
    const double cyc = ms * 1e-3 * 1.90e9;  // approx at ~1.9 GHz
    printf("%-40s %8.3f ms  %6.2f cyc/warp-instr/SM (assuming 1.9 GHz)\n", name, ms, cyc / winst);
}

int main()
{
    const size_t n = (size_t)148 * 2 * 512 * 8;
    uint32_t* h = (uint32_t*)malloc(n * 4);
    uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (uint32_t)(s >> 11); }
    uint32_t *idx, *out;
    cudaMalloc(&idx, n * 4); cudaMalloc(&out, 4096 * 4);
    cudaMemcpy(idx, h, n * 4, cudaMemcpyHostToDevice);
    run<0, 0>("ATOMS.ADD conflict-free", idx, out);
    run<1, 0>("ATOMS.ADD all same address", idx, out);
    run<5, 0>("ATOMS.ADD lane pairs share address", idx, out);
    run<2, 0>("ATOMS.ADD random 256 bins", idx, out);
    run<3, 0>("ATOMS.ADD random 256 bins x4 copies", idx, out);
    run<4, 0>("ATOMS.ADD random 1024 words", idx, out);
    run<0, 4>("ATOMS.ADD (returning) conflict-free", idx, out);
    run<0, 1>("ATOMS.OR conflict-free", idx, out);
    run<4, 1>("ATOMS.OR random 1024 words", idx, out);
    run<0, 2>("LDS conflict-free", idx, out);
    run<2, 2>("LDS random 256", idx, out);
    run<4, 2>("LDS random 1024", idx, out);
    run<0, 3>("STS conflict-free", idx, out);
    run<4, 3>("STS random 1024", idx, out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
