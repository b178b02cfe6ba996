// Microbenchmark (design aid, not product code): do shared-memory loads and warp shuffles share the SM's LSU data
// pipe on B200?  Every variant runs the same loop with a different mix of LDS.32 / SHFL per iteration; the
// wavefront model says t ~ max(issue, LDS + SHFL) if they share the pipe and max(issue, LDS, SHFL) if not.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/lsu_mix scripts/ubench/lsu_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NLDS, int NSHFL, int NALU>
__global__ void __launch_bounds__(288, 4) k(unsigned* out, int iters) {
    __shared__ unsigned sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 2654435761u;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    unsigned addr = (unsigned)__cvta_generic_to_shared(sm) + lane * 4 + (threadIdx.x >> 5) * 256;
    unsigned acc = threadIdx.x, v[8] = {1, 2, 3, 4, 5, 6, 7, 8};
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const unsigned a_it = addr + ((it & 3) << 7);  // another word of the same banks every iteration
#pragma unroll
        for (int j = 0; j < NLDS; ++j) {
            unsigned x;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(a_it + j * 512) : "memory");
            v[j & 7] ^= x;
        }
#pragma unroll
        for (int j = 0; j < NSHFL; ++j) v[j & 7] ^= __shfl_down_sync(0xffffffffu, acc + j, 1);
#pragma unroll
        for (int j = 0; j < NALU; ++j) v[j & 7] = __byte_perm(v[j & 7], v[(j + 3) & 7], 0x5140 + j) + acc;
        acc += v[0] ^ v[3];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ v[0] ^ v[1] ^ v[2] ^ v[3] ^ v[4] ^ v[5] ^ v[6] ^ v[7];
}

template <int NLDS, int NSHFL, int NALU>
void run(const char* name, unsigned* out) {
    const int iters = 4000, grid = 148 * 4;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<NLDS, NSHFL, NALU><<<grid, 288>>>(out, iters);
    cudaEventRecord(a);
    k<NLDS, NSHFL, NALU><<<grid, 288>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    // warp-iterations per SM: 4 CTAs x 9 warps x iters; clocks at the nominal max clock
    const double clk_per_warp_iter = ms * 1e-3 * clk_khz * 1e3 / (4.0 * 9 * iters);
    printf("%-28s LDS %d SHFL %d ALU %2d : %.3f ms, %.2f clk per warp-iteration per SM (at %d MHz nominal)\n", name, NLDS,
           NSHFL, NALU, ms, clk_per_warp_iter, clk_khz / 1000);
}

int main() {
    unsigned* out;
    cudaMalloc(&out, 148 * 4 * 288 * 4);
    run<6, 0, 0>("6 LDS", out);
    run<4, 0, 0>("4 LDS", out);
    run<0, 6, 0>("6 SHFL", out);
    run<0, 2, 0>("2 SHFL", out);
    run<5, 1, 0>("5 LDS + 1 SHFL", out);
    run<4, 2, 0>("4 LDS + 2 SHFL", out);
    run<6, 1, 0>("6 LDS + 1 SHFL", out);
    run<6, 1, 20>("6 LDS + 1 SHFL + 20 ALU", out);
    run<5, 2, 20>("5 LDS + 2 SHFL + 20 ALU", out);
    run<4, 1, 20>("4 LDS + 1 SHFL + 20 ALU", out);
    run<0, 0, 20>("20 ALU", out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
