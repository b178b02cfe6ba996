// Microbenchmark (design aid, not product code): does a warp-wide LDS.64 whose lanes share their 8-byte chunks cost one
// shared-memory wavefront when the unique data fits 128 bytes, like LDS.32 does?  The bilinear tap pattern: lane i needs
// bytes 3 i .. 3 i + 5 of a staged row (a warp's 32 pixels = 96 + 3 contiguous bytes).
//   A: the launched scheme -- word (3 i >> 2), the next word, and a predicated third word for the lanes with (3 i & 3) == 3
//   B: the aligned 8-byte chunk (3 i >> 3) and a predicated second chunk for the lanes with (3 i & 7) >= 3
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/lds64_share scripts/ubench/lds64_share.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int ROWS>
__global__ void __launch_bounds__(288, 4) k(unsigned* out, int iters) {
    __shared__ __align__(16) unsigned sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 2654435761u;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + warp * 1024;
    unsigned acc = threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const unsigned off = 3 * lane + (it & 7) + ((it >> 3) & 3) * 128;  // byte offset of the lane's window, sliding start
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const unsigned row = base + r * 160 + 4096 * (r & 1);
            if (MODE == 0) {
                unsigned a0, a1;
                const unsigned ad = row + (off & ~3u);
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a0) : "r"(ad) : "memory");
                asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(a1) : "r"(ad) : "memory");
                acc ^= __byte_perm(a0, a1, 0x5140 + (off & 3));
                asm volatile("{.reg .pred p; setp.eq.u32 p, %2, 3; @p ld.shared.u32 %0, [%1+8];}" : "+r"(a0) : "r"(ad), "r"(off & 3u) : "memory");
                acc += a0;
            } else {
                unsigned a0, a1, b0 = 0, b1 = 0;
                const unsigned ad = row + (off & ~7u);
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a0), "=r"(a1) : "r"(ad) : "memory");
                acc ^= __byte_perm(a0, a1, 0x5140 + (off & 3));
                asm volatile("{.reg .pred p; setp.ge.u32 p, %3, 3; @p ld.shared.v2.u32 {%0, %1}, [%2+8];}" : "+r"(b0), "+r"(b1) : "r"(ad), "r"(off & 7u) : "memory");
                acc += b0 ^ b1;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE, int ROWS>
void run(const char* name, unsigned* out) {
    const int iters = 4000, grid = 148 * 4;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<MODE, ROWS><<<grid, 288>>>(out, iters);
    cudaEventRecord(a);
    k<MODE, ROWS><<<grid, 288>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double clk = ms * 1e-3 * clk_khz * 1e3 / (4.0 * 9 * iters);
    printf("%-44s rows %d: %.3f ms, %.2f clk per warp-iteration per SM = %.2f per row\n", name, ROWS, ms, clk, clk / ROWS);
}

int main() {
    unsigned* out;
    cudaMalloc(&out, 148 * 4 * 288 * 4);
    run<0, 2>("A: 2 LDS.32 + predicated third word", out);
    run<1, 2>("B: LDS.64 + predicated second LDS.64", out);
    run<0, 4>("A: 2 LDS.32 + predicated third word", out);
    run<1, 4>("B: LDS.64 + predicated second LDS.64", out);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
