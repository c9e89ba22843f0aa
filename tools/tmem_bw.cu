// Micro-benchmark (not part of the library): tcgen05.ld throughput per SM as a function of the number of reading warps.
// Answers whether the long-sequence attention kernel (one 64 KB fp32 S tile read per 128 x 128 scores) sits on a TMEM-read floor.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_bw tools/tmem_bw.cu && /tmp/tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

__global__ void k(int iters, long long *cycles, uint32_t *sink)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t a[32], b[32];
        ld32(base + ((i * 64) & 448), a);
        ld32(base + ((i * 64 + 32) & 448), b);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= a[j] + b[j];
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot));
}

int main()
{
    long long *cyc; uint32_t *sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 1024 * 4);
    const int iters = 2000;
    for (int warps : {4, 8, 12, 16}) {
        k<<<148, warps * 32>>>(iters, cyc, sink);
        k<<<148, warps * 32>>>(iters, cyc, sink);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const double bytes = (double)warps * iters * 2 * 32 * 32 * 4;
        printf("%2d warps: %lld cycles, %.1f B/clk/SM TMEM read (%.1f B/clk per warp)\n", warps, h[0], bytes / h[0], bytes / h[0] / warps);
    }
    return 0;
}
