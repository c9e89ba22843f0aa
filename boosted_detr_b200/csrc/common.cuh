// Shared helpers for libbdetr (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/bdetr.h"

namespace bdetr {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// Checks the launch (not the execution: everything is asynchronous).
#define BDETR_CHECK_LAUNCH(name)                                                     \
    do {                                                                             \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            ::bdetr::set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__)); \
            return BDETR_E_CUDA;                                                     \
        }                                                                            \
        ::bdetr::count_launch();                                                     \
    } while (0)

#define BDETR_CUDA(call)                                                             \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            ::bdetr::set_error("%s failed: %s", #call, cudaGetErrorString(e__));     \
            return BDETR_E_CUDA;                                                     \
        }                                                                            \
    } while (0)

#define BDETR_REQUIRE(cond, code, msg)                                               \
    do {                                                                             \
        if (!(cond)) {                                                               \
            ::bdetr::set_error("%s: %s", __func__, msg);                             \
            return code;                                                             \
        }                                                                            \
    } while (0)

__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}

// Counter-based dropout keep decision shared with the oracle (DESIGN.md "dropout").
__device__ __forceinline__ bool dropout_keep(uint32_t idx, uint32_t key, uint32_t thresh)
{
    return lowbias32(idx ^ key) >= thresh;
}

__host__ inline uint32_t dropout_threshold(float rate)
{
    double t = (double)rate * 4294967296.0;
    if (t <= 0.0) return 0u;
    if (t >= 4294967295.0) return 4294967295u;
    return (uint32_t)t;
}

// round-to-nearest fp32 -> tf32 (10-bit mantissa).  In tensor-core mode every tensor that feeds a tcgen05
// GEMM is stored already rounded, so the MMA's operand truncation is exact (no truncation bias).
__device__ __forceinline__ float tf32_rn(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float maybe_tf32(float x, int on) { return on ? tf32_rn(x) : x; }

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------
// The step is ~800 short dependent kernels, so the drain -> launch -> ramp bubble between two kernels is a
// large share of it.  Every kernel is launched with programmaticStreamSerialization: it signals
// `launch_dependents` as soon as it starts, so the NEXT kernel's CTAs are scheduled onto free SMs and run their
// prologue (barrier init, TMEM allocation, index math) while this one is still computing; the next kernel
// then blocks in `griddepcontrol.wait` until this grid has completed and its writes are visible.  Every global
// access of a kernel sits after its pdl_wait(), so the memory semantics are those of plain stream order.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_trigger(); pdl_wait(); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // error picked up by BDETR_CHECK_LAUNCH
}

// the same for a kernel that runs as thread-block clusters of `cluster` CTAs along x
template <typename... KArgs, typename... Args>
static inline void launch_cluster_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, int cluster, size_t smem, cudaStream_t s, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace bdetr
