// Optimizer step on the flat weight / gradient buffers: Keras SGD with momentum (optionally Nesterov) and PER-VARIABLE
// gradient-norm clipping -- the optimizer the reference trains with,
//   tf.keras.optimizers.SGD(learning_rate=CosineDecayRestarts(1e-3, 4000, m_mul=.95, alpha=.1), momentum=.9 / .95,
//                           nesterov=True, clipnorm=0.1)          (/root/reference/Boosted_DETR_COCO.ipynb cells 26, 30)
// restated from the published TF 2.x semantics:
//   clipnorm   g <- tf.clip_by_norm(g, c) = (g * c) / max(||g||_2, c), one norm per variable
//   momentum   accum <- momentum * accum - lr * g ;  var += nesterov ? momentum * accum - lr * g : accum
// HBM-bound streaming kernels: one read of g for the norms, one read of g / accum / var and one write of accum / var
// for the update (28 bytes per parameter), float4 accesses, grid = 16 K-element chunks of the trainable variables only
// (frozen blocks -- the reference's boosted training regime toggles `trainable` per block -- cost nothing).
#include "common.cuh"

namespace bdetr {

constexpr int OPT_THREADS = 256;

// Sum of squares of one chunk, fixed reduction order (bitwise reproducible run to run).
__global__ void __launch_bounds__(OPT_THREADS)
sgd_sqnorm_kernel(const bdetr_opt_chunk *__restrict__ chunks, const float *__restrict__ grads, float *__restrict__ partial)
{
    pdl_sync();
    const bdetr_opt_chunk ch = chunks[blockIdx.x];
    const float *g = grads + ch.offset;
    float acc = 0.0f;
    if ((ch.offset & 3) == 0) {
        const int n4 = ch.len >> 2;
        const float4 *g4 = reinterpret_cast<const float4 *>(g);
#pragma unroll 4
        for (int i = threadIdx.x; i < n4; i += OPT_THREADS) { const float4 v = g4[i]; acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
        for (int i = (n4 << 2) + threadIdx.x; i < ch.len; i += OPT_THREADS) acc += g[i] * g[i];
    } else {
        for (int i = threadIdx.x; i < ch.len; i += OPT_THREADS) acc += g[i] * g[i];
    }
    __shared__ float red[OPT_THREADS / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < OPT_THREADS / 32; ++w) s += red[w];
        partial[blockIdx.x] = s;
    }
}

__device__ __forceinline__ void sgd_update(float g, float &a, float &w, float lr, float momentum, int nesterov, float clip, float den)
{
    if (clip > 0.0f) g = __fdiv_rn(__fmul_rn(g, clip), den);
    const float step = __fmul_rn(lr, g);
    a = __fsub_rn(__fmul_rn(a, momentum), step);
    w = __fadd_rn(w, nesterov ? __fsub_rn(__fmul_rn(a, momentum), step) : a);
}

__global__ void __launch_bounds__(OPT_THREADS)
sgd_apply_kernel(const bdetr_opt_chunk *__restrict__ chunks, float *__restrict__ weights, const float *__restrict__ grads,
                 float *__restrict__ accum, const float *__restrict__ partial, float lr, const float *__restrict__ lr_dev,
                 float momentum, int nesterov, float clip)
{
    pdl_sync();
    if (lr_dev) lr = *lr_dev;                               // CUDA-graph replays: the step's learning rate lives in device memory
    const bdetr_opt_chunk ch = chunks[blockIdx.x];
    float den = 1.0f;
    if (clip > 0.0f) {
        float n2 = 0.0f;                                    // the variable's chunks, summed in table order by every CTA alike
        for (int i = 0; i < ch.var_chunks; ++i) n2 += partial[ch.var_first + i];
        den = fmaxf(n2 > 0.0f ? sqrtf(n2) : n2, clip);
    }
    float *w = weights + ch.offset, *a = accum + ch.offset;
    const float *g = grads + ch.offset;
    int done = 0;
    if ((ch.offset & 3) == 0) {
        const int n4 = ch.len >> 2;
        for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
            const float4 g4 = reinterpret_cast<const float4 *>(g)[i];
            float4 a4 = reinterpret_cast<float4 *>(a)[i], w4 = reinterpret_cast<float4 *>(w)[i];
            sgd_update(g4.x, a4.x, w4.x, lr, momentum, nesterov, clip, den);
            sgd_update(g4.y, a4.y, w4.y, lr, momentum, nesterov, clip, den);
            sgd_update(g4.z, a4.z, w4.z, lr, momentum, nesterov, clip, den);
            sgd_update(g4.w, a4.w, w4.w, lr, momentum, nesterov, clip, den);
            reinterpret_cast<float4 *>(a)[i] = a4;
            reinterpret_cast<float4 *>(w)[i] = w4;
        }
        done = n4 << 2;
    }
    for (int i = done + threadIdx.x; i < ch.len; i += OPT_THREADS) {
        float ai = a[i], wi = w[i];
        sgd_update(g[i], ai, wi, lr, momentum, nesterov, clip, den);
        a[i] = ai; w[i] = wi;
    }
}

}  // namespace bdetr

using namespace bdetr;

extern "C" __attribute__((visibility("default"))) int bdetr_sgd_step(int n_chunks, const bdetr_opt_chunk *chunks,
                                     float *weights, const float *grads, float *accum, float *partial,
                                     float lr, const float *lr_dev, float momentum, int nesterov, float clipnorm, void *stream)
{
    BDETR_REQUIRE(n_chunks >= 0, BDETR_E_BAD_SHAPE, "n_chunks must be non-negative");
    if (n_chunks == 0) return BDETR_OK;
    BDETR_REQUIRE(chunks && weights && grads && accum && partial, BDETR_E_NULL, "null pointer");
    cudaStream_t s = as_stream(stream);
    if (clipnorm > 0.0f) {
        launch_k(sgd_sqnorm_kernel, n_chunks, OPT_THREADS, 0, s, chunks, grads, partial);
        BDETR_CHECK_LAUNCH("sgd_sqnorm_kernel");
    }
    launch_k(sgd_apply_kernel, n_chunks, OPT_THREADS, 0, s, chunks, weights, grads, accum, partial, lr, lr_dev, momentum, nesterov, clipnorm);
    BDETR_CHECK_LAUNCH("sgd_apply_kernel");
    return BDETR_OK;
}
