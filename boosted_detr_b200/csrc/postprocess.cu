// Numeric half of the reference's InverseTokenization (tokenizers.py:126-137) and the confidence statistic of the
// early-exit path the reference lists as TODO (README.md:9): "final predictions are produced once prediction
// confidence reaches a desired threshold".
//   tokens_categories [M]   = argmax_c cat_pred[m, :]                 (tf.argmax: first index of the maximum)
//   tokens_attributes [M,A] = (attr_pred[m,a] >= .5) * a              (multi-hot indicator times tf.range)
//   confidence [M]          = max_c cat_pred[m, c] * conf_scale       (conf_scale = 1 / number of summed softmaxes: the boosted
//                             running prediction after block i is a sum of i + 2 probability vectors, quirk Q2 / Q3)
//   image_conf [B]          = min over the image's queries of confidence (optional)
// One warp per prediction row; HBM-bound (each input read once).
#include <math_constants.h>
#include "kernels.cuh"

namespace bdetr {

__global__ void __launch_bounds__(256)
inverse_tokenize_kernel(int M, int C, int A, const float *__restrict__ cat_pred, const float *__restrict__ attr_pred,
                        int32_t *__restrict__ tok_cat, int32_t *__restrict__ tok_attr, float *__restrict__ conf, float conf_scale)
{
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const float *p = cat_pred + (size_t)row * C;
    float best = -CUDART_INF_F;
    int bi = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {                    // ascending c per lane: strict > keeps the first maximum
        const float v = p[c];
        if (v > best || (v != v && best == best)) { best = v; bi = c; }   // (a NaN wins like in tf.argmax's max-reduction)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const bool take = (ob > best) || (ob == best && oi < bi) || (ob != ob && best == best);
        if (take) { best = ob; bi = oi; }
    }
    if (lane == 0) {
        if (tok_cat) tok_cat[row] = bi == 0x7fffffff ? 0 : bi;
        if (conf) conf[row] = best * conf_scale;
    }
    if (tok_attr && attr_pred) {
        const float *a = attr_pred + (size_t)row * A;
        int32_t *o = tok_attr + (size_t)row * A;
        for (int k = lane; k < A; k += 32) o[k] = a[k] >= 0.5f ? k : 0;
    }
}

// image_conf[b] = min_q conf[b, q]; grid B
__global__ void __launch_bounds__(128)
image_confidence_kernel(int Q, const float *__restrict__ conf, float *__restrict__ image_conf)
{
    pdl_sync();
    __shared__ float red[4];
    const int b = blockIdx.x;
    float m = CUDART_INF_F;
    for (int q = threadIdx.x; q < Q; q += 128) m = fminf(m, conf[(size_t)b * Q + q]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) image_conf[b] = fminf(fminf(red[0], red[1]), fminf(red[2], red[3]));
}

}  // namespace bdetr

using namespace bdetr;

extern "C" __attribute__((visibility("default"))) int bdetr_inverse_tokenize(int B, int Q, int C, int A, const float *cat_pred, const float *attr_pred,
                                        int32_t *tokens_categories, int32_t *tokens_attributes, float *confidence,
                                        float *image_confidence, float conf_scale, void *stream)
{
    BDETR_REQUIRE(B > 0 && Q > 0 && C > 0 && A >= 0, BDETR_E_BAD_SHAPE, "bad shape");
    BDETR_REQUIRE(cat_pred && (tokens_categories || confidence), BDETR_E_NULL, "null pointer");
    BDETR_REQUIRE(!image_confidence || confidence, BDETR_E_NULL, "image_confidence needs the per-query confidence buffer");
    cudaStream_t s = as_stream(stream);
    const int M = B * Q;
    launch_k(inverse_tokenize_kernel, ceil_div(M, 8), 256, 0, s, M, C, A, cat_pred, attr_pred, tokens_categories, tokens_attributes, confidence, conf_scale);
    BDETR_CHECK_LAUNCH("inverse_tokenize_kernel");
    if (image_confidence) {
        launch_k(image_confidence_kernel, B, 128, 0, s, Q, (const float *)confidence, image_confidence);
        BDETR_CHECK_LAUNCH("image_confidence_kernel");
    }
    return BDETR_OK;
}
