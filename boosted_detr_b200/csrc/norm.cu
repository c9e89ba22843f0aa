// Row-wise / column-wise fp32 kernels of the dense path: dropout + residual + LayerNorm (fwd/bwd),
// positional add and its batch reduction, BatchNorm (+ReLU backward) and the prediction-head
// activations with the boosted running sum.  All HBM-bound: coalesced float4 rows, warp shuffles.
// Reference: transformers.py:139-151,182-193 (Dropout/Add/LayerNorm), :226-227,:441 (Add),
// prediction_heads.py:59-62,126-129,196-199, boosted_model.py:222-229.
#include <math_constants.h>
#include "kernels.cuh"

namespace bdetr {

// ------------------------------------------------------------------------------------------
// dropout + residual + LayerNorm.  One warp per row, NV float4 per lane (D = 128*NV).
// ------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
res_ln_fwd_kernel(int M, const float *__restrict__ resid, float *__restrict__ z, const float *__restrict__ gamma,
                  const float *__restrict__ beta, float eps, float keep_scale, uint32_t thresh, uint32_t key,
                  const uint32_t *__restrict__ seed_dev,
                  float *__restrict__ out, float *__restrict__ mean_o, float *__restrict__ rstd_o, int round_out)
{
    pdl_sync();
    if (seed_dev) key = lowbias32(*seed_dev ^ key);      // per-step seed read from device memory (CUDA-graph replays)
    constexpr int D = NV * 128;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    float v[NV * 4];
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int col = (k * 32 + lane) * 4;
        const size_t off = (size_t)row * D + col;
        const float4 a = *reinterpret_cast<const float4 *>(z + off);
        const float4 r = *reinterpret_cast<const float4 *>(resid + off);
        float av[4] = {a.x, a.y, a.z, a.w};
        const float rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (thresh) av[e] = dropout_keep((uint32_t)(off + e), key, thresh) ? av[e] * keep_scale : 0.0f;
            v[4 * k + e] = rv[e] + av[e];
            sum += v[4 * k + e];
        }
        *reinterpret_cast<float4 *>(z + off) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    }
    const float mean = warp_sum(sum) * (1.0f / D);
    float sq = 0.0f;
#pragma unroll
    for (int e = 0; e < NV * 4; ++e) { const float dlt = v[e] - mean; sq = fmaf(dlt, dlt, sq); }
    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int col = (k * 32 + lane) * 4;
        const float4 g = *reinterpret_cast<const float4 *>(gamma + col);
        const float4 bt = *reinterpret_cast<const float4 *>(beta + col);
        float4 y;
        y.x = (v[4 * k] - mean) * rstd * g.x + bt.x; y.y = (v[4 * k + 1] - mean) * rstd * g.y + bt.y;
        y.z = (v[4 * k + 2] - mean) * rstd * g.z + bt.z; y.w = (v[4 * k + 3] - mean) * rstd * g.w + bt.w;
        if (round_out) { y.x = tf32_rn(y.x); y.y = tf32_rn(y.y); y.z = tf32_rn(y.z); y.w = tf32_rn(y.w); }
        *reinterpret_cast<float4 *>(out + (size_t)row * D + col) = y;
    }
    if (lane == 0) { mean_o[row] = mean; rstd_o[row] = rstd; }
}

template <int NV>
__global__ void __launch_bounds__(256)
res_ln_bwd_kernel(int M, const float *__restrict__ d_out, const float *__restrict__ z, const float *__restrict__ mean_i,
                  const float *__restrict__ rstd_i, const float *__restrict__ gamma, float keep_scale, uint32_t thresh,
                  uint32_t key, const uint32_t *__restrict__ seed_dev, float *__restrict__ d_resid, int acc_resid,
                  float *__restrict__ d_a, float *__restrict__ g_gamma, float *__restrict__ g_beta, float *__restrict__ g_bias,
                  int round_out)
{
    pdl_sync();
    if (seed_dev) key = lowbias32(*seed_dev ^ key);
    constexpr int D = NV * 128;
    __shared__ float red_g[8][D + 4];
    __shared__ float red_b[8][D + 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    float gg[NV * 4], gb[NV * 4], gam[NV * 4], gbias[NV * 4];   // gbias: column sums of d_a = bias gradient of the Dense before
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const float4 g = *reinterpret_cast<const float4 *>(gamma + (k * 32 + lane) * 4);
        gam[4 * k] = g.x; gam[4 * k + 1] = g.y; gam[4 * k + 2] = g.z; gam[4 * k + 3] = g.w;
    }
#pragma unroll
    for (int e = 0; e < NV * 4; ++e) { gg[e] = 0.0f; gb[e] = 0.0f; gbias[e] = 0.0f; }
    for (int row = blockIdx.x * nwarp + warp; row < M; row += gridDim.x * nwarp) {
        const float mean = mean_i[row], rstd = rstd_i[row];
        float xh[NV * 4], dy[NV * 4];
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const size_t off = (size_t)row * D + (k * 32 + lane) * 4;
            const float4 zz = *reinterpret_cast<const float4 *>(z + off);
            const float4 dd = *reinterpret_cast<const float4 *>(d_out + off);
            const float zv[4] = {zz.x, zz.y, zz.z, zz.w}, dv[4] = {dd.x, dd.y, dd.z, dd.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int idx = 4 * k + e;
                xh[idx] = (zv[e] - mean) * rstd;
                gg[idx] = fmaf(dv[e], xh[idx], gg[idx]);
                gb[idx] += dv[e];
                dy[idx] = dv[e] * gam[idx];           // d xhat
                s1 += dy[idx];
                s2 = fmaf(dy[idx], xh[idx], s2);
            }
        }
        s1 = warp_sum(s1) * (1.0f / D);
        s2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const size_t off = (size_t)row * D + (k * 32 + lane) * 4;
            float dz[4], da[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int idx = 4 * k + e;
                dz[e] = rstd * (dy[idx] - s1 - xh[idx] * s2);
                da[e] = thresh ? (dropout_keep((uint32_t)(off + e), key, thresh) ? dz[e] * keep_scale : 0.0f) : dz[e];
                da[e] = maybe_tf32(da[e], round_out);
                gbias[4 * k + e] += da[e];
            }
            *reinterpret_cast<float4 *>(d_a + off) = make_float4(da[0], da[1], da[2], da[3]);
            if (d_resid) {
                float4 r = make_float4(dz[0], dz[1], dz[2], dz[3]);
                if (acc_resid) { const float4 c = *reinterpret_cast<const float4 *>(d_resid + off); r.x += c.x; r.y += c.y; r.z += c.z; r.w += c.w; }
                *reinterpret_cast<float4 *>(d_resid + off) = r;
            }
        }
    }
    // reduce gamma/beta partials over the CTA's warps, one atomic per column per CTA
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            red_g[warp][(k * 32 + lane) * 4 + e] = gg[4 * k + e];
            red_b[warp][(k * 32 + lane) * 4 + e] = gb[4 * k + e];
        }
    __syncthreads();
    if (g_gamma) for (int c = threadIdx.x; c < D; c += blockDim.x) {      // (NULL: frozen layer, no parameter gradients)
        float a = 0.0f, b = 0.0f;
        for (int w = 0; w < nwarp; ++w) { a += red_g[w][c]; b += red_b[w][c]; }
        atomicAdd(&g_gamma[c], a);
        atomicAdd(&g_beta[c], b);
    }
    if (g_bias) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NV; ++k)
#pragma unroll
            for (int e = 0; e < 4; ++e) red_g[warp][(k * 32 + lane) * 4 + e] = gbias[4 * k + e];
        __syncthreads();
        for (int c = threadIdx.x; c < D; c += blockDim.x) {
            float a = 0.0f;
            for (int w = 0; w < nwarp; ++w) a += red_g[w][c];
            atomicAdd(&g_bias[c], a);
        }
    }
}

int launch_res_ln_fwd(int M, int D, const float *resid, float *z, const float *gamma, const float *beta, float eps,
                      float rate, uint32_t key, const uint32_t *seed_dev, float *out, float *mean, float *rstd, int round_out,
                      cudaStream_t s)
{
    BDETR_REQUIRE(M > 0 && (D == 128 || D == 256 || D == 512), BDETR_E_UNSUPPORTED, "LayerNorm width must be 128, 256 or 512");
    const uint32_t thresh = dropout_threshold(rate);
    const float ks = 1.0f / (1.0f - rate);
    const int grid = ceil_div(M, 8);
    if (D == 128) launch_k(res_ln_fwd_kernel<1>, grid, 256, 0, s, M, resid, z, gamma, beta, eps, ks, thresh, key, seed_dev, out, mean, rstd, round_out);
    else if (D == 256) launch_k(res_ln_fwd_kernel<2>, grid, 256, 0, s, M, resid, z, gamma, beta, eps, ks, thresh, key, seed_dev, out, mean, rstd, round_out);
    else launch_k(res_ln_fwd_kernel<4>, grid, 256, 0, s, M, resid, z, gamma, beta, eps, ks, thresh, key, seed_dev, out, mean, rstd, round_out);
    BDETR_CHECK_LAUNCH("res_ln_fwd_kernel");
    return BDETR_OK;
}

int launch_res_ln_bwd(int M, int D, const float *d_out, const float *z, const float *mean, const float *rstd,
                      const float *gamma, float rate, uint32_t key, const uint32_t *seed_dev, float *d_resid, int acc_resid,
                      float *d_a, float *g_gamma, float *g_beta, float *g_bias, int round_out, cudaStream_t s)
{
    BDETR_REQUIRE(M > 0 && (D == 128 || D == 256 || D == 512), BDETR_E_UNSUPPORTED, "LayerNorm width must be 128, 256 or 512");
    const uint32_t thresh = dropout_threshold(rate);
    const float ks = 1.0f / (1.0f - rate);
    const int grid = min(ceil_div(M, 8), 296);
    if (D == 128) launch_k(res_ln_bwd_kernel<1>, grid, 256, 0, s, M, d_out, z, mean, rstd, gamma, ks, thresh, key, seed_dev, d_resid, acc_resid, d_a, g_gamma, g_beta, g_bias, round_out);
    else if (D == 256) launch_k(res_ln_bwd_kernel<2>, grid, 256, 0, s, M, d_out, z, mean, rstd, gamma, ks, thresh, key, seed_dev, d_resid, acc_resid, d_a, g_gamma, g_beta, g_bias, round_out);
    else launch_k(res_ln_bwd_kernel<4>, grid, 256, 0, s, M, d_out, z, mean, rstd, gamma, ks, thresh, key, seed_dev, d_resid, acc_resid, d_a, g_gamma, g_beta, g_bias, round_out);
    BDETR_CHECK_LAUNCH("res_ln_bwd_kernel");
    return BDETR_OK;
}

// ------------------------------------------------------------------------------------------
// Broadcast form for the hoisted decoder self-attention: the attention output a [P,D] and the residual q0 [P,D] are
// batch-invariant, only the dropout mask differs per image:  z[r] = q0[r % P] + dropout_r(a[r % P]), out = LN(z).
// ------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
res_ln_bcast_fwd_kernel(int M, int P, const float *__restrict__ resid, const float *__restrict__ a, float *__restrict__ z,
                        const float *__restrict__ gamma, const float *__restrict__ beta, float eps, float keep_scale, uint32_t thresh,
                        uint32_t key, const uint32_t *__restrict__ seed_dev, float *__restrict__ out, float *__restrict__ mean_o,
                        float *__restrict__ rstd_o, int round_out)
{
    pdl_sync();
    if (seed_dev) key = lowbias32(*seed_dev ^ key);
    constexpr int D = NV * 128;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int prow = row % P;
    float v[NV * 4];
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int col = (k * 32 + lane) * 4;
        const size_t off = (size_t)row * D + col, poff = (size_t)prow * D + col;
        const float4 av4 = *reinterpret_cast<const float4 *>(a + poff);
        const float4 r = *reinterpret_cast<const float4 *>(resid + poff);
        float av[4] = {av4.x, av4.y, av4.z, av4.w};
        const float rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (thresh) av[e] = dropout_keep((uint32_t)(off + e), key, thresh) ? av[e] * keep_scale : 0.0f;
            v[4 * k + e] = rv[e] + av[e];
            sum += v[4 * k + e];
        }
        if (z) *reinterpret_cast<float4 *>(z + off) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    }
    const float mean = warp_sum(sum) * (1.0f / D);
    float sq = 0.0f;
#pragma unroll
    for (int e = 0; e < NV * 4; ++e) { const float dlt = v[e] - mean; sq = fmaf(dlt, dlt, sq); }
    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int col = (k * 32 + lane) * 4;
        const float4 g = *reinterpret_cast<const float4 *>(gamma + col);
        const float4 bt = *reinterpret_cast<const float4 *>(beta + col);
        float4 y;
        y.x = (v[4 * k] - mean) * rstd * g.x + bt.x; y.y = (v[4 * k + 1] - mean) * rstd * g.y + bt.y;
        y.z = (v[4 * k + 2] - mean) * rstd * g.z + bt.z; y.w = (v[4 * k + 3] - mean) * rstd * g.w + bt.w;
        if (round_out) { y.x = tf32_rn(y.x); y.y = tf32_rn(y.y); y.z = tf32_rn(y.z); y.w = tf32_rn(y.w); }
        *reinterpret_cast<float4 *>(out + (size_t)row * D + col) = y;
    }
    if (lane == 0) { mean_o[row] = mean; rstd_o[row] = rstd; }
}

int launch_res_ln_bcast_fwd(int M, int P, int D, const float *resid, const float *a, float *z, const float *gamma, const float *beta,
                            float eps, float rate, uint32_t key, const uint32_t *seed_dev, float *out, float *mean, float *rstd,
                            int round_out, cudaStream_t s)
{
    BDETR_REQUIRE(M > 0 && P > 0 && D == 256, BDETR_E_UNSUPPORTED, "broadcast LayerNorm needs width 256");
    const uint32_t thresh = dropout_threshold(rate);
    launch_k(res_ln_bcast_fwd_kernel<2>, ceil_div(M, 8), 256, 0, s, M, P, resid, a, z, gamma, beta, eps, 1.0f / (1.0f - rate), thresh, key,
             seed_dev, out, mean, rstd, round_out);
    BDETR_CHECK_LAUNCH("res_ln_bcast_fwd_kernel");
    return BDETR_OK;
}

// ------------------------------------------------------------------------------------------
// Batch reduction of up to four [B, rows, D] tensors in one launch: sum[t][l,:] = sum_b src[t][b,l,:] (optional) and
// colsum[t][:] += sum_{b,l} src[t][b,l,:] (optional: the bias gradient of the Dense that produced the tensor).
// CTA = 64 float4 columns x 4 rows; grid (rows / 4, tensors).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
batch_reduce_kernel(BatchReduce a)
{
    pdl_sync();
    __shared__ float4 red[4][64];
    const int t = blockIdx.y;
    const int rows = a.rows[t], D4 = a.D >> 2;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const float4 *src = reinterpret_cast<const float4 *>(a.src[t]);
    float4 *sum = reinterpret_cast<float4 *>(a.sum[t]);
    const int l = blockIdx.x * 4 + ty;                  // one row per thread row: many small CTAs, the loads of a thread independent
    if (blockIdx.x * 4 >= rows) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tx < D4 && l < rows) {
        const size_t stride = (size_t)rows * D4;
        const float4 *p = src + (size_t)l * D4 + tx;
        int b = 0;
        for (; b + 4 <= a.B; b += 4) {
            const float4 v0 = p[(size_t)b * stride], v1 = p[(size_t)(b + 1) * stride], v2 = p[(size_t)(b + 2) * stride], v3 = p[(size_t)(b + 3) * stride];
            acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
            acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
        }
        for (; b < a.B; ++b) { const float4 v = p[(size_t)b * stride]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        if (sum) sum[(size_t)l * D4 + tx] = acc;
    }
    if (a.colsum[t] == nullptr) return;
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && tx < D4) {
        float4 s = red[0][tx];
        for (int k = 1; k < 4; ++k) { const float4 v = red[k][tx]; s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
        float *c = a.colsum[t] + 4 * tx;
        atomicAdd(c, s.x); atomicAdd(c + 1, s.y); atomicAdd(c + 2, s.z); atomicAdd(c + 3, s.w);
    }
}

int launch_batch_reduce(const BatchReduce &a, cudaStream_t s)
{
    BDETR_REQUIRE(a.n >= 1 && a.n <= 4 && a.B > 0 && a.D > 0 && a.D % 4 == 0 && a.D <= 256, BDETR_E_BAD_SHAPE, "bad batch reduction");
    int maxrows = 0;
    for (int t = 0; t < a.n; ++t) { BDETR_REQUIRE(a.src[t] && a.rows[t] > 0, BDETR_E_NULL, "null batch-reduction input"); maxrows = max(maxrows, a.rows[t]); }
    launch_k(batch_reduce_kernel, dim3(ceil_div(maxrows, 4), a.n), 256, 0, s, a);
    BDETR_CHECK_LAUNCH("batch_reduce_kernel");
    return BDETR_OK;
}

// ------------------------------------------------------------------------------------------
// elementwise helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 round4(float4 v, int on)
{
    if (on) { v.x = tf32_rn(v.x); v.y = tf32_rn(v.y); v.z = tf32_rn(v.z); v.w = tf32_rn(v.w); }
    return v;
}
__global__ void add_rows_fwd_kernel(size_t n4, size_t ld4, const float4 *__restrict__ x, const float4 *__restrict__ pos, float4 *__restrict__ out, int round_out)
{
    pdl_sync();
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (size_t)gridDim.x * blockDim.x) {
        const float4 a = x[e], p = pos[e % ld4];
        out[e] = round4(make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w), round_out);
    }
}
__global__ void round_tf32_kernel(size_t n, const float *__restrict__ src, float *__restrict__ dst)
{
    pdl_sync();
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) dst[e] = tf32_rn(src[e]);
}
__global__ void batch_sum_acc_kernel(int B, size_t ld4, const float4 *__restrict__ src, float4 *__restrict__ dst)
{
    pdl_sync();
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < ld4; e += (size_t)gridDim.x * blockDim.x) {
        float4 a = dst[e];
        for (int b = 0; b < B; ++b) { const float4 v = src[(size_t)b * ld4 + e]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
        dst[e] = a;
    }
}
__global__ void tile_rows_kernel(size_t n4, size_t ld4, const float4 *__restrict__ src, float4 *__restrict__ dst, int round_out)
{
    pdl_sync();
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (size_t)gridDim.x * blockDim.x) dst[e] = round4(src[e % ld4], round_out);
}
__global__ void accumulate_kernel(size_t n, const float *__restrict__ x, float *__restrict__ y)
{
    pdl_sync();
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) y[e] += x[e];
}

// buf is [n, len]: buf[i] += buf[i + 1] for i = n - 2 .. 0 (running gradient of the boosted running prediction:
// block i's prediction feeds the losses of blocks i .. n - 1)
__global__ void suffix_sum_kernel(int n, size_t len, float *__restrict__ buf)
{
    pdl_sync();
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < len; e += (size_t)gridDim.x * blockDim.x) {
        float acc = buf[(size_t)(n - 1) * len + e];
        for (int i = n - 2; i >= 0; --i) { acc += buf[(size_t)i * len + e]; buf[(size_t)i * len + e] = acc; }
    }
}

static inline int ew_grid(size_t n) { size_t g = (n + 255) / 256; return (int)(g < 148 * 8 ? (g ? g : 1) : 148 * 8); }

int launch_add_rows_fwd(int B, int L, int D, const float *x, const float *pos, float *out, int round_out, cudaStream_t s)
{
    BDETR_REQUIRE(B > 0 && L > 0 && D > 0 && D % 4 == 0, BDETR_E_BAD_SHAPE, "D must be a multiple of 4");
    const size_t ld4 = (size_t)L * D / 4, n4 = ld4 * B;
    launch_k(add_rows_fwd_kernel, ew_grid(n4), 256, 0, s, n4, ld4, (const float4 *)x, (const float4 *)pos, (float4 *)out, round_out);
    BDETR_CHECK_LAUNCH("add_rows_fwd_kernel");
    return BDETR_OK;
}
int launch_batch_sum_acc(int B, int L, int D, const float *src, float *dst, cudaStream_t s)
{
    BDETR_REQUIRE(B > 0 && L > 0 && D > 0 && D % 4 == 0, BDETR_E_BAD_SHAPE, "D must be a multiple of 4");
    const size_t ld4 = (size_t)L * D / 4;
    launch_k(batch_sum_acc_kernel, ew_grid(ld4), 256, 0, s, B, ld4, (const float4 *)src, (float4 *)dst);
    BDETR_CHECK_LAUNCH("batch_sum_acc_kernel");
    return BDETR_OK;
}
int launch_tile_rows(int B, int L, int D, const float *src, float *dst, int round_out, cudaStream_t s)
{
    BDETR_REQUIRE(B > 0 && L > 0 && D > 0 && D % 4 == 0, BDETR_E_BAD_SHAPE, "D must be a multiple of 4");
    const size_t ld4 = (size_t)L * D / 4, n4 = ld4 * B;
    launch_k(tile_rows_kernel, ew_grid(n4), 256, 0, s, n4, ld4, (const float4 *)src, (float4 *)dst, round_out);
    BDETR_CHECK_LAUNCH("tile_rows_kernel");
    return BDETR_OK;
}
int launch_suffix_sum(int n, size_t len, float *buf, cudaStream_t s)
{
    BDETR_REQUIRE(n > 0 && len > 0 && buf, BDETR_E_BAD_SHAPE, "bad suffix-sum arguments");
    launch_k(suffix_sum_kernel, ew_grid(len), 256, 0, s, n, len, buf);
    BDETR_CHECK_LAUNCH("suffix_sum_kernel");
    return BDETR_OK;
}
int launch_round_tf32(size_t n, const float *src, float *dst, cudaStream_t s)
{
    BDETR_REQUIRE(n > 0 && src && dst, BDETR_E_BAD_SHAPE, "bad round arguments");
    launch_k(round_tf32_kernel, ew_grid(n), 256, 0, s, n, src, dst);
    BDETR_CHECK_LAUNCH("round_tf32_kernel");
    return BDETR_OK;
}
int launch_accumulate(size_t n, const float *x, float *y, cudaStream_t s)
{
    BDETR_REQUIRE(n > 0 && x && y, BDETR_E_BAD_SHAPE, "bad accumulate arguments");
    launch_k(accumulate_kernel, ew_grid(n), 256, 0, s, n, x, y);
    BDETR_CHECK_LAUNCH("accumulate_kernel");
    return BDETR_OK;
}

// ------------------------------------------------------------------------------------------
// BatchNorm over the rows of [M, Dh] (Keras non-fused path: biased variance, eps inside rsqrt).
// CTA = 32 columns x 8 row-lanes.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float col_reduce8(float v, float (*red)[33])
{
    const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
    __syncthreads();
    red[r][c] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][c];
    return t;
}

constexpr int BN_ROWS = 128;     // rows per CTA: grid = (Dh/32, M/128) so the whole chip takes part

// per row-chunk partials (no atomics: the forward stays bit-reproducible):
// acc[chunk][0:Dh] = sum_m (h - pivot), acc[chunk][Dh:2Dh] = sum_m (h - pivot)^2, pivot = h[0, c] (kills the cancellation)
__global__ void __launch_bounds__(256)
bn_stats_kernel(int M, int Dh, const float *__restrict__ h, float *__restrict__ acc)
{
    pdl_sync();
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), r = threadIdx.x >> 5;
    const bool live = c < Dh;
    const int m0 = blockIdx.y * BN_ROWS, m1 = min(M, m0 + BN_ROWS);
    const float pivot = live ? h[c] : 0.0f;
    float s1 = 0.0f, s2 = 0.0f;
    if (live) for (int m = m0 + r; m < m1; m += 8) { const float d = h[(size_t)m * Dh + c] - pivot; s1 += d; s2 = fmaf(d, d, s2); }
    s1 = col_reduce8(s1, red);
    s2 = col_reduce8(s2, red);
    if (live && r == 0) { acc[(size_t)blockIdx.y * 2 * Dh + c] = s1; acc[(size_t)blockIdx.y * 2 * Dh + Dh + c] = s2; }
}

__global__ void __launch_bounds__(256)
bn_apply_kernel(int M, int Dh, const float *__restrict__ h, const float *__restrict__ gamma, const float *__restrict__ beta,
                float *__restrict__ moving_mean, float *__restrict__ moving_var, float eps, float momentum, int training,
                const float *__restrict__ acc, float *__restrict__ hn, float *__restrict__ mean_o, float *__restrict__ rstd_o)
{
    pdl_sync();
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), r = threadIdx.x >> 5;
    if (c >= Dh) return;
    float mean, var;
    if (training) {
        float t1 = 0.0f, t2 = 0.0f;
        for (int k = 0; k < (int)gridDim.y; ++k) { t1 += acc[(size_t)k * 2 * Dh + c]; t2 += acc[(size_t)k * 2 * Dh + Dh + c]; }   // fixed order
        const float a1 = t1 / (float)M, a2 = t2 / (float)M;
        mean = h[c] + a1;
        var = fmaxf(a2 - a1 * a1, 0.0f);
        if (blockIdx.y == 0 && r == 0) {
            moving_mean[c] = moving_mean[c] * momentum + mean * (1.0f - momentum);
            moving_var[c] = moving_var[c] * momentum + var * (1.0f - momentum);
        }
    } else {
        mean = moving_mean[c];
        var = moving_var[c];
    }
    const float rstd = rsqrtf(var + eps);
    if (blockIdx.y == 0 && r == 0) { mean_o[c] = mean; rstd_o[c] = rstd; }
    const float g = gamma[c], b = beta[c];
    const int m0 = blockIdx.y * BN_ROWS, m1 = min(M, m0 + BN_ROWS);
    for (int m = m0 + r; m < m1; m += 8) hn[(size_t)m * Dh + c] = (h[(size_t)m * Dh + c] - mean) * rstd * g + b;
}

// acc[0:Dh] += sum dy, acc[Dh:2Dh] += sum dy * xhat
__global__ void __launch_bounds__(256)
bn_bwd_stats_kernel(int M, int Dh, const float *__restrict__ h, const float *__restrict__ d_hn, const float *__restrict__ mean_i,
                    const float *__restrict__ rstd_i, float *__restrict__ acc)
{
    pdl_sync();
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), r = threadIdx.x >> 5;
    const bool live = c < Dh;
    const int m0 = blockIdx.y * BN_ROWS, m1 = min(M, m0 + BN_ROWS);
    const float mean = live ? mean_i[c] : 0.0f, rstd = live ? rstd_i[c] : 0.0f;
    float s1 = 0.0f, s2 = 0.0f;
    if (live) for (int m = m0 + r; m < m1; m += 8) {
        const float dy = d_hn[(size_t)m * Dh + c];
        s1 += dy;
        s2 = fmaf(dy, (h[(size_t)m * Dh + c] - mean) * rstd, s2);
    }
    s1 = col_reduce8(s1, red);
    s2 = col_reduce8(s2, red);
    if (live && r == 0) { atomicAdd(&acc[c], s1); atomicAdd(&acc[Dh + c], s2); }
}

__global__ void __launch_bounds__(256)
bn_relu_bwd_apply_kernel(int M, int Dh, const float *__restrict__ h, const float *__restrict__ d_hn, const float *__restrict__ gamma,
                         const float *__restrict__ mean_i, const float *__restrict__ rstd_i, const float *__restrict__ acc,
                         float *__restrict__ d_h, float *__restrict__ g_gamma, float *__restrict__ g_beta,
                         float *__restrict__ g_bias, int round_out, int batch_stats)
{
    pdl_sync();
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), r = threadIdx.x >> 5;
    if (c >= Dh) {                      // keep the block-wide reduction below convergent
        if (g_bias) col_reduce8(0.0f, red);
        return;
    }
    const float mean = mean_i[c], rstd = rstd_i[c], g = gamma[c];
    // batch_stats == 0: the layer normalised with its moving statistics (inference mode, or a frozen head: Keras runs a
    // non-trainable BatchNormalization in inference mode) -- mean / variance are constants, so dx = g * rstd * dy
    const float s1 = batch_stats ? acc[c] : 0.0f, s2 = batch_stats ? acc[Dh + c] : 0.0f;
    if (batch_stats && blockIdx.y == 0 && r == 0) { g_gamma[c] += s2; g_beta[c] += s1; }
    const float invM = 1.0f / (float)M;
    const int m0 = blockIdx.y * BN_ROWS, m1 = min(M, m0 + BN_ROWS);
    float bsum = 0.0f;
    for (int m = m0 + r; m < m1; m += 8) {
        const size_t off = (size_t)m * Dh + c;
        const float hv = h[off];
        const float xh = (hv - mean) * rstd;
        const float dx = g * rstd * (d_hn[off] - s1 * invM - xh * s2 * invM);
        const float o = hv > 0.0f ? maybe_tf32(dx, round_out) : 0.0f;  // ReLU backward (h is the post-ReLU activation)
        d_h[off] = o;
        bsum += o;
    }
    if (g_bias) {                       // bias gradient of the Dense in front of the BatchNorm
        bsum = col_reduce8(bsum, red);
        if (r == 0) atomicAdd(&g_bias[c], bsum);
    }
}

int launch_bn_fwd(int M, int Dh, const float *h, const float *gamma, const float *beta, float *moving_mean,
                  float *moving_var, float eps, float momentum, int training, float *acc, float *hn, float *mean, float *rstd,
                  cudaStream_t s)
{
    BDETR_REQUIRE(M > 0 && Dh > 0 && acc, BDETR_E_BAD_SHAPE, "bad BatchNorm arguments");
    dim3 grid(ceil_div(Dh, 32), ceil_div(M, BN_ROWS));
    if (training) {
        launch_k(bn_stats_kernel, grid, 256, 0, s, M, Dh, h, acc);
        BDETR_CHECK_LAUNCH("bn_stats_kernel");
    }
    launch_k(bn_apply_kernel, grid, 256, 0, s, M, Dh, h, gamma, beta, moving_mean, moving_var, eps, momentum, training, acc, hn, mean, rstd);
    BDETR_CHECK_LAUNCH("bn_apply_kernel");
    return BDETR_OK;
}
int launch_bn_relu_bwd(int M, int Dh, const float *h, const float *d_hn, const float *gamma, const float *mean,
                       const float *rstd, float *acc, float *d_h, float *g_gamma, float *g_beta, float *g_bias, int round_out,
                       int batch_stats, cudaStream_t s)
{
    BDETR_REQUIRE(M > 0 && Dh > 0 && acc, BDETR_E_BAD_SHAPE, "bad BatchNorm arguments");
    dim3 grid(ceil_div(Dh, 32), ceil_div(M, BN_ROWS));
    if (batch_stats) {
        BDETR_CUDA(cudaMemsetAsync(acc, 0, sizeof(float) * 2 * Dh, s));
        launch_k(bn_bwd_stats_kernel, grid, 256, 0, s, M, Dh, h, d_hn, mean, rstd, acc);
        BDETR_CHECK_LAUNCH("bn_bwd_stats_kernel");
    }
    launch_k(bn_relu_bwd_apply_kernel, grid, 256, 0, s, M, Dh, h, d_hn, gamma, mean, rstd, acc, d_h, g_gamma, g_beta, g_bias, round_out,
             batch_stats);
    BDETR_CHECK_LAUNCH("bn_relu_bwd_apply_kernel");
    return BDETR_OK;
}

// ------------------------------------------------------------------------------------------
// head activations + boosted running sum.  One warp per row.
// kind 0: softmax, 1: sigmoid, 2: 3*sigmoid(x/100) - 1
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256)
head_act_fwd_kernel(int M, int N, int kind, float mult, float *__restrict__ act, float *__restrict__ cum, int cum_init)
{
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    float *a = act + (size_t)row * N;
    float *c = cum + (size_t)row * N;
    if (kind == 0) {
        float mx = -CUDART_INF_F;
        for (int n = lane; n < N; n += 32) mx = fmaxf(mx, a[n]);
        mx = warp_max(mx);
        float s = 0.0f;
        for (int n = lane; n < N; n += 32) s += expf(a[n] - mx);
        s = warp_sum(s);
        for (int n = lane; n < N; n += 32) {
            const float p = expf(a[n] - mx) / s;
            a[n] = p;
            c[n] = (cum_init ? 0.0f : c[n]) + mult * p;
        }
    } else {
        for (int n = lane; n < N; n += 32) {
            const float v = kind == 1 ? sigmoidf_(a[n]) : 3.0f * sigmoidf_(a[n] / 100.0f) - 1.0f;
            a[n] = v;
            c[n] = (cum_init ? 0.0f : c[n]) + mult * v;
        }
    }
}

__global__ void __launch_bounds__(256)
head_act_bwd_kernel(int M, int N, int kind, float mult, const float *__restrict__ act, const float *__restrict__ d_cum,
                    float *__restrict__ d_logits)
{
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const float *a = act + (size_t)row * N;
    const float *g = d_cum + (size_t)row * N;
    float *d = d_logits + (size_t)row * N;
    if (kind == 0) {
        float dot = 0.0f;
        for (int n = lane; n < N; n += 32) dot = fmaf(g[n], a[n], dot);
        dot = warp_sum(dot);
        for (int n = lane; n < N; n += 32) d[n] = mult * a[n] * (g[n] - dot);
    } else if (kind == 1) {
        for (int n = lane; n < N; n += 32) d[n] = mult * g[n] * a[n] * (1.0f - a[n]);
    } else {
        for (int n = lane; n < N; n += 32) {
            const float sg = (a[n] + 1.0f) / 3.0f;
            d[n] = mult * g[n] * 3.0f * sg * (1.0f - sg) / 100.0f;
        }
    }
}

int launch_head_act_fwd(int M, int N, int kind, float mult, float *act, float *cum, int cum_init, cudaStream_t s)
{
    BDETR_REQUIRE(M > 0 && N > 0 && kind >= 0 && kind <= 2, BDETR_E_BAD_SHAPE, "bad head activation arguments");
    launch_k(head_act_fwd_kernel, ceil_div(M, 8), 256, 0, s, M, N, kind, mult, act, cum, cum_init);
    BDETR_CHECK_LAUNCH("head_act_fwd_kernel");
    return BDETR_OK;
}
int launch_head_act_bwd(int M, int N, int kind, float mult, const float *act, const float *d_cum, float *d_logits, cudaStream_t s)
{
    BDETR_REQUIRE(M > 0 && N > 0 && kind >= 0 && kind <= 2, BDETR_E_BAD_SHAPE, "bad head activation arguments");
    launch_k(head_act_bwd_kernel, ceil_div(M, 8), 256, 0, s, M, N, kind, mult, act, d_cum, d_logits);
    BDETR_CHECK_LAUNCH("head_act_bwd_kernel");
    return BDETR_OK;
}

}  // namespace bdetr
