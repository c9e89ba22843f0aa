// BackboneNeck (reference backbone.py:66-95): BatchNorm -> 1x1 Conv2D(1792 -> 256, lecun_normal, tanh)
// -> BatchNorm on the [B, rows, cols, Cin] feature map of the (out-of-scope, frozen) EfficientNet backbone -- the step
// immediately before the hot path.  A 1x1 convolution over a channels-last map IS a Dense layer over its M = B*rows*cols
// pixels, so the block runs as ONE tcgen05 GEMM [M,Cin] x [Cin,256] with both normalisations folded around it:
//   BN1 is folded INTO the weights:   (x s1 + h1) W + b = x (diag(s1) W) + (h1 W + b),  s1 = rstd1 gamma1, h1 = beta1 - mean1 s1
//   tanh runs in the GEMM epilogue;   BN2 is one fused scale / shift pass that also rounds to tf32 (its output is block
//   0's GEMM operand), so the [M, Cin] input is read exactly once per direction and nothing of size [M, Cin] is written.
// Batch statistics (training): deterministic two-level column reductions.  Backward (the neck trains, the backbone is
// frozen: `detection_model.EncoderBackbone.trainable = False`, notebook cell 30 -> no input gradient):
//   d_t = BN2'(d_y), d_u = d_t (1 - t^2), gW' = x^T d_u (tcgen05 wgrad), gb' = colsum(d_u), then the fold is undone:
//   gW[c,n] = s1[c] gW'[c,n] + h1[c] gb'[n];  gb = gb';  g_s1[c] = sum_n gW'[c,n] W[c,n];  g_h1[c] = sum_n gb'[n] W[c,n];
//   g_gamma1 = rstd1 (g_s1 - mean1 g_h1);  g_beta1 = g_h1.
#include "kernels.cuh"

using namespace bdetr;
#define API extern "C" __attribute__((visibility("default")))
#define TRY(x) do { int rc__ = (x); if (rc__ != BDETR_OK) return rc__; } while (0)

namespace bdetr {

// partial column sums over 128-row chunks, pivoted on row 0: part[chunk][0][c] = sum (x - x0), part[chunk][1][c] = sum (x - x0)^2
__global__ void __launch_bounds__(256)
col_stats_kernel(int M, int C, const float *__restrict__ x, float *__restrict__ part)
{
    pdl_sync();
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), r = threadIdx.x >> 5;
    const bool live = c < C;
    const int m0 = blockIdx.y * 128, m1 = min(M, m0 + 128);
    const float pivot = live ? x[c] : 0.0f;
    float s1 = 0.0f, s2 = 0.0f;
    if (live) for (int m = m0 + r; m < m1; m += 8) { const float d = x[(size_t)m * C + c] - pivot; s1 += d; s2 = fmaf(d, d, s2); }
    __syncthreads(); red[r][threadIdx.x & 31] = s1; __syncthreads();
    float t1 = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t1 += red[j][threadIdx.x & 31];
    __syncthreads(); red[r][threadIdx.x & 31] = s2; __syncthreads();
    float t2 = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t2 += red[j][threadIdx.x & 31];
    if (live && r == 0) { part[((size_t)blockIdx.y * 2) * C + c] = t1; part[((size_t)blockIdx.y * 2 + 1) * C + c] = t2; }
}

// Keras BatchNormalization statistics -> (mean, rstd, scale = rstd gamma, shift = beta - mean scale); training also
// updates the moving statistics (momentum .99, biased variance, eps 1e-3)
__global__ void __launch_bounds__(256)
bn_finalize_kernel(int M, int C, const float *__restrict__ x, const float *__restrict__ part, int chunks, const float *__restrict__ gamma,
                   const float *__restrict__ beta, float *__restrict__ moving_mean, float *__restrict__ moving_var, float eps, float momentum,
                   int training, float *__restrict__ mean_o, float *__restrict__ rstd_o, float *__restrict__ scale, float *__restrict__ shift)
{
    pdl_sync();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= C) return;
    float mean, var;
    if (training) {
        float t1 = 0.0f, t2 = 0.0f;
        for (int j = 0; j < chunks; ++j) { t1 += part[((size_t)j * 2) * C + c]; t2 += part[((size_t)j * 2 + 1) * C + c]; }     // fixed order
        const float a1 = t1 / (float)M, a2 = t2 / (float)M;
        mean = x[c] + a1;
        var = fmaxf(a2 - a1 * a1, 0.0f);
        moving_mean[c] = moving_mean[c] * momentum + mean * (1.0f - momentum);
        moving_var[c] = moving_var[c] * momentum + var * (1.0f - momentum);
    } else {
        mean = moving_mean[c];
        var = moving_var[c];
    }
    const float rstd = rsqrtf(var + eps);
    mean_o[c] = mean; rstd_o[c] = rstd;
    const float sc = rstd * gamma[c];
    scale[c] = sc;
    shift[c] = beta[c] - mean * sc;
}

// W'[c,n] = tf32(s1[c] W[c,n]);  bpart[cta][n] = sum over the CTA's 32 input channels of h1[c] W[c,n]   (N <= 256)
__global__ void __launch_bounds__(256)
neck_fold_kernel(int C, int N, const float *__restrict__ W, const float *__restrict__ s1, const float *__restrict__ h1,
                 float *__restrict__ Wf, float *__restrict__ bpart)
{
    pdl_sync();
    const int n = threadIdx.x;
    const int c0 = blockIdx.x * 32, c1 = min(C, c0 + 32);
    float acc = 0.0f;
    if (n < N) for (int c = c0; c < c1; ++c) {
        const float w = W[(size_t)c * N + n];
        Wf[(size_t)c * N + n] = tf32_rn(s1[c] * w);
        acc = fmaf(h1[c], w, acc);
    }
    if (n < N) bpart[(size_t)blockIdx.x * N + n] = acc;
}
// b'[n] = b[n] + sum of the partials in fixed order
__global__ void __launch_bounds__(256)
neck_bias_kernel(int parts, int N, const float *__restrict__ b, const float *__restrict__ bpart, float *__restrict__ bf)
{
    pdl_sync();
    const int n = threadIdx.x;
    if (n >= N) return;
    float acc = b[n];
    for (int p = 0; p < parts; ++p) acc += bpart[(size_t)p * N + n];
    bf[n] = acc;
}

// y = tf32?(t * scale + shift)   (BatchNorm 2 as one affine pass)
__global__ void __launch_bounds__(256)
affine_cols_kernel(size_t n4, int N4, const float4 *__restrict__ t, const float4 *__restrict__ scale, const float4 *__restrict__ shift,
                   float4 *__restrict__ y, int round_out)
{
    pdl_sync();
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (size_t)gridDim.x * blockDim.x) {
        const float4 v = t[e], sc = scale[e % N4], sh = shift[e % N4];
        float4 o = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
        if (round_out) { o.x = tf32_rn(o.x); o.y = tf32_rn(o.y); o.z = tf32_rn(o.z); o.w = tf32_rn(o.w); }
        y[e] = o;
    }
}

// backward of BN2 (batch statistics) + tanh: partial sums of d_y and d_y * that over 128-row chunks
__global__ void __launch_bounds__(256)
neck_bn2_bwd_stats_kernel(int M, int N, const float *__restrict__ t, const float *__restrict__ d_y, const float *__restrict__ mean,
                          const float *__restrict__ rstd, float *__restrict__ part)
{
    pdl_sync();
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), r = threadIdx.x >> 5;
    const bool live = c < N;
    const int m0 = blockIdx.y * 128, m1 = min(M, m0 + 128);
    const float mu = live ? mean[c] : 0.0f, rs = live ? rstd[c] : 0.0f;
    float s1 = 0.0f, s2 = 0.0f;
    if (live) for (int m = m0 + r; m < m1; m += 8) {
        const float dy = d_y[(size_t)m * N + c];
        s1 += dy;
        s2 = fmaf(dy, (t[(size_t)m * N + c] - mu) * rs, s2);
    }
    __syncthreads(); red[r][threadIdx.x & 31] = s1; __syncthreads();
    float t1 = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t1 += red[j][threadIdx.x & 31];
    __syncthreads(); red[r][threadIdx.x & 31] = s2; __syncthreads();
    float t2 = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t2 += red[j][threadIdx.x & 31];
    if (live && r == 0) { part[((size_t)blockIdx.y * 2) * N + c] = t1; part[((size_t)blockIdx.y * 2 + 1) * N + c] = t2; }
}

// d_u = tf32( gamma2 rstd2 (d_y - S1/M - that S2/M) (1 - t^2) ); g_gamma2 += S2, g_beta2 += S1 (CTA row 0 only)
__global__ void __launch_bounds__(256)
neck_bn2_bwd_apply_kernel(int M, int N, const float *__restrict__ t, const float *__restrict__ d_y, const float *__restrict__ mean,
                          const float *__restrict__ rstd, const float *__restrict__ gamma, const float *__restrict__ part, int chunks,
                          int batch_stats, float *__restrict__ d_u, float *__restrict__ g_gamma, float *__restrict__ g_beta)
{
    pdl_sync();
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), r = threadIdx.x >> 5;
    if (c >= N) return;
    float S1 = 0.0f, S2 = 0.0f;
    for (int j = 0; j < chunks; ++j) { S1 += part[((size_t)j * 2) * N + c]; S2 += part[((size_t)j * 2 + 1) * N + c]; }
    if (blockIdx.y == 0 && r == 0 && g_gamma) { g_gamma[c] += S2; g_beta[c] += S1; }
    const float invM = 1.0f / (float)M;
    const float a1 = batch_stats ? S1 * invM : 0.0f, a2 = batch_stats ? S2 * invM : 0.0f;
    const float mu = mean[c], rs = rstd[c], g = gamma[c];
    const int m0 = blockIdx.y * 128, m1 = min(M, m0 + 128);
    for (int m = m0 + r; m < m1; m += 8) {
        const size_t off = (size_t)m * N + c;
        const float tv = t[off];
        const float th = (tv - mu) * rs;
        d_u[off] = tf32_rn(g * rs * (d_y[off] - a1 - th * a2) * (1.0f - tv * tv));
    }
}

// undo the fold: one warp per input channel c (lane = output column)
__global__ void __launch_bounds__(256)
neck_unfold_grads_kernel(int C, int N, const float *__restrict__ W, const float *__restrict__ gWf, const float *__restrict__ gbf,
                         const float *__restrict__ s1, const float *__restrict__ h1, const float *__restrict__ mean1, const float *__restrict__ rstd1,
                         float *__restrict__ gW, float *__restrict__ g_gamma1, float *__restrict__ g_beta1)
{
    pdl_sync();
    const int lane = threadIdx.x & 31, c = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c >= C) return;
    const float sc = s1[c], sh = h1[c];
    float gs = 0.0f, gh = 0.0f;
    for (int n = lane; n < N; n += 32) {
        const float w = W[(size_t)c * N + n], gwf = gWf[(size_t)c * N + n], gb = gbf[n];
        gW[(size_t)c * N + n] += sc * gwf + sh * gb;
        gs = fmaf(gwf, w, gs);
        gh = fmaf(gb, w, gh);
    }
    gs = warp_sum(gs); gh = warp_sum(gh);
    if (lane == 0) { g_gamma1[c] += rstd1[c] * (gs - mean1[c] * gh); g_beta1[c] += gh; }
}

static inline int neck_grid(size_t n) { size_t g = (n + 255) / 256; return (int)(g < 148 * 8 ? (g ? g : 1) : 148 * 8); }

}  // namespace bdetr

API int bdetr_backbone_neck_fwd(int M, int Cin, int N, const float *x_tc, const bdetr_neck_params *w, float bn_eps, float bn_momentum,
                                int training, float *out, const bdetr_neck_saved *sv, int round_out, void *stream)
{
    BDETR_REQUIRE(M > 0 && Cin >= 32 && Cin % 4 == 0 && N >= 64 && N <= 256 && N % 4 == 0, BDETR_E_BAD_SHAPE, "bad shape (N <= 256)");
    BDETR_REQUIRE(x_tc && w && out && sv && sv->t && sv->wf && sv->bf && sv->part && sv->stat1 && sv->stat2, BDETR_E_NULL, "null pointer");
    ModeScope tc(BDETR_MODE_TF32);
    cudaStream_t s = as_stream(stream);
    const int chunks = ceil_div(M, 128);
    float *mean1 = sv->stat1, *rstd1 = mean1 + Cin, *s1 = rstd1 + Cin, *h1 = s1 + Cin;
    float *mean2 = sv->stat2, *rstd2 = mean2 + N, *s2 = rstd2 + N, *h2 = s2 + N;
    if (training) {
        launch_k(col_stats_kernel, dim3(ceil_div(Cin, 32), chunks), 256, 0, s, M, Cin, x_tc, sv->part);
        BDETR_CHECK_LAUNCH("col_stats_kernel");
    }
    launch_k(bn_finalize_kernel, ceil_div(Cin, 256), 256, 0, s, M, Cin, x_tc, (const float *)sv->part, chunks, (const float *)w->bn1_gamma,
             (const float *)w->bn1_beta, w->bn1_moving_mean, w->bn1_moving_var, bn_eps, bn_momentum, training, mean1, rstd1, s1, h1);
    BDETR_CHECK_LAUNCH("bn_finalize_kernel");
    const int parts = ceil_div(Cin, 32);
    float *bpart = sv->part;                                 // the statistics partials are consumed: the scratch is reused
    launch_k(neck_fold_kernel, parts, 256, 0, s, Cin, N, (const float *)w->conv_w, (const float *)s1, (const float *)h1, sv->wf, bpart);
    BDETR_CHECK_LAUNCH("neck_fold_kernel");
    launch_k(neck_bias_kernel, 1, 256, 0, s, parts, N, (const float *)w->conv_b, (const float *)bpart, sv->bf);
    BDETR_CHECK_LAUNCH("neck_bias_kernel");
    GroupedGemm g;
    g.M = M; g.N = N; g.K = Cin; g.lda = Cin; g.ldb = N; g.ldc = N; g.act = 2;
    g.A[0] = x_tc; g.B[0] = sv->wf; g.bias[0] = sv->bf; g.C[0] = sv->t;
    TRY(launch_gemm_umma_grouped(g, s));
    if (training) {
        launch_k(col_stats_kernel, dim3(ceil_div(N, 32), chunks), 256, 0, s, M, N, (const float *)sv->t, sv->part);
        BDETR_CHECK_LAUNCH("col_stats_kernel");
    }
    launch_k(bn_finalize_kernel, ceil_div(N, 256), 256, 0, s, M, N, (const float *)sv->t, (const float *)sv->part, chunks, (const float *)w->bn2_gamma,
             (const float *)w->bn2_beta, w->bn2_moving_mean, w->bn2_moving_var, bn_eps, bn_momentum, training, mean2, rstd2, s2, h2);
    BDETR_CHECK_LAUNCH("bn_finalize_kernel");
    const size_t n4 = (size_t)M * N / 4;
    launch_k(affine_cols_kernel, neck_grid(n4), 256, 0, s, n4, N / 4, (const float4 *)sv->t, (const float4 *)s2, (const float4 *)h2, (float4 *)out, round_out);
    BDETR_CHECK_LAUNCH("affine_cols_kernel");
    return BDETR_OK;
}

API int bdetr_backbone_neck_bwd(int M, int Cin, int N, const float *x_tc, const bdetr_neck_params *w, int training,
                                const bdetr_neck_saved *sv, const float *d_out, const bdetr_neck_params *gw, float *d_u, float *gwf, void *stream)
{
    BDETR_REQUIRE(M > 0 && Cin >= 32 && Cin % 4 == 0 && N >= 64 && N <= 256 && N % 4 == 0, BDETR_E_BAD_SHAPE, "bad shape (N <= 256)");
    BDETR_REQUIRE(x_tc && w && sv && d_out && gw && d_u && gwf, BDETR_E_NULL, "null pointer");
    ModeScope tc(BDETR_MODE_TF32);
    cudaStream_t s = as_stream(stream);
    const int chunks = ceil_div(M, 128);
    float *mean1 = sv->stat1, *rstd1 = mean1 + Cin, *s1 = rstd1 + Cin, *h1 = s1 + Cin;
    float *mean2 = sv->stat2, *rstd2 = mean2 + N;
    dim3 grid(ceil_div(N, 32), chunks);
    launch_k(neck_bn2_bwd_stats_kernel, grid, 256, 0, s, M, N, (const float *)sv->t, d_out, (const float *)mean2, (const float *)rstd2, sv->part);
    BDETR_CHECK_LAUNCH("neck_bn2_bwd_stats_kernel");
    launch_k(neck_bn2_bwd_apply_kernel, grid, 256, 0, s, M, N, (const float *)sv->t, d_out, (const float *)mean2, (const float *)rstd2,
             (const float *)w->bn2_gamma, (const float *)sv->part, chunks, training, d_u, gw->bn2_gamma, gw->bn2_beta);
    BDETR_CHECK_LAUNCH("neck_bn2_bwd_apply_kernel");
    // gW' = x^T d_u (overwrite) and gb' = colsum(d_u) into the first N floats of sv->bf's gradient twin (gwf tail)
    float *gbf = gwf + (size_t)Cin * N;
    BDETR_CUDA(cudaMemsetAsync(gbf, 0, sizeof(float) * N, s));
    GroupedGemm g;
    g.M = Cin; g.N = N; g.K = M; g.TA = true; g.lda = Cin; g.ldb = N; g.ldc = N;
    g.A[0] = x_tc; g.B[0] = d_u; g.C[0] = gwf;
    TRY(launch_gemm_umma_grouped(g, s));
    BatchReduce r;
    r.B = 1; r.D = N; r.n = 1; r.src[0] = d_u; r.rows[0] = M; r.colsum[0] = gbf;
    TRY(launch_batch_reduce(r, s));
    launch_k(neck_unfold_grads_kernel, ceil_div(Cin, 8), 256, 0, s, Cin, N, (const float *)w->conv_w, (const float *)gwf, (const float *)gbf,
             (const float *)s1, (const float *)h1, (const float *)mean1, (const float *)rstd1, gw->conv_w, gw->bn1_gamma, gw->bn1_beta);
    BDETR_CHECK_LAUNCH("neck_unfold_grads_kernel");
    TRY(launch_accumulate((size_t)N, gbf, gw->conv_b, s));
    return BDETR_OK;
}
