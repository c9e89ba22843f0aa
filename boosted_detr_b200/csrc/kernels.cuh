// Internal launchers shared by the C-ABI entry points in ops.cu.
#pragma once
#include "common.cuh"

namespace bdetr {

int current_mode();
// Entry points that ARE one compute mode (the fused tensor-core path) pin it for the calling thread while they run.
struct ModeScope {
    explicit ModeScope(int mode);
    ~ModeScope();
    int saved;
};

// Fork / join of independent kernel chains onto library-owned auxiliary streams.  The kernels of this workload
// are short and at most ~100 CTAs wide, so independent ones (wgrad next to dgrad, the q/k/v projections) are
// run side by side.  Every entry point joins what it forked before it returns, so callers still see plain
// stream semantics on the stream they passed, and the fork/join is captured into CUDA graphs as parallel branches.
struct Branches {
    explicit Branches(cudaStream_t main_stream);
    cudaStream_t fork(int i);       // aux stream i (0..2), ordered after everything enqueued on main so far
    int join();                     // main waits for every aux stream forked since the last join
    int join_deferrable();          // the same, unless the caller asked to join parameter-gradient chains itself (bdetr_join)
    cudaStream_t main;
    int base, used, rc;
    bool on, shared;
};
bool concurrency_enabled();
int join_pending(cudaStream_t main_stream, cudaStream_t waiter);

// C[M,N] = (beta ? C : 0) + op(A)[M,K] @ op(B)[K,N] (+bias[n]) ; act 1 = relu ;
// relu_mask != NULL: C[m,n] = 0 where relu_mask[m*ldc+n] <= 0 (backward of relu, applied last).
// TA: A stored [K,M] (lda = M-stride of k rows); TB: B stored [N,K].
// round_out: store the result rounded to tf32 (tensor-core mode, outputs that feed another GEMM).
int launch_gemm(int M, int N, int K, const float *A, int lda, bool TA, const float *B, int ldb, bool TB,
                const float *bias, int act, const float *relu_mask, int beta, int round_out, float *C, int ldc,
                cudaStream_t s);

// tcgen05 / TMA / TMEM tensor-core path (gemm_umma.cu), same contract as launch_gemm
bool umma_gemm_eligible(int M, int N, int K, const float *A, int lda, bool TA, const float *B, int ldb, bool TB, int ldc);
int launch_gemm_umma(int M, int N, int K, const float *A, int lda, bool TA, const float *B, int ldb, bool TB,
                     const float *bias, int act, const float *relu_mask, int beta, int round_out, float *C, int ldc,
                     cudaStream_t s);

// Up to three tcgen05 GEMMs of identical shape [M,N,K] in one launch (gemm_umma.cu).
//   sum_groups = false: independent outputs C[g] = op(A[share_a ? 0 : g]) op(B[g]) (+ bias[g]) (+ rowtab[g]) ...
//   sum_groups = true : ONE output C[0] = sum_g op(A[g]) op(B[g])  (the k loop runs over all groups)
// rowtab[g]: out[r,n] += rowtab[g][(r % rowtab_period) * rowtab_ld + n]; colsum[g][n] += column sums of the stored output.
struct GroupedGemm {
    int M = 0, N = 0, K = 0, groups = 1;
    bool TA = false, TB = false, share_a = false, sum_groups = false;
    const float *A[3] = {nullptr, nullptr, nullptr}; int lda = 0;
    const float *B[3] = {nullptr, nullptr, nullptr}; int ldb = 0;
    float *C[3] = {nullptr, nullptr, nullptr}; int ldc = 0;
    const float *bias[3] = {nullptr, nullptr, nullptr};
    const float *rowtab[3] = {nullptr, nullptr, nullptr}; int rowtab_period = 0, rowtab_ld = 0;
    float *colsum[3] = {nullptr, nullptr, nullptr};
    int act = 0; const float *relu_mask = nullptr; int beta = 0; int round_out = 0;
};
int launch_gemm_umma_grouped(const GroupedGemm &g, cudaStream_t s);

// dst[n] += sum_m src[m, n]
int launch_colsum_acc(int M, int N, const float *src, float *dst, cudaStream_t s);

int launch_attention_fwd(int B, int H, int Lq, int Lk, int d, const float *qp, const float *kp, const float *vp,
                         float *o, float *lse, int round_out, cudaStream_t s);
int launch_attention_bwd(int B, int H, int Lq, int Lk, int d, const float *qp, const float *kp, const float *vp,
                         const float *o, const float *lse, const float *d_o, float *delta,
                         float *d_qp, float *d_kp, float *d_vp, int round_out, cudaStream_t s);

// multi-stream variant for long sequences (attention_umma_ms.cu)
bool attention_umma_ms_eligible(int B, int H, int Lq, int Lk);
int launch_attention_fwd_umma_ms(int B, int H, int Lq, int Lk, int d, const float *qp, const float *kp, const float *vp,
                                 float *o, float *lse, int round_out, cudaStream_t s);
extern int g_force_attention_kernel;
// fp16-operand variant (attention_umma_ms_f16.cu): q / k / v are cast into ws16 (attention_f16_workspace_bytes) first
size_t attention_f16_workspace_bytes(int B, int H, int Lq, int Lk, int d);
int launch_attention_fwd_umma_ms_f16(int B, int H, int Lq, int Lk, int d, const float *qp, const float *kp, const float *vp,
                                     void *ws16, float *o, float *lse, int round_out, cudaStream_t s);
bool attention_f16_enabled();
// tcgen05 flash attention forward (attention_umma.cu)
bool attention_umma_eligible(int B, int H, int Lq, int Lk, int d, const float *qp, const float *kp, const float *vp);
int launch_attention_fwd_umma(int B, int H, int Lq, int Lk, int d, const float *qp, const float *kp, const float *vp,
                              float *o, float *lse, int round_out, cudaStream_t s);

bool attention_bwd_umma_eligible(int B, int H, int Lq, int Lk, int d, const float *qp, const float *kp, const float *vp,
                                 const float *d_o);
int launch_attention_bwd_umma(int B, int H, int Lq, int Lk, int d, const float *qp, const float *kp, const float *vp,
                              const float *o, const float *lse, const float *d_o, float *delta,
                              float *d_qp, float *d_kp, float *d_vp, int round_out, cudaStream_t s);

// z = resid + dropout(z) (in place), out = LN(z); saves mean/rstd.
int launch_res_ln_fwd(int M, int D, const float *resid, float *z, const float *gamma, const float *beta, float eps,
                      float rate, uint32_t key, const uint32_t *seed_dev, float *out, float *mean, float *rstd, int round_out,
                      cudaStream_t s);
// d_z -> d_resid (overwrite/accumulate) and d_a = dropout'(d_z); g_gamma/g_beta accumulated; g_bias (nullable)
// accumulates the column sums of d_a = the bias gradient of the Dense layer that produced the LayerNorm input.
int launch_res_ln_bwd(int M, int D, const float *d_out, const float *z, const float *mean, const float *rstd,
                      const float *gamma, float rate, uint32_t key, const uint32_t *seed_dev, float *d_resid, int acc_resid,
                      float *d_a, float *g_gamma, float *g_beta, float *g_bias, int round_out, cudaStream_t s);

int launch_res_ln_bcast_fwd(int M, int P, int D, const float *resid, const float *a, float *z, const float *gamma, const float *beta,
                            float eps, float rate, uint32_t key, const uint32_t *seed_dev, float *out, float *mean, float *rstd,
                            int round_out, cudaStream_t s);
struct BatchReduce {
    int n = 0, B = 0, D = 0;
    const float *src[4] = {nullptr, nullptr, nullptr, nullptr}; int rows[4] = {0, 0, 0, 0};
    float *sum[4] = {nullptr, nullptr, nullptr, nullptr};
    float *colsum[4] = {nullptr, nullptr, nullptr, nullptr};
};
int launch_batch_reduce(const BatchReduce &a, cudaStream_t s);
// Dense + bias + Dropout + residual (+ pos rows) + LayerNorm in one tcgen05 kernel (gemm_ln.cu); z may be NULL (not saved)
int launch_gemm_ln(int M, int K, const float *A, const float *W, const float *bias, const float *resid, const float *pos,
                   int pos_period, const float *gamma, const float *beta, float eps, float rate, uint32_t key,
                   const uint32_t *seed_dev, float *z, float *out, float *mean, float *rstd, int round_out, cudaStream_t s);

int launch_add_rows_fwd(int B, int L, int D, const float *x, const float *pos, float *out, int round_out, cudaStream_t s);
int launch_batch_sum_acc(int B, int L, int D, const float *src, float *dst, cudaStream_t s);
int launch_tile_rows(int B, int L, int D, const float *src, float *dst, int round_out, cudaStream_t s);
int launch_round_tf32(size_t n, const float *src, float *dst, cudaStream_t s);
int launch_accumulate(size_t n, const float *x, float *y, cudaStream_t s);
int launch_suffix_sum(int n, size_t len, float *buf, cudaStream_t s);

// acc: [2*Dh] scratch for the cross-CTA column sums
int launch_bn_fwd(int M, int Dh, const float *h, const float *gamma, const float *beta, float *moving_mean,
                  float *moving_var, float eps, float momentum, int training, float *acc, float *hn, float *mean, float *rstd,
                  cudaStream_t s);
// d_h = relu'(h) * BN_backward(d_hn); g_gamma/g_beta accumulated
int launch_bn_relu_bwd(int M, int Dh, const float *h, const float *d_hn, const float *gamma, const float *mean,
                       const float *rstd, float *acc, float *d_h, float *g_gamma, float *g_beta, float *g_bias, int round_out,
                       int batch_stats, cudaStream_t s);
// act in place on logits [M,N]; cum = (init ? 0 : cum) + mult*act
int launch_head_act_fwd(int M, int N, int kind, float mult, float *act, float *cum, int cum_init, cudaStream_t s);
int launch_head_act_bwd(int M, int N, int kind, float mult, const float *act, const float *d_cum, float *d_logits,
                        cudaStream_t s);

}  // namespace bdetr
