// Dense + bias + Dropout + residual (+ positional rows) + LayerNorm in ONE tcgen05 kernel:
//     z   = resid (+ pos[row % pos_period]) + Dropout(A W + bias)          (reference transformers.py:101,147-148 / :187-190)
//     out = LayerNorm_eps(z) * gamma + beta                                 (:149 / :191)
// A [M,K] K-major activations, W a Keras kernel [K,256] (MN-major B operand), N = 256 = the model width, so one CTA owns
// whole rows: 128 rows x 256 columns, fp32 accumulator in 256 TMEM columns, and the row statistics never leave the CTA.
// Eight warps, all of them epilogue warps; warp 0 first runs the TMA producer loop and warp 1 (the TMEM allocator) the MMA
// issue loop -- the epilogue cannot start before the last MMA anyway.  Epilogue phase 1: every thread moves the 128
// accumulator columns of its TMEM lane (quarter warp & 3, column half warp / 4) into swizzled 32 x 32 boxes in the idle
// pipeline stages.  Phase 2: one warp per row, lane = every 32nd column, exactly the access pattern of the stand-alone
// LayerNorm kernel (coalesced 128-byte reads of the residual / positional rows, two-pass mean / centred variance by
// warp shuffles, coalesced writes) -- but fed from shared memory: the GEMM result never goes to HBM.
#include <cstring>
#include "umma.cuh"

namespace bdetr {

constexpr int LN_BM = 128, LN_N = 256, LN_BK = 32, LN_STAGES = 4, LN_THREADS = 256;
constexpr uint32_t LN_A_STAGE = LN_BM * LN_BK * 4;      // 16 KB
constexpr uint32_t LN_B_STAGE = LN_N * LN_BK * 4;       // 32 KB

struct LnMaps { CUtensorMap a, b; };
struct LnEpilogue {
    int M, num_kb;
    const float *bias, *resid, *pos; int pos_period;
    const float *gamma, *beta; float eps;
    float keep_scale; uint32_t thresh, key; const uint32_t *seed_dev;
    float *z, *out, *mean, *rstd;
    int write_z, round_out;
};

__global__ void __launch_bounds__(LN_THREADS, 1)
gemm_ln_kernel(const __grid_constant__ LnMaps maps, const __grid_constant__ LnEpilogue ep)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem + LN_STAGES * LN_A_STAGE;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_b + LN_STAGES * LN_B_STAGE);
    uint64_t *empty = full + LN_STAGES;
    uint64_t *accum_full = empty + LN_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_full + 1);
    float *s_bias = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~uintptr_t(15));
    float *s_gamma = s_bias + LN_N, *s_beta = s_gamma + LN_N;
    
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * LN_BM;
    const int nkb = ep.num_kb;

    if (threadIdx.x == 0) {
        for (int s = 0; s < LN_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, LN_N);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();
    s_bias[threadIdx.x] = ep.bias ? ep.bias[threadIdx.x] : 0.0f;
    s_gamma[threadIdx.x] = ep.gamma[threadIdx.x];
    s_beta[threadIdx.x] = ep.beta[threadIdx.x];
    __syncthreads();

    if (warp == 0) {
        for (int i = 0; i < nkb; ++i) {
            const int s = i % LN_STAGES;
            if (i >= LN_STAGES) mbar_wait(&empty[s], ((i / LN_STAGES) - 1) & 1);
            const int k0 = i * LN_BK;
            uint8_t *a_dst = smem_a + s * LN_A_STAGE, *b_dst = smem_b + s * LN_B_STAGE;
            if (elect_one()) {
                mbar_expect_tx(&full[s], LN_A_STAGE + LN_B_STAGE);
                tma_load_2d(a_dst, &maps.a, k0, m0, &full[s]);                                    // box {32 k, 128 rows}
#pragma unroll
                for (int j = 0; j < LN_N / 32; ++j) tma_load_2d(b_dst + j * 4096, &maps.b, 32 * j, k0, &full[s]);   // box {32 n, 32 k}
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc_tf32(LN_BM, LN_N, 0, 1);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % LN_STAGES;
            mbar_wait(&full[s], (i / LN_STAGES) & 1);
            tc_fence_after();
            const uint32_t a_base = smem_u32(smem_a + s * LN_A_STAGE), b_base = smem_u32(smem_b + s * LN_B_STAGE);
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < LN_BK / 8; ++j) {
                    const uint64_t a_desc = make_smem_desc(a_base + j * 32, 16, 1024, 2);
                    const uint64_t b_desc = make_smem_desc(b_base + j * 1024, 4096, 512, 1);
                    umma_tf32(tmem_base, a_desc, b_desc, idesc, (i | j) != 0);
                }
                umma_commit(&empty[s]);
                if (i == nkb - 1) umma_commit(accum_full);
            }
            __syncwarp();
        }
    }
    {
        // ---- epilogue phase 1: accumulator rows (one TMEM lane = one row per thread) -> shared memory, as 32 x 32 boxes
        // with 128-byte swizzled rows inside the now idle pipeline stages (box (c, q) = columns 32c.., rows 32q..)
        const int q = warp & 3, hf = warp >> 2;
        mbar_wait(accum_full, 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + hf * 128;
        uint32_t ra[32], rb[32];
        tmem_ld32_issue(taddr, ra);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t *cur = (c & 1) ? rb : ra, *nxt = (c & 1) ? ra : rb;
            tmem_ld32_wait(cur);
            if (c + 1 < 4) tmem_ld32_issue(taddr + (c + 1) * 32, nxt);
            uint8_t *box = smem + (size_t)((hf * 4 + c) * 4 + q) * 4096;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4 *>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(cur[4 * j], cur[4 * j + 1], cur[4 * j + 2], cur[4 * j + 3]);
        }
    }
    __syncthreads();
    {
        // ---- epilogue phase 2: one warp per row (16 rows per warp), lane = column 32c + lane for c = 0..7: coalesced
        // 128-byte reads of the residual / positional rows and 128-byte writes of z / out; row statistics by warp shuffles
        uint32_t key = ep.key;
        if (ep.seed_dev) key = lowbias32(*ep.seed_dev ^ key);
        float bias[8], gam[8], bet[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { bias[c] = s_bias[c * 32 + lane]; gam[c] = s_gamma[c * 32 + lane]; bet[c] = s_beta[c * 32 + lane]; }
        constexpr int RPW = LN_BM / 8;                  // rows per warp
        float rs[2][8], ps[2][8];                       // residual / positional rows, double-buffered one row ahead
        auto load_row = [&](int rloc, float *r8, float *p8) {
            const int row = m0 + rloc;
            if (row < ep.M) {
                const float *rrow = ep.resid + (size_t)row * LN_N + lane;
#pragma unroll
                for (int c = 0; c < 8; ++c) r8[c] = rrow[c * 32];
                if (ep.pos) {
                    const float *prow = ep.pos + (size_t)(row % ep.pos_period) * LN_N + lane;
#pragma unroll
                    for (int c = 0; c < 8; ++c) p8[c] = prow[c * 32];
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) p8[c] = 0.0f;
                }
            }
        };
        load_row(warp * RPW, rs[0], ps[0]);
#pragma unroll 2
        for (int i = 0; i < RPW; ++i) {
            const int rloc = warp * RPW + i, row = m0 + rloc;
            if (i + 1 < RPW) load_row(rloc + 1, rs[(i + 1) & 1], ps[(i + 1) & 1]);
            if (row >= ep.M) break;                     // rows are processed in order: the rest of this warp's rows are out of range too
            const float *r8 = rs[i & 1], *p8 = ps[i & 1];
            const int qq = rloc >> 5, rr = rloc & 31;
            float v[8];
            float sum = 0.0f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float acc = *reinterpret_cast<const float *>(smem + (size_t)(c * 4 + qq) * 4096 + rr * 128 + ((((lane >> 2) ^ (rr & 7)) << 4) | ((lane & 3) << 2)));
                float a = acc + bias[c];
                if (ep.thresh) a = dropout_keep((uint32_t)((size_t)row * LN_N + c * 32 + lane), key, ep.thresh) ? a * ep.keep_scale : 0.0f;
                v[c] = (r8[c] + p8[c]) + a;
                sum += v[c];
            }
            const float mean = warp_sum(sum) * (1.0f / LN_N);
            float sq = 0.0f;
#pragma unroll
            for (int c = 0; c < 8; ++c) { const float d = v[c] - mean; sq = fmaf(d, d, sq); }
            const float rstd = rsqrtf(warp_sum(sq) * (1.0f / LN_N) + ep.eps);
            if (ep.write_z) {
                float *zrow = ep.z + (size_t)row * LN_N + lane;
#pragma unroll
                for (int c = 0; c < 8; ++c) zrow[c * 32] = v[c];
            }
            float *orow = ep.out + (size_t)row * LN_N + lane;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float y = (v[c] - mean) * rstd * gam[c] + bet[c];
                if (ep.round_out) y = tf32_rn(y);
                orow[c * 32] = y;
            }
            if (lane == 0) { ep.mean[row] = mean; ep.rstd[row] = rstd; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, LN_N);
}

int launch_gemm_ln(int M, int K, const float *A, const float *W, const float *bias, const float *resid, const float *pos,
                   int pos_period, const float *gamma, const float *beta, float eps, float rate, uint32_t key,
                   const uint32_t *seed_dev, float *z, float *out, float *mean, float *rstd, int round_out, cudaStream_t s)
{
    BDETR_REQUIRE(M > 0 && K >= LN_BK && K % 4 == 0 && A && W && resid && gamma && beta && out && mean && rstd, BDETR_E_BAD_SHAPE,
                  "bad fused Dense + LayerNorm arguments");
    LnMaps maps;
    memset(&maps, 0, sizeof(maps));
    bool ok = encode_tensor_map_2d(&maps.a, A, M, K, K, LN_BK, LN_BM, false);
    ok = ok && encode_tensor_map_2d(&maps.b, W, K, LN_N, LN_N, 32, LN_BK, true);
    BDETR_REQUIRE(ok, BDETR_E_CUDA, "cuTensorMapEncodeTiled failed");
    LnEpilogue ep;
    memset(&ep, 0, sizeof(ep));
    ep.M = M; ep.num_kb = ceil_div(K, LN_BK);
    ep.bias = bias; ep.resid = resid; ep.pos = pos; ep.pos_period = pos_period > 0 ? pos_period : 1;
    ep.gamma = gamma; ep.beta = beta; ep.eps = eps;
    ep.thresh = dropout_threshold(rate); ep.keep_scale = 1.0f / (1.0f - rate); ep.key = key; ep.seed_dev = seed_dev;
    ep.z = z; ep.out = out; ep.mean = mean; ep.rstd = rstd; ep.write_z = z != nullptr; ep.round_out = round_out;
    const size_t smem = (size_t)LN_STAGES * (LN_A_STAGE + LN_B_STAGE) + (2 * LN_STAGES + 1) * 8 + 64 + (3 * LN_N + 512) * 4 + 1024;
    static bool optin = false;
    if (!optin) {
        BDETR_CUDA(cudaFuncSetAttribute(gemm_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        optin = true;
    }
    launch_k(gemm_ln_kernel, dim3(ceil_div(M, LN_BM)), LN_THREADS, smem, s, maps, ep);
    BDETR_CHECK_LAUNCH("gemm_ln_kernel");
    return BDETR_OK;
}

}  // namespace bdetr
