// Dense + bias + Dropout + residual (+ positional rows) + LayerNorm in ONE tcgen05 kernel:
//     z   = resid (+ pos[row % pos_period]) + Dropout(A W + bias)          (reference transformers.py:101,147-148 / :187-190)
//     out = LayerNorm_eps(z) * gamma + beta                                 (:149 / :191)
// A [M,K] K-major activations, W a Keras kernel [K,256] (MN-major B operand), N = 256 = the model width, so one CTA owns
// whole rows: 128 rows x 256 columns, fp32 accumulator in 256 TMEM columns, and the row statistics never leave the CTA.
// 16 warps, all of them epilogue warps; warp 0 first runs the TMA producer loop and warp 1 (the TMEM allocator) the MMA
// issue loop -- the epilogue cannot start before the last MMA anyway.  Epilogue phase 1: warp w moves two 32 x 32
// accumulator blocks (TMEM lane quarter w & 3, column groups 2 (w / 4), + 1) into swizzled boxes in the idle pipeline stages.
// Phase 2: one warp per row, lane = every 32nd column, exactly the access pattern of the stand-alone
// LayerNorm kernel (coalesced 128-byte reads of the residual / positional rows, two-pass mean / centred variance by
// warp shuffles, coalesced writes) -- but fed from shared memory: the GEMM result never goes to HBM.
#include <cstdlib>
#include <cstring>
#include "umma.cuh"

namespace bdetr {

extern long long *g_umma_timeline;

constexpr int LN_BM = 128, LN_N = 256, LN_BK = 32, LN_STAGES = 4, LN_THREADS = 512;
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
    long long *dbg;       // optional clock64 stamps of CTA 0 (bdetr_debug_set_timeline)
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
    const bool dbg_on = ep.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 64;      // warp 2, lane 0
#define LN_STAMP(slot) do { if (dbg_on) ep.dbg[slot] = clock64(); } while (0)
    LN_STAMP(0);

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
    LN_STAMP(1);
    if (threadIdx.x < LN_N) {
        s_bias[threadIdx.x] = ep.bias ? ep.bias[threadIdx.x] : 0.0f;
        s_gamma[threadIdx.x] = ep.gamma[threadIdx.x];
        s_beta[threadIdx.x] = ep.beta[threadIdx.x];
    }
    __syncthreads();
    LN_STAMP(2);

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
        // ---- epilogue phase 1: 32 warps, warp w moves the 32 x 32 accumulator block (TMEM lane quarter w & 3, column
        // group w >> 2) into its swizzled box in the now idle pipeline stages (box (c, q) = columns 32c.., rows 32q..)
        const int q = warp & 3, cg2 = warp >> 2;        // 16 warps: TMEM lane quarter x pair of 32-column groups
        mbar_wait(accum_full, 0);
        tc_fence_after();
        LN_STAMP(3);
        uint32_t ra[32], rb[32];
        tmem_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (2 * cg2) * 32, ra);
        tmem_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (2 * cg2 + 1) * 32, rb);
        tmem_ld32_wait(ra);
        tmem_ld32_wait(rb);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t *cur = h ? rb : ra;
            uint8_t *box = smem + (size_t)((2 * cg2 + h) * 4 + q) * 4096;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4 *>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(cur[4 * j], cur[4 * j + 1], cur[4 * j + 2], cur[4 * j + 3]);
        }
        LN_STAMP(4);
    }
    __syncthreads();
    LN_STAMP(5);
    {
        // ---- epilogue phase 2: one warp per row (4 rows per warp, two at a time), lane = column 32c + lane for c = 0..7:
        // coalesced 128-byte reads of the residual / positional rows and 128-byte writes of z / out, row statistics by
        // warp shuffles.  A row is a pure latency chain (loads, two 5-step shuffle reductions, rsqrt): 32 resident warps
        // per SM are what hides it -- with 8 warps this phase took 25 000 cycles, 2.5x the GEMM itself.
        uint32_t key = ep.key;
        if (ep.seed_dev) key = lowbias32(*ep.seed_dev ^ key);
        constexpr int RPW = LN_BM / (LN_THREADS / 32);  // rows per warp
        // lane owns columns [4 lane, 4 lane + 4) and [128 + 4 lane, ...): every global / shared access is a 16-byte vector
        // (a warp moves 512 contiguous bytes per instruction; with 4-byte accesses the 50-CTA kernel was bound by the
        // per-SM load / store instruction rate: 51 B/clk)
        const int col0 = 4 * lane;
        const float4 b0 = *reinterpret_cast<const float4 *>(s_bias + col0), b1 = *reinterpret_cast<const float4 *>(s_bias + 128 + col0);
        const float4 g0 = *reinterpret_cast<const float4 *>(s_gamma + col0), g1 = *reinterpret_cast<const float4 *>(s_gamma + 128 + col0);
        const float4 t0 = *reinterpret_cast<const float4 *>(s_beta + col0), t1 = *reinterpret_cast<const float4 *>(s_beta + 128 + col0);
        const float bias[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const float gam[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bet[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
        constexpr int RI = 4;                           // rows in flight per warp: four independent latency chains interleaved
#pragma unroll 1
        for (int i = 0; i < RPW; i += RI) {
            float v[RI][8];
            float sum[RI], sq[RI], mean[RI];
            bool live[RI];
            float4 r0[RI], r1[RI];
#pragma unroll
            for (int u = 0; u < RI; ++u) {              // all global loads of the batch first
                const int row = m0 + warp * RPW + i + u;
                live[u] = row < ep.M;
                r0[u] = r1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live[u]) {
                    const size_t goff = (size_t)row * LN_N + col0;
                    r0[u] = *reinterpret_cast<const float4 *>(ep.resid + goff); r1[u] = *reinterpret_cast<const float4 *>(ep.resid + goff + 128);
                    if (ep.pos) {
                        const float *prow = ep.pos + (size_t)(row % ep.pos_period) * LN_N + col0;
                        const float4 p0 = *reinterpret_cast<const float4 *>(prow), p1 = *reinterpret_cast<const float4 *>(prow + 128);
                        r0[u].x += p0.x; r0[u].y += p0.y; r0[u].z += p0.z; r0[u].w += p0.w;
                        r1[u].x += p1.x; r1[u].y += p1.y; r1[u].z += p1.z; r1[u].w += p1.w;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < RI; ++u) {
                const int rloc = warp * RPW + i + u, row = m0 + rloc;
                const int qq = rloc >> 5, rr = rloc & 31;
                const uint8_t *sp = smem + (size_t)((lane >> 3) * 4 + qq) * 4096 + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4);
                const float4 a0 = *reinterpret_cast<const float4 *>(sp), a1 = *reinterpret_cast<const float4 *>(sp + 16 * 4096);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float rv[8] = {r0[u].x, r0[u].y, r0[u].z, r0[u].w, r1[u].x, r1[u].y, r1[u].z, r1[u].w};
                sum[u] = 0.0f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float a = av[c] + bias[c];
                    const int col = (c < 4 ? col0 : 128 + col0) + (c & 3);
                    if (ep.thresh) a = dropout_keep((uint32_t)((size_t)row * LN_N + col), key, ep.thresh) ? a * ep.keep_scale : 0.0f;
                    v[u][c] = live[u] ? rv[c] + a : 0.0f;
                    sum[u] += v[u][c];
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int u = 0; u < RI; ++u) sum[u] += __shfl_xor_sync(0xffffffffu, sum[u], o);
#pragma unroll
            for (int u = 0; u < RI; ++u) {
                mean[u] = sum[u] * (1.0f / LN_N);
                sq[u] = 0.0f;
#pragma unroll
                for (int c = 0; c < 8; ++c) { const float d = v[u][c] - mean[u]; sq[u] = fmaf(d, d, sq[u]); }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int u = 0; u < RI; ++u) sq[u] += __shfl_xor_sync(0xffffffffu, sq[u], o);
#pragma unroll
            for (int u = 0; u < RI; ++u) {
                if (!live[u]) continue;
                const int row = m0 + warp * RPW + i + u;
                const size_t goff = (size_t)row * LN_N + col0;
                const float rstd = rsqrtf(sq[u] * (1.0f / LN_N) + ep.eps);
                if (ep.write_z) {
                    *reinterpret_cast<float4 *>(ep.z + goff) = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
                    *reinterpret_cast<float4 *>(ep.z + goff + 128) = make_float4(v[u][4], v[u][5], v[u][6], v[u][7]);
                }
                float y[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    y[c] = (v[u][c] - mean[u]) * rstd * gam[c] + bet[c];
                    if (ep.round_out) y[c] = tf32_rn(y[c]);
                }
                *reinterpret_cast<float4 *>(ep.out + goff) = make_float4(y[0], y[1], y[2], y[3]);
                *reinterpret_cast<float4 *>(ep.out + goff + 128) = make_float4(y[4], y[5], y[6], y[7]);
                if (lane == 0) { ep.mean[row] = mean[u]; ep.rstd[row] = rstd; }
            }
        }
    }
    LN_STAMP(6);
    tc_fence_before();
    __syncthreads();
    LN_STAMP(7);
    if (warp == 1) tmem_dealloc(tmem_base, LN_N);
}


// ------------------------------------------------------------------------------------------------------------------
// Column-split variant: a cluster of CN CTAs shares one 128-row tile, CTA r owns columns [r NC, (r + 1) NC), NC = 256 / CN.
// The single-CTA kernel above is bound INSIDE the SM: 128 x 256 outputs at ~35 instructions each (the dropout hash
// alone is 13) are ~36 000 warp instructions on four schedulers, and 384 KB of fp32 operands through one SM's shared
// memory -- 17 500 + 6 100 cycles with only M / 128 = 50 SMs busy.  Splitting the row over a cluster puts 4x the SMs on
// the same tile; the LayerNorm statistics cross the cluster once, through distributed shared memory: every thread
// PUSHES the (mean, M2) of its 32 columns into all CN CTAs (a remote read costs ~1 500 cycles, a remote write is fire
// and forget), one cluster barrier, Chan's merge (numerically a two-pass variance).
//   * warp 0 = TMA producer, warp 1 = MMA issuer; the other warps meanwhile pull the residual (+ positional) tile into
//     swizzled shared-memory boxes with coalesced 16-byte loads (per-thread row-strided 256-bit loads were tried: they
//     delayed the operand TMA loads by ~4 000 cycles), and every warp hashes its dropout keep-mask in the GEMM's shadow;
//   * epilogue: thread = (TMEM lane = row, 32-column group): accumulator by tcgen05.ld, residual from its box, no
//     shuffles; z goes back into its box and leaves by TMA store BEFORE the cluster barrier, out after it.
// ------------------------------------------------------------------------------------------------------------------
struct LnSplitMaps { CUtensorMap a, b, z, out; };

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_dsmem_f2(uint32_t local_addr, uint32_t rank, float x, float y)
{
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int CN>
__global__ void __launch_bounds__(4 * (LN_N / CN / 32) * 32, CN == 4 ? 2 : 1)
gemm_ln_split_kernel(const __grid_constant__ LnSplitMaps maps, const __grid_constant__ LnEpilogue ep)
{
    constexpr int NC = LN_N / CN;                        // columns of this CTA
    constexpr int CG = NC / 32;                          // 32-column groups per CTA
    constexpr int WARPS = 4 * CG, THREADS = WARPS * 32;
    constexpr int NP = CN;                               // partial statistics per row: one per CTA
    constexpr int ST = CN == 4 ? 3 : 4;                  // 107 KB (two CTAs per SM) / 193 KB
    constexpr uint32_t B_STAGE = NC * LN_BK * 4;
    constexpr uint32_t RBUF = LN_BM * NC * 4;            // residual tile, later the z boxes
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *smem_a = smem;                              // after the main loop: the out boxes
    uint8_t *smem_b = smem + ST * LN_A_STAGE;
    uint8_t *rbuf = smem_b + ST * B_STAGE;
    float2 *s_part = reinterpret_cast<float2 *>(rbuf + RBUF);          // [CN][128] (mean, M2) of a CTA's NC columns of a row
    float2 *s_loc = s_part + NP * LN_BM;                               // [CG][128] the same per 32-column group, CTA-local
    float *s_bias = reinterpret_cast<float *>(s_loc + CG * LN_BM);
    float *s_gamma = s_bias + NC, *s_beta = s_gamma + NC;
    uint64_t *full = reinterpret_cast<uint64_t *>(s_beta + NC);
    uint64_t *empty = full + ST;
    uint64_t *accum_full = empty + ST;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int m0 = (blockIdx.x / CN) * LN_BM, n0 = (int)rank * NC;
    const int nkb = ep.num_kb;
    const bool dbg_on = ep.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 64;
#define LNS_STAMP(slot) do { if (dbg_on) ep.dbg[slot] = clock64(); } while (0)
    LNS_STAMP(0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < ST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, NC);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_arrive_relaxed();                            // "this CTA is running" (waited for before the first remote write)
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();
    LNS_STAMP(1);

    if (warp == 0) {
        for (int i = 0; i < nkb; ++i) {
            const int s = i % ST;
            if (i >= ST) mbar_wait(&empty[s], ((i / ST) - 1) & 1);
            const int k0 = i * LN_BK;
            uint8_t *a_dst = smem_a + s * LN_A_STAGE, *b_dst = smem_b + s * B_STAGE;
            if (elect_one()) {
                mbar_expect_tx(&full[s], LN_A_STAGE + B_STAGE);
                tma_load_2d(a_dst, &maps.a, k0, m0, &full[s]);                                    // box {32 k, 128 rows}
#pragma unroll
                for (int j = 0; j < CG; ++j) tma_load_2d(b_dst + j * 4096, &maps.b, n0 + 32 * j, k0, &full[s]);   // box {32 n, 32 k}
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc_tf32(LN_BM, NC, 0, 1);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % ST;
            mbar_wait(&full[s], (i / ST) & 1);
            tc_fence_after();
            const uint32_t a_base = smem_u32(smem_a + s * LN_A_STAGE), b_base = smem_u32(smem_b + s * B_STAGE);
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
    } else {
        // ---- the idle epilogue warps stage the parameters and the residual (+ positional) tile while the GEMM runs.
        // box (g, q) = columns 32 g.., rows 32 q.. of the tile: [32 rows][128 B], 16-byte chunk j of row r at j ^ (r & 7)
        const int t = threadIdx.x - 64;
        constexpr int LOADERS = THREADS - 64;
        for (int c = t; c < NC; c += LOADERS) {
            s_bias[c] = ep.bias ? ep.bias[n0 + c] : 0.0f;
            s_gamma[c] = ep.gamma[n0 + c];
            s_beta[c] = ep.beta[n0 + c];
        }
        constexpr int F4 = NC / 4, ITEMS = LN_BM * F4, BATCH = 4;
        for (int it0 = t; it0 < ITEMS; it0 += LOADERS * BATCH) {
            float4 r[BATCH];
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
                const int it = it0 + u * LOADERS;
                r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (it < ITEMS) {
                    const int rl = it / F4, f = it - rl * F4, rw = m0 + rl;
                    if (rw < ep.M) {
                        r[u] = *reinterpret_cast<const float4 *>(ep.resid + (size_t)rw * LN_N + n0 + 4 * f);
                        if (ep.pos) {
                            const float4 p = *reinterpret_cast<const float4 *>(ep.pos + (size_t)(rw % ep.pos_period) * LN_N + n0 + 4 * f);
                            r[u].x += p.x; r[u].y += p.y; r[u].z += p.z; r[u].w += p.w;
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
                const int it = it0 + u * LOADERS;
                if (it < ITEMS) {
                    const int rl = it / F4, f = it - rl * F4;
                    const int gb = f >> 3, j = f & 7, qb = rl >> 5, rr = rl & 31;
                    *reinterpret_cast<float4 *>(rbuf + (size_t)(gb * 4 + qb) * 4096 + rr * 128 + ((j ^ (rr & 7)) << 4)) = r[u];
                }
            }
        }
    }
    // ---- epilogue: thread = row (TMEM lane quarter q, lane) x 32-column group g
    const int q = warp & 3, g = warp >> 2;
    const int rloc = q * 32 + lane, row = m0 + rloc;
    const uint32_t zbox = smem_u32(rbuf) + (uint32_t)(g * 4 + q) * 4096u;
    const uint32_t obox = smem_u32(smem) + (uint32_t)(g * 4 + q) * 4096u;
    // in the shadow of the GEMM: the dropout keep-mask of the thread's 32 outputs
    uint32_t keep = 0xFFFFFFFFu;
    if (ep.thresh) {
        uint32_t key = ep.key;
        if (ep.seed_dev) key = lowbias32(*ep.seed_dev ^ key);
        const uint32_t idx0 = (uint32_t)row * LN_N + n0 + g * 32;
        keep = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) keep |= (dropout_keep(idx0 + i, key, ep.thresh) ? 1u : 0u) << i;
    }
    __syncthreads();                                    // residual tile and parameters staged
    LNS_STAMP(2);
    float v[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 r4 = lds128(zbox + lane * 128 + ((j ^ (lane & 7)) << 4));
        v[4 * j] = r4.x; v[4 * j + 1] = r4.y; v[4 * j + 2] = r4.z; v[4 * j + 3] = r4.w;
    }
    mbar_wait(accum_full, 0);
    tc_fence_after();
    LNS_STAMP(3);
    {
        uint32_t acc[32];
        tmem_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + g * 32, acc);
        tmem_ld32_wait(acc);
        const uint32_t bias_addr = smem_u32(s_bias + g * 32);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const float4 b4 = lds128(bias_addr + 4 * i);                                   // broadcast read
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float a = __uint_as_float(acc[i + e]) + bb[e];
                if (ep.thresh) a = ((keep >> (i + e)) & 1u) ? a * ep.keep_scale : 0.0f;
                v[i + e] += a;
            }
        }
    }
    // statistics of the thread's 32 values, pushed to every CTA of the cluster
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 4) { s0 += v[i]; s1 += v[i + 1]; s2 += v[i + 2]; s3 += v[i + 3]; }
    const float m_loc = ((s0 + s1) + (s2 + s3)) * (1.0f / 32.0f);
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const float d0 = v[i] - m_loc, d1 = v[i + 1] - m_loc, d2 = v[i + 2] - m_loc, d3 = v[i + 3] - m_loc;
        q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
    }
    float m_cta = m_loc, m2_cta = (q0 + q1) + (q2 + q3);
    if (CG > 1) {
        // merge the CG column groups of the row inside the CTA first (the warps q, q + 4, .. of one TMEM lane quarter meet
        // at named barrier 1 + q): one partial per CTA crosses the cluster
        s_loc[g * LN_BM + rloc] = make_float2(m_cta, m2_cta);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(32 * CG) : "memory");
        float ms = 0.f, m2s = 0.f, pl[CG];
#pragma unroll
        for (int i = 0; i < CG; ++i) { const float2 t = s_loc[i * LN_BM + rloc]; pl[i] = t.x; ms += t.x; m2s += t.y; }
        m_cta = ms * (1.0f / CG);
        float dv = 0.f;
#pragma unroll
        for (int i = 0; i < CG; ++i) { const float d = pl[i] - m_cta; dv = fmaf(d, d, dv); }
        m2_cta = m2s + 32.0f * dv;
    }
    cluster_wait();                                     // (start-of-kernel barrier) the peers are running
    {
        const uint32_t slot = smem_u32(&s_part[(int)rank * LN_BM + rloc]);
#pragma unroll
        for (int r = g; r < CN; r += CG) st_dsmem_f2(slot, (uint32_t)r, m_cta, m2_cta);     // the row's CG threads share the pushes
    }
    cluster_arrive();
    LNS_STAMP(7);
    const bool store_rows = m0 + q * 32 < ep.M;
    if (ep.write_z) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sts128(zbox + lane * 128 + ((j ^ (lane & 7)) << 4), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        fence_proxy_async_smem();
        __syncwarp();
        if (store_rows && elect_one()) {
            tma_store_2d(&maps.z, rbuf + (size_t)(g * 4 + q) * 4096, n0 + g * 32, m0 + q * 32);
            tma_store_commit();
        }
        __syncwarp();
    }
    cluster_wait();
    LNS_STAMP(4);
    float msum = 0.f, m2 = 0.f, pm[NP];
    {
        const uint32_t part_addr = smem_u32(&s_part[rloc]);
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            float x, y;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(part_addr + (uint32_t)i * LN_BM * 8u));
            pm[i] = x; msum += x; m2 += y;
        }
    }
    const float mean = msum * (1.0f / NP);
    float dev = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) { const float d = pm[i] - mean; dev = fmaf(d, d, dev); }
    const float rstd = rsqrtf((m2 + (float)NC * dev) * (1.0f / LN_N) + ep.eps);
    LNS_STAMP(8);
    {
        const uint32_t gamma_addr = smem_u32(s_gamma + g * 32), beta_addr = smem_u32(s_beta + g * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 g4 = lds128(gamma_addr + 16 * j), t4 = lds128(beta_addr + 16 * j);
            float4 y;
            y.x = (v[4 * j] - mean) * rstd * g4.x + t4.x; y.y = (v[4 * j + 1] - mean) * rstd * g4.y + t4.y;
            y.z = (v[4 * j + 2] - mean) * rstd * g4.z + t4.z; y.w = (v[4 * j + 3] - mean) * rstd * g4.w + t4.w;
            if (ep.round_out) { y.x = tf32_rn(y.x); y.y = tf32_rn(y.y); y.z = tf32_rn(y.z); y.w = tf32_rn(y.w); }
            sts128(obox + lane * 128 + ((j ^ (lane & 7)) << 4), y);
        }
    }
    if (rank == 0 && g == 0 && row < ep.M) { ep.mean[row] = mean; ep.rstd[row] = rstd; }
    fence_proxy_async_smem();
    __syncwarp();
    if (store_rows && elect_one()) {
        tma_store_2d(&maps.out, smem + (size_t)(g * 4 + q) * 4096, n0 + g * 32, m0 + q * 32);
        tma_store_commit();
    }
    tma_store_wait_read_all();                          // every lane: whichever lane committed the z / out groups waits for them
    __syncwarp();
    LNS_STAMP(5);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, NC);
    LNS_STAMP(6);
}

template <int CN>
static int launch_gemm_ln_split(const LnSplitMaps &maps, const LnEpilogue &ep, int M, cudaStream_t s)
{
    constexpr int NC = LN_N / CN, THREADS = 4 * (NC / 32) * 32, ST = CN == 4 ? 3 : 4;
    const size_t smem = (size_t)ST * (LN_A_STAGE + NC * LN_BK * 4) + (size_t)LN_BM * NC * 4 + (CN + NC / 32) * LN_BM * 8 + 3 * NC * 4 +
                        (2 * ST + 1) * 8 + 64 + 1024;
    static bool optin = false;
    if (!optin) {
        BDETR_CUDA(cudaFuncSetAttribute(gemm_ln_split_kernel<CN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        optin = true;
    }
    launch_cluster_k(gemm_ln_split_kernel<CN>, dim3(CN * ceil_div(M, LN_BM)), THREADS, CN, smem, s, maps, ep);
    BDETR_CHECK_LAUNCH("gemm_ln_split_kernel");
    return BDETR_OK;
}

// column split of the fused kernel: BDETR_LN_SPLIT = 1 (single CTA per row tile), 2 or 4 (default)
static int ln_split()
{
    static const int v = [] { const char *e = getenv("BDETR_LN_SPLIT"); const int n = e ? atoi(e) : 4; return (n == 1 || n == 2) ? n : 4; }();
    return v;
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
    ep.dbg = g_umma_timeline; ep.z = z; ep.out = out; ep.mean = mean; ep.rstd = rstd; ep.write_z = z != nullptr; ep.round_out = round_out;
    if (ln_split() > 1) {
        LnSplitMaps sm;
        memset(&sm, 0, sizeof(sm));
        sm.a = maps.a; sm.b = maps.b;
        ok = encode_tensor_map_2d(&sm.out, out, M, LN_N, LN_N, 32, 32, false);
        if (z) ok = ok && encode_tensor_map_2d(&sm.z, z, M, LN_N, LN_N, 32, 32, false);
        BDETR_REQUIRE(ok, BDETR_E_CUDA, "cuTensorMapEncodeTiled failed");
        return ln_split() == 2 ? launch_gemm_ln_split<2>(sm, ep, M, s) : launch_gemm_ln_split<4>(sm, ep, M, s);
    }
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
