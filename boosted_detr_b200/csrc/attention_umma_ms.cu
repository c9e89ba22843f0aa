// Flash-style attention forward for LONG sequences (BASELINE config 5: ~20k encoder tokens): the same math and
// output layout as attention_umma.cu, re-scheduled for latency hiding.
// Reference: MultiheadAttention.call, /root/reference/ModelComponents/transformers.py:77-100.
//
// With head dim 32 a 128x128 score tile is only 2 x 1 MFLOP of MMA against 16 384 exponentials, and one tile's
// dependency chain (S = QKt -> row max -> exp2 -> P -> O += PV -> next S) is several thousand cycles long, so the
// one-tile-per-CTA kernel leaves the MUFU and tensor pipes idle most of the time (ncu: XU 40 %, tensor 20 % at
// two CTAs per SM).  Here ONE CTA per SM runs NS = 3 independent query tiles ("streams") against the SAME K/V
// ring:
//   warp 0        TMA producer: 3 Q tiles once, then a 3-stage ring of (K tile, V tile), 128 keys each
//   warp 1        TMEM allocator (all 512 columns) + single-thread tcgen05.mma issuer; per K/V tile it serves the
//                 streams round-robin:  wait P_j -> O_j = P_j V  -> S_j(next tile) = Q_j K'^T
//                 so a stream's next scores are already being computed while its softmax warps read O
//   warps 2..13   three softmax warpgroups (one per stream, one thread per query row / TMEM lane)
// TMEM columns: stream j owns [160 j, 160 j + 128) for S / P (P overwrites S in place, tf32-masked) and
// [160 j + 128, 160 j + 160) for the per-tile O.  K/V tiles are read from L2 once per 384 query rows instead of
// once per 128.
#include <math_constants.h>
#include "umma.cuh"

namespace bdetr {

extern long long *g_umma_timeline;      // bdetr_debug_set_timeline: >= 16 slots

constexpr int MS_NS = 3;           // query tiles (streams) per CTA
constexpr int MS_BM = 128;         // query rows per stream
constexpr int MS_KT = 128;         // keys per tile
constexpr int MS_HD = 32;
constexpr int MS_STAGES = 3;
constexpr int MS_THREADS = 64 + 128 * MS_NS + 32 * (MS_NS - 1);      // TMA warp, 3 MMA issuer warps (1, 14, 15), 12 softmax warps
constexpr uint32_t MS_TILE_BYTES = MS_KT * MS_HD * 4;     // 16 KB (Q, K and V tiles all have this size)
constexpr uint32_t MS_TMEM_COLS = 512;
constexpr uint32_t MS_STREAM_COLS = 160;
constexpr uint32_t MS_O_COL = 128;
constexpr float MS_TH = 8.0f;      // log2 slack before the softmax reference maximum is raised
constexpr int MS_POLY_DEFAULT = 2;  // of every 8 groups of exponentials, how many run on the FMA pipe

// exp2 on the FMA pipe for a share of the elements.  With head dim 32 the softmax needs one exponential per 128
// tensor-core flops; the MUFU does 16 per clock per SM, the tcgen05 pipe ~30 score elements per clock in tf32, so the
// MUFU -- not the tensor pipe -- is the bound (ncu r1b: XU 65 %, tensor 33 %).  Groups of four elements selected by
// POLY_MASK (bit g = group g of every 8) are evaluated as 2^x = 2^round(x) * p(x - round(x)) with a degree-3 minimax p
// (max relative error 8.0e-5 = 2^-13.6, below the 2^-11 of the tf32 weights P is truncated to) in packed FADD2 / FFMA2
// plus one LEA per element for the exponent; the rest still goes through MUFU.EX2.
__device__ __forceinline__ void exp2_poly_x2(float x0, float x1, uint32_t &r0, uint32_t &r1)
{
    x0 = fmaxf(x0, -125.0f); x1 = fmaxf(x1, -125.0f);                  // masked (-inf) / far-away scores: ~2^-125, never a wrapped exponent
    const uint64_t magic = pack_f32x2(12582912.0f, 12582912.0f);       // 1.5 * 2^23: x + magic has round(x) in its low mantissa bits
    const uint64_t X = pack_f32x2(x0, x1);
    const uint64_t T = add_f32x2(X, magic);
    const uint64_t F = sub_f32x2(X, sub_f32x2(T, magic));              // x - round(x) in [-0.5, 0.5]
    uint64_t P = fma_f32x2(F, pack_f32x2(0.05519810691475868f, 0.05519810691475868f), pack_f32x2(0.24267712235450745f, 0.24267712235450745f));
    P = fma_f32x2(P, F, pack_f32x2(0.6932618021965027f, 0.6932618021965027f));
    P = fma_f32x2(P, F, pack_f32x2(0.9999227523803711f, 0.9999227523803711f));
    float p0, p1, t0, t1;
    unpack_f32x2(P, p0, p1); unpack_f32x2(T, t0, t1);
    r0 = __float_as_uint(p0) + (__float_as_uint(t0) << 23);              // low bits of t = round(x) (two's complement): add to the exponent
    r1 = __float_as_uint(p1) + (__float_as_uint(t1) << 23);
}

template <uint32_t POLY_MASK>
__global__ void __launch_bounds__(MS_THREADS, 1)
attention_fwd_umma_ms_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                             const __grid_constant__ CUtensorMap map_v, int H, int Lq, int Lk,
                             float *__restrict__ o, float *__restrict__ lse, float scale_log2, int round_out,
                             long long *__restrict__ dbg)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *smem_q = smem;
    uint8_t *smem_k = smem_q + MS_NS * MS_TILE_BYTES;
    uint8_t *smem_v = smem_k + MS_STAGES * MS_TILE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_v + MS_STAGES * MS_TILE_BYTES);
    uint64_t *q_full = bars, *kv_full = bars + 1, *kv_empty = kv_full + MS_STAGES;
    uint64_t *s_full = kv_empty + MS_STAGES, *p_full = s_full + MS_NS, *o_full = p_full + MS_NS;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(o_full + MS_NS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, h = blockIdx.y, q_base = blockIdx.x * (MS_NS * MS_BM);
    const int ntiles = (Lk + MS_KT - 1) / MS_KT;
    const int nact = min(MS_NS, (Lq - q_base + MS_BM - 1) / MS_BM);        // streams with at least one valid row
    // optional cycle accounting of CTA (0,0,0) (bdetr_debug_set_timeline): where a softmax warp and the MMA thread wait
    const bool dbg_on = dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
    long long dt[6] = {0, 0, 0, 0, 0, 0}, tk = 0;
#define MS_TICK() do { if (dbg_on) tk = clock64(); } while (0)
#define MS_TOCK(slot) do { if (dbg_on) { const long long n__ = clock64(); dt[slot] += n__ - tk; tk = n__; } } while (0)

    if (threadIdx.x == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < MS_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], nact); }   // every live stream's issuer releases the stage
        for (int j = 0; j < MS_NS; ++j) { mbar_init(&s_full[j], 1); mbar_init(&p_full[j], 128); mbar_init(&o_full[j], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, MS_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();

    if (warp == 0) {
        // TMA producer: warp-convergent loop, one elected lane issues (see elect_one)
        if (elect_one()) {
            mbar_expect_tx(q_full, nact * MS_TILE_BYTES);
            for (int j = 0; j < nact; ++j)
                tma_load_2d(smem_q + j * MS_TILE_BYTES, &map_q, h * MS_HD, b * Lq + q_base + j * MS_BM, q_full);
        }
        __syncwarp();
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % MS_STAGES;
            if (t >= MS_STAGES) mbar_wait(&kv_empty[s], ((t / MS_STAGES) - 1) & 1);
            if (elect_one()) {
                mbar_expect_tx(&kv_full[s], 2 * MS_TILE_BYTES);
                tma_load_2d(smem_k + s * MS_TILE_BYTES, &map_k, h * MS_HD, b * Lk + t * MS_KT, &kv_full[s]);
                tma_load_2d(smem_v + s * MS_TILE_BYTES, &map_v, h * MS_HD, b * Lk + t * MS_KT, &kv_full[s]);
            }
            __syncwarp();
        }
    } else if (warp == 1 || warp >= 2 + 4 * MS_NS) {
        // MMA issuers: ONE WARP PER STREAM (warp 1 -> stream 0, warps 14 / 15 -> streams 1 / 2).  Each walks its own
        // stream's tiles with blocking barrier waits: wait P_j(t) -> O_j += P_j V(t) -> S_j(t+1) = Q_j K(t+1)^T.  A single
        // issuer polling three streams put its polling period and the other streams' ~20 MMA issues (one elected lane,
        // ~25 cycles each) into every stream's critical path: the softmax warps waited ~2 000 cycles per tile for the
        // next S.  The tensor pipe still executes in issue order, so S_j(t+1) cannot overwrite P_j(t) before PV has read it.
        const int j = warp == 1 ? 0 : warp - (2 + 4 * MS_NS) + 1;
        if (j < nact) {
            constexpr uint32_t idesc_s = make_idesc_tf32(MS_BM, MS_KT, 0, 0);      // both K-major
            constexpr uint32_t idesc_o = make_idesc_tf32(MS_BM, MS_HD, 0, 1);      // A from TMEM, B (V) MN-major
            const uint32_t sbase = tmem_base + j * MS_STREAM_COLS;
            auto issue_s = [&](int stage) {                                         // call inside elect_one()
                const uint32_t q_addr = smem_u32(smem_q + j * MS_TILE_BYTES), k_addr = smem_u32(smem_k + stage * MS_TILE_BYTES);
#pragma unroll
                for (int i = 0; i < MS_HD / 8; ++i)
                    umma_tf32(sbase, make_smem_desc(q_addr + i * 32, 16, 1024, 2), make_smem_desc(k_addr + i * 32, 16, 1024, 2), idesc_s, i != 0);
                umma_commit(&s_full[j]);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            if (elect_one()) issue_s(0);
            __syncwarp();
            for (int t = 0; t < ntiles; ++t) {
                const bool more = t + 1 < ntiles;
                if (more) mbar_wait(&kv_full[(t + 1) % MS_STAGES], ((t + 1) / MS_STAGES) & 1);      // next K tile landed (usually long ago)
                mbar_wait(&p_full[j], t & 1);
                tc_fence_after();
                const int s = t % MS_STAGES;
                const uint32_t v_addr = smem_u32(smem_v + s * MS_TILE_BYTES);
                if (elect_one()) {
                    // O_j (+)= P_j V: accumulated in TMEM across tiles (the softmax rescales it in place on the rare
                    // occasions the reference maximum is raised)
#pragma unroll
                    for (int i = 0; i < MS_KT / 8; ++i)
                        umma_tf32_ts(sbase + MS_O_COL, sbase + i * 8, make_smem_desc(v_addr + i * 1024, 4096, 512, 1), idesc_o, (t | i) != 0);
                    umma_commit(&o_full[j]);
                    if (more) issue_s((t + 1) % MS_STAGES);
                    umma_commit(&kv_empty[s]);                   // this stream is done with K/V tile t (count = live streams)
                }
                __syncwarp();
            }
        }
    } else {
        const int j = (warp - 2) >> 2;                  // stream (softmax warps 2 .. 13)
        const int q = warp & 3;                         // TMEM lane quadrant this warp may access
        if (j < nact) {
            const int q0 = q_base + j * MS_BM;
            const int row = q0 + q * 32 + lane;
            const uint32_t lane_addr = tmem_base + j * MS_STREAM_COLS + ((uint32_t)(q * 32) << 16);
            // Online softmax with ONE pass over the scores.  TMEM reads run at 64 B/clk per SM -- the same 16 values per
            // clock as the MUFU -- so a separate row-max pass over S would double the binding traffic.  Instead the
            // reference maximum m is only raised when a 32-column chunk exceeds it by more than 2^MS_TH (then the
            // running sums, and the few chunks of this tile already written as P, are rescaled: rare after the first
            // tile), otherwise P = exp2(s - m) simply uses the stale m: P <= 2^MS_TH, exact in the final o = acc / l.
            float m = -CUDART_INF_F, l = 0.0f;
            for (int t = 0; t < ntiles; ++t) {
                const int valid = min(MS_KT, Lk - t * MS_KT);
                MS_TICK();
                mbar_wait(&s_full[j], t & 1);
                MS_TOCK(0);
                tc_fence_after();
                const bool full_tile = valid == MS_KT;            // only the last tile of a ragged Lk needs masking
                uint32_t ra[32], rb[32];
                float psum = 0.0f;
                tmem_ld32_issue(lane_addr, ra);
#pragma unroll
                for (int c = 0; c < MS_KT / 32; ++c) {
                    uint32_t *cur = (c & 1) ? rb : ra, *nxt = (c & 1) ? ra : rb;
                    tmem_ld32_wait(cur);
                    if (c + 1 < MS_KT / 32) tmem_ld32_issue(lane_addr + (c + 1) * 32, nxt);
                    if (!full_tile) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (c * 32 + i >= valid) cur[i] = 0xFF800000u;      // -inf: exp2 -> 0, never the max
                    }
                    const float cmx = max32_tree(cur) * scale_log2;
                    const bool raise = cmx > m + MS_TH;               // also true for the first finite chunk (m = -inf)
                    if (__any_sync(0xffffffffu, raise)) {             // warp-uniform: the TMEM fix-ups below are warp-collective
                        const float m_new = raise ? cmx : m;
                        const float f = raise ? ex2_approx(m - m_new) : 1.0f;      // exp2(-inf) = 0 on the first chunk
                        psum *= f; l *= f;
                        if (c > 0) tmem_st_wait();                    // this thread's earlier P stores must have landed
                        for (int e = 0; e < c * 4; ++e) {             // P chunks of this tile already written with the old m
                            uint32_t w8[8];
                            tmem_ld8(lane_addr + e * 8, w8);
#pragma unroll
                            for (int i = 0; i < 8; ++i) w8[i] = __float_as_uint(__uint_as_float(w8[i]) * f);   // power of two: exact
                            tmem_st8(lane_addr + e * 8, w8);
                        }
                        if (t > 0) {
                            // O (sum over the earlier tiles, in TMEM) carries the old reference: wait until the last PV
                            // has landed, then rescale this thread's row in place.  PV(t) cannot start before p_full.
                            mbar_wait(&o_full[j], (t - 1) & 1);
                            tc_fence_after();
                            for (int e = 0; e < MS_HD / 8; ++e) {
                                uint32_t w8[8];
                                tmem_ld8(lane_addr + MS_O_COL + e * 8, w8);
#pragma unroll
                                for (int i = 0; i < 8; ++i) w8[i] = __float_as_uint(__uint_as_float(w8[i]) * f);
                                tmem_st8(lane_addr + MS_O_COL + e * 8, w8);
                            }
                        }
                        m = m_new;
                    }
                    // P = exp2(s * scale - m): packed FFMA2 for the argument, MUFU.EX2, then keep the 10 mantissa bits the
                    // tf32 MMA reads (same masking rule as attention_umma.cu) and sum exactly those weights with packed
                    // adds into four independent partial sums, so the truncation cancels in the normalisation.
                    const uint64_t sc2 = pack_f32x2(scale_log2, scale_log2), nm2 = pack_f32x2(-m, -m);
                    uint64_t ps_a = pack_f32x2(0.0f, 0.0f), ps_b = ps_a;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        float x0, x1, x2, x3;
                        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])), sc2, nm2), x0, x1);
                        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(cur[i + 2]), __uint_as_float(cur[i + 3])), sc2, nm2), x2, x3);
                        if ((POLY_MASK >> ((i >> 2) & 7)) & 1u) {         // compile-time choice per group: FMA-pipe exponential
                            uint32_t e0, e1, e2, e3;
                            exp2_poly_x2(x0, x1, e0, e1);
                            exp2_poly_x2(x2, x3, e2, e3);
                            cur[i] = e0 & 0xFFFFE000u; cur[i + 1] = e1 & 0xFFFFE000u;
                            cur[i + 2] = e2 & 0xFFFFE000u; cur[i + 3] = e3 & 0xFFFFE000u;
                        } else {
                            cur[i] = __float_as_uint(ex2_approx(x0)) & 0xFFFFE000u;
                            cur[i + 1] = __float_as_uint(ex2_approx(x1)) & 0xFFFFE000u;
                            cur[i + 2] = __float_as_uint(ex2_approx(x2)) & 0xFFFFE000u;
                            cur[i + 3] = __float_as_uint(ex2_approx(x3)) & 0xFFFFE000u;
                        }
                        ps_a = add_f32x2(ps_a, pack_f32x2(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])));
                        ps_b = add_f32x2(ps_b, pack_f32x2(__uint_as_float(cur[i + 2]), __uint_as_float(cur[i + 3])));
                    }
                    {
                        float s0, s1;
                        unpack_f32x2(add_f32x2(ps_a, ps_b), s0, s1);
                        psum += s0 + s1;
                    }
                    tmem_st32_u(lane_addr + c * 32, cur);
                }
                MS_TOCK(1);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_full[j]);
                l += psum;
                MS_TOCK(2);
            }
            // all tiles accumulated: O = TMEM accumulator / l
            MS_TICK();
            mbar_wait(&o_full[j], (ntiles - 1) & 1);
            MS_TOCK(3);
            tc_fence_after();
            float acc[MS_HD];
            tmem_ld32(lane_addr + MS_O_COL, acc);
            if (dbg_on && warp == 2) { dbg[0] = dt[0]; dbg[1] = dt[1]; dbg[2] = dt[2]; dbg[3] = dt[3]; dbg[4] = dt[4]; }
            if (row < Lq) {
                const float inv = 1.0f / l;
                float *dst = o + (((size_t)b * H + h) * Lq + row) * MS_HD;
#pragma unroll
                for (int i = 0; i < MS_HD; i += 4) {
                    float4 w = make_float4(acc[i] * inv, acc[i + 1] * inv, acc[i + 2] * inv, acc[i + 3] * inv);
                    if (round_out) { w.x = tf32_rn(w.x); w.y = tf32_rn(w.y); w.z = tf32_rn(w.z); w.w = tf32_rn(w.w); }
                    *reinterpret_cast<float4 *>(dst + i) = w;
                }
                lse[((size_t)b * H + h) * Lq + row] = m + log2f(l);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, MS_TMEM_COLS);
}

// Long query sequences only: a CTA covers 384 query rows and owns a whole SM, so short sequences (L = 400) keep the
// one-tile-per-CTA kernel (more, smaller CTAs).
bool attention_umma_ms_eligible(int B, int H, int Lq, int Lk)
{
    const long long ctas = (long long)((Lq + MS_NS * MS_BM - 1) / (MS_NS * MS_BM)) * H * B;
    return Lq >= 4 * MS_NS * MS_BM && Lk >= 8 * MS_KT && ctas >= 148;
}

int launch_attention_fwd_umma_ms(int B, int H, int Lq, int Lk, int d, const float *qp, const float *kp, const float *vp,
                                 float *o, float *lse, int round_out, cudaStream_t s)
{
    BDETR_REQUIRE(d == MS_HD, BDETR_E_UNSUPPORTED, "head dim must be 32 (D/H)");
    const int D = H * d;
    CUtensorMap mq, mk, mv;
    bool ok = encode_tensor_map_2d(&mq, qp, (long long)B * Lq, D, D, MS_HD, MS_BM, false);
    ok = ok && encode_tensor_map_2d(&mk, kp, (long long)B * Lk, D, D, MS_HD, MS_KT, false);
    ok = ok && encode_tensor_map_2d(&mv, vp, (long long)B * Lk, D, D, MS_HD, MS_KT, true);
    BDETR_REQUIRE(ok, BDETR_E_CUDA, "cuTensorMapEncodeTiled failed");
    const size_t smem = (size_t)(MS_NS + 2 * MS_STAGES) * MS_TILE_BYTES + 32 * 8 + 16 + 1024;
    // share of exponentials evaluated on the FMA pipe: groups of 4 elements per 32 (bdetr_debug_force_attention_kernel
    // 20 + n selects n of 8 for experiments; the default was chosen from gpurun_out/bench_attention.json)
    const int share = (g_force_attention_kernel >= 20 && g_force_attention_kernel <= 28) ? g_force_attention_kernel - 20 : MS_POLY_DEFAULT;
    auto kern = share == 0 ? attention_fwd_umma_ms_kernel<0x00u> : share == 1 ? attention_fwd_umma_ms_kernel<0x10u>
              : share == 2 ? attention_fwd_umma_ms_kernel<0x22u> : share == 3 ? attention_fwd_umma_ms_kernel<0x4Au>
              : share == 4 ? attention_fwd_umma_ms_kernel<0xAAu> : share == 5 ? attention_fwd_umma_ms_kernel<0xB5u>
              : share == 6 ? attention_fwd_umma_ms_kernel<0xDDu> : share == 7 ? attention_fwd_umma_ms_kernel<0xEFu>
              : attention_fwd_umma_ms_kernel<0xFFu>;
    static bool optin[9] = {false};
    if (!optin[share]) {
        BDETR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        optin[share] = true;
    }
    const float scale_log2 = (1.0f / sqrtf((float)d)) * 1.4426950408889634f;
    dim3 grid(ceil_div(Lq, MS_NS * MS_BM), H, B);
    launch_k(kern, grid, MS_THREADS, smem, s, mq, mk, mv, H, Lq, Lk, o, lse, scale_log2, round_out, g_umma_timeline);
    BDETR_CHECK_LAUNCH("attention_fwd_umma_ms_kernel");
    return BDETR_OK;
}

}  // namespace bdetr
