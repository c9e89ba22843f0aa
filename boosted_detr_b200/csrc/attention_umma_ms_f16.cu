// fp16-operand variant of the long-sequence attention forward (attention_umma_ms.cu): the reference's own GPU
// arithmetic is Keras `mixed_float16` (/root/reference/ModelComponents/parameters.py:73) -- fp16 operands, fp32
// accumulation -- and this kernel is that policy for MultiheadAttention.call (transformers.py:77-100):
//   * Q / K / V tiles are fp16 copies of the projected tensors ([B*L, 256] halves, written by cast_f16x3_kernel):
//     8 KB per 128 x 32 tile instead of 16 KB, 64-byte rows (TMA / UMMA SWIZZLE_64B), six K/V stages in flight;
//   * S = Q K^T is tcgen05.mma kind::f16 (UMMA_K = 16: two instructions per tile instead of four), fp32 S in TMEM;
//   * P = exp2(S - m) is rounded to fp16 (11-bit significand, the precision the tf32 variant keeps by masking) with one
//     cvt.rn.f16x2.f32 per PAIR and written back over S as 64 packed columns: half the TMEM write traffic, half the
//     conversion instructions; O += P V is kind::f16 with P read from TMEM (8 instructions per tile instead of 16);
//   * the row sum l adds the unrounded fp32 exponentials (round-to-nearest errors of P are zero-mean, 2^-12).
// Same schedule otherwise: three query tiles ("streams") per CTA against one K/V ring, one MMA issuer warp per stream,
// single-pass lazy-max softmax, O accumulated in TMEM.  Output o [B,H,Lq,32] / lse stay fp32.
#include <cuda_fp16.h>
#include <math_constants.h>
#include "umma.cuh"

namespace bdetr {

extern long long *g_umma_timeline;      // bdetr_debug_set_timeline: >= 16 slots

constexpr int MH_NS = 3;           // query tiles (streams) per CTA
constexpr int MH_BM = 128;         // query rows per stream
constexpr int MH_KT = 128;         // keys per tile
constexpr int MH_HD = 32;
constexpr int MH_STAGES = 6;
constexpr int MH_THREADS = 64 + 128 * MH_NS + 32 * (MH_NS - 1);      // TMA warp, 3 MMA issuer warps (1, 14, 15), 12 softmax warps
constexpr uint32_t MH_TILE_BYTES = MH_KT * MH_HD * 2;     // 8 KB fp16 (Q, K and V tiles all have this size)
constexpr uint32_t MH_TMEM_COLS = 512;
constexpr uint32_t MH_STREAM_COLS = 160;
constexpr uint32_t MH_O_COL = 128;
constexpr float MH_TH = 8.0f;      // log2 slack before the softmax reference maximum is raised
constexpr int MH_POLY_DEFAULT = 2;  // of every 8 groups of exponentials, how many run on the FMA pipe

// exp2 on the FMA pipe for a share of the elements.  With head dim 32 the softmax needs one exponential per 128
// tensor-core flops; the MUFU does 16 per clock per SM, the tcgen05 pipe ~30 score elements per clock in tf32, so the
// MUFU -- not the tensor pipe -- is the bound (ncu r1b: XU 65 %, tensor 33 %).  Groups of four elements selected by
// POLY_MASK (bit g = group g of every 8) are evaluated as 2^x = 2^round(x) * p(x - round(x)) with a degree-3 minimax p
// (max relative error 8.0e-5 = 2^-13.6, below the 2^-11 of the tf32 weights P is truncated to) in packed FADD2 / FFMA2
// plus one LEA per element for the exponent; the rest still goes through MUFU.EX2.
__device__ __forceinline__ void exp2_poly_x2_h(float x0, float x1, uint32_t &r0, uint32_t &r1)
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
__global__ void __launch_bounds__(MH_THREADS, 1)
attention_fwd_umma_ms_f16_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                             const __grid_constant__ CUtensorMap map_v, int H, int Lq, int Lk,
                             float *__restrict__ o, float *__restrict__ lse, float scale_log2, int round_out,
                             long long *__restrict__ dbg)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *smem_q = smem;
    uint8_t *smem_k = smem_q + MH_NS * MH_TILE_BYTES;
    uint8_t *smem_v = smem_k + MH_STAGES * MH_TILE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_v + MH_STAGES * MH_TILE_BYTES);
    uint64_t *q_full = bars, *kv_full = bars + 1, *kv_empty = kv_full + MH_STAGES;
    uint64_t *s_full = kv_empty + MH_STAGES, *p_full = s_full + MH_NS, *o_full = p_full + MH_NS;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(o_full + MH_NS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, h = blockIdx.y, q_base = blockIdx.x * (MH_NS * MH_BM);
    const int ntiles = (Lk + MH_KT - 1) / MH_KT;
    const int nact = min(MH_NS, (Lq - q_base + MH_BM - 1) / MH_BM);        // streams with at least one valid row
    // optional cycle accounting of CTA (0,0,0) (bdetr_debug_set_timeline): where a softmax warp and the MMA thread wait
    const bool dbg_on = dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
    long long dt[6] = {0, 0, 0, 0, 0, 0}, tk = 0;
#define MH_TICK() do { if (dbg_on) tk = clock64(); } while (0)
#define MH_TOCK(slot) do { if (dbg_on) { const long long n__ = clock64(); dt[slot] += n__ - tk; tk = n__; } } while (0)

    if (threadIdx.x == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < MH_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], nact); }   // every live stream's issuer releases the stage
        for (int j = 0; j < MH_NS; ++j) { mbar_init(&s_full[j], 1); mbar_init(&p_full[j], 128); mbar_init(&o_full[j], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, MH_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();

    if (warp == 0) {
        // TMA producer: warp-convergent loop, one elected lane issues (see elect_one)
        if (elect_one()) {
            mbar_expect_tx(q_full, nact * MH_TILE_BYTES);
            for (int j = 0; j < nact; ++j)
                tma_load_2d(smem_q + j * MH_TILE_BYTES, &map_q, h * MH_HD, b * Lq + q_base + j * MH_BM, q_full);
        }
        __syncwarp();
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % MH_STAGES;
            if (t >= MH_STAGES) mbar_wait(&kv_empty[s], ((t / MH_STAGES) - 1) & 1);
            if (elect_one()) {
                mbar_expect_tx(&kv_full[s], 2 * MH_TILE_BYTES);
                tma_load_2d(smem_k + s * MH_TILE_BYTES, &map_k, h * MH_HD, b * Lk + t * MH_KT, &kv_full[s]);
                tma_load_2d(smem_v + s * MH_TILE_BYTES, &map_v, h * MH_HD, b * Lk + t * MH_KT, &kv_full[s]);
            }
            __syncwarp();
        }
    } else if (warp == 1 || warp >= 2 + 4 * MH_NS) {
        // MMA issuers: ONE WARP PER STREAM (warp 1 -> stream 0, warps 14 / 15 -> streams 1 / 2).  Each walks its own
        // stream's tiles with blocking barrier waits: wait P_j(t) -> O_j += P_j V(t) -> S_j(t+1) = Q_j K(t+1)^T.  A single
        // issuer polling three streams put its polling period and the other streams' ~20 MMA issues (one elected lane,
        // ~25 cycles each) into every stream's critical path: the softmax warps waited ~2 000 cycles per tile for the
        // next S.  The tensor pipe still executes in issue order, so S_j(t+1) cannot overwrite P_j(t) before PV has read it.
        const int j = warp == 1 ? 0 : warp - (2 + 4 * MH_NS) + 1;
        if (j < nact) {
            constexpr uint32_t idesc_s = make_idesc_f16(MH_BM, MH_KT, 0, 0);       // both K-major
            constexpr uint32_t idesc_o = make_idesc_f16(MH_BM, MH_HD, 0, 1);       // A (P, fp16 pairs) from TMEM, B (V) MN-major
            const uint32_t sbase = tmem_base + j * MH_STREAM_COLS;
            auto issue_s = [&](int stage) {                                         // call inside elect_one()
                const uint32_t q_addr = smem_u32(smem_q + j * MH_TILE_BYTES), k_addr = smem_u32(smem_k + stage * MH_TILE_BYTES);
#pragma unroll
                for (int i = 0; i < MH_HD / 16; ++i)               // UMMA_K = 16 halves = 32 B inside the 64-byte swizzled row; 8-row atoms 512 B apart
                    umma_f16(sbase, make_smem_desc(q_addr + i * 32, 16, 512, 4), make_smem_desc(k_addr + i * 32, 16, 512, 4), idesc_s, i != 0);
                umma_commit(&s_full[j]);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            if (elect_one()) issue_s(0);
            __syncwarp();
            for (int t = 0; t < ntiles; ++t) {
                const bool more = t + 1 < ntiles;
                if (more) mbar_wait(&kv_full[(t + 1) % MH_STAGES], ((t + 1) / MH_STAGES) & 1);      // next K tile landed (usually long ago)
                mbar_wait(&p_full[j], t & 1);
                tc_fence_after();
                const int s = t % MH_STAGES;
                const uint32_t v_addr = smem_u32(smem_v + s * MH_TILE_BYTES);
                if (elect_one()) {
                    // O_j (+)= P_j V: accumulated in TMEM across tiles (the softmax rescales it in place on the rare
                    // occasions the reference maximum is raised)
#pragma unroll
                    for (int i = 0; i < MH_KT / 16; ++i)               // 16 keys per instruction: 8 packed P columns, two 8-key atoms (512 B) of V
                        umma_f16_ts(sbase + MH_O_COL, sbase + i * 8, make_smem_desc(v_addr + i * 1024, 512, 512, 4), idesc_o, (t | i) != 0);
                    umma_commit(&o_full[j]);
                    if (more) issue_s((t + 1) % MH_STAGES);
                    umma_commit(&kv_empty[s]);                   // this stream is done with K/V tile t (count = live streams)
                }
                __syncwarp();
            }
        }
    } else {
        const int j = (warp - 2) >> 2;                  // stream (softmax warps 2 .. 13)
        const int q = warp & 3;                         // TMEM lane quadrant this warp may access
        if (j < nact) {
            const int q0 = q_base + j * MH_BM;
            const int row = q0 + q * 32 + lane;
            const uint32_t lane_addr = tmem_base + j * MH_STREAM_COLS + ((uint32_t)(q * 32) << 16);
            // Online softmax with ONE pass over the scores: the reference maximum m is only raised when a 32-column chunk
            // exceeds it by more than 2^MH_TH (then the running sums, and the few chunks of this tile already written as
            // P, are rescaled: rare after the first tile), otherwise P = exp2(s - m) simply uses the stale m: P <= 2^MH_TH,
            // exact in the final o = acc / l.  (A two-pass form -- row maximum of the whole tile first, one vote per tile,
            // then one branch-free 128-element block -- was measured at the same speed, 400-418 vs 403-411 TFLOP/s: the
            // per-chunk vote is not what stalls these warps.  tcgen05.ld itself is cheap, 60 B/clk per warp.)
            float m = -CUDART_INF_F, l = 0.0f;
            for (int t = 0; t < ntiles; ++t) {
                const int valid = min(MH_KT, Lk - t * MH_KT);
                MH_TICK();
                mbar_wait(&s_full[j], t & 1);
                MH_TOCK(0);
                tc_fence_after();
                const bool full_tile = valid == MH_KT;            // only the last tile of a ragged Lk needs masking
                uint32_t ra[32], rb[32];
                float psum = 0.0f;
                tmem_ld32_issue(lane_addr, ra);
#pragma unroll
                for (int c = 0; c < MH_KT / 32; ++c) {
                    uint32_t *cur = (c & 1) ? rb : ra, *nxt = (c & 1) ? ra : rb;
                    tmem_ld32_wait(cur);
                    if (c + 1 < MH_KT / 32) tmem_ld32_issue(lane_addr + (c + 1) * 32, nxt);
                    if (!full_tile) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (c * 32 + i >= valid) cur[i] = 0xFF800000u;      // -inf: exp2 -> 0, never the max
                    }
                    const float cmx = max32_tree(cur) * scale_log2;
                    const bool raise = cmx > m + MH_TH;               // also true for the first finite chunk (m = -inf)
                    if (__any_sync(0xffffffffu, raise)) {             // warp-uniform: the TMEM fix-ups below are warp-collective
                        const float m_new = raise ? cmx : m;
                        const float f = raise ? ex2_approx(m - m_new) : 1.0f;      // exp2(-inf) = 0 on the first chunk
                        psum *= f; l *= f;
                        if (c > 0) tmem_st_wait();                    // this thread's earlier P stores must have landed
                        const uint32_t f2 = pack_f16x2_rn(f, f);
                        for (int e = 0; e < c * 2; ++e) {             // P chunks of this tile already written with the old m (16 halves per 8 columns)
                            uint32_t w8[8];
                            tmem_ld8(lane_addr + e * 8, w8);
#pragma unroll
                            for (int i = 0; i < 8; ++i) w8[i] = mul_f16x2(w8[i], f2);
                            tmem_st8(lane_addr + e * 8, w8);
                        }
                        if (t > 0) {
                            // O (sum over the earlier tiles, in TMEM) carries the old reference: wait until the last PV
                            // has landed, then rescale this thread's row in place.  PV(t) cannot start before p_full.
                            mbar_wait(&o_full[j], (t - 1) & 1);
                            tc_fence_after();
                            for (int e = 0; e < MH_HD / 8; ++e) {
                                uint32_t w8[8];
                                tmem_ld8(lane_addr + MH_O_COL + e * 8, w8);
#pragma unroll
                                for (int i = 0; i < 8; ++i) w8[i] = __float_as_uint(__uint_as_float(w8[i]) * f);
                                tmem_st8(lane_addr + MH_O_COL + e * 8, w8);
                            }
                        }
                        m = m_new;
                    }
                    // P = exp2(s * scale - m): packed FFMA2 for the argument, MUFU.EX2 (or the FMA-pipe polynomial for the
                    // POLY_MASK groups), the fp32 values summed with packed adds, then one cvt.rn.f16x2.f32 per pair
                    const uint64_t sc2 = pack_f32x2(scale_log2, scale_log2), nm2 = pack_f32x2(-m, -m);
                    uint64_t ps_a = pack_f32x2(0.0f, 0.0f), ps_b = ps_a;
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        float x0, x1, x2, x3;
                        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])), sc2, nm2), x0, x1);
                        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(cur[i + 2]), __uint_as_float(cur[i + 3])), sc2, nm2), x2, x3);
                        float p0, p1, p2, p3;
                        if ((POLY_MASK >> ((i >> 2) & 7)) & 1u) {         // compile-time choice per group: FMA-pipe exponential
                            uint32_t e0, e1, e2, e3;
                            exp2_poly_x2_h(x0, x1, e0, e1);
                            exp2_poly_x2_h(x2, x3, e2, e3);
                            p0 = __uint_as_float(e0); p1 = __uint_as_float(e1); p2 = __uint_as_float(e2); p3 = __uint_as_float(e3);
                        } else {
                            p0 = ex2_approx(x0); p1 = ex2_approx(x1); p2 = ex2_approx(x2); p3 = ex2_approx(x3);
                        }
                        ps_a = add_f32x2(ps_a, pack_f32x2(p0, p1));
                        ps_b = add_f32x2(ps_b, pack_f32x2(p2, p3));
                        pk[i >> 1] = pack_f16x2_rn(p0, p1);
                        pk[(i >> 1) + 1] = pack_f16x2_rn(p2, p3);
                    }
                    {
                        float s0, s1;
                        unpack_f32x2(add_f32x2(ps_a, ps_b), s0, s1);
                        psum += s0 + s1;
                    }
                    tmem_st16_u(lane_addr + c * 16, pk);            // P chunk c: 16 packed columns, over S columns that were read long ago
                }
                MH_TOCK(1);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_full[j]);
                l += psum;
                MH_TOCK(2);
            }
            // all tiles accumulated: O = TMEM accumulator / l
            MH_TICK();
            mbar_wait(&o_full[j], (ntiles - 1) & 1);
            MH_TOCK(3);
            tc_fence_after();
            float acc[MH_HD];
            tmem_ld32(lane_addr + MH_O_COL, acc);
            if (dbg_on && warp == 2) { dbg[0] = dt[0]; dbg[1] = dt[1]; dbg[2] = dt[2]; dbg[3] = dt[3]; dbg[4] = dt[4]; }
            if (row < Lq) {
                const float inv = 1.0f / l;
                float *dst = o + (((size_t)b * H + h) * Lq + row) * MH_HD;
#pragma unroll
                for (int i = 0; i < MH_HD; i += 4) {
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
    if (warp == 1) tmem_dealloc(tmem_base, MH_TMEM_COLS);
}

// fp32 -> fp16 copies of the projected q / k / v (one launch; blockIdx.y = tensor).  At config 5 this moves 0.37 GB per
// block = ~60 us beside a ~3 ms attention; writing fp16 straight from the projection GEMM's epilogue would remove it.
__global__ void __launch_bounds__(256)
cast_f16x3_kernel(const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v,
                  __half *__restrict__ q16, __half *__restrict__ k16, __half *__restrict__ v16, size_t nq8, size_t nk8)
{
    pdl_sync();
    const int which = blockIdx.y;
    const float *src = which == 0 ? q : which == 1 ? k : v;
    __half *dst = which == 0 ? q16 : which == 1 ? k16 : v16;
    const size_t n8 = which == 0 ? nq8 : nk8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = reinterpret_cast<const float4 *>(src)[2 * i], b = reinterpret_cast<const float4 *>(src)[2 * i + 1];
        uint4 w;
        w.x = pack_f16x2_rn(a.x, a.y); w.y = pack_f16x2_rn(a.z, a.w); w.z = pack_f16x2_rn(b.x, b.y); w.w = pack_f16x2_rn(b.z, b.w);
        reinterpret_cast<uint4 *>(dst)[i] = w;
    }
}

// halves of workspace the fp16 path needs for these shapes (q16 | k16 | v16), 0 when the shapes do not take it
size_t attention_f16_workspace_bytes(int B, int H, int Lq, int Lk, int d)
{
    const bool forced = g_force_attention_kernel == 2 || (g_force_attention_kernel >= 20 && g_force_attention_kernel <= 28);      // tests: any shape
    if (d != MH_HD || !(forced || attention_umma_ms_eligible(B, H, Lq, Lk))) return 0;
    return ((size_t)B * Lq + 2 * (size_t)B * Lk) * H * d * sizeof(__half);
}

int launch_attention_fwd_umma_ms_f16(int B, int H, int Lq, int Lk, int d, const float *qp, const float *kp, const float *vp,
                                     void *ws16, float *o, float *lse, int round_out, cudaStream_t s)
{
    BDETR_REQUIRE(d == MH_HD, BDETR_E_UNSUPPORTED, "head dim must be 32 (D/H)");
    BDETR_REQUIRE(ws16 && (reinterpret_cast<uintptr_t>(ws16) & 127) == 0, BDETR_E_NULL, "fp16 attention needs a 128-byte aligned workspace");
    const int D = H * d;
    const size_t nq = (size_t)B * Lq * D, nk = (size_t)B * Lk * D;
    __half *q16 = static_cast<__half *>(ws16), *k16 = q16 + nq, *v16 = k16 + nk;
    {
        const size_t n8 = (nq > nk ? nq : nk) / 8;
        const int blocks = (int)((n8 + 255) / 256 < 148 * 8 ? (n8 + 255) / 256 : 148 * 8);
        launch_k(cast_f16x3_kernel, dim3(blocks, 3), 256, 0, s, qp, kp, vp, q16, k16, v16, nq / 8, nk / 8);
        BDETR_CHECK_LAUNCH("cast_f16x3_kernel");
    }
    CUtensorMap mq, mk, mv;
    bool ok = encode_tensor_map_2d_f16(&mq, q16, (long long)B * Lq, D, D, MH_HD, MH_BM);
    ok = ok && encode_tensor_map_2d_f16(&mk, k16, (long long)B * Lk, D, D, MH_HD, MH_KT);
    ok = ok && encode_tensor_map_2d_f16(&mv, v16, (long long)B * Lk, D, D, MH_HD, MH_KT);
    BDETR_REQUIRE(ok, BDETR_E_CUDA, "cuTensorMapEncodeTiled failed");
    const size_t smem = (size_t)(MH_NS + 2 * MH_STAGES) * MH_TILE_BYTES + 32 * 8 + 16 + 1024;
    // share of exponentials on the FMA pipe (bdetr_debug_force_attention_kernel 20 + n: n = 0, 2 or 4 of 8 here)
    const int want = (g_force_attention_kernel >= 20 && g_force_attention_kernel <= 28) ? g_force_attention_kernel - 20 : MH_POLY_DEFAULT;
    const int share = want <= 0 ? 0 : want <= 2 ? 1 : 2;
    auto kern = share == 0 ? attention_fwd_umma_ms_f16_kernel<0x00u> : share == 1 ? attention_fwd_umma_ms_f16_kernel<0x22u>
              : attention_fwd_umma_ms_f16_kernel<0xAAu>;
    static bool optin[3] = {false};
    if (!optin[share]) {
        BDETR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        optin[share] = true;
    }
    const float scale_log2 = (1.0f / sqrtf((float)d)) * 1.4426950408889634f;
    dim3 grid(ceil_div(Lq, MH_NS * MH_BM), H, B);
    launch_k(kern, grid, MH_THREADS, smem, s, mq, mk, mv, H, Lq, Lk, o, lse, scale_log2, round_out, g_umma_timeline);
    BDETR_CHECK_LAUNCH("attention_fwd_umma_ms_f16_kernel");
    return BDETR_OK;
}

}  // namespace bdetr
