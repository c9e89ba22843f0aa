// K7 cost matrix, K8 per-image assignment, K9 matched loss (+ backward) for sm_100a.
//
// Reference: /root/reference/ModelComponents/losses_and_metrics.py
//   CostArray :215-225, CategoryLoss :44-49, AttributeLoss :51-57, BoxLoss :68-72, coco_to_tf :59-66,
//   MatchingAssignment :228-251, MatchingMask :195-212, MatchingLoss.call :111-161, ExistLoss :33-37.
// These are HBM-bound fp32 / latency-bound fp64 kernels: plain CUDA cores, coalesced stores,
// shared-memory staging, no tensor cores (DESIGN.md "matcher kernels").
#include <math_constants.h>
#include "common.cuh"

namespace bdetr {

// ---------------------------------------------------------------------------------------------
// Per-pair box math shared by the cost, loss and gradient kernels.  Unfused (_rn) arithmetic on
// purpose: the reference evaluates op by op, and near-tied assignments flip on 1-ulp differences.
// ---------------------------------------------------------------------------------------------
struct BoxTF { float ymin, xmin, ymax, xmax; };

__device__ __forceinline__ BoxTF coco_to_tf(float x, float y, float w, float h)
{
    BoxTF b; b.ymin = y; b.xmin = x; b.ymax = __fadd_rn(y, h); b.xmax = __fadd_rn(x, w);
    return b;
}
__device__ __forceinline__ float div_no_nan(float a, float b) { return b == 0.0f ? 0.0f : __fdiv_rn(a, b); }

// tf.maximum / tf.minimum propagate NaN; fmaxf/fminf do not.  Padded rows never hold NaN and a NaN
// prediction must poison the cost (scipy then raises), so propagate explicitly.
__device__ __forceinline__ float tf_max(float a, float b) { return (a != a || b != b) ? CUDART_NAN_F : fmaxf(a, b); }
__device__ __forceinline__ float tf_min(float a, float b) { return (a != a || b != b) ? CUDART_NAN_F : fminf(a, b); }

// t = target (tfa's b1), p = prediction (b2).  Returns 2*(1-giou) + 5*mean((10t-10p)^2); *iou_out = IoU.
__device__ __forceinline__ float box_pair_cost(const BoxTF &t, const BoxTF &p, float *iou_out)
{
    const float tw = tf_max(0.0f, __fsub_rn(t.xmax, t.xmin)), th = tf_max(0.0f, __fsub_rn(t.ymax, t.ymin));
    const float pw = tf_max(0.0f, __fsub_rn(p.xmax, p.xmin)), ph = tf_max(0.0f, __fsub_rn(p.ymax, p.ymin));
    const float at = __fmul_rn(tw, th), ap = __fmul_rn(pw, ph);
    const float iy0 = tf_max(t.ymin, p.ymin), ix0 = tf_max(t.xmin, p.xmin);
    const float iy1 = tf_min(t.ymax, p.ymax), ix1 = tf_min(t.xmax, p.xmax);
    const float iw = tf_max(0.0f, __fsub_rn(ix1, ix0)), ih = tf_max(0.0f, __fsub_rn(iy1, iy0));
    const float ai = __fmul_rn(iw, ih);
    const float un = __fsub_rn(__fadd_rn(at, ap), ai);
    const float iou = div_no_nan(ai, un);
    const float ey0 = tf_min(t.ymin, p.ymin), ex0 = tf_min(t.xmin, p.xmin);
    const float ey1 = tf_max(t.ymax, p.ymax), ex1 = tf_max(t.xmax, p.xmax);
    const float ew = tf_max(0.0f, __fsub_rn(ex1, ex0)), eh = tf_max(0.0f, __fsub_rn(ey1, ey0));
    const float ae = __fmul_rn(ew, eh);
    const float giou = __fsub_rn(iou, div_no_nan(__fsub_rn(ae, un), ae));
    const float d0 = __fsub_rn(__fmul_rn(10.0f, t.ymin), __fmul_rn(10.0f, p.ymin));
    const float d1 = __fsub_rn(__fmul_rn(10.0f, t.xmin), __fmul_rn(10.0f, p.xmin));
    const float d2 = __fsub_rn(__fmul_rn(10.0f, t.ymax), __fmul_rn(10.0f, p.ymax));
    const float d3 = __fsub_rn(__fmul_rn(10.0f, t.xmax), __fmul_rn(10.0f, p.xmax));
    const float ss = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)), __fmul_rn(d3, d3));
    const float l2 = __fdiv_rn(ss, 4.0f);
    if (iou_out) *iou_out = iou;
    return __fadd_rn(__fmul_rn(2.0f, __fsub_rn(1.0f, giou)), __fmul_rn(5.0f, l2));
}

// Gradient of box_pair_cost w.r.t. the prediction in COCO (x,y,w,h) form, times `g`.
// tf.maximum/minimum gradient convention: ties go to the FIRST argument (zero, or the target).
__device__ __forceinline__ void box_pair_grad(const BoxTF &t, const BoxTF &p, float g, float out[4])
{
    const float pwr = p.xmax - p.xmin, phr = p.ymax - p.ymin;
    const float tw = fmaxf(0.0f, t.xmax - t.xmin), th = fmaxf(0.0f, t.ymax - t.ymin);
    const float pw = fmaxf(0.0f, pwr), ph = fmaxf(0.0f, phr);
    const float at = tw * th, ap = pw * ph;
    const float iy0 = fmaxf(t.ymin, p.ymin), ix0 = fmaxf(t.xmin, p.xmin);
    const float iy1 = fminf(t.ymax, p.ymax), ix1 = fminf(t.xmax, p.xmax);
    const float iwr = ix1 - ix0, ihr = iy1 - iy0;
    const float iw = fmaxf(0.0f, iwr), ih = fmaxf(0.0f, ihr);
    const float ai = iw * ih;
    const float un = at + ap - ai;
    const float ey0 = fminf(t.ymin, p.ymin), ex0 = fminf(t.xmin, p.xmin);
    const float ey1 = fmaxf(t.ymax, p.ymax), ex1 = fmaxf(t.xmax, p.xmax);
    const float ewr = ex1 - ex0, ehr = ey1 - ey0;
    const float ew = fmaxf(0.0f, ewr), eh = fmaxf(0.0f, ehr);
    const float ae = ew * eh;

    const float d_giou = -2.0f * g;
    // giou = iou - r, r = div_no_nan(ae - un, ae)
    float d_ae = 0.0f, d_un = 0.0f, d_ai = 0.0f;
    if (ae != 0.0f) {
        const float d_r = -d_giou;
        const float d_num = d_r / ae;
        d_ae = d_num - d_r * (ae - un) / (ae * ae);
        d_un = -d_num;
    }
    if (un != 0.0f) {                       // iou = ai / un
        d_ai = d_giou / un;
        d_un += -d_giou * ai / (un * un);
    }
    const float d_ap = d_un;                // un = at + ap - ai
    d_ai -= d_un;

    float d_ymin = 0.0f, d_xmin = 0.0f, d_ymax = 0.0f, d_xmax = 0.0f;
    // enclosing box
    if (ewr > 0.0f) { const float d = d_ae * eh; if (p.xmax > t.xmax) d_xmax += d; if (p.xmin < t.xmin) d_xmin -= d; }
    if (ehr > 0.0f) { const float d = d_ae * ew; if (p.ymax > t.ymax) d_ymax += d; if (p.ymin < t.ymin) d_ymin -= d; }
    // intersection
    if (iwr > 0.0f) { const float d = d_ai * ih; if (p.xmax < t.xmax) d_xmax += d; if (p.xmin > t.xmin) d_xmin -= d; }
    if (ihr > 0.0f) { const float d = d_ai * iw; if (p.ymax < t.ymax) d_ymax += d; if (p.ymin > t.ymin) d_ymin -= d; }
    // prediction area
    if (pwr > 0.0f) { const float d = d_ap * ph; d_xmax += d; d_xmin -= d; }
    if (phr > 0.0f) { const float d = d_ap * pw; d_ymax += d; d_ymin -= d; }
    // l2 term: 5 * mean_4((10t-10p)^2)  ->  d/dp_k = -25 * (10 t_k - 10 p_k)
    const float k = -25.0f * g;
    d_ymin += k * (10.0f * t.ymin - 10.0f * p.ymin);
    d_xmin += k * (10.0f * t.xmin - 10.0f * p.xmin);
    d_ymax += k * (10.0f * t.ymax - 10.0f * p.ymax);
    d_xmax += k * (10.0f * t.xmax - 10.0f * p.xmax);
    out[0] = d_xmin + d_xmax;   // x
    out[1] = d_ymin + d_ymax;   // y
    out[2] = d_xmax;            // w
    out[3] = d_ymax;            // h
}

__device__ __forceinline__ float safe_clip(float v)
{
    return (v != v) ? v : fminf(fmaxf(v, 0.001f), 0.999f);
}
__device__ __forceinline__ bool in_clip(float v) { return v >= 0.001f && v <= 0.999f; }

// -log(clip(p) + 1e-7)
__device__ __forceinline__ float neg_log_clip(float p) { return -logf(__fadd_rn(safe_clip(p), 1e-7f)); }
// focal terms of tfa.SigmoidFocalCrossEntropy(alpha .25, gamma 2) for y = 1 / y = 0
__device__ __forceinline__ float focal1(float pc)
{
    const float om = __fsub_rn(1.0f, pc);
    return __fmul_rn(__fmul_rn(0.25f, __fmul_rn(om, om)), -logf(__fadd_rn(pc, 1e-7f)));
}
__device__ __forceinline__ float focal0(float pc)
{
    return __fmul_rn(__fmul_rn(0.75f, __fmul_rn(pc, pc)), -logf(__fadd_rn(__fsub_rn(1.0f, pc), 1e-7f)));
}

// ---------------------------------------------------------------------------------------------
// K7: cost matrix.  CTA = (image b, tile of QT prediction columns); thread = one column, looping
// over a quarter of the target rows, so each warp stores 128 contiguous bytes per row.
// ---------------------------------------------------------------------------------------------
// a / b for normal, non-zero b: reciprocal estimate refined by Newton steps with exact FMA residuals -- the fast
// path every IEEE-division expansion takes, without its range checks and slow-path call (the operands here are
// box areas / class counts).  Rounds like a true division except in vanishingly rare double-rounding cases.
__device__ __forceinline__ float div_rn_fast(float a, float b)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    r = fmaf(fmaf(-b, r, 1.0f), r, r);
    const float q = __fmul_rn(a, r);
    return fmaf(fmaf(-b, q, a), r, q);
}
__device__ __forceinline__ float div_no_nan_fast(float a, float b) { return b == 0.0f ? 0.0f : div_rn_fast(a, b); }

constexpr int CM_QT = 64;
constexpr int CM_THREADS = 512;
constexpr int CM_TB = 12;        // floats kept per target box: ymin,xmin,ymax,xmax | area, 10*ymin,10*xmin,10*ymax | 10*xmax, nan flag, pad

struct CostSmemLayout {
    int Cs, As, CW, AW;
    uint32_t magicC, magicA;      // e / C == __umulhi(e, magicC) for e < 2^20 (exact: magic = ceil(2^32 / C))
    size_t off_nlc, off_df, off_s0, off_tbox, off_cstar, off_cbits, off_abits, bytes;
};

static CostSmemLayout cost_smem_layout(int T, int C, int A, bool has_attr)
{
    CostSmemLayout L;
    L.Cs = C; L.As = A; L.CW = (C + 31) / 32; L.AW = (A + 31) / 32;      // flat copies: the gather stride C is at worst 2-way conflicted
    L.magicC = (uint32_t)((0x100000000ull + C - 1) / C); L.magicA = (uint32_t)((0x100000000ull + A - 1) / A);
    size_t o = 0;
    L.off_nlc = o; o += sizeof(float) * CM_QT * L.Cs;
    L.off_df = o; if (has_attr) o += sizeof(float) * CM_QT * L.As;
    L.off_s0 = o; o += sizeof(float) * CM_QT;
    L.off_tbox = o; o += sizeof(float) * CM_TB * T;
    L.off_cstar = o; o += sizeof(int) * T;
    L.off_cbits = o; o += sizeof(uint32_t) * T * L.CW;
    L.off_abits = o; if (has_attr) o += sizeof(uint32_t) * T * L.AW;
    L.bytes = o;
    return L;
}

// Box term with the per-box invariants (areas, 10x coordinates) hoisted; same unfused arithmetic as
// box_pair_cost, so the bits are identical.  NaN inputs are handled by the caller (flag), which lets the
// min/max use the plain NaN-suppressing instructions.
__device__ __forceinline__ float box_pair_cost_pre(const float (&t)[CM_TB], const float (&p)[9])
{
    const float iy0 = fmaxf(t[0], p[0]), ix0 = fmaxf(t[1], p[1]);
    const float iy1 = fminf(t[2], p[2]), ix1 = fminf(t[3], p[3]);
    const float iw = fmaxf(0.0f, __fsub_rn(ix1, ix0)), ih = fmaxf(0.0f, __fsub_rn(iy1, iy0));
    const float ai = __fmul_rn(iw, ih);
    const float un = __fsub_rn(__fadd_rn(t[4], p[4]), ai);
    const float iou = div_no_nan_fast(ai, un);
    const float ey0 = fminf(t[0], p[0]), ex0 = fminf(t[1], p[1]);
    const float ey1 = fmaxf(t[2], p[2]), ex1 = fmaxf(t[3], p[3]);
    const float ew = fmaxf(0.0f, __fsub_rn(ex1, ex0)), eh = fmaxf(0.0f, __fsub_rn(ey1, ey0));
    const float ae = __fmul_rn(ew, eh);
    const float giou = __fsub_rn(iou, div_no_nan_fast(__fsub_rn(ae, un), ae));
    const float d0 = __fsub_rn(t[5], p[5]), d1 = __fsub_rn(t[6], p[6]), d2 = __fsub_rn(t[7], p[7]), d3 = __fsub_rn(t[8], p[8]);
    const float ss = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)), __fmul_rn(d3, d3));
    const float l2 = __fmul_rn(ss, 0.25f);                      // == ss / 4 exactly
    return __fadd_rn(__fmul_rn(2.0f, __fsub_rn(1.0f, giou)), __fmul_rn(5.0f, l2));
}
__device__ __forceinline__ void box_invariants(float x, float y, float w, float h, float out[CM_TB])
{
    const BoxTF b = coco_to_tf(x, y, w, h);
    out[0] = b.ymin; out[1] = b.xmin; out[2] = b.ymax; out[3] = b.xmax;
    const float bw = fmaxf(0.0f, __fsub_rn(b.xmax, b.xmin)), bh = fmaxf(0.0f, __fsub_rn(b.ymax, b.ymin));
    out[4] = __fmul_rn(bw, bh);
    out[5] = __fmul_rn(10.0f, b.ymin); out[6] = __fmul_rn(10.0f, b.xmin);
    out[7] = __fmul_rn(10.0f, b.ymax); out[8] = __fmul_rn(10.0f, b.xmax);
    out[9] = (x != x || y != y || w != w || h != h) ? 1.0f : 0.0f;
    out[10] = 0.0f; out[11] = 0.0f;
}

template <bool HAS_ATTR>
__global__ void __launch_bounds__(CM_THREADS)
cost_matrix_kernel(int T, int Q, int C, int A,
                   const float *__restrict__ cat_true, const float *__restrict__ attr_true,
                   const float *__restrict__ box_true, const float *__restrict__ cat_pred,
                   const float *__restrict__ attr_pred, const float *__restrict__ box_pred,
                   float w_cat, float w_box, float w_attr, float *__restrict__ cost, CostSmemLayout L)
{
    pdl_sync();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *nlc = reinterpret_cast<float *>(smem_raw + L.off_nlc);
    float *dfs = reinterpret_cast<float *>(smem_raw + L.off_df);
    float *s0 = reinterpret_cast<float *>(smem_raw + L.off_s0);
    float *tbox = reinterpret_cast<float *>(smem_raw + L.off_tbox);
    int *cstar = reinterpret_cast<int *>(smem_raw + L.off_cstar);
    uint32_t *cbits = reinterpret_cast<uint32_t *>(smem_raw + L.off_cbits);
    uint32_t *abits = reinterpret_cast<uint32_t *>(smem_raw + L.off_abits);

    const int b = blockIdx.y, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int Cs = L.Cs, As = L.As, CW = L.CW, AW = L.AW;

    // ---- target side: flat, division-free scans (independent loads, unrolled by the compiler) ----
    for (int e = tid; e < T * CW; e += CM_THREADS) cbits[e] = 0u;
    if (HAS_ATTR) for (int e = tid; e < T * AW; e += CM_THREADS) abits[e] = 0u;
    __syncthreads();
    const float *ct = cat_true + (size_t)b * T * C;
#pragma unroll 4
    for (int e = tid; e < T * C; e += CM_THREADS) {
        if (ct[e] != 0.0f) { const int t = __umulhi((uint32_t)e, L.magicC), c = e - t * C; atomicOr(&cbits[t * CW + (c >> 5)], 1u << (c & 31)); }
    }
    if (HAS_ATTR) {
        const float *atp = attr_true + (size_t)b * T * A;
#pragma unroll 4
        for (int e = tid; e < T * A; e += CM_THREADS) {
            if (atp[e] != 0.0f) { const int t = __umulhi((uint32_t)e, L.magicA), a = e - t * A; atomicOr(&abits[t * AW + (a >> 5)], 1u << (a & 31)); }
        }
    }
    for (int t = tid; t < T; t += CM_THREADS) {
        const float4 bx = reinterpret_cast<const float4 *>(box_true)[(size_t)b * T + t];
        float inv[CM_TB];
        box_invariants(bx.x, bx.y, bx.z, bx.w, inv);
#pragma unroll
        for (int k = 0; k < CM_TB; k += 4)
            *reinterpret_cast<float4 *>(tbox + t * CM_TB + k) = make_float4(inv[k], inv[k + 1], inv[k + 2], inv[k + 3]);
    }
    const float fC = (float)C, fA = (float)A;
    const bool small_attr = HAS_ATTR && A <= 8;                 // branch-free attribute sum for tiny vocabularies

    // ---- column tiles of this image handled by this CTA ----
    for (int q0 = blockIdx.x * CM_QT; q0 < Q; q0 += gridDim.x * CM_QT) {
        const int nq = min(CM_QT, Q - q0);
        __syncthreads();                                        // bit sets complete / previous tile's readers done
        if (q0 == blockIdx.x * CM_QT) {                         // first tile: single class of one-hot rows
            for (int t = tid; t < T; t += CM_THREADS) {
                int nset = 0, first = -1;
                for (int w = 0; w < CW; ++w) {
                    const uint32_t bits = cbits[t * CW + w];
                    if (bits && first < 0) first = (w << 5) + __ffs(bits) - 1;
                    nset += __popc(bits);
                }
                cstar[t] = nset == 1 ? first : -1;
            }
        }
        // prediction side: flat elementwise transforms of the tile (contiguous in HBM)
        const float *cp = cat_pred + ((size_t)b * Q + q0) * C;
#pragma unroll 4
        for (int e = tid; e < nq * C; e += CM_THREADS) nlc[e] = div_rn_fast(neg_log_clip(cp[e]), fC);
        if (HAS_ATTR) {
            const float *ap = attr_pred + ((size_t)b * Q + q0) * A;
#pragma unroll 4
            for (int e = tid; e < nq * A; e += CM_THREADS) {
                const float pc = safe_clip(ap[e]);
                dfs[e] = __fsub_rn(focal1(pc), focal0(pc));
            }
            // row sums of the y = 0 focal term: one warp per column
            for (int q = warp; q < nq; q += CM_THREADS / 32) {
                float sacc = 0.0f;
                for (int a = lane; a < A; a += 32) sacc += focal0(safe_clip(ap[q * A + a]));
                sacc = warp_sum(sacc);
                if (lane == 0) s0[q] = sacc;
            }
        }
        __syncthreads();

        const int qi = tid % CM_QT, tg = tid / CM_QT;
        if (qi < nq) {
            const int q = q0 + qi;
            const float4 pb = reinterpret_cast<const float4 *>(box_pred)[(size_t)b * Q + q];
            float pinv[CM_TB];
            box_invariants(pb.x, pb.y, pb.z, pb.w, pinv);
            float p9[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) p9[k] = pinv[k];
            const bool pnan = pinv[9] != 0.0f;
            const float *my_nl = nlc + qi * Cs;
            const float *my_df = dfs + qi * As;
            const float my_s0 = HAS_ATTR ? s0[qi] : 0.0f;
            float dfr[8];
            if (small_attr) {
#pragma unroll
                for (int a = 0; a < 8; ++a) dfr[a] = a < A ? my_df[a] : 0.0f;
            }
            float *out = cost + (size_t)b * T * Q + q;

#pragma unroll 2
            for (int t = tg; t < T; t += CM_THREADS / CM_QT) {
                float cat;
                const int cs = cstar[t];
                if (cs >= 0) {
                    cat = __fadd_rn(0.0f, my_nl[cs]);               // one-hot row: a single gather
                } else {
                    cat = 0.0f;
                    for (int w = 0; w < CW; ++w) {
                        uint32_t bits = cbits[t * CW + w];
                        while (bits) { const int c = __ffs(bits) - 1; bits &= bits - 1; cat = __fadd_rn(cat, my_nl[(w << 5) + c]); }
                    }
                }
                float tb[CM_TB];
#pragma unroll
                for (int k = 0; k < CM_TB; k += 4) {
                    const float4 v4 = *reinterpret_cast<const float4 *>(tbox + t * CM_TB + k);
                    tb[k] = v4.x; tb[k + 1] = v4.y; tb[k + 2] = v4.z; tb[k + 3] = v4.w;
                }
                float box = box_pair_cost_pre(tb, p9);
                if (pnan || tb[9] != 0.0f) box = CUDART_NAN_F;      // tf.maximum / minimum propagate NaN
                float v = __fadd_rn(__fmul_rn(w_cat, cat), __fmul_rn(w_box, box));
                if (HAS_ATTR) {
                    float sa = my_s0;
                    if (small_attr) {
                        const uint32_t bits = abits[t * AW];
#pragma unroll
                        for (int a = 0; a < 8; ++a) sa = __fadd_rn(sa, (bits >> a) & 1u ? dfr[a] : 0.0f);   // + 0.0f is exact
                    } else {
                        for (int w = 0; w < AW; ++w) {
                            uint32_t bits = abits[t * AW + w];
                            while (bits) { const int a = __ffs(bits) - 1; bits &= bits - 1; sa = __fadd_rn(sa, my_df[(w << 5) + a]); }
                        }
                    }
                    v = __fadd_rn(v, __fmul_rn(w_attr, div_rn_fast(sa, fA)));
                } else {
                    v = __fadd_rn(v, 0.0f);
                }
                out[(size_t)t * Q] = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K8: validation + per-image shortest-augmenting-path solver (one warp per image).
// ---------------------------------------------------------------------------------------------
__global__ void lsap_validate_kernel(int B, int T, int Q, const float *__restrict__ cost,
                                     const int32_t *__restrict__ num_objects, int32_t *__restrict__ status)
{
    pdl_sync();
    const size_t per = (size_t)T * Q, total = (size_t)B * per;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(e / per);
        const int t = (int)((e - (size_t)b * per) / Q);
        int n = num_objects[b]; n = n < 0 ? 0 : (n > T ? T : n);
        if (t < n) {
            const float v = cost[e];
            if (v != v || v == -CUDART_INF_F) status[b] = BDETR_E_INVALID_COST;
        }
    }
}

constexpr int RANK_FREE = 0x40000000;
constexpr int RANK_USED = 0x3FFFFFFF;

__global__ void __launch_bounds__(32)
lsap_kernel(int T, int Q, const float *__restrict__ cost, const int32_t *__restrict__ num_objects,
            int32_t *__restrict__ col4row_out, int32_t *__restrict__ row4col_out, int32_t *__restrict__ status, int stage_cost)
{
    pdl_sync();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, lane = threadIdx.x;
    const int N = max(T, Q);
    double *v = reinterpret_cast<double *>(smem_raw);
    double *spc = v + N;
    double *u = spc + N;
    int *path = reinterpret_cast<int *>(u + N);
    int *row4col = path + N;
    int *col4row = row4col + N;
    int *remaining = col4row + N;
    int *srlist = remaining + N;
    // [n, Q] copy of the image's live rows (stage_cost only), 16-byte aligned behind the solver state
    float *cost_s = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(srlist + N) + 15) & ~uintptr_t(15));

    int n = num_objects[b]; n = n < 0 ? 0 : (n > T ? T : n);
    int32_t *c4r_o = col4row_out + (size_t)b * T;
    int32_t *r4c_o = row4col_out + (size_t)b * Q;
    for (int t = lane; t < T; t += 32) c4r_o[t] = -1;
    for (int q = lane; q < Q; q += 32) r4c_o[q] = -1;
    if (n == 0 || status[b] != 0) return;

    const bool tr = n > Q;                 // scipy transposes when there are more rows than columns
    const int R = tr ? Q : n, Cn = tr ? n : Q;
    const float *cb = cost + (size_t)b * T * Q;
    // Small problems (BASELINE config 2: 20 x 100) are pure latency: every column scan depends on the previous one, so
    // a global load per scan costs an L2 round trip each time.  The live rows are copied to shared memory once (all
    // loads independent, coalesced); the scans then run at shared-memory latency.  Same values, same order.
    const bool staged = stage_cost && !tr;
    if (staged) {
        const int total = n * Q;
        if ((reinterpret_cast<uintptr_t>(cb) & 15) == 0 && (total & 3) == 0) {
            for (int e = lane; e < (total >> 2); e += 32) reinterpret_cast<float4 *>(cost_s)[e] = reinterpret_cast<const float4 *>(cb)[e];
        } else {
            for (int e = lane; e < total; e += 32) cost_s[e] = cb[e];
        }
        cb = cost_s;
    }

    for (int i = lane; i < R; i += 32) { u[i] = 0.0; col4row[i] = -1; }
    for (int j = lane; j < Cn; j += 32) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }
    __syncwarp();

    bool infeasible = false;
    for (int cur = 0; cur < R; ++cur) {
        double minVal = 0.0;
        int i = cur, nrem = Cn, sink = -1, nsr = 0;
        for (int it = lane; it < Cn; it += 32) { remaining[it] = Cn - it - 1; spc[it] = CUDART_INF; }
        __syncwarp();
        while (sink == -1) {
            if (lane == 0) srlist[nsr] = i;
            ++nsr;
            const double ui = u[i];
            double best_s = CUDART_INF;
            int best_rank = -1;
            const float *crow = tr ? (cb + i) : (cb + (size_t)i * Q);
            // The column scan, 8 strided positions per lane at a time: all cost loads of a group are issued
            // before any of them is used (one L2 round trip per group instead of one per column).
            for (int base = lane; base < nrem; base += 32 * 8) {
                int jj[8];
                float cc[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int it = base + 32 * k;
                    jj[k] = it < nrem ? remaining[it] : -1;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) cc[k] = jj[k] >= 0 ? (tr ? crow[(size_t)jj[k] * Q] : crow[jj[k]]) : 0.0f;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int j = jj[k];
                    if (j < 0) continue;
                    const int it = base + 32 * k;
                    const double r = ((minVal + (double)cc[k]) - ui) - v[j];
                    double s = spc[j];
                    if (r < s) { path[j] = i; spc[j] = r; s = r; }
                    const int rank = (row4col[j] == -1) ? (RANK_FREE + it) : (RANK_USED - it);
                    if (s < best_s || (s == best_s && rank > best_rank)) { best_s = s; best_rank = rank; }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double os = __shfl_xor_sync(0xffffffffu, best_s, o);
                const int orank = __shfl_xor_sync(0xffffffffu, best_rank, o);
                if (os < best_s || (os == best_s && orank > best_rank)) { best_s = os; best_rank = orank; }
            }
            minVal = best_s;
            if (minVal == CUDART_INF) { infeasible = true; break; }
            const int index = best_rank >= RANK_FREE ? best_rank - RANK_FREE : RANK_USED - best_rank;
            const int j = remaining[index];
            const int owner = row4col[j];
            __syncwarp();
            if (lane == 0) { remaining[index] = remaining[nrem - 1]; remaining[nrem - 1] = j; }
            --nrem;
            if (owner == -1) sink = j; else i = owner;
            __syncwarp();
        }
        if (infeasible) break;
        // dual updates (same fp64 expressions as the sequential solver)
        for (int k = lane; k < nsr; k += 32) {
            const int r = srlist[k];
            if (r == cur) u[r] = u[r] + minVal;
            else u[r] = u[r] + (minVal - spc[col4row[r]]);
        }
        for (int k = nrem + lane; k < Cn; k += 32) {
            const int j = remaining[k];
            v[j] = v[j] - (minVal - spc[j]);
        }
        __syncwarp();
        if (lane == 0) {
            int j = sink;
            for (;;) {
                const int r = path[j];
                row4col[j] = r;
                const int tmp = col4row[r]; col4row[r] = j; j = tmp;
                if (r == cur) break;
            }
        }
        __syncwarp();
    }
    if (infeasible) { if (lane == 0) status[b] = BDETR_E_INFEASIBLE; return; }
    if (!tr) {
        for (int i = lane; i < R; i += 32) { const int j = col4row[i]; c4r_o[i] = j; r4c_o[j] = i; }
    } else {
        for (int i = lane; i < R; i += 32) { const int j = col4row[i]; c4r_o[j] = i; r4c_o[i] = j; }
    }
}

// mask [B,T,Q] and assigned [B,Q] from the index form: the bandwidth-bound part of K8.
__global__ void lsap_mask_kernel(int B, int T, int Q, const int32_t *__restrict__ col4row,
                                 const int32_t *__restrict__ row4col, float *__restrict__ mask,
                                 float *__restrict__ assigned)
{
    pdl_sync();
    const size_t rows = (size_t)B * T;
    if (mask) {
        if ((Q & 3) == 0) {
            const int Q4 = Q >> 2;
            const size_t total = rows * Q4;
            for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
                const size_t row = e / Q4;
                const int q = (int)(e - row * Q4) << 2;
                const int c = col4row[row];
                float4 o; o.x = (c == q) ? 1.0f : 0.0f; o.y = (c == q + 1) ? 1.0f : 0.0f;
                o.z = (c == q + 2) ? 1.0f : 0.0f; o.w = (c == q + 3) ? 1.0f : 0.0f;
                reinterpret_cast<float4 *>(mask)[e] = o;
            }
        } else {
            const size_t total = rows * Q;
            for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
                const size_t row = e / Q;
                const int q = (int)(e - row * Q);
                mask[e] = (col4row[row] == q) ? 1.0f : 0.0f;
            }
        }
    }
    if (assigned) {
        const size_t total = (size_t)B * Q;
        for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x)
            assigned[e] = row4col[e] >= 0 ? 1.0f : 0.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// K9: matched loss forward / backward.
// ---------------------------------------------------------------------------------------------
constexpr int ML_THREADS = 128;

__device__ __forceinline__ float block_sum(float v, float *red)
{
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float s = 0.0f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += red[k];
    return s;
}

__device__ __forceinline__ float total_objects(const int32_t *num_objects, int B)
{
    // 1 + sum_b num_objects (losses_and_metrics.py:144); exact in fp32 for any realistic batch
    int s = 0;
    for (int k = 0; k < B; ++k) s += num_objects[k];
    return 1.0f + (float)s;
}

__global__ void __launch_bounds__(ML_THREADS)
matched_loss_fwd_kernel(int B, int T, int Q, int C, int A,
                        const float *__restrict__ cat_true, const float *__restrict__ attr_true,
                        const float *__restrict__ box_true, const int32_t *__restrict__ num_objects,
                        const float *__restrict__ cat_pred, const float *__restrict__ attr_pred,
                        const float *__restrict__ box_pred, const int32_t *__restrict__ col4row,
                        const int32_t *__restrict__ row4col,
                        float w_cat, float w_box, float w_attr, float w_exist,
                        float *__restrict__ losses, float *__restrict__ iou)
{
    pdl_sync();
    __shared__ float red[ML_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float total_n = total_objects(num_objects, B);
    float cat_s = 0.0f, attr_s = 0.0f, box_s = 0.0f, ex_s = 0.0f;

    for (int t = tid; t < T; t += ML_THREADS) {
        const int q = col4row[(size_t)b * T + t];
        if (q < 0) continue;
        const float *ct = cat_true + ((size_t)b * T + t) * C;
        const float *cp = cat_pred + ((size_t)b * Q + q) * C;
        float cat = 0.0f;
        for (int c = 0; c < C; ++c) if (ct[c] != 0.0f) cat += ct[c] * neg_log_clip(cp[c]);
        cat_s += w_cat * (cat / (float)C);
        if (w_attr != 0.0f) {
            const float *at = attr_true + ((size_t)b * T + t) * A;
            const float *ap = attr_pred + ((size_t)b * Q + q) * A;
            float s = 0.0f;
            for (int a = 0; a < A; ++a) { const float pc = safe_clip(ap[a]); s += (at[a] != 0.0f) ? focal1(pc) : focal0(pc); }
            attr_s += w_attr * (s / (float)A);
        }
        const float4 tb4 = reinterpret_cast<const float4 *>(box_true)[(size_t)b * T + t];
        const float4 pb4 = reinterpret_cast<const float4 *>(box_pred)[(size_t)b * Q + q];
        float iou_v;
        const float bc = box_pair_cost(coco_to_tf(tb4.x, tb4.y, tb4.z, tb4.w), coco_to_tf(pb4.x, pb4.y, pb4.z, pb4.w), &iou_v);
        box_s += w_box * bc;
        // IOU metric = 1 - (1 - iou), summed over (b,t) per prediction column (quirk Q7)
        atomicAdd(&iou[q], (1.0f - (1.0f - iou_v)) / total_n);
    }
    for (int q = tid; q < Q; q += ML_THREADS) {
        const float y = row4col[(size_t)b * Q + q] >= 0 ? 0.0f : 1.0f;      // 1 - assigned
        const float pc = safe_clip(cat_pred[((size_t)b * Q + q) * C]);
        const float bce = -(y * logf(pc + 1e-7f) + (1.0f - y) * logf(1.0f - pc + 1e-7f));
        ex_s += w_exist * bce;
    }
    cat_s = block_sum(cat_s, red);
    attr_s = block_sum(attr_s, red);
    box_s = block_sum(box_s, red);
    ex_s = block_sum(ex_s, red);
    if (tid == 0) {
        const float cat_l = cat_s / total_n, attr_l = attr_s / total_n, box_l = box_s / total_n;
        const float ex_l = (ex_s / (float)Q) / (1.0f + (float)Q);
        losses[0 * B + b] = ((cat_l + attr_l) + box_l) + ex_l;
        losses[1 * B + b] = cat_l;
        losses[2 * B + b] = attr_l;
        losses[3 * B + b] = box_l;
        losses[4 * B + b] = ex_l;
    }
}

// one thread per (b, q): writes/accumulates the whole gradient rows of that prediction
__global__ void __launch_bounds__(ML_THREADS)
matched_loss_bwd_kernel(int B, int T, int Q, int C, int A,
                        const float *__restrict__ cat_true, const float *__restrict__ attr_true,
                        const float *__restrict__ box_true, const int32_t *__restrict__ num_objects,
                        const float *__restrict__ cat_pred, const float *__restrict__ attr_pred,
                        const float *__restrict__ box_pred, const int32_t *__restrict__ row4col,
                        float w_cat, float w_box, float w_attr, float w_exist, float gscale,
                        float *__restrict__ d_cat, float *__restrict__ d_attr, float *__restrict__ d_box)
{
    pdl_sync();
    const int e = blockIdx.x * ML_THREADS + threadIdx.x;
    if (e >= B * Q) return;
    const int b = e / Q;
    const float total_n = total_objects(num_objects, B);
    const int t = row4col[e];
    const float *cp = cat_pred + (size_t)e * C;
    float *dc = d_cat + (size_t)e * C;
    // existence term on class 0
    {
        const float p0 = cp[0];
        if (in_clip(p0)) {
            const float y = t >= 0 ? 0.0f : 1.0f;
            const float dbce = -(y / (p0 + 1e-7f) - (1.0f - y) / (1.0f - p0 + 1e-7f));
            dc[0] += gscale * w_exist * dbce / ((float)Q * (1.0f + (float)Q));
        }
    }
    if (t < 0) return;
    const float gs = gscale / total_n;
    const float *ct = cat_true + ((size_t)b * T + t) * C;
    for (int c = 0; c < C; ++c) {
        const float y = ct[c];
        if (y != 0.0f && in_clip(cp[c])) dc[c] += gs * w_cat * y * (-1.0f / (cp[c] + 1e-7f)) / (float)C;
    }
    if (w_attr != 0.0f) {
        const float *at = attr_true + ((size_t)b * T + t) * A;
        const float *ap = attr_pred + (size_t)e * A;
        float *da = d_attr + (size_t)e * A;
        for (int a = 0; a < A; ++a) {
            const float p = ap[a];
            if (!in_clip(p)) continue;
            float d;
            if (at[a] != 0.0f) {
                const float om = 1.0f - p, nl = -logf(p + 1e-7f);
                d = 0.25f * (-2.0f * om * nl - om * om / (p + 1e-7f));
            } else {
                const float nl = -logf(1.0f - p + 1e-7f);
                d = 0.75f * (2.0f * p * nl + p * p / (1.0f - p + 1e-7f));
            }
            da[a] += gs * w_attr * d / (float)A;
        }
    }
    if (w_box != 0.0f) {
        const float4 tb4 = reinterpret_cast<const float4 *>(box_true)[(size_t)b * T + t];
        const float4 pb4 = reinterpret_cast<const float4 *>(box_pred)[e];
        float g4[4];
        box_pair_grad(coco_to_tf(tb4.x, tb4.y, tb4.z, tb4.w), coco_to_tf(pb4.x, pb4.y, pb4.z, pb4.w), gs * w_box, g4);
        float4 *db = reinterpret_cast<float4 *>(d_box) + e;
        float4 cur = *db;
        cur.x += g4[0]; cur.y += g4[1]; cur.z += g4[2]; cur.w += g4[3];
        *db = cur;
    }
}

}  // namespace bdetr

using namespace bdetr;

extern "C" __attribute__((visibility("default"))) int bdetr_cost_matrix_fwd(int B, int T, int Q, int C, int A,
                                     const float *cat_true, const float *attr_true, const float *box_true,
                                     const float *cat_pred, const float *attr_pred, const float *box_pred,
                                     float w_cat, float w_box, float w_attr, float *cost, void *stream)
{
    BDETR_REQUIRE(B > 0 && T > 0 && Q > 0 && C > 0 && A > 0, BDETR_E_BAD_SHAPE, "B,T,Q,C,A must be positive");
    BDETR_REQUIRE(cat_true && attr_true && box_true && cat_pred && attr_pred && box_pred && cost, BDETR_E_NULL, "null pointer");
    const bool has_attr = (w_attr != 0.0f);
    const CostSmemLayout L = cost_smem_layout(T, C, A, has_attr);
    BDETR_REQUIRE(L.bytes <= 227 * 1024, BDETR_E_UNSUPPORTED, "C/A/T too large for the shared-memory tile");
    dim3 grid(ceil_div(Q, CM_QT), B);
    // opt in to large dynamic shared memory once per size class (never during a later graph capture)
    static size_t optin[2] = {0, 0};
    if (L.bytes > optin[has_attr]) {
        if (has_attr) BDETR_CUDA(cudaFuncSetAttribute(cost_matrix_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.bytes));
        else BDETR_CUDA(cudaFuncSetAttribute(cost_matrix_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.bytes));
        optin[has_attr] = L.bytes;
    }
    if (has_attr) {
        launch_k(cost_matrix_kernel<true>, grid, CM_THREADS, L.bytes, as_stream(stream), 
            T, Q, C, A, cat_true, attr_true, box_true, cat_pred, attr_pred, box_pred, w_cat, w_box, w_attr, cost, L);
    } else {
        launch_k(cost_matrix_kernel<false>, grid, CM_THREADS, L.bytes, as_stream(stream), 
            T, Q, C, A, cat_true, attr_true, box_true, cat_pred, attr_pred, box_pred, w_cat, w_box, w_attr, cost, L);
    }
    BDETR_CHECK_LAUNCH("cost_matrix_kernel");
    return BDETR_OK;
}

extern "C" __attribute__((visibility("default"))) size_t bdetr_lsap_smem_bytes(int T, int Q)
{
    const size_t N = (size_t)(T > Q ? T : Q);
    return N * (3 * sizeof(double) + 5 * sizeof(int));
}

extern "C" __attribute__((visibility("default"))) int bdetr_lsap_assign(int B, int T, int Q, const float *cost, const int32_t *num_objects,
                                 int32_t *col4row, int32_t *row4col, float *mask, float *assigned,
                                 int32_t *status, void *stream)
{
    BDETR_REQUIRE(B > 0 && T > 0 && Q > 0, BDETR_E_BAD_SHAPE, "B,T,Q must be positive");
    BDETR_REQUIRE(cost && num_objects && col4row && row4col && status, BDETR_E_NULL, "null pointer");
    size_t smem = bdetr_lsap_smem_bytes(T, Q);
    BDETR_REQUIRE(smem <= 227 * 1024, BDETR_E_UNSUPPORTED, "T/Q too large for the shared-memory solver");
    // small cost matrices are staged in shared memory (latency-bound regime); large ones stay in L2 so that many
    // images fit on an SM
    const int stage = ((size_t)T * Q * sizeof(float) <= 48 * 1024) ? 1 : 0;
    if (stage) smem += (size_t)T * Q * sizeof(float) + 16;
    cudaStream_t s = as_stream(stream);
    BDETR_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t) * B, s));
    {
        const size_t total = (size_t)B * T * Q;
        const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
        launch_k(lsap_validate_kernel, blocks, 256, 0, s, B, T, Q, cost, num_objects, status);
        BDETR_CHECK_LAUNCH("lsap_validate_kernel");
    }
    static size_t lsap_optin = 0;
    if (smem > lsap_optin) {
        BDETR_CUDA(cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lsap_optin = smem;
    }
    launch_k(lsap_kernel, B, 32, smem, s, T, Q, cost, num_objects, col4row, row4col, status, stage);
    BDETR_CHECK_LAUNCH("lsap_kernel");
    if (mask || assigned) {
        const size_t total = (size_t)B * T * Q / 4 + 1;
        const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
        launch_k(lsap_mask_kernel, blocks, 256, 0, s, B, T, Q, col4row, row4col, mask, assigned);
        BDETR_CHECK_LAUNCH("lsap_mask_kernel");
    }
    return BDETR_OK;
}

extern "C" __attribute__((visibility("default"))) int bdetr_matched_loss_fwd(int B, int T, int Q, int C, int A,
                                      const float *cat_true, const float *attr_true, const float *box_true,
                                      const int32_t *num_objects,
                                      const float *cat_pred, const float *attr_pred, const float *box_pred,
                                      const int32_t *col4row, const int32_t *row4col,
                                      float w_cat, float w_box, float w_attr, float w_exist,
                                      float *losses, float *iou, void *stream)
{
    BDETR_REQUIRE(B > 0 && T > 0 && Q > 0 && C > 0 && A > 0, BDETR_E_BAD_SHAPE, "B,T,Q,C,A must be positive");
    BDETR_REQUIRE(cat_true && attr_true && box_true && num_objects && cat_pred && attr_pred && box_pred &&
                  col4row && row4col && losses && iou, BDETR_E_NULL, "null pointer");
    cudaStream_t s = as_stream(stream);
    BDETR_CUDA(cudaMemsetAsync(iou, 0, sizeof(float) * Q, s));
    launch_k(matched_loss_fwd_kernel, B, ML_THREADS, 0, s, B, T, Q, C, A, cat_true, attr_true, box_true, num_objects,
                                                    cat_pred, attr_pred, box_pred, col4row, row4col,
                                                    w_cat, w_box, w_attr, w_exist, losses, iou);
    BDETR_CHECK_LAUNCH("matched_loss_fwd_kernel");
    return BDETR_OK;
}

extern "C" __attribute__((visibility("default"))) int bdetr_matched_loss_bwd(int B, int T, int Q, int C, int A,
                                      const float *cat_true, const float *attr_true, const float *box_true,
                                      const int32_t *num_objects,
                                      const float *cat_pred, const float *attr_pred, const float *box_pred,
                                      const int32_t *col4row, const int32_t *row4col,
                                      float w_cat, float w_box, float w_attr, float w_exist, float gscale,
                                      float *d_cat_pred, float *d_attr_pred, float *d_box_pred, void *stream)
{
    (void)col4row;
    BDETR_REQUIRE(B > 0 && T > 0 && Q > 0 && C > 0 && A > 0, BDETR_E_BAD_SHAPE, "B,T,Q,C,A must be positive");
    BDETR_REQUIRE(cat_true && attr_true && box_true && num_objects && cat_pred && attr_pred && box_pred &&
                  row4col && d_cat_pred && d_attr_pred && d_box_pred, BDETR_E_NULL, "null pointer");
    launch_k(matched_loss_bwd_kernel, ceil_div(B * Q, ML_THREADS), ML_THREADS, 0, as_stream(stream), 
        B, T, Q, C, A, cat_true, attr_true, box_true, num_objects, cat_pred, attr_pred, box_pred, row4col,
        w_cat, w_box, w_attr, w_exist, gscale, d_cat_pred, d_attr_pred, d_box_pred);
    BDETR_CHECK_LAUNCH("matched_loss_bwd_kernel");
    return BDETR_OK;
}
