// K7 cost matrix, K8 per-image assignment, K9 matched loss (+ backward) for sm_100a.
//
// Reference: /root/reference/ModelComponents/losses_and_metrics.py
//   CostArray :215-225, CategoryLoss :44-49, AttributeLoss :51-57, BoxLoss :68-72, coco_to_tf :59-66,
//   MatchingAssignment :228-251, MatchingMask :195-212, MatchingLoss.call :111-161, ExistLoss :33-37.
// These are HBM-bound fp32 / latency-bound fp64 kernels: plain CUDA cores, coalesced stores,
// shared-memory staging, no tensor cores (DESIGN.md "matcher kernels").
#include <math_constants.h>
#include "common.cuh"
#include "umma.cuh"

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
// K7: cost matrix.  CTA = (image b, tile of 64 prediction columns), 8 warps.  A thread owns TWO columns (lane and
// lane + 32) and walks the target rows of its warp (t = warp, warp + 8, ...), so
//   * the per-row data (box, hoisted invariants, class slot, attribute bits) is one 64-byte shared-memory record read
//     with four broadcast LDS.128 and shared by both columns,
//   * all add / sub / mul / fma work runs as packed f32x2 instructions (FADD2 / FFMA2: one issue slot for both
//     columns; every lane is IEEE round-to-nearest, so the bits equal the scalar op-by-op evaluation),
//   * each warp stores two full 128-byte lines per row.
// The kernel is issue-bound, not HBM-bound (~45 instructions per pair after packing against 12.7 bytes per pair), so
// the design removes instructions: -log(clip(p)+1e-7)/C is evaluated only for the classes that occur in the image's
// targets (compacted "slots"), the attribute term of small vocabularies (A <= 4) is a 2^A-entry table per column, and
// divisions are the Newton fast path.
// ---------------------------------------------------------------------------------------------
typedef uint64_t f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// Product rounded once and NEVER contracted with a following add: ptxas (12.9) fuses mul.rn.f32x2 + add.rn.f32x2 into
// one FFMA2 even with --fmad=false, which would change the bits.  An FMA whose addend is a -0.0 pair that only exists
// at run time (kernel argument `nz`) is the exact product and cannot be folded.
#define MUL2(a, b) fma2((a), (b), nz)

// a / b lane-wise, the same Newton sequence as the scalar fast path: b == 0 lanes return 0 (divide_no_nan)
__device__ __forceinline__ f32x2 div2_no_nan(f32x2 a, f32x2 b, f32x2 nz)
{
    float b0, b1, r0, r1;
    upk2(b, b0, b1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(b0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(b1));
    const f32x2 one = pk2(1.0f, 1.0f), nb = sub2(nz, b);              // -0 - b == -b exactly
    f32x2 r = pk2(r0, r1);
    r = fma2(fma2(nb, r, one), r, r);
    const f32x2 q = MUL2(a, r);
    float q0, q1;
    upk2(fma2(fma2(nb, q, a), r, q), q0, q1);
    return pk2(b0 == 0.0f ? 0.0f : q0, b1 == 0.0f ? 0.0f : q1);
}
// a / c for a per-thread constant c != 0 with rc = refined reciprocal of c (hoisted)
__device__ __forceinline__ float rcp_refined(float c)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(c));
    return fmaf(fmaf(-c, r, 1.0f), r, r);
}
__device__ __forceinline__ float div_by_const(float a, float c, float rc)
{
    const float q = __fmul_rn(a, rc);
    return fmaf(fmaf(-c, q, a), rc, q);
}
__device__ __forceinline__ f32x2 div2_by_const(f32x2 a, float c, float rc, f32x2 nz)
{
    const f32x2 R = pk2(rc, rc), NC = pk2(-c, -c);
    const f32x2 q = MUL2(a, R);
    return fma2(fma2(NC, q, a), R, q);
}

constexpr int CM_QT = 64;          // prediction columns per CTA (two per lane)
constexpr int CM_THREADS = 256;
constexpr int CM_WARPS = CM_THREADS / 32;
constexpr int CM_ROW = 16;         // floats per target-row record
constexpr int CM_SMALL_A = 4;      // attribute vocabularies up to this size use the 2^A table
constexpr int CM_TAB = 17;         // row stride of that table (odd: conflict-free for a warp-uniform mask)

// Prepared targets of one image ("blob", built once per batch by cost_targets_kernel, read by every column tile of
// every boosted block through one TMA bulk copy):
//   header  4 x u32   [0] number of class slots (distinct classes that occur), [1] magic divisor ceil(2^32 / slots)
//   rows    T x 16 floats: [0..3] ymin xmin ymax xmax | [4,5] area twice | [6..13] 10*ymin, 10*xmin, 10*ymax, 10*xmax,
//           each twice (ready-made f32x2 operands) | [14] meta: bits 0..15 class slot + 1 (0 = row is not one-hot ->
//           general path), bit 30 = row identical to the last row (padding), bit 31 = NaN box | [15] attribute bit word 0
//   cls_of  slot -> class (i16), slot_of class -> slot (i16), cbits [T][CW], abits [T][AW]
struct CostBlobLayout {
    int CW, AW;
    uint32_t off_rows, off_clsof, off_slotof, off_cbits, off_abits, bytes;      // every offset and the size are multiples of 16
};
__host__ __device__ inline uint32_t cm_align16(uint32_t v) { return (v + 15u) & ~15u; }
__host__ __device__ inline CostBlobLayout cost_blob_layout(int T, int C, int A)
{
    CostBlobLayout L;
    L.CW = (C + 31) / 32; L.AW = (A + 31) / 32;
    uint32_t o = 16;
    L.off_rows = o; o += (uint32_t)T * CM_ROW * 4u;
    L.off_clsof = o; o = cm_align16(o + (uint32_t)C * 2u);
    L.off_slotof = o; o = cm_align16(o + (uint32_t)C * 2u);
    L.off_cbits = o; o = cm_align16(o + (uint32_t)T * L.CW * 4u);
    L.off_abits = o; o = cm_align16(o + (uint32_t)T * L.AW * 4u);
    L.bytes = o;
    return L;
}

struct CostSmemLayout {
    int Cs, As;
    size_t off_bar, off_raw, off_blob, off_nlc, off_attr, off_s0, bytes;
};

static CostSmemLayout cost_smem_layout(int T, int C, int A, bool has_attr)
{
    CostSmemLayout L;
    L.Cs = C | 1; L.As = A | 1;                                        // odd strides: column-strided gathers are conflict-free
    const bool small_a = A <= CM_SMALL_A;
    size_t o = 0;
    L.off_bar = o; o += 16;
    L.off_raw = o; o += sizeof(float) * CM_QT * C;                     // TMA destination: the tile's class probabilities as they sit in HBM
    o = (o + 15) & ~size_t(15);
    L.off_blob = o; o += cost_blob_layout(T, C, A).bytes;              // TMA destination: the image's prepared targets
    L.off_nlc = o; o += sizeof(float) * CM_QT * L.Cs;
    L.off_attr = o; if (has_attr) o += sizeof(float) * CM_QT * (small_a ? CM_TAB : L.As);
    L.off_s0 = o; o += sizeof(float) * CM_QT;
    L.bytes = o;
    return L;
}

struct PredBox2 {                  // two prediction columns: scalars for min / max, packed invariants for the rest
    float y0a, x0a, y1a, x1a, y0b, x0b, y1b, x1b;
    f32x2 area, ty0, tx0, ty1, tx1;
};

__device__ __forceinline__ void pred_invariants(float x, float y, float w, float h, float out[9], bool &isnan_)
{
    const BoxTF b = coco_to_tf(x, y, w, h);
    out[0] = b.ymin; out[1] = b.xmin; out[2] = b.ymax; out[3] = b.xmax;
    const float bw = fmaxf(0.0f, __fsub_rn(b.xmax, b.xmin)), bh = fmaxf(0.0f, __fsub_rn(b.ymax, b.ymin));
    out[4] = __fmul_rn(bw, bh);
    out[5] = __fmul_rn(10.0f, b.ymin); out[6] = __fmul_rn(10.0f, b.xmin);
    out[7] = __fmul_rn(10.0f, b.ymax); out[8] = __fmul_rn(10.0f, b.xmax);
    isnan_ = (x != x || y != y || w != w || h != h);
}

// Box term of both columns against one target row; op order is box_pair_cost's (NaN inputs are flagged by the caller,
// which lets min / max be the plain NaN-suppressing instructions).
__device__ __forceinline__ f32x2 box_pair_cost_x2(const float4 tb, const f32x2 t_area, const f32x2 t0, const f32x2 t1,
                                                  const f32x2 t2, const f32x2 t3, const PredBox2 &p, const f32x2 nz)
{
    const f32x2 IX0 = pk2(fmaxf(tb.y, p.x0a), fmaxf(tb.y, p.x0b)), IY0 = pk2(fmaxf(tb.x, p.y0a), fmaxf(tb.x, p.y0b));
    const f32x2 IX1 = pk2(fminf(tb.w, p.x1a), fminf(tb.w, p.x1b)), IY1 = pk2(fminf(tb.z, p.y1a), fminf(tb.z, p.y1b));
    float wa, wb, ha, hb;
    upk2(sub2(IX1, IX0), wa, wb); upk2(sub2(IY1, IY0), ha, hb);
    const f32x2 AI = MUL2(pk2(fmaxf(0.0f, wa), fmaxf(0.0f, wb)), pk2(fmaxf(0.0f, ha), fmaxf(0.0f, hb)));
    const f32x2 UN = sub2(add2(t_area, p.area), AI);
    const f32x2 IOU = div2_no_nan(AI, UN, nz);
    const f32x2 EX0 = pk2(fminf(tb.y, p.x0a), fminf(tb.y, p.x0b)), EY0 = pk2(fminf(tb.x, p.y0a), fminf(tb.x, p.y0b));
    const f32x2 EX1 = pk2(fmaxf(tb.w, p.x1a), fmaxf(tb.w, p.x1b)), EY1 = pk2(fmaxf(tb.z, p.y1a), fmaxf(tb.z, p.y1b));
    upk2(sub2(EX1, EX0), wa, wb); upk2(sub2(EY1, EY0), ha, hb);
    const f32x2 AE = MUL2(pk2(fmaxf(0.0f, wa), fmaxf(0.0f, wb)), pk2(fmaxf(0.0f, ha), fmaxf(0.0f, hb)));
    const f32x2 GIOU = sub2(IOU, div2_no_nan(sub2(AE, UN), AE, nz));
    const f32x2 D0 = sub2(t0, p.ty0), D1 = sub2(t1, p.tx0), D2 = sub2(t2, p.ty1), D3 = sub2(t3, p.tx1);
    const f32x2 SS = add2(add2(add2(MUL2(D0, D0), MUL2(D1, D1)), MUL2(D2, D2)), MUL2(D3, D3));
    const f32x2 L2 = MUL2(SS, pk2(0.25f, 0.25f));                      // == ss / 4 exactly
    return add2(MUL2(pk2(2.0f, 2.0f), sub2(pk2(1.0f, 1.0f), GIOU)), MUL2(pk2(5.0f, 5.0f), L2));
}

// 1-D bulk copy global -> shared through the TMA engine, completion on an mbarrier (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr uint32_t CM_META_TAIL = 1u << 30, CM_META_NAN = 1u << 31, CM_META_SLOT = 0xffffu;

// K7a: target side, one CTA per image.  The one-hot / multi-hot blocks are staged in shared memory by TMA, turned into
// bit words by warp ballots (no atomics, no index arithmetic), and reduced to the blob described above.
__global__ void __launch_bounds__(CM_THREADS)
cost_targets_kernel(int T, int C, int A, const float *__restrict__ cat_true, const float *__restrict__ attr_true,
                    const float *__restrict__ box_true, unsigned char *__restrict__ blobs, CostBlobLayout BL)
{
    pdl_sync();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *present = reinterpret_cast<uint32_t *>(smem_raw + 16);           // CW words (<= 1024 classes)
    unsigned char *blob = smem_raw + 16 + 128;
    float *st_c = reinterpret_cast<float *>(blob + BL.bytes);
    float *st_a = st_c + (((size_t)T * C + 3) & ~size_t(3));
    uint32_t *hdr = reinterpret_cast<uint32_t *>(blob);
    float *rows = reinterpret_cast<float *>(blob + BL.off_rows);
    int16_t *cls_of = reinterpret_cast<int16_t *>(blob + BL.off_clsof);
    int16_t *slot_of = reinterpret_cast<int16_t *>(blob + BL.off_slotof);
    uint32_t *cbits = reinterpret_cast<uint32_t *>(blob + BL.off_cbits);
    uint32_t *abits = reinterpret_cast<uint32_t *>(blob + BL.off_abits);

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int CW = BL.CW, AW = BL.AW;
    const float *ct = cat_true + (size_t)b * T * C;
    const float *atp = attr_true + (size_t)b * T * A;
    const uint32_t cb = (uint32_t)(T * C) * 4u, ab = (uint32_t)(T * A) * 4u;
    const bool tma_c = ((reinterpret_cast<uintptr_t>(ct) | cb) & 15) == 0, tma_a = ((reinterpret_cast<uintptr_t>(atp) | ab) & 15) == 0;
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar, (tma_c ? cb : 0u) + (tma_a ? ab : 0u));
        if (tma_c) bulk_load_1d(st_c, ct, cb, bar);
        if (tma_a) bulk_load_1d(st_a, atp, ab, bar);
    }
    if (tid < CW) present[tid] = 0u;
    if (!tma_c) {
#pragma unroll 8
        for (int e = tid; e < T * C; e += CM_THREADS) st_c[e] = ct[e];
    }
    if (!tma_a) {
#pragma unroll 8
        for (int e = tid; e < T * A; e += CM_THREADS) st_a[e] = atp[e];
    }
    for (int t = tid; t < T; t += CM_THREADS) {
        const float4 bx = reinterpret_cast<const float4 *>(box_true)[(size_t)b * T + t];
        float inv[9]; bool tn;
        pred_invariants(bx.x, bx.y, bx.z, bx.w, inv, tn);
        float4 *r4 = reinterpret_cast<float4 *>(rows + t * CM_ROW);
        r4[0] = make_float4(inv[0], inv[1], inv[2], inv[3]);
        r4[1] = make_float4(inv[4], inv[4], inv[5], inv[5]);
        r4[2] = make_float4(inv[6], inv[6], inv[7], inv[7]);
        r4[3] = make_float4(inv[8], inv[8], __uint_as_float(tn ? CM_META_NAN : 0u), 0.0f);
    }
    __syncthreads();
    mbar_wait(bar, 0);

    {
        // lane k keeps bit word k of the current row (one coalesced store per row) and the OR over this warp's rows
        uint32_t seen = 0u;
#pragma unroll 2
        for (int t = warp; t < T; t += CM_WARPS) {
            uint32_t mine = 0u;
            for (int k = 0; k < CW; ++k) {
                const int c = (k << 5) + lane;
                const uint32_t bits = __ballot_sync(0xffffffffu, c < C && st_c[t * C + c] != 0.0f);
                if (lane == k) mine = bits;
            }
            if (lane < CW) cbits[t * CW + lane] = mine;
            seen |= mine;
            mine = 0u;
            for (int k = 0; k < AW; ++k) {
                const int a = (k << 5) + lane;
                const uint32_t bits = __ballot_sync(0xffffffffu, a < A && st_a[t * A + a] != 0.0f);
                if (lane == k) mine = bits;
            }
            if (lane < AW) abits[t * AW + lane] = mine;
        }
        if (lane < CW && seen) atomicOr(&present[lane], seen);
    }
    __syncthreads();
    // class -> slot (rank among the classes that occur in this image), slot -> class
    for (int c = tid; c < C; c += CM_THREADS) {
        const int w = c >> 5;
        const uint32_t word = present[w];
        int s = -1;
        if ((word >> (c & 31)) & 1u) {
            s = __popc(word & ((1u << (c & 31)) - 1u));
            for (int k = 0; k < w; ++k) s += __popc(present[k]);
            cls_of[s] = (int16_t)c;
        }
        slot_of[c] = (int16_t)s;
    }
    if (tid == 0) {
        int n = 0;
        for (int k = 0; k < CW; ++k) n += __popc(present[k]);
        hdr[0] = (uint32_t)n;
        hdr[1] = n > 0 ? (uint32_t)((0x100000000ull + n - 1) / n) : 0u;
        hdr[2] = 0u; hdr[3] = 0u;
    }
    __syncthreads();
    uint32_t meta_new[1];                                                // T <= CM_THREADS is not assumed: loop, one row per pass
    for (int t0 = 0; t0 < T; t0 += CM_THREADS) {
        const int t = t0 + tid;
        if (t < T) {
            int nset = 0, first = -1;
            bool same = true;                                            // identical to the last row, bit for bit?
            for (int w = 0; w < CW; ++w) {
                const uint32_t bits = cbits[t * CW + w];
                if (bits && first < 0) first = (w << 5) + __ffs(bits) - 1;
                nset += __popc(bits);
                same = same && bits == cbits[(T - 1) * CW + w];
            }
            for (int w = 0; w < AW; ++w) same = same && abits[t * AW + w] == abits[(T - 1) * AW + w];
            const uint4 mine = *reinterpret_cast<const uint4 *>(rows + t * CM_ROW), last = *reinterpret_cast<const uint4 *>(rows + (T - 1) * CM_ROW);
            same = same && mine.x == last.x && mine.y == last.y && mine.z == last.z && mine.w == last.w;
            uint32_t meta = __float_as_uint(rows[t * CM_ROW + 14]);
            same = same && meta == __float_as_uint(rows[(T - 1) * CM_ROW + 14]);   // NaN flags (nothing else is set yet)
            if (nset == 1) meta |= (uint32_t)(slot_of[first] + 1);
            if (same) meta |= CM_META_TAIL;
            meta_new[0] = meta;
        }
        __syncthreads();                                                 // every comparison against the last row's old meta is done
        if (t < T) {
            rows[t * CM_ROW + 14] = __uint_as_float(meta_new[0]);
            rows[t * CM_ROW + 15] = __uint_as_float(abits[t * AW]);
        }
        __syncthreads();
    }
    uint4 *dst = reinterpret_cast<uint4 *>(blobs + (size_t)b * BL.bytes);
    for (uint32_t e = tid; e < BL.bytes / 16; e += CM_THREADS) dst[e] = reinterpret_cast<const uint4 *>(blob)[e];
}

// K7b: the pairs.  CTA = (image b, tile of 64 prediction columns).
template <bool HAS_ATTR, bool SMALL_A>
__global__ void __launch_bounds__(CM_THREADS)
cost_matrix_kernel(int T, int Q, int C, int A, const unsigned char *__restrict__ blobs,
                   const float *__restrict__ cat_pred, const float *__restrict__ attr_pred, const float *__restrict__ box_pred,
                   float w_cat, float w_box, float w_attr, float *__restrict__ cost, CostSmemLayout L, CostBlobLayout BL, f32x2 nz)
{
    pdl_sync();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + L.off_bar);
    float *raw = reinterpret_cast<float *>(smem_raw + L.off_raw);
    unsigned char *blob = smem_raw + L.off_blob;
    float *nlc = reinterpret_cast<float *>(smem_raw + L.off_nlc);
    float *attr_s = reinterpret_cast<float *>(smem_raw + L.off_attr);        // 2^A table (SMALL_A) or F1-F0 per attribute
    float *s0 = reinterpret_cast<float *>(smem_raw + L.off_s0);
    const uint32_t *hdr = reinterpret_cast<const uint32_t *>(blob);
    const float *rows = reinterpret_cast<const float *>(blob + BL.off_rows);
    const int16_t *cls_of = reinterpret_cast<const int16_t *>(blob + BL.off_clsof);
    const int16_t *slot_of = reinterpret_cast<const int16_t *>(blob + BL.off_slotof);
    const uint32_t *cbits = reinterpret_cast<const uint32_t *>(blob + BL.off_cbits);
    const uint32_t *abits = reinterpret_cast<const uint32_t *>(blob + BL.off_abits);

    const int b = blockIdx.y, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int Cs = L.Cs, As = L.As, CW = BL.CW, AW = BL.AW;
    const int q0 = blockIdx.x * CM_QT, nq = min(CM_QT, Q - q0);

    // ---- phase 0: everything this CTA reads more than once moves HBM/L2 -> shared memory with two TMA bulk copies
    // (prepared targets, the tile's class probabilities); per-thread operands are loaded into registers meanwhile ----
    const float *cp = cat_pred + ((size_t)b * Q + q0) * C;
    const uint32_t tile_bytes = (uint32_t)(nq * C) * 4u;
    const bool use_tma = ((reinterpret_cast<uintptr_t>(cp) | tile_bytes) & 15) == 0;
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar, BL.bytes + (use_tma ? tile_bytes : 0u));
        bulk_load_1d(blob, blobs + (size_t)b * BL.bytes, BL.bytes, bar);
        if (use_tma) bulk_load_1d(raw, cp, tile_bytes, bar);
    }
    const int qa = lane, qb = lane + 32;
    const bool va = qa < nq, vb = qb < nq;
    const int qbl = vb ? qb : (va ? qa : 0);                             // column b clamped for loads, never stored
    const float4 pa4 = reinterpret_cast<const float4 *>(box_pred)[(size_t)b * Q + q0 + (va ? qa : 0)];
    const float4 pb4 = reinterpret_cast<const float4 *>(box_pred)[(size_t)b * Q + q0 + qbl];
    float pattr[CM_SMALL_A];
    if (HAS_ATTR && SMALL_A && tid < nq) {
#pragma unroll
        for (int a = 0; a < CM_SMALL_A; ++a) pattr[a] = a < A ? attr_pred[((size_t)b * Q + q0 + tid) * A + a] : 0.5f;
    }
    if (!use_tma) {                                                      // unaligned tile: plain coalesced copy
#pragma unroll 8
        for (int e = tid; e < nq * C; e += CM_THREADS) raw[e] = cp[e];
    }
    __syncthreads();                                                     // barrier initialised (and the plain copy complete)
    mbar_wait(bar, 0);
    const int ns = (int)hdr[0];
    const uint32_t magicS = hdr[1];

    // ---- phase 4: prediction side of this column tile: -log(clip(p) + 1e-7) / C for the classes that occur only ----
    const float fC = (float)C, fA = (float)A;
    const float rC = rcp_refined(fC), rA = rcp_refined(fA);
#pragma unroll 4
    for (int e = tid; e < nq * ns; e += CM_THREADS) {
        const int q = __umulhi((uint32_t)e, magicS), s = e - q * ns;
        nlc[q * Cs + s] = div_by_const(neg_log_clip(raw[q * C + cls_of[s]]), fC, rC);
    }
    if (HAS_ATTR) {
        if (SMALL_A) {
            // table[q][m] = w_attr * ((s0 + sum_{a in m} (F1 - F0)(p_a)) / A), summed in attribute order; one thread per column
            if (tid < nq) {
                float df[CM_SMALL_A], sa0 = 0.0f;
#pragma unroll
                for (int a = 0; a < CM_SMALL_A; ++a) if (a < A) {
                    const float pc = safe_clip(pattr[a]);
                    const float f0 = focal0(pc);
                    df[a] = __fsub_rn(focal1(pc), f0);
                    sa0 = __fadd_rn(sa0, f0);
                }
                for (int m = 0; m < (1 << A); ++m) {
                    float sa = sa0;
#pragma unroll
                    for (int a = 0; a < CM_SMALL_A; ++a) if (a < A) sa = __fadd_rn(sa, (m >> a) & 1 ? df[a] : 0.0f);
                    attr_s[tid * CM_TAB + m] = __fmul_rn(w_attr, div_by_const(sa, fA, rA));
                }
            }
        } else {
            // one warp per column: F1 - F0 per attribute and the row sum of F0 (lane-strided partial sums, shuffle tree)
            const float *ap = attr_pred + ((size_t)b * Q + q0) * A;
            for (int q = warp; q < nq; q += CM_WARPS) {
                float sacc = 0.0f;
#pragma unroll 4
                for (int a = lane; a < A; a += 32) {
                    const float pc = safe_clip(ap[q * A + a]);
                    const float f0 = focal0(pc);
                    attr_s[q * As + a] = __fsub_rn(focal1(pc), f0);
                    sacc += f0;
                }
                sacc = warp_sum(sacc);
                if (lane == 0) s0[q] = sacc;
            }
        }
    }
    __syncthreads();

    // ---- phase 5: the pairs.  lane -> columns (lane, lane + 32), warp -> rows ----
    if (!va) return;
    PredBox2 p;
    uint32_t nan_a, nan_b;                                               // 0 or the quiet-NaN bits, OR-ed into the result
    {
        float ia[9], ib[9];
        bool na, nb_;
        pred_invariants(pa4.x, pa4.y, pa4.z, pa4.w, ia, na);
        pred_invariants(pb4.x, pb4.y, pb4.z, pb4.w, ib, nb_);
        p.y0a = ia[0]; p.x0a = ia[1]; p.y1a = ia[2]; p.x1a = ia[3];
        p.y0b = ib[0]; p.x0b = ib[1]; p.y1b = ib[2]; p.x1b = ib[3];
        p.area = pk2(ia[4], ib[4]); p.ty0 = pk2(ia[5], ib[5]); p.tx0 = pk2(ia[6], ib[6]); p.ty1 = pk2(ia[7], ib[7]); p.tx1 = pk2(ia[8], ib[8]);
        nan_a = na ? 0x7fc00000u : 0u; nan_b = nb_ ? 0x7fc00000u : 0u;
    }
    // word offsets of this thread's columns inside the staged tables
    int nl_a = qa * Cs, nl_b = qbl * Cs;
    int at_a = qa * (SMALL_A ? CM_TAB : As), at_b = qbl * (SMALL_A ? CM_TAB : As);
    f32x2 S0 = (HAS_ATTR && !SMALL_A) ? pk2(s0[qa], s0[qbl]) : 0ull;
    f32x2 WCAT = pk2(w_cat, w_cat), WBOX = pk2(w_box, w_box);
    const f32x2 WATTR = pk2(w_attr, w_attr);
    // Everything above is loop-invariant and cheap to recompute, so the compiler would rematerialise it (S2R, x != x,
    // 10 * x, q * Cs ...) inside the row loop: ~40 extra instructions per iteration.  Pin the values in registers.
#define CM_KEEP_F(x) asm volatile("" : "+f"(x))
#define CM_KEEP_R(x) asm volatile("" : "+r"(x))
#define CM_KEEP_L(x) asm volatile("" : "+l"(x))
    CM_KEEP_F(p.y0a); CM_KEEP_F(p.x0a); CM_KEEP_F(p.y1a); CM_KEEP_F(p.x1a);
    CM_KEEP_F(p.y0b); CM_KEEP_F(p.x0b); CM_KEEP_F(p.y1b); CM_KEEP_F(p.x1b);
    CM_KEEP_L(p.area); CM_KEEP_L(p.ty0); CM_KEEP_L(p.tx0); CM_KEEP_L(p.ty1); CM_KEEP_L(p.tx1);
    CM_KEEP_R(nan_a); CM_KEEP_R(nan_b); CM_KEEP_R(nl_a); CM_KEEP_R(nl_b); CM_KEEP_R(at_a); CM_KEEP_R(at_b);
    CM_KEEP_L(S0); CM_KEEP_L(WCAT); CM_KEEP_L(WBOX);

    // cost of both columns against row t (bits of the two results, NaN propagation included)
    auto pair_cost = [&](int t, const float4 *r4, uint32_t &o0, uint32_t &o1) {
        const float4 tb = r4[0];
        const ulonglong2 ra = reinterpret_cast<const ulonglong2 *>(r4)[1], rb = reinterpret_cast<const ulonglong2 *>(r4)[2];
        const float4 rc = r4[3];
        const uint32_t meta = __float_as_uint(rc.z), abw = __float_as_uint(rc.w);
        const int slot1 = (int)(meta & CM_META_SLOT);
        f32x2 CAT;
        if (slot1 > 0) {                                                 // one-hot row: one gather per column
            CAT = pk2(nlc[nl_a + slot1 - 1], nlc[nl_b + slot1 - 1]);
        } else {                                                         // general row: sum over its classes, ascending
            CAT = pk2(0.0f, 0.0f);
            for (int w = 0; w < CW; ++w) {
                uint32_t bits = cbits[t * CW + w];
                while (bits) { const int s = slot_of[(w << 5) + __ffs(bits) - 1]; bits &= bits - 1; CAT = add2(CAT, pk2(nlc[nl_a + s], nlc[nl_b + s])); }
            }
        }
        const f32x2 BOX = box_pair_cost_x2(tb, ra.x, ra.y, rb.x, rb.y, pk2(rc.x, rc.y), p, nz);
        f32x2 V = add2(MUL2(WCAT, CAT), MUL2(WBOX, BOX));
        if (HAS_ATTR) {
            if (SMALL_A) {
                V = add2(V, pk2(attr_s[at_a + abw], attr_s[at_b + abw]));
            } else {
                f32x2 SA = S0;
                for (int w = 0; w < AW; ++w) {
                    uint32_t bits = w == 0 ? abw : abits[t * AW + w];
                    while (bits) { const int a = (w << 5) + __ffs(bits) - 1; bits &= bits - 1; SA = add2(SA, pk2(attr_s[at_a + a], attr_s[at_b + a])); }
                }
                V = add2(V, MUL2(WATTR, div2_by_const(SA, fA, rA, nz)));
            }
        }
        float v0, v1;
        upk2(V, v0, v1);
        const uint32_t tnan = (uint32_t)((int32_t)meta >> 31) & 0x7fc00000u;   // tf.maximum / minimum propagate NaN
        o0 = __float_as_uint(v0) | tnan | nan_a;
        o1 = __float_as_uint(v1) | tnan | nan_b;
    };

    const float4 *rows4 = reinterpret_cast<const float4 *>(rows);
    const float4 *r4 = rows4 + warp * (CM_ROW / 4);
    uint32_t *out = reinterpret_cast<uint32_t *>(cost) + (size_t)b * T * Q + (size_t)warp * Q + q0 + qa;
    const size_t out_step = (size_t)CM_WARPS * Q;
    bool have_tail = false;
    uint32_t tail0 = 0u, tail1 = 0u;
#pragma unroll 1
    for (int t = warp; t < T; t += CM_WARPS, r4 += CM_WARPS * (CM_ROW / 4), out += out_step) {
        uint32_t o0, o1;
        if (__float_as_uint(r4[3].z) & CM_META_TAIL) {                   // warp-uniform: rows equal to the last row share its cost
            if (!have_tail) { pair_cost(T - 1, rows4 + (T - 1) * (CM_ROW / 4), tail0, tail1); have_tail = true; }
            o0 = tail0; o1 = tail1;
        } else {
            pair_cost(t, r4, o0, o1);
        }
        out[0] = o0;
        if (vb) out[32] = o1;
    }
#undef CM_KEEP_F
#undef CM_KEEP_R
#undef CM_KEEP_L
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

// ---------------------------------------------------------------------------------------------
// K8, register-resident form (n <= Q <= 32 * KMAX): the same shortest-augmenting-path solver, same fp64 expressions,
// same tie rule, but the per-column state lives in REGISTERS -- lane l owns columns l, l + 32, ... and keeps their dual
// v, shortest-path cost, owner row and position in scipy's swap-removed `remaining` list -- so a column scan has no
// dependent shared-memory round trips, and the warp arg-min is three hardware warp reductions (REDUX on the high word
// of an order-preserving 64-bit key of the cost, on its low word, on the tie rank) instead of fifteen shuffles.
// The solver is a serial chain of scans (early-training predictions are near-tied in a way all targets agree on, which
// drives it towards its n^2 / 2 worst case), so scan latency is the whole cost: ~1800 -> ~300 cycles per scan.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t f64_order_key(double x)
{
    const uint64_t b = (uint64_t)__double_as_longlong(x + 0.0);          // + 0.0: -0 and +0 compare equal, give them one key
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

template <int KMAX>
__global__ void __launch_bounds__(32)
lsap_reg_kernel(int T, int Q, const float *__restrict__ cost, const int32_t *__restrict__ num_objects,
                int32_t *__restrict__ col4row_out, int32_t *__restrict__ row4col_out, int32_t *__restrict__ status,
                int stage_cost, int validate)
{
    pdl_sync();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, lane = threadIdx.x;
    double *u = reinterpret_cast<double *>(smem_raw);                       // [T]
    int *path = reinterpret_cast<int *>(u + T);                             // [Q]
    int *col4row = path + Q;                                                // [T]
    float *cost_s = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(col4row + T) + 15) & ~uintptr_t(15));

    int n = num_objects[b]; n = n < 0 ? 0 : (n > T ? T : n);
    int32_t *c4r_o = col4row_out + (size_t)b * T;
    int32_t *r4c_o = row4col_out + (size_t)b * Q;
    for (int t = lane; t < T; t += 32) c4r_o[t] = -1;
    for (int q = lane; q < Q; q += 32) r4c_o[q] = -1;
    const float *cb = cost + (size_t)b * T * Q;
    if (validate) {
        // scipy's input check on the live rows, fused with the staging copy (staged problems only): NaN / -inf -> error
        bool bad = false;
        const int total = n * Q;
        if ((reinterpret_cast<uintptr_t>(cb) & 15) == 0 && (total & 3) == 0) {
            for (int e = lane; e < (total >> 2); e += 32) {
                const float4 c4 = reinterpret_cast<const float4 *>(cb)[e];
                reinterpret_cast<float4 *>(cost_s)[e] = c4;
                bad = bad || c4.x != c4.x || c4.y != c4.y || c4.z != c4.z || c4.w != c4.w ||
                      c4.x == -CUDART_INF_F || c4.y == -CUDART_INF_F || c4.z == -CUDART_INF_F || c4.w == -CUDART_INF_F;
            }
        } else {
            for (int e = lane; e < total; e += 32) { const float c = cb[e]; cost_s[e] = c; bad = bad || c != c || c == -CUDART_INF_F; }
        }
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) status[b] = bad ? BDETR_E_INVALID_COST : 0;
        if (bad || n == 0) return;
        cb = cost_s;
    } else {
        if (n == 0 || status[b] != 0) return;
        if (stage_cost) {
            const int total = n * Q;
            for (int e = lane; e < total; e += 32) cost_s[e] = cb[e];
            cb = cost_s;
        }
    }
    for (int i = lane; i < n; i += 32) { u[i] = 0.0; col4row[i] = -1; }

    double v[KMAX], spc[KMAX];
    int r4c[KMAX], pos[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) { v[k] = 0.0; r4c[k] = -1; }
    __syncwarp();

    bool infeasible = false;
    for (int cur = 0; cur < n; ++cur) {
        double minVal = 0.0;
        int i = cur, nrem = Q, sink = -1;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) { const int j = lane + 32 * k; pos[k] = j < Q ? Q - 1 - j : -1; spc[k] = CUDART_INF; }
        while (sink == -1) {
            const double ui = u[i];
            const float *crow = cb + (size_t)i * Q;
            float cc[KMAX];
#pragma unroll
            for (int k = 0; k < KMAX; ++k) cc[k] = pos[k] >= 0 ? crow[lane + 32 * k] : 0.0f;
            double best_s = CUDART_INF;
            int best_rank = -1, best_k = 0;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (pos[k] >= 0) {
                    const double r = ((minVal + (double)cc[k]) - ui) - v[k];
                    if (r < spc[k]) { path[lane + 32 * k] = i; spc[k] = r; }
                    const double sk = spc[k];
                    const int rank = (r4c[k] == -1) ? (RANK_FREE + pos[k]) : (RANK_USED - pos[k]);
                    if (sk < best_s || (sk == best_s && rank > best_rank)) { best_s = sk; best_rank = rank; best_k = k; }
                }
            }
            // warp arg-min of (cost ascending, rank descending): ranks are distinct, so exactly one lane wins
            const uint64_t key = f64_order_key(best_s);
            const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
            const uint32_t mhi = __reduce_min_sync(0xffffffffu, hi);
            const bool c1 = hi == mhi;
            const uint32_t mlo = __reduce_min_sync(0xffffffffu, c1 ? lo : 0xffffffffu);
            const bool c2 = c1 && lo == mlo;
            const int mrank = __reduce_max_sync(0xffffffffu, c2 ? best_rank : -1);
            if (mrank < 0) { infeasible = true; break; }                 // every remaining column is +inf (rank -1 = no column)
            const uint32_t win = __ballot_sync(0xffffffffu, c2 && best_rank == mrank);
            const int wl = __ffs(win) - 1;
            // the winner's column: its cost, owner row, position and index k
            int w_owner = 0, w_pos = 0;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) if (k == best_k) { w_owner = r4c[k]; w_pos = pos[k]; }
            minVal = __shfl_sync(0xffffffffu, best_s, wl);
            if (minVal == CUDART_INF) { infeasible = true; break; }
            const int packed = __shfl_sync(0xffffffffu, (best_k << 16) | w_pos, wl);
            const int owner = __shfl_sync(0xffffffffu, w_owner, wl);
            const int index = packed & 0xffff, wk = packed >> 16;
            const int j = wl + 32 * wk;
            // swap-remove: the column at the end of the list takes the freed position, the winner leaves the list
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (pos[k] == nrem - 1) pos[k] = index;
                if (lane == wl && k == wk) pos[k] = -2;                  // -2 = scanned and removed in this search (the set SC)
            }
            --nrem;
            if (owner == -1) sink = j; else i = owner;
        }
        if (infeasible) break;
        // dual updates (same fp64 expressions as the sequential solver).  Every visited row other than `cur` is the owner
        // of exactly one removed column, so the row update is done from the column side.
        if (lane == 0) u[cur] = u[cur] + minVal;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (pos[k] == -2) {
                const int r = r4c[k];
                if (r >= 0) u[r] = u[r] + (minVal - spc[k]);
                v[k] = v[k] - (minVal - spc[k]);
            }
        }
        __syncwarp();
        // augment along the stored path (uniform walk; the owner registers are patched by their lanes)
        {
            int j = sink;
            for (;;) {
                const int r = path[j];
#pragma unroll
                for (int k = 0; k < KMAX; ++k) if (lane + 32 * k == j) r4c[k] = r;
                const int tmp = col4row[r];
                __syncwarp();
                if (lane == 0) col4row[r] = j;
                __syncwarp();
                j = tmp;
                if (r == cur) break;
            }
        }
        __syncwarp();
    }
    if (infeasible) { if (lane == 0) status[b] = BDETR_E_INFEASIBLE; return; }
    for (int i = lane; i < n; i += 32) { const int j = col4row[i]; c4r_o[i] = j; r4c_o[j] = i; }
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
constexpr int ML_THREADS = 256;

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

// 1 + sum_b num_objects (losses_and_metrics.py:144): integer sum over the batch by the lanes of the calling warp (every
// warp computes it for itself: B / 32 loads per lane instead of a B-long serial loop per thread); exact in fp32
__device__ __forceinline__ float total_objects(const int32_t *num_objects, int B)
{
    int s = 0;
    for (int k = threadIdx.x & 31; k < B; k += 32) s += num_objects[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return 1.0f + (float)s;
}

// forward: one CTA per image.  Phase 1: one THREAD per matched target row for the scalar work (class term, box term, IoU
// metric); the row's one-hot / multi-hot class entries are fetched sixteen at a time as independent loads, not as a
// serial load -> test -> branch chain (25 us for 16 images), and without a warp's 32 lanes repeating one row's box math
// (the warp-per-row form executed 5x the instructions).  Phase 2, large attribute vocabularies only: one warp per row,
// lanes striding the attributes (two logarithms each).
__global__ void __launch_bounds__(1024)
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
    __shared__ float red[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float total_n = total_objects(num_objects, B);
    float cat_s = 0.0f, attr_s = 0.0f, box_s = 0.0f, ex_s = 0.0f;
    const bool attr_inline = w_attr != 0.0f && A <= 8, attr_warp = w_attr != 0.0f && A > 8;

    if (T <= 32) {
        // few targets per image (the training configs: T = 20): one WARP per row, launched with T warps -- every row of the
        // image in flight at once, lanes striding the classes / attributes (14 us vs 25 us for 16 images)
        const int t = warp;
        const int q = t < T ? col4row[(size_t)b * T + t] : -1;              // warp-uniform
        if (q >= 0) {
            const float *ct = cat_true + ((size_t)b * T + t) * C;
            const float *cp = cat_pred + ((size_t)b * Q + q) * C;
            float cat = 0.0f;
            for (int c = lane; c < C; c += 32) { const float y = ct[c]; if (y != 0.0f) cat += y * neg_log_clip(cp[c]); }
            cat = warp_sum(cat);
            float sa = 0.0f;
            if (w_attr != 0.0f) {
                const float *at = attr_true + ((size_t)b * T + t) * A;
                const float *ap = attr_pred + ((size_t)b * Q + q) * A;
                for (int a = lane; a < A; a += 32) { const float pc = safe_clip(ap[a]); sa += (at[a] != 0.0f) ? focal1(pc) : focal0(pc); }
                sa = warp_sum(sa);
            }
            if (lane == 0) {
                cat_s += w_cat * (cat / (float)C);
                if (w_attr != 0.0f) attr_s += w_attr * (sa / (float)A);
                const float4 tb4 = reinterpret_cast<const float4 *>(box_true)[(size_t)b * T + t];
                const float4 pb4 = reinterpret_cast<const float4 *>(box_pred)[(size_t)b * Q + q];
                float iou_v;
                const float bc = box_pair_cost(coco_to_tf(tb4.x, tb4.y, tb4.z, tb4.w), coco_to_tf(pb4.x, pb4.y, pb4.z, pb4.w), &iou_v);
                box_s += w_box * bc;
                atomicAdd(&iou[q], (1.0f - (1.0f - iou_v)) / total_n);
            }
        }
    } else {
    for (int t = tid; t < T; t += blockDim.x) {
        const int q = col4row[(size_t)b * T + t];
        if (q < 0) continue;
        const float *ct = cat_true + ((size_t)b * T + t) * C;
        const float *cp = cat_pred + ((size_t)b * Q + q) * C;
        const float4 tb4 = reinterpret_cast<const float4 *>(box_true)[(size_t)b * T + t];
        const float4 pb4 = reinterpret_cast<const float4 *>(box_pred)[(size_t)b * Q + q];
        float cat = 0.0f;
        for (int c0 = 0; c0 < C; c0 += 16) {
            float y[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) y[k] = c0 + k < C ? ct[c0 + k] : 0.0f;
#pragma unroll
            for (int k = 0; k < 16; ++k) if (y[k] != 0.0f) cat += y[k] * neg_log_clip(cp[c0 + k]);
        }
        cat_s += w_cat * (cat / (float)C);
        if (attr_inline) {
            const float *at = attr_true + ((size_t)b * T + t) * A;
            const float *ap = attr_pred + ((size_t)b * Q + q) * A;
            float sa = 0.0f;
            for (int a = 0; a < A; ++a) { const float pc = safe_clip(ap[a]); sa += (at[a] != 0.0f) ? focal1(pc) : focal0(pc); }
            attr_s += w_attr * (sa / (float)A);
        }
        float iou_v;
        const float bc = box_pair_cost(coco_to_tf(tb4.x, tb4.y, tb4.z, tb4.w), coco_to_tf(pb4.x, pb4.y, pb4.z, pb4.w), &iou_v);
        box_s += w_box * bc;
        // IOU metric = 1 - (1 - iou), summed over (b,t) per prediction column (quirk Q7)
        atomicAdd(&iou[q], (1.0f - (1.0f - iou_v)) / total_n);
    }
    if (attr_warp) {
        for (int t = warp; t < T; t += (int)(blockDim.x >> 5)) {
            const int q = col4row[(size_t)b * T + t];                    // warp-uniform
            if (q < 0) continue;
            const float *at = attr_true + ((size_t)b * T + t) * A;
            const float *ap = attr_pred + ((size_t)b * Q + q) * A;
            float sa = 0.0f;
            for (int a = lane; a < A; a += 32) { const float pc = safe_clip(ap[a]); sa += (at[a] != 0.0f) ? focal1(pc) : focal0(pc); }
            sa = warp_sum(sa);
            if (lane == 0) attr_s += w_attr * (sa / (float)A);
        }
    }
    }
    for (int q = tid; q < Q; q += blockDim.x) {
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

// backward: one WARP per (b, q): accumulates the whole gradient rows of that prediction, lanes striding the classes / attributes
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
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * (ML_THREADS / 32) + (threadIdx.x >> 5);
    if (e >= B * Q) return;
    const int b = e / Q;
    const float total_n = total_objects(num_objects, B);
    const int t = row4col[e];
    const float *cp = cat_pred + (size_t)e * C;
    float *dc = d_cat + (size_t)e * C;
    // existence term on class 0
    if (lane == 0) {
        const float p0 = cp[0];
        if (in_clip(p0)) {
            const float y = t >= 0 ? 0.0f : 1.0f;
            const float dbce = -(y / (p0 + 1e-7f) - (1.0f - y) / (1.0f - p0 + 1e-7f));
            dc[0] += gscale * w_exist * dbce / ((float)Q * (1.0f + (float)Q));
        }
    }
    if (t < 0) return;
    __syncwarp();                                                        // lane 0's update of dc[0] precedes the class term on dc[0]
    const float gs = gscale / total_n;
    const float *ct = cat_true + ((size_t)b * T + t) * C;
    for (int c = lane; c < C; c += 32) {
        const float y = ct[c];
        if (y != 0.0f && in_clip(cp[c])) dc[c] += gs * w_cat * y * (-1.0f / (cp[c] + 1e-7f)) / (float)C;
    }
    if (w_attr != 0.0f) {
        const float *at = attr_true + ((size_t)b * T + t) * A;
        const float *ap = attr_pred + (size_t)e * A;
        float *da = d_attr + (size_t)e * A;
        for (int a = lane; a < A; a += 32) {
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
    if (w_box != 0.0f && lane == 0) {
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

extern "C" __attribute__((visibility("default"))) size_t bdetr_cost_targets_bytes(int B, int T, int C, int A)
{
    if (B <= 0 || T <= 0 || C <= 0 || A <= 0) return 0;
    return (size_t)B * cost_blob_layout(T, C, A).bytes;
}

static size_t cost_targets_smem(int T, int C, int A, const CostBlobLayout &BL)
{
    return 16 + 128 + BL.bytes + sizeof(float) * ((((size_t)T * C + 3) & ~size_t(3)) + (((size_t)T * A + 3) & ~size_t(3)));
}

extern "C" __attribute__((visibility("default"))) int bdetr_cost_targets_prepare(int B, int T, int C, int A,
                                     const float *cat_true, const float *attr_true, const float *box_true,
                                     void *prepared, void *stream)
{
    BDETR_REQUIRE(B > 0 && T > 0 && C > 0 && A > 0, BDETR_E_BAD_SHAPE, "B,T,C,A must be positive");
    BDETR_REQUIRE(cat_true && attr_true && box_true && prepared, BDETR_E_NULL, "null pointer");
    BDETR_REQUIRE((reinterpret_cast<uintptr_t>(prepared) & 15) == 0, BDETR_E_BAD_SHAPE, "prepared-targets buffer must be 16-byte aligned");
    BDETR_REQUIRE(C <= 1024 && A <= 1024 && (long long)T * C < (1 << 22) && (long long)T * A < (1 << 22), BDETR_E_UNSUPPORTED, "C / A / T*C / T*A too large");
    const CostBlobLayout BL = cost_blob_layout(T, C, A);
    const size_t smem = cost_targets_smem(T, C, A, BL);
    BDETR_REQUIRE(smem <= 227 * 1024, BDETR_E_UNSUPPORTED, "T*(C+A) too large for the shared-memory staging of the targets");
    static size_t optin = 0;        // opt in to large dynamic shared memory once per size class (never during a later graph capture)
    if (smem > optin) {
        BDETR_CUDA(cudaFuncSetAttribute(cost_targets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        optin = smem;
    }
    launch_k(cost_targets_kernel, B, CM_THREADS, smem, as_stream(stream), T, C, A, cat_true, attr_true, box_true,
             reinterpret_cast<unsigned char *>(prepared), BL);
    BDETR_CHECK_LAUNCH("cost_targets_kernel");
    return BDETR_OK;
}

extern "C" __attribute__((visibility("default"))) int bdetr_cost_matrix_prepared(int B, int T, int Q, int C, int A, const void *prepared,
                                     const float *cat_pred, const float *attr_pred, const float *box_pred,
                                     float w_cat, float w_box, float w_attr, float *cost, void *stream)
{
    BDETR_REQUIRE(B > 0 && T > 0 && Q > 0 && C > 0 && A > 0, BDETR_E_BAD_SHAPE, "B,T,Q,C,A must be positive");
    BDETR_REQUIRE(prepared && cat_pred && attr_pred && box_pred && cost, BDETR_E_NULL, "null pointer");
    BDETR_REQUIRE((reinterpret_cast<uintptr_t>(prepared) & 15) == 0, BDETR_E_BAD_SHAPE, "prepared-targets buffer must be 16-byte aligned");
    BDETR_REQUIRE(C <= 1024 && CM_QT * (long long)C < (1 << 20), BDETR_E_UNSUPPORTED, "C too large");
    const bool has_attr = (w_attr != 0.0f);
    const bool small_a = A <= CM_SMALL_A;
    const CostSmemLayout L = cost_smem_layout(T, C, A, has_attr);
    const CostBlobLayout BL = cost_blob_layout(T, C, A);
    BDETR_REQUIRE(L.bytes <= 227 * 1024, BDETR_E_UNSUPPORTED, "C/A/T too large for the shared-memory tile");
    dim3 grid(ceil_div(Q, CM_QT), B);
    const int variant = has_attr ? (small_a ? 1 : 2) : 0;
    auto kern = variant == 0 ? cost_matrix_kernel<false, true> : variant == 1 ? cost_matrix_kernel<true, true> : cost_matrix_kernel<true, false>;
    static size_t optin[3] = {0, 0, 0};
    if (L.bytes > optin[variant]) {
        BDETR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.bytes));
        optin[variant] = L.bytes;
    }
    const f32x2 neg_zero_pair = 0x8000000080000000ull;
    launch_k(kern, grid, CM_THREADS, L.bytes, as_stream(stream),
             T, Q, C, A, reinterpret_cast<const unsigned char *>(prepared), cat_pred, attr_pred, box_pred, w_cat, w_box, w_attr, cost, L, BL, neg_zero_pair);
    BDETR_CHECK_LAUNCH("cost_matrix_kernel");
    return BDETR_OK;
}

// MatchingMetric.call (losses_and_metrics.py:176-192): pairwise IoU [B,T,Q] of COCO boxes (targets = rows), times the
// assignment mask when one is given.  IOU_Metric (:17-18) = 1 - tfa.giou_loss(mode='iou') = 1 - (1 - iou).
__global__ void __launch_bounds__(256)
pairwise_iou_kernel(int T, int Q, const float *__restrict__ box_true, const float *__restrict__ box_pred, const float *__restrict__ mask,
                    float *__restrict__ out)
{
    pdl_sync();
    const int b = blockIdx.y;
    const size_t per = (size_t)T * Q;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < per; e += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(e / Q), q = (int)(e - (size_t)t * Q);
        const float4 tb4 = reinterpret_cast<const float4 *>(box_true)[(size_t)b * T + t];
        const float4 pb4 = reinterpret_cast<const float4 *>(box_pred)[(size_t)b * Q + q];
        float iou_v;
        (void)box_pair_cost(coco_to_tf(tb4.x, tb4.y, tb4.z, tb4.w), coco_to_tf(pb4.x, pb4.y, pb4.z, pb4.w), &iou_v);
        const float v = 1.0f - (1.0f - iou_v);
        out[(size_t)b * per + e] = mask ? mask[(size_t)b * per + e] * v : v;
    }
}

extern "C" __attribute__((visibility("default"))) int bdetr_pairwise_iou(int B, int T, int Q, const float *box_true, const float *box_pred,
                                    const float *mask, float *out, void *stream)
{
    BDETR_REQUIRE(B > 0 && T > 0 && Q > 0, BDETR_E_BAD_SHAPE, "B,T,Q must be positive");
    BDETR_REQUIRE(box_true && box_pred && out, BDETR_E_NULL, "null pointer");
    const int gx = (int)(((size_t)T * Q + 255) / 256);
    launch_k(pairwise_iou_kernel, dim3(gx < 64 ? gx : 64, B), 256, 0, as_stream(stream), T, Q, box_true, box_pred, mask, out);
    BDETR_CHECK_LAUNCH("pairwise_iou_kernel");
    return BDETR_OK;
}

// Convenience form of the two calls above with a library-owned, grow-only scratch buffer for the prepared targets.
// The buffer is allocated on first use / growth (do that outside CUDA-graph capture) and shared by all calls on the
// device: NOT for concurrent use from several streams -- callers that overlap matchers (the model does) hold their own
// buffer and use bdetr_cost_targets_prepare + bdetr_cost_matrix_prepared.
extern "C" __attribute__((visibility("default"))) int bdetr_cost_matrix_fwd(int B, int T, int Q, int C, int A,
                                     const float *cat_true, const float *attr_true, const float *box_true,
                                     const float *cat_pred, const float *attr_pred, const float *box_pred,
                                     float w_cat, float w_box, float w_attr, float *cost, void *stream)
{
    BDETR_REQUIRE(B > 0 && T > 0 && Q > 0 && C > 0 && A > 0, BDETR_E_BAD_SHAPE, "B,T,Q,C,A must be positive");
    static void *scratch[16] = {nullptr};
    static size_t scratch_bytes[16] = {0};
    int dev = 0;
    BDETR_CUDA(cudaGetDevice(&dev));
    BDETR_REQUIRE(dev >= 0 && dev < 16, BDETR_E_UNSUPPORTED, "device index out of range");
    const size_t need = bdetr_cost_targets_bytes(B, T, C, A);
    if (need > scratch_bytes[dev]) {
        BDETR_CUDA(cudaStreamSynchronize(as_stream(stream)));
        if (scratch[dev]) BDETR_CUDA(cudaFree(scratch[dev]));
        scratch[dev] = nullptr; scratch_bytes[dev] = 0;
        BDETR_CUDA(cudaMalloc(&scratch[dev], need));
        scratch_bytes[dev] = need;
    }
    const int rc = bdetr_cost_targets_prepare(B, T, C, A, cat_true, attr_true, box_true, scratch[dev], stream);
    if (rc != BDETR_OK) return rc;
    return bdetr_cost_matrix_prepared(B, T, Q, C, A, scratch[dev], cat_pred, attr_pred, box_pred, w_cat, w_box, w_attr, cost, stream);
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
    cudaStream_t s = as_stream(stream);
    // Register-resident solver whenever no image can need scipy's transposed path (T <= Q) and a lane's share of the
    // columns fits its registers; staged problems also validate their own rows (no memset / validate launches).
    const bool reg_form = T <= Q && Q <= 320;
    const bool fused_validate = reg_form && stage;
    if (!fused_validate) {
        BDETR_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t) * B, s));
        const size_t total = (size_t)B * T * Q;
        const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
        launch_k(lsap_validate_kernel, blocks, 256, 0, s, B, T, Q, cost, num_objects, status);
        BDETR_CHECK_LAUNCH("lsap_validate_kernel");
    }
    if (reg_form) {
        size_t rsmem = (size_t)T * (sizeof(double) + sizeof(int)) + (size_t)Q * sizeof(int) + 16;
        if (stage) rsmem += (size_t)T * Q * sizeof(float);
        auto kern = Q <= 128 ? lsap_reg_kernel<4> : lsap_reg_kernel<10>;
        static size_t reg_optin[2] = {0, 0};
        const int which = Q <= 128 ? 0 : 1;
        if (rsmem > 48 * 1024 && rsmem > reg_optin[which]) {
            BDETR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
            reg_optin[which] = rsmem;
        }
        launch_k(kern, B, 32, rsmem, s, T, Q, cost, num_objects, col4row, row4col, status, stage, fused_validate ? 1 : 0);
        BDETR_CHECK_LAUNCH("lsap_reg_kernel");
    } else {
        if (stage) smem += (size_t)T * Q * sizeof(float) + 16;
        static size_t lsap_optin = 0;
        if (smem > lsap_optin) {
            BDETR_CUDA(cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            lsap_optin = smem;
        }
        launch_k(lsap_kernel, B, 32, smem, s, T, Q, cost, num_objects, col4row, row4col, status, stage);
        BDETR_CHECK_LAUNCH("lsap_kernel");
    }
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
    launch_k(matched_loss_fwd_kernel, B, T <= 32 ? 32 * T : ML_THREADS, 0, s, B, T, Q, C, A, cat_true, attr_true, box_true, num_objects,
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
    launch_k(matched_loss_bwd_kernel, ceil_div(B * Q, ML_THREADS / 32), ML_THREADS, 0, as_stream(stream), 
        B, T, Q, C, A, cat_true, attr_true, box_true, num_objects, cat_pred, attr_pred, box_pred, row4col,
        w_cat, w_box, w_attr, w_exist, gscale, d_cat_pred, d_attr_pred, d_box_pred);
    BDETR_CHECK_LAUNCH("matched_loss_bwd_kernel");
    return BDETR_OK;
}
