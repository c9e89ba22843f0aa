// Tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory ->
// tcgen05.mma kind::tf32 with the fp32 accumulator in TMEM -> tcgen05.ld epilogue.
//
// Operands are the fp32 activations / weights exactly as they sit in HBM (no conversion pass):
// TF32 keeps a 10-bit mantissa (>= the reference's mixed_float16 compute type), accumulation is fp32.
// Both operand majors are supported through the UMMA descriptors, so the three GEMMs of a Dense
// layer need no transposes:
//     forward   Y[M,N]  = X[M,K]   W[K,N]      A K-major,  B MN-major
//     dgrad     dX[M,K] = dY[M,N]  W[K,N]^T    A K-major,  B K-major
//     wgrad     dW[K,N] += X[M,K]^T dY[M,N]    A MN-major, B MN-major   (split-K, fp32 red.add)
// One CTA computes one 128 x BN output tile (x one K split): warp 0 = TMA producer, warp 1 = TMEM
// allocator + single-thread MMA issuer, warps 2-5 = epilogue (one TMEM lane = one output row each).
#include <cstdlib>
#include <cstring>
#include "umma.cuh"

namespace bdetr {

constexpr int UM_BM = 128;
constexpr int UM_BK = 32;              // fp32 elements per stage along the contraction = one 128B swizzle row
constexpr int UM_THREADS = 192;
constexpr int UM_MAX_GROUPS = 3;
// Up to three GEMMs in one launch.  N-groups: independent outputs C[g] = op(A[g or 0]) op(B[g]) (the q / k / v
// projections of one input, the hidden layers of the three heads, their weight gradients); K-groups: ONE output
// C = sum_g op(A[g]) op(B[g]) (dX = dQ Wq^T + dK Wk^T + dV Wv^T), the k loop simply runs over all groups.
struct alignas(64) UmmaMaps { CUtensorMap a[UM_MAX_GROUPS], b[UM_MAX_GROUPS], c[UM_MAX_GROUPS]; };
struct UmmaEpilogue {
    int M, N;                                   // N = output columns per group
    const float *bias[UM_MAX_GROUPS];
    // row addend: out[r, n] += rowtab[g][(r % rowtab_period) * rowtab_ld + n]  (positional term folded into a projection:
    // (x + pos) W = x W + pos W, with pos W a batch-invariant [L, N] table; or a plain [M, N] addend with period = M)
    const float *rowtab[UM_MAX_GROUPS]; int rowtab_period, rowtab_ld;
    float *colsum[UM_MAX_GROUPS];               // colsum[g][n] += sum over rows of the stored output (bias gradient of the Dense before)
    int act; const float *relu_mask; int beta; int atomic_out; int round_out;
    int ldc;
    int num_kb, kb_per_split;                   // k-blocks per K-group / per split
    int n_groups, k_groups, share_a, tiles_n;   // tiles_n = column tiles per N-group
    long long *dbg;       // optional clock64 timeline of CTA (0,0,0) (bdetr_debug_set_timeline)
};

template <int BN, int UM_STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(UM_THREADS, 1)
gemm_umma_kernel(const __grid_constant__ UmmaMaps maps, const __grid_constant__ UmmaEpilogue ep)
{
    constexpr uint32_t A_STAGE = UM_BM * UM_BK * 4;      // 16 KB
    constexpr uint32_t B_STAGE = BN * UM_BK * 4;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // SWIZZLE_128B atoms need 1024B alignment
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem + UM_STAGES * A_STAGE;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_b + UM_STAGES * B_STAGE);
    uint64_t *empty = full + UM_STAGES;
    uint64_t *accum_full = empty + UM_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_full + 1);
    float *smem_bias = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~uintptr_t(15));   // BN floats
    float *smem_colsum = smem_bias + BN;                                                                                // BN floats

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool dbg_on = ep.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#define UMMA_STAMP(slot) do { if (dbg_on) ep.dbg[slot] = clock64(); } while (0)
    if (threadIdx.x == 0) UMMA_STAMP(0);
    const int grp = blockIdx.x / ep.tiles_n;                       // N-group of this CTA (0 when there is one group)
    const int m0 = blockIdx.y * UM_BM, n0 = (blockIdx.x - grp * ep.tiles_n) * BN;
    const int kb_beg = blockIdx.z * ep.kb_per_split;
    const int kb_end = min(ep.num_kb, kb_beg + ep.kb_per_split);
    const int nkb1 = kb_end - kb_beg;                              // k-blocks per K-group
    const int nkb = nkb1 * ep.k_groups;
    const CUtensorMap *map_c = &maps.c[grp];

    if (threadIdx.x == 0) {
        for (int s = 0; s < UM_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // PDL: dependents are released only after this CTA owns its TMEM columns (a dependent that grabbed TMEM first while
    // blocked in griddepcontrol.wait could starve a late CTA of this grid); global data is touched below the wait.
    pdl_sync();
    if (threadIdx.x == 0) UMMA_STAMP(1);

    if (warp == 0) {
        // ===== TMA producer (warp-convergent loop, one elected lane issues) =====
        for (int i = 0; i < nkb; ++i) {
            const int s = i % UM_STAGES;
            if (i >= UM_STAGES) mbar_wait(&empty[s], ((i / UM_STAGES) - 1) & 1);
            const int kg = i / nkb1;                                   // K-group of this k-block
            const int k0 = (kb_beg + (i - kg * nkb1)) * UM_BK;
            const CUtensorMap *map_a = &maps.a[ep.k_groups > 1 ? kg : (ep.share_a ? 0 : grp)];
            const CUtensorMap *map_b = &maps.b[ep.k_groups > 1 ? kg : grp];
            uint8_t *a_dst = smem_a + s * A_STAGE, *b_dst = smem_b + s * B_STAGE;
            if (elect_one()) {
                if (i == 1) UMMA_STAMP(2);
                mbar_expect_tx(&full[s], A_STAGE + B_STAGE);
                if (!A_MN) {
                    tma_load_2d(a_dst, map_a, k0, m0, &full[s]);                      // box {32 k, 128 rows}
                } else {
#pragma unroll
                    for (int j = 0; j < UM_BM / 32; ++j) tma_load_2d(a_dst + j * 4096, map_a, m0 + 32 * j, k0, &full[s]);   // box {32 m, 32 k}
                }
                if (!B_MN) {
                    tma_load_2d(b_dst, map_b, k0, n0, &full[s]);                      // box {32 k, BN rows}
                } else {
#pragma unroll
                    for (int j = 0; j < BN / 32; ++j) tma_load_2d(b_dst + j * 4096, map_b, n0 + 32 * j, k0, &full[s]);      // box {32 n, 32 k}
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===== MMA issuer (warp-convergent loop, one elected lane issues) =====
        constexpr uint32_t idesc = make_idesc_tf32(UM_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % UM_STAGES;
            mbar_wait(&full[s], (i / UM_STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_base = smem_u32(smem_a + s * A_STAGE), b_base = smem_u32(smem_b + s * B_STAGE);
            if (elect_one()) {
                if (i == 0) UMMA_STAMP(3);
#pragma unroll
                for (int j = 0; j < UM_BK / 8; ++j) {                                   // UMMA_K = 8 for tf32
                    // K-major: 8 rows x 128B atoms, 1024B apart (SBO); advance 32B per k-step inside the swizzled row.
                    // MN-major: [mn-block of 32][k row][128B]; mn-blocks 4096B apart (LBO); 4-k-row atoms 512B apart
                    // (SBO); one k-step = 8 k-rows = 1024B.
                    const uint64_t a_desc = A_MN ? make_smem_desc(a_base + j * 1024, 4096, 512, 1) : make_smem_desc(a_base + j * 32, 16, 1024, 2);
                    const uint64_t b_desc = B_MN ? make_smem_desc(b_base + j * 1024, 4096, 512, 1) : make_smem_desc(b_base + j * 32, 16, 1024, 2);
                    umma_tf32(tmem_base, a_desc, b_desc, idesc, (i | j) != 0);
                }
                umma_commit(&empty[s]);            // frees the smem stage when these MMAs retire
                if (i == nkb - 1) {
                    umma_commit(accum_full);       // accumulator complete
                    UMMA_STAMP(4);
                }
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: warp w owns TMEM lanes 32*(w%4)..  Each lane holds one output row; a 32x32 block goes to
        // shared memory in the 128B-swizzled box layout (conflict-free 16-byte stores) and leaves through one TMA
        // store -- or TMA reduce-add for accumulate / split-K -- which also clips the M and N tails.  The pipeline
        // stages are free once the accumulator is complete, so the staging boxes reuse them. =====
        const int q = warp & 3;
        // bias -> shared memory while the mainloop runs (a global load per chunk would sit on the critical path of the
        // epilogue: measured +1.1 us per GEMM)
        const float *bias_g = ep.bias[grp];
        const float *rowtab_g = ep.rowtab[grp];
        float *colsum_g = ep.colsum[grp];
        const bool add_bias = bias_g && (!ep.atomic_out || blockIdx.z == 0);
        if (add_bias || colsum_g) {
            const int e = threadIdx.x - 64;                    // 0..127
            if (e < BN) {
                smem_bias[e] = (add_bias && n0 + e < ep.N) ? bias_g[n0 + e] : 0.0f;
                smem_colsum[e] = 0.0f;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");      // the four epilogue warps only
        }
        if (nkb > 0) {
            mbar_wait(accum_full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (warp == 2 && lane == 0) UMMA_STAMP(5);
        const int row = m0 + q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t ra[32], rb[32];
        if (nkb > 0) tmem_ld32_issue(lane_addr, ra);
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
            const int c0 = c * 32;
            uint32_t *cur = (c & 1) ? rb : ra, *nxt = (c & 1) ? ra : rb;
            float v[32];
            if (nkb > 0) {
                tmem_ld32_wait(cur);
                if (c + 1 < BN / 32) tmem_ld32_issue(lane_addr + c0 + 32, nxt);     // next chunk in flight while this one is written out
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(cur[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.0f;
            }
            const int ncol = ep.N - (n0 + c0);                 // columns of this chunk inside the matrix
            if (ncol > 0) {
                if (add_bias) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 b4 = *reinterpret_cast<const float4 *>(smem_bias + c0 + i);     // broadcast read
                        v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                    }
                }
                if (rowtab_g && row < ep.M && (!ep.atomic_out || blockIdx.z == 0)) {
                    const float *trow = rowtab_g + (size_t)(row % ep.rowtab_period) * ep.rowtab_ld + n0 + c0;
                    if (ncol >= 32) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 t4 = *reinterpret_cast<const float4 *>(trow + i);
                            v[i] += t4.x; v[i + 1] += t4.y; v[i + 2] += t4.z; v[i + 3] += t4.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (i < ncol) v[i] += trow[i];
                    }
                }
                if (ep.act == 1) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
                } else if (ep.act == 2) {                          // tanh (BackboneNeck's 1x1 convolution, backbone.py:76-78)
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = tanhf(v[i]);
                }
                if (ep.relu_mask && row < ep.M) {
                    const float *mrow = ep.relu_mask + (size_t)row * ep.ldc + n0 + c0;
                    if (ncol >= 32) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 k4 = *reinterpret_cast<const float4 *>(mrow + i);
                            if (!(k4.x > 0.f)) v[i] = 0.f; if (!(k4.y > 0.f)) v[i + 1] = 0.f;
                            if (!(k4.z > 0.f)) v[i + 2] = 0.f; if (!(k4.w > 0.f)) v[i + 3] = 0.f;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (i < ncol && !(mrow[i] > 0.f)) v[i] = 0.f;
                    }
                }
                if (ep.round_out) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = tf32_rn(v[i]);
                }
                uint8_t *box = smem + (size_t)(c * 4 + q) * 4096;      // [32 rows][128 B], swizzle = chunk ^ (row & 7)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4 *>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                if (colsum_g) {
                    // column sums of the staged 32 x 32 block: lane = column, walking the rows (conflict-free: for a fixed
                    // row the 32 lanes read the 32 words of one swizzled 128-byte line)
                    __syncwarp();
                    const int nrow = min(32, ep.M - (m0 + q * 32));
                    float cs = 0.0f;
                    for (int r = 0; r < nrow; ++r)
                        cs += *reinterpret_cast<const float *>(box + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
                    if (nrow > 0) atomicAdd(&smem_colsum[c0 + lane], cs);
                }
            }
        }
        if (colsum_g) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int e = threadIdx.x - 64;
            if (e < BN && n0 + e < ep.N) atomicAdd(&colsum_g[n0 + e], smem_colsum[e]);
        }
        // One generic->async proxy fence for all staged boxes of this warp (it costs a few hundred cycles: once, not per
        // chunk), then the elected lane hands the boxes to the TMA engine.
        fence_proxy_async_smem();
        __syncwarp();
        if (m0 + q * 32 < ep.M) {
            if (elect_one()) {
#pragma unroll
                for (int c = 0; c < BN / 32; ++c) {
                    if (ep.N - (n0 + c * 32) > 0) {
                        uint8_t *box = smem + (size_t)(c * 4 + q) * 4096;
                        if (ep.atomic_out || ep.beta) tma_reduce_add_2d(map_c, box, n0 + c * 32, m0 + q * 32);
                        else tma_store_2d(map_c, box, n0 + c * 32, m0 + q * 32);
                    }
                }
                tma_store_commit();
                // the staging boxes are not reused: it is enough that the TMA engine has READ them before the CTA exits
                // (the global writes complete asynchronously and are ordered before grid completion)
                tma_store_wait_read_all();
            }
            __syncwarp();
        }
    }
    if (warp == 2 && lane == 0) UMMA_STAMP(6);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) UMMA_STAMP(7);
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN));
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
long long *g_umma_timeline = nullptr;
// cuTensorMapEncodeTiled is fetched through the runtime so that libbdetr.so has no link-time dependency on
// libcuda.so.1 (the library must also load on GPU-less build hosts).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

bool encode_tensor_map_2d(CUtensorMap *map, const float *base, long long rows, int cols, int ld, int box_cols, int box_rows,
                          bool mn_major)
{
    EncodeTiledFn cuTensorMapEncodeTiled = encode_tiled_fn();
    if (!cuTensorMapEncodeTiled) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

bool encode_tensor_map_2d_f16(CUtensorMap *map, const void *base, long long rows, int cols, int ld, int box_cols, int box_rows)
{
    EncodeTiledFn cuTensorMapEncodeTiled = encode_tiled_fn();
    if (!cuTensorMapEncodeTiled) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

constexpr bool UMMA_SHALLOW_DEFAULT = true;

template <int BN, int STAGES>
static size_t umma_smem_bytes() { return (size_t)STAGES * (UM_BM * UM_BK * 4 + BN * UM_BK * 4) + (2 * STAGES + 1) * 8 + 64 + 2 * BN * 4 + 1024; }

bool umma_gemm_eligible(int M, int N, int K, const float *A, int lda, bool TA, const float *B, int ldb, bool TB, int ldc)
{
    (void)TA; (void)TB;
    auto ok = [](const void *p, int ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 4 == 0; };
    return ok(A, lda) && ok(B, ldb) && ldc % 4 == 0 && K >= 32 && N >= 64 && M >= 128;
}

// Pipeline depth.  Deep = the whole K = 256 extent (8 k-blocks) in flight: lowest latency for ONE kernel (TMA round trips
// are ~1000 cycles), but 192 KB of stages means one CTA per SM, so GEMMs of concurrently running chains (encoder /
// decoder / heads backward on separate streams) queue for SMs.  Shallow = 96 KB (two CTAs per SM).  Selected once from
// the environment (BDETR_UMMA_STAGES=deep|shallow); the default is what the whole-step benchmark favours.
static bool umma_shallow()
{
    static const bool v = [] { const char *e = getenv("BDETR_UMMA_STAGES"); return e ? (e[0] == 's' || e[0] == 'S') : UMMA_SHALLOW_DEFAULT; }();
    return v;
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
static int launch_umma_stages(dim3 grid, const UmmaMaps &maps, const UmmaEpilogue &ep, cudaStream_t s)
{
    static bool optin = false;
    const size_t smem = umma_smem_bytes<BN, STAGES>();
    if (!optin) {
        BDETR_CUDA(cudaFuncSetAttribute(gemm_umma_kernel<BN, STAGES, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        optin = true;
    }
    launch_k(gemm_umma_kernel<BN, STAGES, A_MN, B_MN>, grid, UM_THREADS, smem, s, maps, ep);
    BDETR_CHECK_LAUNCH("gemm_umma_kernel");
    return BDETR_OK;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_umma_inst(dim3 grid, const UmmaMaps &maps, const UmmaEpilogue &ep, cudaStream_t s)
{
    if (umma_shallow()) return launch_umma_stages<BN, BN == 64 ? 4 : 3, A_MN, B_MN>(grid, maps, ep, s);
    return launch_umma_stages<BN, BN == 64 ? 8 : 6, A_MN, B_MN>(grid, maps, ep, s);
}

// General form: `groups` GEMMs of identical shape in one launch (see UmmaMaps).  TA: A stored [K,M]; TB: B stored [N,K].
int launch_gemm_umma_grouped(const GroupedGemm &g, cudaStream_t s)
{
    const int M = g.M, N = g.N, K = g.K, G = g.groups;
    BDETR_REQUIRE(G >= 1 && G <= UM_MAX_GROUPS && M > 0 && N > 0 && K > 0, BDETR_E_BAD_SHAPE, "bad grouped GEMM");
    const bool kmode = g.sum_groups && G > 1;
    // operand majors: A K-major when stored [M,K]; MN-major when stored [K,M].  B K-major when stored [N,K];
    // MN-major when stored [K,N] (Keras kernels and dY).
    const bool A_MN = g.TA, B_MN = !g.TB;
    const int n_groups = kmode ? 1 : G;
    // 128x64 tiles while they fit in one wave (two CTAs per SM), else 128x128
    const int BN = (N >= 128 && ceil_div(M, UM_BM) * ceil_div(N, 64) * n_groups > 148) ? 128 : 64;
    UmmaMaps maps;
    memset(&maps, 0, sizeof(maps));
    bool ok = true;
    auto ptr_ok = [](const void *p, int ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 4 == 0; };
    for (int i = 0; i < G; ++i) {
        const float *A = g.A[g.share_a ? 0 : i];
        BDETR_REQUIRE(A && g.B[i] && ptr_ok(A, g.lda) && ptr_ok(g.B[i], g.ldb), BDETR_E_BAD_SHAPE, "grouped GEMM operands must be 16-byte aligned");
        if (!g.share_a || i == 0) {
            if (!A_MN) ok = ok && encode_tensor_map_2d(&maps.a[i], A, M, K, g.lda, UM_BK, UM_BM, false);          // [M rows, K cols], box {32 k, 128 m}
            else ok = ok && encode_tensor_map_2d(&maps.a[i], A, K, M, g.lda, 32, UM_BK, true);                   // [K rows, M cols], box {32 m, 32 k}
        }
        if (!B_MN) ok = ok && encode_tensor_map_2d(&maps.b[i], g.B[i], N, K, g.ldb, UM_BK, BN, false);           // [N rows, K cols], box {32 k, BN n}
        else ok = ok && encode_tensor_map_2d(&maps.b[i], g.B[i], K, N, g.ldb, 32, UM_BK, true);                 // [K rows, N cols], box {32 n, 32 k}
        if (!kmode || i == 0) {
            BDETR_REQUIRE(g.C[i] && ptr_ok(g.C[i], g.ldc), BDETR_E_BAD_SHAPE, "grouped GEMM output must be 16-byte aligned");
            ok = ok && encode_tensor_map_2d(&maps.c[i], g.C[i], M, N, g.ldc, 32, 32, false);                    // output boxes: 32 x 32 per epilogue warp
        }
    }
    BDETR_REQUIRE(ok, BDETR_E_CUDA, "cuTensorMapEncodeTiled failed");

    UmmaEpilogue ep;
    memset(&ep, 0, sizeof(ep));
    ep.dbg = g_umma_timeline;
    ep.M = M; ep.N = N; ep.act = g.act; ep.relu_mask = g.relu_mask; ep.beta = g.beta; ep.round_out = g.round_out; ep.ldc = g.ldc;
    for (int i = 0; i < n_groups; ++i) { ep.bias[i] = g.bias[i]; ep.rowtab[i] = g.rowtab[i]; ep.colsum[i] = g.colsum[i]; }
    ep.rowtab_period = g.rowtab_period > 0 ? g.rowtab_period : M;
    ep.rowtab_ld = g.rowtab_ld > 0 ? g.rowtab_ld : N;
    ep.num_kb = ceil_div(K, UM_BK);
    ep.n_groups = n_groups; ep.k_groups = kmode ? G : 1; ep.share_a = g.share_a ? 1 : 0;
    ep.tiles_n = ceil_div(N, BN);
    const int tiles = ceil_div(M, UM_BM) * ep.tiles_n * n_groups;
    int splits = 1;
    const bool plain = g.act == 0 && g.relu_mask == nullptr && !g.round_out && !kmode;
    bool any_cs = false;
    for (int i = 0; i < n_groups; ++i) any_cs = any_cs || g.colsum[i] != nullptr;
    // deep contractions with few output tiles (weight gradients): cut K across CTAs, up to two resident CTAs per SM
    // Target CTA count of a split-K launch.  These are the parameter-gradient GEMMs: they run on side chains beside a
    // critical chain that is bound by SM slots, and nothing waits for them before the end of the block, so they are kept
    // THIN (80 / 96 CTAs, each walking a longer K range) instead of filling the GPU (148 / 296): step 2.59 -> 2.48 ms.
    // BDETR_WGRAD_CTAS=single[,grouped] overrides the targets for A/B runs.
    static const int wgrad_ctas[2] = {
        [] { const char *e = getenv("BDETR_WGRAD_CTAS"); return e ? atoi(e) : 0; }(),
        [] { const char *e = getenv("BDETR_WGRAD_CTAS"); const char *c = e ? strchr(e, ',') : nullptr; return c ? atoi(c + 1) : (e ? atoi(e) : 0); }()};
    const int wc = wgrad_ctas[n_groups > 1 ? 1 : 0];
    const int target = wc > 0 ? wc : (n_groups > 1 ? 96 : 80);
    if (plain && !any_cs && tiles < 74 && ep.num_kb >= 16) splits = min(ep.num_kb / 4, max(1, target / tiles));
    ep.kb_per_split = ceil_div(ep.num_kb, splits);
    splits = ceil_div(ep.num_kb, ep.kb_per_split);
    ep.atomic_out = splits > 1;
    if (ep.atomic_out && !g.beta)
        for (int i = 0; i < n_groups; ++i) BDETR_CUDA(cudaMemset2DAsync(g.C[i], (size_t)g.ldc * 4, 0, (size_t)N * 4, M, s));
    dim3 grid(ep.tiles_n * n_groups, ceil_div(M, UM_BM), splits);
#define UMMA_DISPATCH(BN_)                                                                        \
    do {                                                                                          \
        if (!A_MN && B_MN) return launch_umma_inst<BN_, false, true>(grid, maps, ep, s);              \
        if (!A_MN && !B_MN) return launch_umma_inst<BN_, false, false>(grid, maps, ep, s);            \
        if (A_MN && B_MN) return launch_umma_inst<BN_, true, true>(grid, maps, ep, s);                \
        return launch_umma_inst<BN_, true, false>(grid, maps, ep, s);                                 \
    } while (0)
    if (BN == 128) UMMA_DISPATCH(128);
    UMMA_DISPATCH(64);
#undef UMMA_DISPATCH
}

// Same contract as launch_gemm (gemm_simt.cu).
int launch_gemm_umma(int M, int N, int K, const float *A, int lda, bool TA, const float *B, int ldb, bool TB,
                     const float *bias, int act, const float *relu_mask, int beta, int round_out, float *C, int ldc,
                     cudaStream_t s)
{
    GroupedGemm g;
    g.M = M; g.N = N; g.K = K; g.TA = TA; g.TB = TB; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
    g.A[0] = A; g.B[0] = B; g.C[0] = C; g.bias[0] = bias;
    g.act = act; g.relu_mask = relu_mask; g.beta = beta; g.round_out = round_out;
    return launch_gemm_umma_grouped(g, s);
}

}  // namespace bdetr
