// (c) feature GEMM on the 5th-generation tensor cores: tcgen05.mma (kind::tf32) with
// the accumulator AND the A operand in TMEM, 3xTF32 error compensation.
//
// ref: t.matmul(AtXt, W)  ehf:222 / 330 / 344; backward dP = dY . W^T (autograd).
//
//   C[R, NO] = act( A[R, KR] . B )      B[k][n] = W[k*NO + n]  (fwd,  W is KR x NO)
//                                        B[k][n] = W[n*KR + k]  (dP,   W is NO x KR)
//
// The reference multiplies in true fp32.  tcgen05 has no fp32-input MMA, so every
// operand is split x = hi + lo with hi = tf32(x), lo = tf32(x - hi) and the product is
// accumulated in fp32 as  hi*lo' + lo*hi' + hi*hi'  (dropped term ~2^-22): three
// kind::tf32 MMAs per K step, ||err||/||ref|| ~ 1e-6 < the 1e-5 parity bar.
//
// The kernel is HBM-bound (AI = 32 flop/B at K = N = 128), so the design goal is to
// stream A once and C once with everything else hidden:
//   * persistent, one CTA per SM, static tile striding; a tile = 128 rows;
//   * W is split once per CTA into W_hi / W_lo and parked in shared memory in the
//     canonical K-major SWIZZLE_128B UMMA layout (2 x 64 KB at K = N = 128);
//   * warp 0 streams raw A with TMA tensor copies (cp.async.bulk.tensor.2d, 32 rows x 32
//     columns per box, SWIZZLE_128B) into a multi-stage ring of 32-row chunks guarded by
//     mbarriers; the hardware swizzle makes the converters' by-row reads bank-conflict free
//     (plain row-major chunks cost a 32-way conflict: measured 6.2K cycles per tile);
//   * warps 4-7 (thread == row == TMEM lane) each own one 32-row chunk of the tile, split the
//     raw rows and tcgen05.st hi / lo into a K-half double-buffered TMEM A operand, so the
//     conversion of one K-half overlaps the MMAs of the other;
//   * warp 1 (one elected thread) issues the MMAs into a double-buffered TMEM
//     accumulator and tcgen05.commit's the mbarriers that recycle A and publish D;
//   * warps 8-11 tcgen05.ld the accumulator, apply the activation, transpose 32x32 blocks
//     through a swizzled staging tile and store full 128-byte row segments.
// TMEM columns: D0 [0,128) D1 [128,256) A buffer h: hi [256+128h, +64) lo [320+128h, +64).
#include <cuda.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace tmgcn {

namespace tc {

constexpr int TILE_M = 128;
constexpr int CHUNK_ROWS = 32;               // rows per raw stage = one converter warp's share of a tile
constexpr int BOX_COLS = 32;                 // 32 fp32 = 128 B = the swizzle span
constexpr int KH = 64;                       // K elements per A buffer
constexpr int MAX_RAW_STAGES = 8;
constexpr int NUM_THREADS = 384;

constexpr uint32_t TMEM_D0 = 0, TMEM_D1 = 128, TMEM_A0 = 256;

// byte offset of element (n, k) inside a K-major SW128 operand of NO rows
__device__ __forceinline__ uint32_t b_offset(int n, int k, int NO) {
    const int chunk = k >> 5, kk = k & 31;
    return (uint32_t)(chunk * NO * 128 + (n >> 3) * 1024 + (n & 7) * 128 + ((((kk >> 2) ^ (n & 7))) << 4) +
                      ((kk & 3) << 2));
}

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

struct Params {
    const float *w;
    float *c;
    int64_t R;
    int KR, NO;
    int act;
    int trans_w;
    int raw_stages;
    int64_t n_tiles;
    unsigned long long *prof;   // debug: per-role cycle counters of CTA 0 (TMGCN_TC_PROF=1), else nullptr
};

// park this lane's 32 accumulator values (one row of a 32x32 block) in the warp's staging tile
template <int ACT>
__device__ __forceinline__ void stage_block(uint8_t *stage, const uint32_t *v, int lane, int units) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        if (u < units) {
            float4 o;
            o.x = act_apply<ACT>(__uint_as_float(v[4 * u + 0]));
            o.y = act_apply<ACT>(__uint_as_float(v[4 * u + 1]));
            o.z = act_apply<ACT>(__uint_as_float(v[4 * u + 2]));
            o.w = act_apply<ACT>(__uint_as_float(v[4 * u + 3]));
            *reinterpret_cast<float4 *>(stage + lane * 128 + ((u ^ (lane & 7)) << 4)) = o;
        }
    }
}

#define PROF_ADD(var)                         \
    do {                                      \
        if (p.prof) {                         \
            const long long t1__ = clock64(); \
            var += t1__ - tl__;               \
            tl__ = t1__;                      \
        }                                     \
    } while (0)

__global__ void __launch_bounds__(NUM_THREADS, 1)
    gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap a_map, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B atoms / TMA boxes must sit on 1024-byte boundaries of the shared window
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int KR = p.KR, NO = p.NO, S = p.raw_stages;
    const int NH = KR / KH;                                            // K-halves per tile (1 or 2)
    const int n_boxes = KR / BOX_COLS;                                 // TMA boxes per chunk
    const uint32_t w_bytes = (uint32_t)KR * NO * 4;
    const uint32_t chunk_bytes = (uint32_t)CHUNK_ROWS * KR * 4;        // n_boxes x 4 KB
    uint8_t *w_hi = smem;
    uint8_t *w_lo = smem + w_bytes;
    uint8_t *raw = smem + 2 * w_bytes;                                 // S x chunk_bytes
    uint8_t *epi_stage = raw + S * chunk_bytes;                        // 4 epilogue warps x 4 KB
    uint64_t *bars = reinterpret_cast<uint64_t *>(epi_stage + 4 * 4096);
    uint64_t *raw_full = bars, *raw_empty = bars + MAX_RAW_STAGES;
    uint64_t *a_full = bars + 2 * MAX_RAW_STAGES, *a_empty = a_full + 2, *d_full = a_full + 4, *d_empty = a_full + 6;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_full + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&a_map) : "memory");
        for (int i = 0; i < S; ++i) {
            mbar_init(&raw_full[i], 1);
            mbar_init(&raw_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 4);
            mbar_init(&a_empty[i], 1);
            mbar_init(&d_full[i], 1);
            mbar_init(&d_empty[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    // split W once: generic-proxy stores into the swizzled UMMA layout
    for (int idx = threadIdx.x; idx < KR * NO; idx += NUM_THREADS) {
        int k, n;
        if (p.trans_w) {            // W is (NO x KR): idx = n*KR + k
            n = idx / KR;
            k = idx - n * KR;
        } else {                    // W is (KR x NO): idx = k*NO + n
            k = idx / NO;
            n = idx - k * NO;
        }
        uint32_t hi, lo;
        split_tf32(__ldg(p.w + idx), hi, lo);
        const uint32_t off = b_offset(n, k, NO);
        *reinterpret_cast<uint32_t *>(w_hi + off) = hi;
        *reinterpret_cast<uint32_t *>(w_lo + off) = lo;
    }
    fence_proxy_async();               // make W visible to the tensor-core (async) proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t first = blockIdx.x, stride = gridDim.x;

    if (warp == 0) {
        // ================= TMA producer: 32-row chunks, one swizzled 32x32 box per 128-byte column block ===
        if (lane == 0) {
            int64_t g = 0;   // global chunk counter
            long long tl__ = p.prof ? clock64() : 0, w_empty = 0, w_issue = 0;
            for (int64_t tile = first; tile < p.n_tiles; tile += stride) {
                for (int q = 0; q < TILE_M / CHUNK_ROWS; ++q, ++g) {
                    const int s = (int)(g % S);
                    PROF_ADD(w_issue);
                    mbar_wait(&raw_empty[s], (uint32_t)(((g / S) & 1) ^ 1));
                    PROF_ADD(w_empty);
                    const int64_t row0 = tile * TILE_M + q * CHUNK_ROWS;
                    if (row0 < p.R) {
                        // rows past R are zero-filled by the TMA unit and still counted in the transaction bytes
                        mbar_arrive_expect_tx(&raw_full[s], chunk_bytes);
                        for (int b = 0; b < n_boxes; ++b)
                            tma_load_2d(raw + s * chunk_bytes + b * 4096, &a_map, b * BOX_COLS, (int)row0,
                                        &raw_full[s]);
                    } else {
                        mbar_arrive(&raw_full[s]);      // nothing to load: just release the consumer
                    }
                }
            }
            if (p.prof && blockIdx.x == 0) {
                p.prof[0] = w_empty;
                p.prof[1] = w_issue;
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one elected thread) =================
        const uint32_t idesc = make_idesc_tf32(TILE_M, NO, false);
        const uint32_t whi = smem_u32(w_hi), wlo = smem_u32(w_lo);
        int64_t it = 0, gh = 0;
        long long tl__ = p.prof ? clock64() : 0, w_dempty = 0, w_afull = 0, w_issue = 0;
        for (int64_t tile = first; tile < p.n_tiles; tile += stride, ++it) {
            const int acc = (int)(it & 1);
            PROF_ADD(w_issue);
            mbar_wait(&d_empty[acc], (uint32_t)(((it >> 1) & 1) ^ 1));
            PROF_ADD(w_dempty);
            const uint32_t d_tmem = tmem_base + (acc ? TMEM_D1 : TMEM_D0);
            for (int h = 0; h < NH; ++h, ++gh) {
                const int ab = (int)(gh & 1);
                mbar_wait(&a_full[ab], (uint32_t)((gh >> 1) & 1));
                PROF_ADD(w_afull);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_hi = tmem_base + TMEM_A0 + ab * 128, a_lo = a_hi + 64;
#pragma unroll
                    for (int j = 0; j < KH / 8; ++j) {
                        const int k = h * KH + j * 8;
                        const uint32_t boff = (uint32_t)((k >> 5) * NO * 128 + ((k & 31) >> 3) * 32);
                        const uint64_t b_hi = make_desc_sw128(whi + boff, 16, 1024);
                        const uint64_t b_lo = make_desc_sw128(wlo + boff, 16, 1024);
                        mma_tf32_ts(d_tmem, a_hi + j * 8, b_lo, idesc, (h | j) ? 1u : 0u);
                        mma_tf32_ts(d_tmem, a_lo + j * 8, b_hi, idesc, 1u);
                        mma_tf32_ts(d_tmem, a_hi + j * 8, b_hi, idesc, 1u);
                    }
                    tc_commit(&a_empty[ab]);            // A buffer reusable once these MMAs retire
                    if (h == NH - 1) tc_commit(&d_full[acc]);
                }
                __syncwarp();
                PROF_ADD(w_issue);
            }
        }
        if (p.prof && blockIdx.x == 0 && lane == 0) {
            p.prof[2] = w_dempty;
            p.prof[3] = w_afull;
            p.prof[4] = w_issue;
        }
    } else if (warp >= 4 && warp < 8) {
        // ================= converters: swizzled raw row -> split -> TMEM A (thread == row == lane) ========
        const int q = warp & 3;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const uint32_t row_off = (uint32_t)lane * 128;
        const int sw = lane & 7;
        int64_t it = 0, gh = 0;
        long long tl__ = p.prof ? clock64() : 0, w_raw = 0, w_aempty = 0, w_work = 0;
        for (int64_t tile = first; tile < p.n_tiles; tile += stride, ++it) {
            const int64_t g = it * (TILE_M / CHUNK_ROWS) + q;
            const int s = (int)(g % S);
            PROF_ADD(w_work);
            mbar_wait(&raw_full[s], (uint32_t)((g / S) & 1));
            PROF_ADD(w_raw);
            const uint8_t *chunk = raw + s * chunk_bytes;
            for (int h = 0; h < NH; ++h, ++gh) {
                const int ab = (int)(gh & 1);
                mbar_wait(&a_empty[ab], (uint32_t)(((gh >> 1) & 1) ^ 1));
                PROF_ADD(w_aempty);
                tc_fence_after();
                const uint32_t t_hi = tmem_base + lane_base + TMEM_A0 + ab * 128;
#pragma unroll
                for (int c = 0; c < KH / 16; ++c) {                    // 16 columns = 4 x 16-byte units
                    const int k0 = h * KH + c * 16;
                    const uint8_t *box = chunk + (k0 >> 5) * 4096 + row_off;
                    const int ub = (k0 & 31) >> 2;
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const float4 x = *reinterpret_cast<const float4 *>(box + (((ub + v) ^ sw) << 4));
                        split_tf32(x.x, hi[4 * v + 0], lo[4 * v + 0]);
                        split_tf32(x.y, hi[4 * v + 1], lo[4 * v + 1]);
                        split_tf32(x.z, hi[4 * v + 2], lo[4 * v + 2]);
                        split_tf32(x.w, hi[4 * v + 3], lo[4 * v + 3]);
                    }
                    tmem_st16(t_hi + c * 16, hi);
                    tmem_st16(t_hi + 64 + c * 16, lo);
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&a_full[ab]);
                    if (h == NH - 1) mbar_arrive(&raw_empty[s]);
                }
                PROF_ADD(w_work);
            }
        }
        if (p.prof && blockIdx.x == 0 && warp == 4 && lane == 0) {
            p.prof[5] = w_raw;
            p.prof[6] = w_aempty;
            p.prof[7] = w_work;
        }
    } else if (warp >= 8) {
        // ================= epilogue: TMEM D -> act -> smem transpose -> coalesced global rows =========
        // A thread owns a ROW of the accumulator (TMEM lane), so storing straight from registers would
        // touch 32 different 128-byte lines per instruction.  Each warp instead parks a 32-row x 32-column
        // block in its private 4 KB staging tile (16-byte units XOR-swizzled by row: conflict-free both
        // ways) and writes it back as 4 full 128-byte row segments per instruction.
        const int q = warp & 3;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        uint8_t *stage = epi_stage + q * 4096;
        int64_t it = 0;
        long long tl__ = p.prof ? clock64() : 0, w_dfull = 0, w_work = 0;
        for (int64_t tile = first; tile < p.n_tiles; tile += stride, ++it) {
            const int acc = (int)(it & 1);
            PROF_ADD(w_work);
            mbar_wait(&d_full[acc], (uint32_t)((it >> 1) & 1));
            PROF_ADD(w_dfull);
            tc_fence_after();
            const int64_t row_base = tile * TILE_M + q * 32;
            const uint32_t t_d = tmem_base + lane_base + (acc ? TMEM_D1 : TMEM_D0);
            for (int c0 = 0; c0 < NO; c0 += 32) {
                const int units = min(32, NO - c0) >> 2;           // 16-byte units in this column block (4 or 8)
                uint32_t v[32];
                tmem_ld16(t_d + c0, v);
                if (units > 4) tmem_ld16(t_d + c0 + 16, v + 16);
                tmem_wait_ld();
                // the activation is selected OUTSIDE the element loop: a per-element runtime switch gets
                // if-converted and evaluates expm1f for every value (measured: 12K of 17K cycles per tile)
                switch (p.act) {
                    case TMGCN_ACT_RELU: stage_block<TMGCN_ACT_RELU>(stage, v, lane, units); break;
                    case TMGCN_ACT_LEAKY: stage_block<TMGCN_ACT_LEAKY>(stage, v, lane, units); break;
                    case TMGCN_ACT_SELU: stage_block<TMGCN_ACT_SELU>(stage, v, lane, units); break;
                    default: stage_block<TMGCN_ACT_NONE>(stage, v, lane, units); break;
                }
                __syncwarp();
                const int u = lane & 7;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = i * 4 + (lane >> 3);
                    const int64_t row = row_base + r;
                    if (u < units && row < p.R) {
                        const float4 o = *reinterpret_cast<const float4 *>(stage + r * 128 + ((u ^ (r & 7)) << 4));
                        st_stream_f4(reinterpret_cast<float4 *>(p.c + row * NO + c0 + u * 4), o);
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[acc]);
        }
        if (p.prof && blockIdx.x == 0 && warp == 8 && lane == 0) {
            p.prof[8] = w_dfull;
            p.prof[9] = w_work;
            p.prof[10] = it;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

static size_t smem_fixed(int KR, int NO) {
    return (size_t)2 * KR * NO * 4 + 4 * 4096 + (2 * MAX_RAW_STAGES + 10) * 8 + 1024 /* alignment slack */;
}
static int raw_stages_for(int KR, int NO) {
    const size_t budget = 227 * 1024;
    const size_t fixed = smem_fixed(KR, NO);
    if (fixed >= budget) return 0;
    int s = (int)((budget - fixed) / ((size_t)CHUNK_ROWS * KR * 4));
    return s > MAX_RAW_STAGES ? MAX_RAW_STAGES : s;
}
static size_t smem_bytes(int KR, int NO) {
    return smem_fixed(KR, NO) + (size_t)raw_stages_for(KR, NO) * CHUNK_ROWS * KR * 4;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time dependency on libcuda
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

}  // namespace tc

bool gemm_tc_eligible(int64_t R, int KR, int NO) {
    if (R < 1 || R > 0x7fffffffLL) return false;          // TMA coordinates are 32-bit
    if (KR != 64 && KR != 128) return false;
    if (NO < 16 || NO > 128 || (NO % 16) != 0) return false;
    return tc::raw_stages_for(KR, NO) >= 4;     // at least one whole tile of raw rows in flight
}

// C[R, NO] = act(A[R, KR] . B), see the header comment for B.  A and C must be 16-byte aligned.
int gemm_tc_fwd(const float *a, const float *w, float *c, int64_t R, int KR, int NO, int act, bool trans_w,
                const float *yaux, cudaStream_t st) {
    TMGCN_REQUIRE(yaux == nullptr, "gemm_tc: fused activation gradient is not supported on the tensor-core path");
    TMGCN_REQUIRE(((uintptr_t)a % 16 == 0) && ((uintptr_t)c % 16 == 0), "gemm_tc: operands must be 16-byte aligned");
    tc::EncodeTiledFn enc = tc::encode_tiled();
    TMGCN_REQUIRE(enc != nullptr, "gemm_tc: cuTensorMapEncodeTiled is unavailable in this driver");
    CUtensorMap a_map;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)KR, (cuuint64_t)R};
        const cuuint64_t gstride[1] = {(cuuint64_t)KR * 4};
        const cuuint32_t box[2] = {(cuuint32_t)tc::BOX_COLS, (cuuint32_t)tc::CHUNK_ROWS};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&a_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)a, gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TMGCN_REQUIRE(r == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
    tc::Params p;
    p.w = w;
    p.c = c;
    p.R = R;
    p.KR = KR;
    p.NO = NO;
    p.act = act;
    p.trans_w = trans_w ? 1 : 0;
    p.raw_stages = tc::raw_stages_for(KR, NO);
    p.n_tiles = ceil_div(R, tc::TILE_M);
    const size_t smem = tc::smem_bytes(KR, NO);
    TMGCN_CUDA(cudaFuncSetAttribute(tc::gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    227 * 1024));
    int64_t grid = sm_count();
    if (grid > p.n_tiles) grid = p.n_tiles;
    static int prof_mode = -1;
    if (prof_mode < 0) {
        const char *e = getenv("TMGCN_TC_PROF");
        prof_mode = (e && e[0] == '1') ? 1 : 0;
    }
    p.prof = nullptr;
    if (prof_mode) {
        TMGCN_CUDA(cudaMalloc(&p.prof, 16 * sizeof(unsigned long long)));
        TMGCN_CUDA(cudaMemsetAsync(p.prof, 0, 16 * sizeof(unsigned long long), st));
    }
    tc::gemm_tf32x3_kernel<<<(unsigned)grid, tc::NUM_THREADS, smem, st>>>(a_map, p);
    if (after_launch("gemm_tf32x3")) return 1;
    if (prof_mode) {   // debug only: synchronous read-back of CTA 0's role counters
        unsigned long long h[16];
        TMGCN_CUDA(cudaStreamSynchronize(st));
        TMGCN_CUDA(cudaMemcpy(h, p.prof, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(p.prof);
        const double t = h[10] ? (double)h[10] : 1.0;
        fprintf(stderr,
                "[gemm_tf32x3 prof, cycles/tile over %llu tiles] producer: wait_empty %.0f issue %.0f | mma: wait_d_empty "
                "%.0f wait_a_full %.0f issue %.0f | conv: wait_raw %.0f wait_a_empty %.0f work %.0f | epi: wait_d_full "
                "%.0f work %.0f\n",
                h[10], h[0] / t, h[1] / t, h[2] / t, h[3] / t, h[4] / t, h[5] / t, h[6] / t, h[7] / t, h[8] / t, h[9] / t);
    }
    return 0;
}

}  // namespace tmgcn
