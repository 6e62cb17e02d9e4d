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
//   * warp 0 streams the raw A rows with TMA bulk copies (cp.async.bulk, one 256 B
//     K-half of a row per copy, padded pitch => conflict-free row reads) into a
//     2-stage shared-memory ring guarded by mbarriers;
//   * warps 4-7 (thread == row == TMEM lane) split the raw rows and tcgen05.st the
//     hi / lo K-halves into a double-buffered TMEM A operand;
//   * warp 1 (one elected thread) issues the MMAs into a double-buffered TMEM
//     accumulator and tcgen05.commit's the mbarriers that recycle A and publish D;
//   * warps 8-11 tcgen05.ld the accumulator, apply the activation, transpose 32x32 blocks through a
//     swizzled staging tile and store full 128-byte row segments.
// TMEM columns: D0 [0,128) D1 [128,256) A_hi0 [256,320) A_lo0 [320,384) A_hi1 [384,448)
// A_lo1 [448,512).
#include "tc_common.cuh"

namespace tmgcn {

namespace tc {

constexpr int TILE_M = 128;
constexpr int KH = 64;                       // K elements per pipeline step (one A buffer)
constexpr int RAW_PITCH = KH * 4 + 16;       // bytes; 17 x 16 B => conflict-free LDS.128 by row
constexpr int RAW_STAGE_BYTES = TILE_M * RAW_PITCH;
constexpr int RAW_STAGES = 2;
constexpr int NUM_THREADS = 384;

constexpr uint32_t TMEM_D0 = 0, TMEM_D1 = 128, TMEM_A0 = 256;  // A buffers: 128 columns each (hi 64 | lo 64)

// byte offset of element (n, k) inside a K-major SW128 operand of NO rows
__device__ __forceinline__ uint32_t b_offset(int n, int k, int NO) {
    const int chunk = k >> 5, kk = k & 31;
    return (uint32_t)(chunk * NO * 128 + (n >> 3) * 1024 + (n & 7) * 128 + ((((kk >> 2) ^ (n & 7))) << 4) +
                      ((kk & 3) << 2));
}

struct Params {
    const float *a;
    const float *w;
    float *c;
    int64_t R;
    int KR, NO;
    int act;
    int trans_w;
    int64_t n_tiles;
};

__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tf32x3_kernel(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B atoms must sit on 1024-byte boundaries of the shared window
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int KR = p.KR, NO = p.NO;
    const int NH = KR / KH;                               // K-halves per tile
    const uint32_t w_bytes = (uint32_t)KR * NO * 4;
    uint8_t *w_hi = smem;
    uint8_t *w_lo = smem + w_bytes;
    uint8_t *raw = smem + 2 * w_bytes;                    // RAW_STAGES x RAW_STAGE_BYTES
    uint8_t *epi_stage = raw + RAW_STAGES * RAW_STAGE_BYTES;          // 4 epilogue warps x 4 KB
    uint64_t *bars = reinterpret_cast<uint64_t *>(epi_stage + 4 * 4096);
    uint64_t *raw_full = bars, *raw_empty = bars + 2, *a_full = bars + 4, *a_empty = bars + 6, *d_full = bars + 8,
             *d_empty = bars + 10;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&raw_full[i], 1);
            mbar_init(&raw_empty[i], 4);
            mbar_init(&a_full[i], 4);
            mbar_init(&a_empty[i], 1);
            mbar_init(&d_full[i], 1);
            mbar_init(&d_empty[i], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
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
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // make W visible to the tensor-core proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t first = blockIdx.x, stride = gridDim.x;

    if (warp == 0) {
        // ================= TMA producer: raw A rows -> smem ring =================
        int64_t g = 0;   // global K-half counter
        for (int64_t tile = first; tile < p.n_tiles; tile += stride) {
            const int64_t row0 = tile * TILE_M;
            const int rows = (int)min((int64_t)TILE_M, p.R - row0);
            for (int h = 0; h < NH; ++h, ++g) {
                const int s = (int)(g & 1);
                mbar_wait(&raw_empty[s], (uint32_t)(((g >> 1) & 1) ^ 1));
                if (lane == 0) mbar_arrive_expect_tx(&raw_full[s], (uint32_t)rows * KH * 4);
                __syncwarp();
                uint8_t *dst = raw + s * RAW_STAGE_BYTES;
                for (int r = lane; r < rows; r += 32)
                    bulk_g2s(dst + r * RAW_PITCH, p.a + (row0 + r) * KR + h * KH, KH * 4, &raw_full[s]);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one elected thread) =================
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NO >> 3) << 17) |
                               ((uint32_t)(TILE_M >> 4) << 24);
        const uint32_t whi = smem_u32(w_hi), wlo = smem_u32(w_lo);
        int64_t g = 0, it = 0;
        for (int64_t tile = first; tile < p.n_tiles; tile += stride, ++it) {
            const int acc = (int)(it & 1);
            mbar_wait(&d_empty[acc], (uint32_t)(((it >> 1) & 1) ^ 1));
            const uint32_t d_tmem = tmem_base + (acc ? TMEM_D1 : TMEM_D0);
            for (int h = 0; h < NH; ++h, ++g) {
                const int ab = (int)(g & 1);
                mbar_wait(&a_full[ab], (uint32_t)((g >> 1) & 1));
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_hi = tmem_base + TMEM_A0 + ab * 128;
                    const uint32_t a_lo = a_hi + 64;
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
                    tc_commit(&a_empty[ab]);               // A buffer reusable once these MMAs retire
                    if (h == NH - 1) tc_commit(&d_full[acc]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ================= converters: raw smem row -> split -> TMEM A =================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        int64_t g = 0;
        for (int64_t tile = first; tile < p.n_tiles; tile += stride) {
            for (int h = 0; h < NH; ++h, ++g) {
                const int s = (int)(g & 1);
                const uint32_t par = (uint32_t)((g >> 1) & 1);
                mbar_wait(&raw_full[s], par);
                mbar_wait(&a_empty[s], par ^ 1);
                tc_fence_after();
                const float4 *src = reinterpret_cast<const float4 *>(raw + s * RAW_STAGE_BYTES + row * RAW_PITCH);
                const uint32_t t_hi = tmem_base + lane_base + TMEM_A0 + s * 128;
#pragma unroll
                for (int c = 0; c < KH / 16; ++c) {
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const float4 x = src[c * 4 + v];
                        split_tf32(x.x, hi[4 * v + 0], lo[4 * v + 0]);
                        split_tf32(x.y, hi[4 * v + 1], lo[4 * v + 1]);
                        split_tf32(x.z, hi[4 * v + 2], lo[4 * v + 2]);
                        split_tf32(x.w, hi[4 * v + 3], lo[4 * v + 3]);
                    }
                    tmem_st16(t_hi + c * 16, hi);
                    tmem_st16(t_hi + 64 + c * 16, lo);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&a_full[s]);
                    mbar_arrive(&raw_empty[s]);
                }
            }
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
        for (int64_t tile = first; tile < p.n_tiles; tile += stride, ++it) {
            const int acc = (int)(it & 1);
            mbar_wait(&d_full[acc], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const int64_t row_base = tile * TILE_M + q * 32;
            const uint32_t t_d = tmem_base + lane_base + (acc ? TMEM_D1 : TMEM_D0);
            for (int c0 = 0; c0 < NO; c0 += 32) {
                const int units = min(32, NO - c0) >> 2;           // 16-byte units in this column block (4 or 8)
                uint32_t v[32];
                tmem_ld16(t_d + c0, v);
                if (units > 4) tmem_ld16(t_d + c0 + 16, v + 16);
                tmem_wait_ld();
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (u < units) {
                        float4 o;
                        o.x = act_apply_rt(__uint_as_float(v[4 * u + 0]), p.act);
                        o.y = act_apply_rt(__uint_as_float(v[4 * u + 1]), p.act);
                        o.z = act_apply_rt(__uint_as_float(v[4 * u + 2]), p.act);
                        o.w = act_apply_rt(__uint_as_float(v[4 * u + 3]), p.act);
                        *reinterpret_cast<float4 *>(stage + lane * 128 + ((u ^ (lane & 7)) << 4)) = o;
                    }
                }
                __syncwarp();
                const int u = lane & 7;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = i * 4 + (lane >> 3);
                    const int64_t row = row_base + r;
                    if (u < units && row < p.R) {
                        const float4 o = *reinterpret_cast<const float4 *>(stage + r * 128 + ((u ^ (r & 7)) << 4));
                        *reinterpret_cast<float4 *>(p.c + row * NO + c0 + u * 4) = o;
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

static size_t smem_bytes(int KR, int NO) {
    return (size_t)2 * KR * NO * 4 + (size_t)RAW_STAGES * RAW_STAGE_BYTES + 4 * 4096 + 16 * 8 +
           1024 /* alignment slack */;
}

}  // namespace tc

bool gemm_tc_eligible(int64_t R, int KR, int NO) {
    if (R < 1) return false;
    if (KR != 64 && KR != 128) return false;
    if (NO < 16 || NO > 128 || (NO % 16) != 0) return false;
    return tc::smem_bytes(KR, NO) <= 227 * 1024;
}

// C[R, NO] = act(A[R, KR] . B), see the header comment for B.  A and C must be 16-byte aligned.
int gemm_tc_fwd(const float *a, const float *w, float *c, int64_t R, int KR, int NO, int act, bool trans_w,
                const float *yaux, cudaStream_t st) {
    TMGCN_REQUIRE(yaux == nullptr, "gemm_tc: fused activation gradient is not supported on the tensor-core path");
    TMGCN_REQUIRE(((uintptr_t)a % 16 == 0) && ((uintptr_t)c % 16 == 0), "gemm_tc: operands must be 16-byte aligned");
    tc::Params p;
    p.a = a;
    p.w = w;
    p.c = c;
    p.R = R;
    p.KR = KR;
    p.NO = NO;
    p.act = act;
    p.trans_w = trans_w ? 1 : 0;
    p.n_tiles = ceil_div(R, tc::TILE_M);
    const size_t smem = tc::smem_bytes(KR, NO);
    TMGCN_CUDA(cudaFuncSetAttribute(tc::gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    227 * 1024));
    int64_t grid = sm_count();
    if (grid > p.n_tiles) grid = p.n_tiles;
    tc::gemm_tf32x3_kernel<<<(unsigned)grid, tc::NUM_THREADS, smem, st>>>(p);
    return after_launch("gemm_tf32x3");
}

}  // namespace tmgcn
