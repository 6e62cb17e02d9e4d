// Slice-summed weight gradient  dW[KI, NO] = sum_r P[r, :]^T dY[r, :]  on tcgen05.
//
// ref: autograd of t.matmul(AtXt, W) (ehf:222 / 344) -- dW = sum over all T*N rows.
//
// One MMA problem per CTA: D (M = KI, N = NO) += A (M x 8) . B (8 x N) marching over the
// rows of this CTA's row range, 3xTF32 compensated like the forward GEMM
// (A_hi.B_lo + A_lo.B_hi + A_hi.B_hi).  The reduction dimension is the ROW index, so
//   A[m = ki][k = r] = P[r][ki]   -- the transpose of the row-major tile: it is produced for
//       free by letting TMEM lane ki (= converter thread ki) read column ki of the raw
//       tile in shared memory (conflict-free) and tcgen05.st it as a K-major TMEM operand;
//   B[k = r][n = no] = dY[r][no]  -- converter thread n reads column n of the raw dY tile the same
//       way and writes row n (32 k-values = one 128-byte swizzle row) of a K-major SWIZZLE_128B
//       shared-memory operand, split into hi and lo.
// Raw 32-row chunks of P and dY are contiguous 16 KB blocks: one TMA bulk copy each.
// The tensor core's fp32 accumulation is not round-to-nearest: over hundreds of thousands of rows the
// error grows linearly (measured 1e-4 at 13K rows per CTA, i.e. ~3e-3 at the 64M-row benchmark shard), so a
// TMEM accumulator only ever sums DW_FLUSH chunks (2048 rows); four epilogue warps then add it into an fp32
// shared-memory accumulator with ordinary rounded adds while the MMAs continue into the second TMEM
// accumulator.  Each CTA writes one partial (KI x NO) and a second kernel sums the partials in fixed
// order (deterministic dW).
// The kernel streams P and dY exactly once: 8*N*F bytes per slice.
#include "tc_common.cuh"

namespace tmgcn {
namespace tc {

constexpr int DW_KC = 32;                        // rows per chunk (MMA K = 8 => 4 steps)
constexpr int DW_STAGES = 2;
constexpr int DW_FLUSH = 64;                     // chunks (2048 rows) accumulated in TMEM before a flush
constexpr int DW_THREADS = 512;
// two accumulators (one being flushed while the other accumulates), then the A stages
constexpr uint32_t DW_TMEM_D = 0, DW_TMEM_A = 256;   // A stage s: [256 + 64 s, +32) hi, [+32, +64) lo

struct DwParams {
    const float *p;     // (R, KI)
    const float *dy;    // (R, NO)
    float *partial;     // (grid, KI, NO)
    int64_t R;
    int KI, NO;
    int64_t chunks_per_cta;
};

__global__ void __launch_bounds__(DW_THREADS, 1) gemm_dw_tf32x3_kernel(const DwParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int KI = p.KI, NO = p.NO;
    const uint32_t rawp_bytes = DW_KC * KI * 4, rawd_bytes = DW_KC * NO * 4;
    const uint32_t b_bytes = DW_KC * NO * 4;           // one of B_hi / B_lo, one stage
    uint8_t *b_hi = smem;                                  // DW_STAGES x b_bytes   (1024-aligned atoms)
    uint8_t *b_lo = b_hi + DW_STAGES * b_bytes;
    uint8_t *raw_p = b_lo + DW_STAGES * b_bytes;
    uint8_t *raw_d = raw_p + DW_STAGES * rawp_bytes;
    float *accs = reinterpret_cast<float *>(raw_d + DW_STAGES * rawd_bytes);   // [NO][128] column-major fp32 sum
    uint64_t *bars = reinterpret_cast<uint64_t *>(accs + 128 * NO);
    uint64_t *raw_full = bars;                     // [S]  TMA landed
    uint64_t *rawp_empty = bars + DW_STAGES;       // [S]  A converters done reading raw P
    uint64_t *rawd_empty = bars + 2 * DW_STAGES;   // [S]  B converters done reading raw dY
    uint64_t *ab_full = bars + 3 * DW_STAGES;      // [S]  A (TMEM) and B (smem) operands ready (8 warp arrivals)
    uint64_t *ab_empty = bars + 4 * DW_STAGES;     // [S]  MMAs that read stage s retired
    uint64_t *d_full = bars + 5 * DW_STAGES;       // [2]  a TMEM accumulator holds a finished group of chunks
    uint64_t *d_empty = bars + 5 * DW_STAGES + 2;  // [2]  ... and has been added into accs
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 5 * DW_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < DW_STAGES; ++s) {
            mbar_init(&raw_full[s], 1);
            mbar_init(&rawp_empty[s], 4);
            mbar_init(&rawd_empty[s], 4);
            mbar_init(&ab_full[s], 8);
            mbar_init(&ab_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&d_full[i], 1);
            mbar_init(&d_empty[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t n_chunks_total = (p.R + DW_KC - 1) / DW_KC;
    const int64_t c_begin = (int64_t)blockIdx.x * p.chunks_per_cta;
    const int64_t c_end = min(n_chunks_total, c_begin + p.chunks_per_cta);
    const int64_t n_my = c_end > c_begin ? c_end - c_begin : 0;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            for (int64_t i = 0; i < n_my; ++i) {
                const int s = (int)(i % DW_STAGES);
                const uint32_t par = (uint32_t)((i / DW_STAGES) & 1);
                mbar_wait(&rawp_empty[s], par ^ 1);
                mbar_wait(&rawd_empty[s], par ^ 1);
                const int64_t row0 = (c_begin + i) * DW_KC;
                const uint32_t rows = (uint32_t)min((int64_t)DW_KC, p.R - row0);
                mbar_arrive_expect_tx(&raw_full[s], rows * (uint32_t)(KI + NO) * 4);
                bulk_g2s(raw_p + s * rawp_bytes, p.p + row0 * KI, rows * KI * 4, &raw_full[s]);
                bulk_g2s(raw_d + s * rawd_bytes, p.dy + row0 * NO, rows * NO * 4, &raw_full[s]);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc_tf32(128, NO, /*b_mn_major=*/false);
        const uint32_t bhi0 = smem_u32(b_hi), blo0 = smem_u32(b_lo);
        for (int64_t i = 0; i < n_my; ++i) {
            const int s = (int)(i % DW_STAGES);
            const int64_t grp = i / DW_FLUSH;
            const int in_grp = (int)(i % DW_FLUSH);
            const int acc = (int)(grp & 1);
            if (in_grp == 0) mbar_wait(&d_empty[acc], (uint32_t)(((grp >> 1) & 1) ^ 1));
            mbar_wait(&ab_full[s], (uint32_t)((i / DW_STAGES) & 1));
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_hi = tmem_base + DW_TMEM_A + s * 64, a_lo = a_hi + 32;
                const uint32_t d_tmem = tmem_base + DW_TMEM_D + acc * 128;
#pragma unroll
                for (int j = 0; j < DW_KC / 8; ++j) {
                    const uint64_t dhi = make_desc_sw128(bhi0 + s * b_bytes + j * 32, 16, 1024);
                    const uint64_t dlo = make_desc_sw128(blo0 + s * b_bytes + j * 32, 16, 1024);
                    mma_tf32_ts(d_tmem, a_hi + j * 8, dlo, idesc, (in_grp | j) ? 1u : 0u);
                    mma_tf32_ts(d_tmem, a_lo + j * 8, dhi, idesc, 1u);
                    mma_tf32_ts(d_tmem, a_hi + j * 8, dhi, idesc, 1u);
                }
                tc_commit(&ab_empty[s]);
                if (in_grp == DW_FLUSH - 1 || i == n_my - 1) tc_commit(&d_full[acc]);
            }
            __syncwarp();
        }
    } else if (warp >= 4 && warp < 8) {
        // ================= A converters: column ki of raw P -> TMEM lane ki =================
        const int q = warp & 3;
        const int ki = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        for (int64_t i = 0; i < n_my; ++i) {
            const int s = (int)(i % DW_STAGES);
            const uint32_t par = (uint32_t)((i / DW_STAGES) & 1);
            const int64_t row0 = (c_begin + i) * DW_KC;
            const int rows = (int)min((int64_t)DW_KC, p.R - row0);
            mbar_wait(&raw_full[s], par);
            mbar_wait(&ab_empty[s], par ^ 1);
            tc_fence_after();
            const float *src = reinterpret_cast<const float *>(raw_p + s * rawp_bytes) + ki;
            uint32_t hi[DW_KC], lo[DW_KC];
#pragma unroll
            for (int r = 0; r < DW_KC; ++r) {
                const float x = (ki < KI && r < rows) ? src[r * KI] : 0.f;
                split_tf32(x, hi[r], lo[r]);
            }
            const uint32_t t_a = tmem_base + lane_base + DW_TMEM_A + s * 64;
            tmem_st32(t_a, hi);
            tmem_st32(t_a + 32, lo);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&rawp_empty[s]);
                mbar_arrive(&ab_full[s]);
            }
        }
    } else if (warp >= 8 && warp < 12) {
        // ================= B converters: column n of raw dY -> row n of a K-major SWIZZLE_128B tile =====
        // (tf32 MN-major operands would need the SW128_32B layout; transposing here keeps the layout the
        //  forward kernel uses: row n = 32 k-values = one 128-byte swizzle row, 8-row groups 1024 B apart)
        const int n = threadIdx.x - 256;                   // 0..127
        for (int64_t i = 0; i < n_my; ++i) {
            const int s = (int)(i % DW_STAGES);
            const uint32_t par = (uint32_t)((i / DW_STAGES) & 1);
            const int64_t row0 = (c_begin + i) * DW_KC;
            const int rows = (int)min((int64_t)DW_KC, p.R - row0);
            mbar_wait(&raw_full[s], par);
            mbar_wait(&ab_empty[s], par ^ 1);
            if (n < NO) {
                const float *src = reinterpret_cast<const float *>(raw_d + s * rawd_bytes) + n;
                uint8_t *dhi = b_hi + s * b_bytes + (n >> 3) * 1024 + (n & 7) * 128;
                uint8_t *dlo = b_lo + s * b_bytes + (n >> 3) * 1024 + (n & 7) * 128;
#pragma unroll
                for (int u = 0; u < DW_KC / 4; ++u) {      // 16-byte unit = 4 consecutive k (= rows)
                    uint4 h, l;
                    split_tf32(4 * u + 0 < rows ? src[(4 * u + 0) * NO] : 0.f, h.x, l.x);
                    split_tf32(4 * u + 1 < rows ? src[(4 * u + 1) * NO] : 0.f, h.y, l.y);
                    split_tf32(4 * u + 2 < rows ? src[(4 * u + 2) * NO] : 0.f, h.z, l.z);
                    split_tf32(4 * u + 3 < rows ? src[(4 * u + 3) * NO] : 0.f, h.w, l.w);
                    const uint32_t off = (uint32_t)((u ^ (n & 7)) << 4);
                    *reinterpret_cast<uint4 *>(dhi + off) = h;
                    *reinterpret_cast<uint4 *>(dlo + off) = l;
                }
            }
            fence_proxy_async();                           // generic-proxy stores -> visible to the MMA
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&rawd_empty[s]);
                mbar_arrive(&ab_full[s]);
            }
        }
    }

    else if (warp >= 12) {
        // ================= flush warps: TMEM accumulator -> rounded fp32 adds into accs -> partial =========
        const int q = warp & 3;
        const int ki = q * 32 + lane;                      // TMEM lane == row of dW
        for (int c = 0; c < NO; ++c) accs[c * 128 + ki] = 0.f;     // column-major: conflict-free by lane
        const int64_t n_groups = (n_my + DW_FLUSH - 1) / DW_FLUSH;
        for (int64_t grp = 0; grp < n_groups; ++grp) {
            const int acc = (int)(grp & 1);
            mbar_wait(&d_full[acc], (uint32_t)((grp >> 1) & 1));
            tc_fence_after();
            const uint32_t t_d = tmem_base + ((uint32_t)(q * 32) << 16) + DW_TMEM_D + acc * 128;
            for (int c = 0; c < NO / 16; ++c) {
                uint32_t v[16];
                tmem_ld16(t_d + c * 16, v);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 16; ++k) accs[(c * 16 + k) * 128 + ki] += __uint_as_float(v[k]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[acc]);
        }
        if (ki < KI) {
            float *dst = p.partial + ((int64_t)blockIdx.x * KI + ki) * NO;
            for (int c = 0; c < NO; c += 4)
                *reinterpret_cast<float4 *>(dst + c) = make_float4(accs[c * 128 + ki], accs[(c + 1) * 128 + ki],
                                                                  accs[(c + 2) * 128 + ki], accs[(c + 3) * 128 + ki]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

static size_t dw_smem_bytes(int KI, int NO) {
    return (size_t)DW_STAGES * (2 * DW_KC * NO * 4 + DW_KC * KI * 4 + DW_KC * NO * 4) + (size_t)128 * NO * 4 +
           32 * 8 + 1024;
}

}  // namespace tc

// fixed-order sum of the per-CTA partials: deterministic dW
static __global__ void dw_reduce_partials(const float *__restrict__ partial, float *__restrict__ out, int n_parts,
                                          int64_t n_elem) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_elem) return;
    float s = 0.f;
    for (int c = 0; c < n_parts; ++c) s += partial[(int64_t)c * n_elem + i];
    out[i] = s;
}

bool gemm_dw_tc_eligible(int64_t R, int KI, int NO) {
    if (R < 1) return false;
    if (KI < 16 || KI > 128 || (KI % 4) != 0) return false;          // M = 128 lanes; rows beyond KI are zero
    if (NO < 16 || NO > 128 || (NO % 16) != 0) return false;         // UMMA N for M = 128
    return tc::dw_smem_bytes(KI, NO) <= 227 * 1024;
}

int gemm_dw_tc_max_ctas() { return sm_count(); }

// dw (KI x NO); ws holds sm_count() partials of KI*NO floats
int gemm_dw_tc(const float *p, const float *dy, float *dw, int64_t R, int KI, int NO, float *ws, cudaStream_t st) {
    TMGCN_REQUIRE(((uintptr_t)p % 16 == 0) && ((uintptr_t)dy % 16 == 0) && ((uintptr_t)ws % 16 == 0),
                  "gemm_dw_tc: operands must be 16-byte aligned");
    tc::DwParams q;
    q.p = p;
    q.dy = dy;
    q.partial = ws;
    q.R = R;
    q.KI = KI;
    q.NO = NO;
    const int64_t n_chunks = ceil_div(R, tc::DW_KC);
    int64_t grid = sm_count();
    if (grid > n_chunks) grid = n_chunks;
    q.chunks_per_cta = ceil_div(n_chunks, grid);
    grid = ceil_div(n_chunks, q.chunks_per_cta);
    TMGCN_CUDA(cudaFuncSetAttribute(tc::gemm_dw_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    227 * 1024));
    tc::gemm_dw_tf32x3_kernel<<<(unsigned)grid, tc::DW_THREADS, tc::dw_smem_bytes(KI, NO), st>>>(q);
    if (after_launch("gemm_dw_tf32x3")) return 1;
    const int64_t n_elem = (int64_t)KI * NO;
    dw_reduce_partials<<<(unsigned)ceil_div(n_elem, 256), 256, 0, st>>>(ws, dw, (int)grid, n_elem);
    return after_launch("dw_reduce_partials");
}

}  // namespace tmgcn
