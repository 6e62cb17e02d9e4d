// (e) edge-endpoint gather readout and its backward.
//
// ref: flat ids ehf:196-198; gather + concat ehf:228-230 / 351-353 / 491-493;
// classifier ehf:232 / 355 / 495; backward = autograd index_put_(accumulate)
// (SURVEY.md section 8a rows a9-a11).
//
// HBM-bound row gathers.  The classifier (2F x C, C <= 8) is folded in so the (E, 2F)
// concat never round-trips HBM.  The backward scatter-add is made deterministic by an
// incidence CSR over ALL T*N rows (inc_ptr[n_rows+1], perm[2E] = e*2+half grouped by
// endpoint row, built once per edge set) instead of atomics, and it is a single pass:
//   S_h[row, c] = sum over incident (e, h) of dOut[e, c]            (tiny, per row)
//   dY[row, f]  = sum_h sum_c S_h[row, c] * U[hF + f, c]            (written once, zeros included)
//   dU[hF+f, c] = sum_rows Y[row, f] * S_h[row, c]                  (register partials per warp,
//                                                                    block partials, fixed-order sum)
// so Y is read once per touched row (not once per edge) and dY is written exactly once.
#include "common.cuh"

namespace tmgcn {

constexpr int MAXC = 8;

__global__ void flat_ids_kernel(const int64_t *__restrict__ edges, int64_t E, int64_t N, int64_t t_offset,
                                int64_t *__restrict__ src, int64_t *__restrict__ dst) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int64_t t = edges[e] - t_offset;
    src[e] = t * N + edges[E + e];
    dst[e] = t * N + edges[2 * E + e];
}

// z[e] = [ y[src[e]] || y[dst[e]] ]   -- one thread per output float4 / float
template <int VEC>
__global__ void gather_fwd_kernel(const float *__restrict__ y, const int64_t *__restrict__ src,
                                  const int64_t *__restrict__ dst, float *__restrict__ z, int64_t E, int F) {
    const int Fv = F / VEC;
    const int64_t total = E * 2 * Fv;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t e = i / (2 * Fv);
        const int r = (int)(i - e * 2 * Fv);
        const int half = r / Fv, f = r - half * Fv;
        const int64_t row = half ? dst[e] : src[e];
        if (VEC == 4)
            reinterpret_cast<float4 *>(z)[i] = __ldg(reinterpret_cast<const float4 *>(y) + row * Fv + f);
        else
            z[i] = __ldg(y + row * Fv + f);
    }
}

// out[e, c] = sum_f y[src[e], f] * u[f, c] + y[dst[e], f] * u[F + f, c]
// one warp per edge; u staged in shared memory transposed as us[c][2F]
template <int C, int VEC>
__global__ void __launch_bounds__(256) readout_fwd_kernel(const float *__restrict__ y, const int64_t *__restrict__ src,
                                                          const int64_t *__restrict__ dst, const float *__restrict__ u,
                                                          float *__restrict__ out, int64_t E, int F) {
    extern __shared__ float us[];  // [C][2F]
    for (int i = threadIdx.x; i < 2 * F * C; i += blockDim.x) {
        const int f = i / C, c = i % C;
        us[c * 2 * F + f] = u[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); e < E; e += warps_total) {
        const float *ys = y + src[e] * F;
        const float *yd = y + dst[e] * F;
        float acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 0.f;
        if (VEC == 4) {
            const int Fv = F >> 2;
            for (int f4 = lane; f4 < 2 * Fv; f4 += 32) {
                const float4 v = f4 < Fv ? __ldg(reinterpret_cast<const float4 *>(ys) + f4)
                                         : __ldg(reinterpret_cast<const float4 *>(yd) + f4 - Fv);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float4 w = *reinterpret_cast<const float4 *>(us + c * 2 * F + 4 * f4);
                    acc[c] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc[c]))));
                }
            }
        } else {
            for (int f = lane; f < 2 * F; f += 32) {
                const float v = f < F ? __ldg(ys + f) : __ldg(yd + f - F);
#pragma unroll
                for (int c = 0; c < C; ++c) acc[c] = fmaf(v, us[c * 2 * F + f], acc[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) out[e * C + c] = acc[c];
        }
    }
}

// dy[row, f] = sum over incident (e, half): dz[e, half*F + f]; every row written (zeros included)
__global__ void __launch_bounds__(256) gather_bwd_kernel(const float *__restrict__ dz,
                                                         const int64_t *__restrict__ inc_ptr,
                                                         const int64_t *__restrict__ perm, int64_t n_rows,
                                                         float *__restrict__ dy, int F) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < n_rows; row += warps_total) {
        const int64_t s = inc_ptr[row], e = inc_ptr[row + 1];
        for (int f = lane; f < F; f += 32) {
            float acc = 0.f;
            for (int64_t q = s; q < e; ++q) {
                const int64_t code = perm[q];
                acc += __ldg(dz + (code >> 1) * 2 * F + (code & 1) * F + f);
            }
            dy[row * F + f] = acc;
        }
    }
}

// fused classifier backward, see the header comment.  Lane owns features f = VEC*(lane + 32 k) .. +VEC, k < NCH.
template <int VEC, int NCH, int CM>
__global__ void __launch_bounds__(256) readout_bwd_kernel(const float *__restrict__ y, const float *__restrict__ u,
                                                          const float *__restrict__ dout,
                                                          const int64_t *__restrict__ inc_ptr,
                                                          const int64_t *__restrict__ perm, int64_t n_rows,
                                                          float *__restrict__ dy, float *__restrict__ du_partial,
                                                          int F, int C) {
    extern __shared__ float sm[];          // us[2F*C] then block reduction scratch red[8 warps][2F*C]
    float *us = sm;
    float *red = sm + 2 * F * C;
    for (int i = threadIdx.x; i < 2 * F * C; i += blockDim.x) us[i] = u[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float acc[2][NCH][VEC][CM];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < NCH; ++k)
#pragma unroll
            for (int v = 0; v < VEC; ++v)
#pragma unroll
                for (int c = 0; c < CM; ++c) acc[h][k][v][c] = 0.f;

    // A warp takes RB consecutive rows per iteration so the dependent chain inc_ptr -> perm -> dOut is paid
    // once per RB rows: lanes load the RB+1 row pointers and then one incidence each (coalesced), and the
    // per-row class sums are assembled with shuffles.
    constexpr int RB = 8;
    const int64_t n_blocks = (n_rows + RB - 1) / RB;
    for (int64_t blk = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); blk < n_blocks; blk += warps_total) {
        const int64_t row0 = blk * RB;
        const int64_t ipl = inc_ptr[min(row0 + (int64_t)min(lane, RB), n_rows)];
        const int64_t s0 = __shfl_sync(0xffffffffu, ipl, 0);
        const int64_t e0 = __shfl_sync(0xffffffffu, ipl, RB);
        const int n_inc = (int)(e0 - s0);
        // incidence owned by this lane (first 32 of the block; the rare longer tail is walked serially below)
        int myh = 0;
        float myd[CM];
#pragma unroll
        for (int c = 0; c < CM; ++c) myd[c] = 0.f;
        if (lane < n_inc) {
            const int64_t code = perm[s0 + lane];
            myh = (int)(code & 1);
            const float *d = dout + (code >> 1) * C;
#pragma unroll
            for (int c = 0; c < CM; ++c)
                if (c < C) myd[c] = __ldg(d + c);
        }
#pragma unroll 1
        for (int r = 0; r < RB; ++r) {
            const int64_t row = row0 + r;
            if (row >= n_rows) break;
            const int lo = (int)(__shfl_sync(0xffffffffu, ipl, r) - s0);
            const int hi = (int)(__shfl_sync(0xffffffffu, ipl, r + 1) - s0);
            float S[2][CM];
#pragma unroll
            for (int c = 0; c < CM; ++c) S[0][c] = S[1][c] = 0.f;
            for (int j = lo; j < hi; ++j) {
                if (j < 32) {
                    const int h = __shfl_sync(0xffffffffu, myh, j);
#pragma unroll
                    for (int c = 0; c < CM; ++c) {
                        const float v = __shfl_sync(0xffffffffu, myd[c], j);
                        if (h) S[1][c] += v; else S[0][c] += v;
                    }
                } else {                       // hub rows: uniform loads
                    const int64_t code = perm[s0 + j];
                    const float *d = dout + (code >> 1) * C;
                    const int h = (int)(code & 1);
#pragma unroll
                    for (int c = 0; c < CM; ++c)
                        if (c < C) {
                            const float v = __ldg(d + c);
                            if (h) S[1][c] += v; else S[0][c] += v;
                        }
                }
            }
            const bool touched = hi > lo;
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int f0 = VEC * (lane + 32 * k);
                if (f0 < F) {
                    float o[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; ++v) o[v] = 0.f;
                    if (touched) {
                        float yv[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) yv[v] = 0.f;
                        if (du_partial) {
                            if (VEC == 4) {
                                const float4 t4 = __ldg(reinterpret_cast<const float4 *>(y + row * F + f0));
                                yv[0] = t4.x; yv[1 % VEC] = t4.y; yv[2 % VEC] = t4.z; yv[3 % VEC] = t4.w;
                            } else {
                                yv[0] = __ldg(y + row * F + f0);
                            }
                        }
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
#pragma unroll
                            for (int c = 0; c < CM; ++c)
                                if (c < C) {
                                    o[v] = fmaf(S[0][c], us[(f0 + v) * C + c], o[v]);
                                    o[v] = fmaf(S[1][c], us[(F + f0 + v) * C + c], o[v]);
                                    if (du_partial) {
                                        acc[0][k][v][c] = fmaf(yv[v], S[0][c], acc[0][k][v][c]);
                                        acc[1][k][v][c] = fmaf(yv[v], S[1][c], acc[1][k][v][c]);
                                    }
                                }
                        }
                    }
                    if (dy) {
                        if (VEC == 4)
                            st_stream_f4(reinterpret_cast<float4 *>(dy + row * F + f0),
                                         make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]));
                        else
                            dy[row * F + f0] = o[0];
                    }
                }
            }
        }
    }
    if (!du_partial) return;
    // block reduction in fixed warp order, then one partial per block
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < NCH; ++k)
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const int f = VEC * (lane + 32 * k) + v;
                if (f < F) {
#pragma unroll
                    for (int c = 0; c < CM; ++c)
                        if (c < C) red[warp * 2 * F * C + (h * F + f) * C + c] = acc[h][k][v][c];
                }
            }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < 2 * F * C; i += blockDim.x) {
        float sum = 0.f;
        for (int w = 0; w < nw; ++w) sum += red[w * 2 * F * C + i];
        du_partial[(int64_t)blockIdx.x * 2 * F * C + i] = sum;
    }
}

__global__ void reduce_partials_edge(const float *__restrict__ partial, float *__restrict__ out, int n_chunks,
                                     int n_elem) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_elem) return;
    float s = 0.f;
    for (int c = 0; c < n_chunks; ++c) s += partial[(int64_t)c * n_elem + i];
    out[i] = s;
}

static int du_blocks() { return sm_count() * 4; }

static int warp_grid(int64_t n_warps, int cap_mult = 32) {
    int64_t blocks = ceil_div(n_warps, 8);
    const int64_t cap = (int64_t)sm_count() * cap_mult;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace tmgcn

using namespace tmgcn;

extern "C" {

int tmgcn_flat_edge_ids(const int64_t *edges, int64_t E, int64_t N, int64_t t_offset, int64_t *src, int64_t *dst,
                        void *stream) {
    TMGCN_REQUIRE(E >= 0 && N >= 0, "flat_edge_ids: bad sizes");
    if (E == 0) return 0;
    TMGCN_REQUIRE(edges && src && dst, "flat_edge_ids: null pointer");
    flat_ids_kernel<<<(unsigned)ceil_div(E, 256), 256, 0, (cudaStream_t)stream>>>(edges, E, N, t_offset, src, dst);
    return after_launch("flat_ids");
}

int tmgcn_edge_gather_fwd(const float *y, const int64_t *src, const int64_t *dst, float *z, int64_t E, int F,
                          void *stream) {
    TMGCN_REQUIRE(E >= 0 && F >= 1, "edge_gather_fwd: bad sizes");
    if (E == 0) return 0;
    TMGCN_REQUIRE(y && src && dst && z, "edge_gather_fwd: null pointer");
    const bool v4 = F % 4 == 0 && (uintptr_t)y % 16 == 0 && (uintptr_t)z % 16 == 0;
    const int64_t total = E * 2 * (v4 ? F / 4 : F);
    int64_t blocks = ceil_div(total, 256);
    if (blocks > (int64_t)sm_count() * 32) blocks = (int64_t)sm_count() * 32;
    if (v4)
        gather_fwd_kernel<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(y, src, dst, z, E, F);
    else
        gather_fwd_kernel<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(y, src, dst, z, E, F);
    return after_launch("gather_fwd");
}

int tmgcn_edge_readout_fwd(const float *y, const int64_t *src, const int64_t *dst, const float *u, float *out,
                           int64_t E, int F, int C, void *stream) {
    TMGCN_REQUIRE(E >= 0 && F >= 1, "edge_readout_fwd: bad sizes");
    TMGCN_REQUIRE(C >= 1 && C <= MAXC, "edge_readout_fwd: C=%d outside [1, %d]", C, MAXC);
    if (E == 0) return 0;
    TMGCN_REQUIRE(y && src && dst && u && out, "edge_readout_fwd: null pointer");
    const size_t smem = (size_t)2 * F * C * sizeof(float);
    TMGCN_REQUIRE(smem <= 48 * 1024, "edge_readout_fwd: 2*F*C too large");
    const int grid = warp_grid(E);
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = F % 4 == 0 && (uintptr_t)y % 16 == 0;
#define TMGCN_RC(CC)                                                                       \
    case CC:                                                                               \
        if (v4)                                                                            \
            readout_fwd_kernel<CC, 4><<<grid, 256, smem, st>>>(y, src, dst, u, out, E, F); \
        else                                                                               \
            readout_fwd_kernel<CC, 1><<<grid, 256, smem, st>>>(y, src, dst, u, out, E, F); \
        break;
    switch (C) {
        TMGCN_RC(1) TMGCN_RC(2) TMGCN_RC(3) TMGCN_RC(4) TMGCN_RC(5) TMGCN_RC(6) TMGCN_RC(7) TMGCN_RC(8)
    }
#undef TMGCN_RC
    return after_launch("readout_fwd");
}

int tmgcn_edge_gather_bwd(const float *dz, const int64_t *inc_ptr, const int64_t *perm, float *dy, int64_t n_rows,
                          int F, void *stream) {
    TMGCN_REQUIRE(n_rows >= 0 && F >= 1, "edge_gather_bwd: bad sizes");
    if (n_rows == 0) return 0;
    TMGCN_REQUIRE(inc_ptr && dy, "edge_gather_bwd: null pointer");
    gather_bwd_kernel<<<warp_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(dz, inc_ptr, perm, n_rows, dy, F);
    return after_launch("gather_bwd");
}

size_t tmgcn_edge_du_ws_bytes(int F, int C) { return (size_t)du_blocks() * 2 * F * C * sizeof(float); }

int tmgcn_edge_readout_bwd(const float *y, const float *u, const float *dout, const int64_t *inc_ptr,
                           const int64_t *perm, float *dy, float *du, int64_t n_rows, int F, int C, void *du_ws,
                           void *stream) {
    TMGCN_REQUIRE(n_rows >= 0 && F >= 1, "edge_readout_bwd: bad sizes");
    TMGCN_REQUIRE(C >= 1 && C <= MAXC, "edge_readout_bwd: C=%d outside [1, %d]", C, MAXC);
    cudaStream_t st = (cudaStream_t)stream;
    const int n_u = 2 * F * C;
    if (n_rows == 0) {
        if (du) TMGCN_CUDA(cudaMemsetAsync(du, 0, (size_t)n_u * sizeof(float), st));
        return 0;
    }
    TMGCN_REQUIRE(u && dout && inc_ptr && (dy || du), "edge_readout_bwd: null pointer");
    TMGCN_REQUIRE(!du || (y && du_ws), "edge_readout_bwd: y and du_ws are required for dU");
    const bool v4 = F % 4 == 0 && (!dy || (uintptr_t)dy % 16 == 0) && (!y || (uintptr_t)y % 16 == 0);
    const int per_lane = v4 ? 4 : 1;
    const int nch = (F + 32 * per_lane - 1) / (32 * per_lane);
    TMGCN_REQUIRE(nch <= 4, "edge_readout_bwd: F=%d too large", F);
    const size_t smem = (size_t)n_u * sizeof(float) * (du ? 9 : 1);
    TMGCN_REQUIRE(smem <= 200 * 1024, "edge_readout_bwd: 2*F*C too large");
    int grid = warp_grid(n_rows, 4);
    if (grid > du_blocks()) grid = du_blocks();
    float *partial = du ? (float *)du_ws : nullptr;
#define TMGCN_LAUNCH(V, K, CMX)                                                                                   \
    {                                                                                                             \
        auto kern = readout_bwd_kernel<V, K, CMX>;                                                                \
        if (smem > 48 * 1024)                                                                                     \
            TMGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
        kern<<<grid, 256, smem, st>>>(y, u, dout, inc_ptr, perm, n_rows, dy, partial, F, C);                      \
    }
#define TMGCN_BY_C(V, K)                                  \
    if (C <= 2) TMGCN_LAUNCH(V, K, 2)                     \
    else if (C <= 4) TMGCN_LAUNCH(V, K, 4)                \
    else TMGCN_LAUNCH(V, K, 8)
    if (v4) {
        if (nch == 1) { TMGCN_BY_C(4, 1) } else { TMGCN_BY_C(4, 4) }
    } else {
        if (nch == 1) { TMGCN_BY_C(1, 1) } else { TMGCN_BY_C(1, 4) }
    }
#undef TMGCN_BY_C
#undef TMGCN_LAUNCH
    if (after_launch("readout_bwd")) return 1;
    if (du) {
        reduce_partials_edge<<<(unsigned)ceil_div(n_u, 256), 256, 0, st>>>(partial, du, grid, n_u);
        if (after_launch("reduce_partials_edge")) return 1;
    }
    return 0;
}
}
