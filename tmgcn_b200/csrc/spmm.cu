// (d) facewise SpMM  y[t] = act( A~[t] . x[t] ), all T slices in one launch.
//
// ref: the per-slice loop `AtXt[k] = t.sparse.mm(At[k], Xt[k])` (ehf:205-207,
// ehf:309-311) and compute_AX (ehf:301-305, ehf:469-473).  The backward
// dX[t] = A~[t]^T . dY[t] is this same kernel on the transposed CSR.
//
// HBM-bound gather: per stored entry one 4*F-byte row of x is fetched.  A group
// of G lanes (G*VEC >= F, VEC-wide vector loads) owns one CSR row, so a gathered
// feature row is one fully coalesced request (512 B at F=128 with G=32,
// VEC=4).  The group loads G (col, val) pairs at once (coalesced) and
// broadcasts them with shuffles; the gathers are issued UNR at a time and bypass
// L1 allocation (no reuse at L1: measured hit rate < 1 %); the kernel is kept at
// <= 48 registers so 5 CTAs (40 warps) per SM hide the DRAM latency of the random
// 512 B requests.  Rows are taken grid-stride by a grid that is a multiple of the
// SM count.
#include <stdlib.h>

#include "common.cuh"

namespace tmgcn {

template <int VEC>
struct VecT;
template <>
struct VecT<4> {
    using T = float4;
};
template <>
struct VecT<2> {
    using T = float2;
};
template <>
struct VecT<1> {
    using T = float;
};

__device__ __forceinline__ void axpy(float4 &a, float s, const float4 &x) {
    a.x = fmaf(s, x.x, a.x);
    a.y = fmaf(s, x.y, a.y);
    a.z = fmaf(s, x.z, a.z);
    a.w = fmaf(s, x.w, a.w);
}
__device__ __forceinline__ void axpy(float2 &a, float s, const float2 &x) {
    a.x = fmaf(s, x.x, a.x);
    a.y = fmaf(s, x.y, a.y);
}
__device__ __forceinline__ void axpy(float &a, float s, const float &x) { a = fmaf(s, x, a); }
__device__ __forceinline__ void vzero(float4 &a) { a = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void vzero(float2 &a) { a = make_float2(0.f, 0.f); }
__device__ __forceinline__ void vzero(float &a) { a = 0.f; }
__device__ __forceinline__ float4 ld_gather(const float4 *p) { return ld_stream_f4(p); }
__device__ __forceinline__ float2 ld_gather(const float2 *p) { return __ldg(p); }
__device__ __forceinline__ float ld_gather(const float *p) { return __ldg(p); }
template <int ACT>
__device__ __forceinline__ void vact(float4 &a) {
    a.x = act_apply<ACT>(a.x);
    a.y = act_apply<ACT>(a.y);
    a.z = act_apply<ACT>(a.z);
    a.w = act_apply<ACT>(a.w);
}
template <int ACT>
__device__ __forceinline__ void vact(float2 &a) {
    a.x = act_apply<ACT>(a.x);
    a.y = act_apply<ACT>(a.y);
}
template <int ACT>
__device__ __forceinline__ void vact(float &a) {
    a = act_apply<ACT>(a);
}

// UNR gathers in flight per warp, MINB resident CTAs per SM, NOALLOC: gathers bypass L1 allocation.
// Tuned on B200 (N = 2M, F = 128, ~20 nnz/row): occupancy beats unroll depth -- (8, 1, false) 4.3 TB/s,
// (8, 4, true) 6.0 TB/s, (4, 5, true) 6.4 TB/s of 6.54 TB/s copy peak; register spills (8, >=5) collapse it.
template <int VEC, int G, int ACT, int UNR = 4, int MINB = 5, bool NOALLOC = true>
__global__ void __launch_bounds__(256, MINB) spmm_rows(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                 const float *__restrict__ val, const float *__restrict__ x,
                                                 float *__restrict__ y, int64_t n_rows, int64_t N, int F) {
    using VT = typename VecT<VEC>::T;
    constexpr int UNROLL = UNR < G ? UNR : G;
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);  // lane inside the group
    const int64_t groups_total = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int Fv = F / VEC;
    // iterate in lock-step per warp so the shuffles stay convergent
    const int64_t n_iter = ceil_div_dev(n_rows, groups_total);
    for (int64_t it = 0; it < n_iter; ++it) {
        const int64_t row = g0 + it * groups_total;
        const bool live = row < n_rows;
        int64_t s = 0, e = 0, xbase = 0;
        if (live) {
            s = rowptr[row];
            e = rowptr[row + 1];
            xbase = (row / N) * N * (int64_t)Fv;  // first vector of this slice of x
        }
        const int len = (int)(e - s);
        const int maxlen = G == 32 ? len : __reduce_max_sync(0xffffffffu, len);
        for (int fb = 0; fb < Fv; fb += G) {  // uniform trip count: shuffles stay convergent
            const int f0 = fb + gl;
            const bool fact = f0 < Fv;
            VT acc;
            vzero(acc);
            const VT *xv = reinterpret_cast<const VT *>(x) + xbase + f0;
            for (int base = 0; base < maxlen; base += G) {
                int c = 0;
                float v = 0.f;
                if (base + gl < len) {
                    c = col[s + base + gl];
                    v = val[s + base + gl];
                }
                const int cnt = min(G, maxlen - base);
                for (int j0 = 0; j0 < cnt; j0 += UNROLL) {
                    VT xs[UNROLL];
                    float vs[UNROLL];
#pragma unroll
                    for (int u = 0; u < UNROLL; ++u) {
                        const int j = j0 + u;
                        const int cj = __shfl_sync(0xffffffffu, c, j & (G - 1), G);
                        vs[u] = __shfl_sync(0xffffffffu, v, j & (G - 1), G);
                        const bool ok = fact && (base + j < len) && (j < cnt);
                        if (ok)
                            xs[u] = NOALLOC ? ld_gather(xv + (int64_t)cj * Fv) : __ldg(xv + (int64_t)cj * Fv);
                        else {
                            vzero(xs[u]);
                            vs[u] = 0.f;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < UNROLL; ++u) axpy(acc, vs[u], xs[u]);
                }
            }
            if (live && fact) {
                vact<ACT>(acc);
                reinterpret_cast<VT *>(y)[row * (int64_t)Fv + f0] = acc;
            }
        }
    }
}

// Skinny operand (W = 4, 6 or 8 floats per row: the 2C-column factor of the low-rank backward for C = 2, 3, 4
// classes): a row gather is one or two small vector loads, so lanes are spread over the NONZEROS of a row
// instead of over features.  LPR lanes share a row and a warp takes 32/LPR consecutive rows per iteration, so
// the dependent chain rowptr -> (col, val) -> gather is paid once per 32/LPR rows (warp-per-row was
// latency-bound: 20 ms for 1.2 G nonzeros); the rows' nonzeros are contiguous, so the (col, val) loads stay
// coalesced.  The x slice (N * 4W bytes) is L2-resident.
template <int W>
struct SkinnyRow {                       // W floats as W/2 float2 (rows are 8-byte aligned for every even W)
    float2 v[W / 2];
};
template <int W>
__device__ __forceinline__ SkinnyRow<W> skinny_load(const float *x, int64_t row) {
    SkinnyRow<W> r;
    if (W % 4 == 0) {
        const float4 *p = reinterpret_cast<const float4 *>(x) + row * (W / 4);
#pragma unroll
        for (int i = 0; i < W / 4; ++i) {
            const float4 q = __ldg(p + i);
            r.v[2 * i] = make_float2(q.x, q.y);
            r.v[2 * i + 1] = make_float2(q.z, q.w);
        }
    } else {
        const float2 *p = reinterpret_cast<const float2 *>(x) + row * (W / 2);
#pragma unroll
        for (int i = 0; i < W / 2; ++i) r.v[i] = __ldg(p + i);
    }
    return r;
}

template <int ACT, int LPR, int W>
__global__ void __launch_bounds__(256, 6) spmm_skinny(const int64_t *__restrict__ rowptr,
                                                      const int32_t *__restrict__ col, const float *__restrict__ val,
                                                      const float *__restrict__ x, float *__restrict__ y,
                                                      int64_t n_rows, int64_t N) {
    constexpr int RPW = 32 / LPR;                      // rows per warp iteration
    constexpr int UN = W <= 4 ? 4 : 2;                 // passes in flight (register budget)
    const int lane = threadIdx.x & 31;
    const int g = lane / LPR, gl = lane % LPR;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_blk = (n_rows + RPW - 1) / RPW;
    for (int64_t blk = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); blk < n_blk; blk += warps_total) {
        const int64_t row = blk * RPW + g;
        const bool live = row < n_rows;
        int64_t s = 0, e = 0, xbase = 0;
        if (live) {
            s = rowptr[row];
            e = rowptr[row + 1];
            xbase = (row / N) * N;
        }
        float acc[W];
#pragma unroll
        for (int i = 0; i < W; ++i) acc[i] = 0.f;
        // UN passes at a time: all (col, val) loads first, then all gathers -- two dependent latencies per
        // UN*LPR nonzeros of a row instead of two per LPR
        for (int64_t k = s + gl; k < e; k += UN * LPR) {
            int c[UN];
            float v[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int64_t kk = k + u * LPR;
                c[u] = kk < e ? col[kk] : -1;
                v[u] = kk < e ? val[kk] : 0.f;
            }
            SkinnyRow<W> xv[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                if (c[u] >= 0) {
                    xv[u] = skinny_load<W>(x, xbase + c[u]);
                } else {
#pragma unroll
                    for (int i = 0; i < W / 2; ++i) xv[u].v[i] = make_float2(0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int i = 0; i < W / 2; ++i) {
                    acc[2 * i] = fmaf(v[u], xv[u].v[i].x, acc[2 * i]);
                    acc[2 * i + 1] = fmaf(v[u], xv[u].v[i].y, acc[2 * i + 1]);
                }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
            for (int i = 0; i < W; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        if (live && gl == 0) {
#pragma unroll
            for (int i = 0; i < W / 2; ++i)
                reinterpret_cast<float2 *>(y)[row * (W / 2) + i] =
                    make_float2(act_apply<ACT>(acc[2 * i]), act_apply<ACT>(acc[2 * i + 1]));
        }
    }
}

template <int W>
static int launch_skinny(const int64_t *rowptr, const int32_t *col, const float *val, const float *x, float *y, int T,
                         int64_t N, int act, cudaStream_t st) {
    // One launch per group of slices whose operand (N * 4W B each) fits comfortably in L2: inside a single
    // grid-stride launch over all T slices the warps drift apart (measured 138 -> 58 Gnnz/s from T = 6 to
    // T = 32) until several operand slices compete for L2 and the small gathers go to DRAM.
    const int64_t slice_bytes = N * 4 * W;
    int64_t per_launch = (int64_t)(l2_bytes() / 3) / (slice_bytes > 0 ? slice_bytes : 1);
    if (per_launch < 1) per_launch = 1;
    for (int64_t t0 = 0; t0 < T; t0 += per_launch) {
        const int64_t rows = (t0 + per_launch < T ? per_launch : T - t0) * N;
        const int64_t *rp = rowptr + t0 * N;
        const float *xs = x + t0 * N * W;
        float *ys = y + t0 * N * W;
        int64_t blocks = ceil_div(rows, 8 * 4);
        const int64_t cap = (int64_t)sm_count() * 6 * 8;
        if (blocks > cap) blocks = cap;
        switch (act) {
            case TMGCN_ACT_NONE: spmm_skinny<TMGCN_ACT_NONE, 8, W><<<(unsigned)blocks, 256, 0, st>>>(rp, col, val, xs, ys, rows, N); break;
            case TMGCN_ACT_RELU: spmm_skinny<TMGCN_ACT_RELU, 8, W><<<(unsigned)blocks, 256, 0, st>>>(rp, col, val, xs, ys, rows, N); break;
            case TMGCN_ACT_LEAKY: spmm_skinny<TMGCN_ACT_LEAKY, 8, W><<<(unsigned)blocks, 256, 0, st>>>(rp, col, val, xs, ys, rows, N); break;
            case TMGCN_ACT_SELU: spmm_skinny<TMGCN_ACT_SELU, 8, W><<<(unsigned)blocks, 256, 0, st>>>(rp, col, val, xs, ys, rows, N); break;
            default: set_error("spmm: unknown activation %d", act); return 1;
        }
        if (after_launch("spmm_skinny")) return 1;
    }
    return 0;
}

template <int VEC, int G>
static int launch_spmm_act(const int64_t *rowptr, const int32_t *col, const float *val, const float *x, float *y,
                           int64_t n_rows, int64_t N, int F, int act, cudaStream_t st) {
    const int threads = 256;
    const int64_t groups_per_block = threads / G;
    int64_t blocks = ceil_div(n_rows, groups_per_block);
    const int64_t cap = (int64_t)sm_count() * 8 * 8;  // 8 resident CTAs/SM x 8 waves, then grid-stride
    if (blocks > cap) blocks = cap;
    const unsigned grid = (unsigned)blocks;
#define TMGCN_LAUNCH(A)                                                                                  \
    spmm_rows<VEC, G, A><<<grid, threads, 0, st>>>(rowptr, col, val, x, y, n_rows, N, F);               \
    break;
    switch (act) {
        case TMGCN_ACT_NONE: TMGCN_LAUNCH(TMGCN_ACT_NONE)
        case TMGCN_ACT_RELU: TMGCN_LAUNCH(TMGCN_ACT_RELU)
        case TMGCN_ACT_LEAKY: TMGCN_LAUNCH(TMGCN_ACT_LEAKY)
        case TMGCN_ACT_SELU: TMGCN_LAUNCH(TMGCN_ACT_SELU)
        default: set_error("spmm: unknown activation %d", act); return 1;
    }
#undef TMGCN_LAUNCH
    return after_launch("spmm_rows");
}

template <int VEC>
static int launch_spmm_g(const int64_t *rowptr, const int32_t *col, const float *val, const float *x, float *y,
                         int64_t n_rows, int64_t N, int F, int act, cudaStream_t st) {
    const int fv = F / VEC;
    if (fv <= 1) return launch_spmm_act<VEC, 1>(rowptr, col, val, x, y, n_rows, N, F, act, st);
    if (fv <= 2) return launch_spmm_act<VEC, 2>(rowptr, col, val, x, y, n_rows, N, F, act, st);
    if (fv <= 4) return launch_spmm_act<VEC, 4>(rowptr, col, val, x, y, n_rows, N, F, act, st);
    if (fv <= 8) return launch_spmm_act<VEC, 8>(rowptr, col, val, x, y, n_rows, N, F, act, st);
    if (fv <= 16) return launch_spmm_act<VEC, 16>(rowptr, col, val, x, y, n_rows, N, F, act, st);
    return launch_spmm_act<VEC, 32>(rowptr, col, val, x, y, n_rows, N, F, act, st);
}

}  // namespace tmgcn

extern "C" int tmgcn_spmm_fwd(const int64_t *rowptr, const int32_t *col, const float *val, const float *x, float *y,
                              int T, int64_t N, int F, int act, void *stream) {
    using namespace tmgcn;
    TMGCN_REQUIRE(T >= 0 && N >= 0 && F >= 1, "spmm: bad sizes T=%d N=%lld F=%d", T, (long long)N, F);
    const int64_t n_rows = (int64_t)T * N;
    if (n_rows == 0) return 0;
    TMGCN_REQUIRE(rowptr && x && y, "spmm: null pointer");
    TMGCN_REQUIRE(x != y, "spmm: in-place operation is not supported");
    cudaStream_t st = (cudaStream_t)stream;
    const bool a16 = ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0);
    const bool a8 = ((uintptr_t)x % 8 == 0) && ((uintptr_t)y % 8 == 0);
    if (F == 4 && a16) return launch_skinny<4>(rowptr, col, val, x, y, T, N, act, st);
    if (F == 6 && a8) return launch_skinny<6>(rowptr, col, val, x, y, T, N, act, st);
    if (F == 8 && a16) return launch_skinny<8>(rowptr, col, val, x, y, T, N, act, st);
    if (F % 4 == 0 && a16) return launch_spmm_g<4>(rowptr, col, val, x, y, n_rows, N, F, act, st);
    if (F % 2 == 0 && a8) return launch_spmm_g<2>(rowptr, col, val, x, y, n_rows, N, F, act, st);
    return launch_spmm_g<1>(rowptr, col, val, x, y, n_rows, N, F, act, st);
}
