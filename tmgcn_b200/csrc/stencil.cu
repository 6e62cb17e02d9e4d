// (b) dense M-transform  X~ = X x_3 M  as a streamed time stencil.
//
// ref: Xt = t.matmul(M, X.reshape(T, -1))  (ehf:204, ehf:308, ehf:346) -- the
// reference multiplies by the dense T x T matrix; M is banded, so every output
// slice is a weighted sum of the b previous input slices.
//
// Layout: a thread owns one float4 column of the (T, N*F) matrix and marches
// through time keeping the last B-1+C inputs in a register ring, so each input
// element is read from HBM exactly once and each output written once
// (8*N*F bytes per slice, the algorithmic minimum).  Adjacent threads own
// adjacent float4s: every load/store is a fully coalesced 512 B per warp.  The
// band weights sit in shared memory (broadcast reads).  The time loop is unrolled
// over the ring period R so every ring index is a compile-time constant, and the
// C loads of a chunk are issued before they are consumed (C independent 16 B
// loads in flight per thread).
#include "common.cuh"

namespace tmgcn {

template <int V>
struct Vec;
template <>
struct Vec<4> {
    using T = float4;
};
template <>
struct Vec<1> {
    using T = float;
};

__device__ __forceinline__ void fma_v(float4 &a, float w, const float4 &x) {
    a.x = fmaf(w, x.x, a.x);
    a.y = fmaf(w, x.y, a.y);
    a.z = fmaf(w, x.z, a.z);
    a.w = fmaf(w, x.w, a.w);
}
__device__ __forceinline__ void fma_v(float &a, float w, const float &x) { a = fmaf(w, x, a); }
__device__ __forceinline__ void zero_v(float4 &a) { a = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void zero_v(float &a) { a = 0.f; }
__device__ __forceinline__ float4 ld_v(const float4 *p) { return ld_stream_f4(p); }
__device__ __forceinline__ float ld_v(const float *p) { return __ldg(p); }
__device__ __forceinline__ void st_v(float4 *p, const float4 &v) { st_stream_f4(p, v); }
__device__ __forceinline__ void st_v(float *p, const float &v) { *p = v; }

constexpr int STENCIL_C = 8;  // loads in flight per thread
constexpr int ring_size(int B) { return STENCIL_C * ((B - 1 + STENCIL_C + STENCIL_C - 1) / STENCIL_C); }

// Shared-memory weight table with B rows of zero padding on both sides so the
// kernels index it without bounds checks: sw[(t + B) * B + i] = M[t, t-i].
template <int B>
__device__ __forceinline__ void load_weights(float *sw, const float *__restrict__ band_w, int T_out, int b) {
    const int total = (T_out + 2 * B) * B;
    for (int k = threadIdx.x; k < total; k += blockDim.x) {
        const int t = k / B - B, i = k % B;
        sw[k] = (t >= 0 && t < T_out && i < b) ? band_w[t * b + i] : 0.f;
    }
    __syncthreads();
}

// FWD : y[t]  = sum_i w[t][i]       * x[halo + t - i]          (REVERSE = false)
// BWD : gx[s] = sum_i w[s-halo+i][i] * gy[s - halo + i]         (REVERSE = true)
// Both march over the T_in = halo + T_out slices of the LONG tensor (x / gx); the
// forward walks it upwards reading x and writing y, the backward walks it
// downwards reading gy and writing gx.
template <int B, int V, bool REVERSE>
__global__ void __launch_bounds__(256) stencil_kernel(const float *__restrict__ src_halo,
                                                      const float *__restrict__ src, float *__restrict__ dst,
                                                      int T_out, int halo, int64_t n_vec,
                                                      const float *__restrict__ band_w, int b, int s_begin, int s_end,
                                                      int acc_begin) {
    using VT = typename Vec<V>::T;
    constexpr int C = STENCIL_C, R = ring_size(B);
    extern __shared__ float sw[];
    load_weights<B>(sw, band_w, T_out, b);
    const int T_in = halo + T_out;
    // one column per thread; a capped grid (the peer-memory boundary launch) strides over the columns
    for (int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pos < n_vec;
         pos += (int64_t)gridDim.x * blockDim.x) {
    // forward: the first `halo` input slices come from src_halo (which may be a peer GPU's memory mapped
    // over NVLink: the halo exchange is then fused into this kernel), the rest from src
    const VT *in_halo = reinterpret_cast<const VT *>(src_halo) + pos;
    const VT *in = reinterpret_cast<const VT *>(src) + pos - (REVERSE ? 0 : (int64_t)halo * n_vec);
    VT *out = reinterpret_cast<VT *>(dst) + pos;
    VT ring[R];
#pragma unroll
    for (int k = 0; k < R; ++k) zero_v(ring[k]);

    for (int u0 = 0; u0 < T_in; u0 += R) {
#pragma unroll
        for (int q = 0; q < R; q += C) {
            if (u0 + q < T_in) {
#pragma unroll
                for (int k = 0; k < C; ++k) {
                    const int u = u0 + q + k;  // step number
                    if (!REVERSE) {
                        // entering element: x[u]
                        if (u < T_in) ring[q + k] = ld_v((u < halo ? in_halo : in) + (int64_t)u * n_vec);
                    } else {
                        // entering element: gy[t0], t0 = (T_in-1-u) - halo  (absent for the halo slices)
                        const int t0 = T_in - 1 - u - halo;
                        if (u < T_in) {
                            if (t0 >= 0)
                                ring[q + k] = ld_v(in + (int64_t)t0 * n_vec);
                            else
                                zero_v(ring[q + k]);
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < C; ++k) {
                    const int u = u0 + q + k;
                    if (u < T_in) {
                        VT acc;
                        zero_v(acc);
                        if (!REVERSE) {
                            const int t = u - halo;
                            if (t >= 0) {
                                const float *w = sw + (t + B) * B;
#pragma unroll
                                for (int i = 0; i < B; ++i) fma_v(acc, w[i], ring[(q + k - i + 2 * R) % R]);
                                st_v(out + (int64_t)t * n_vec, acc);
                            }
                        } else {
                            const int s = T_in - 1 - u;
                            const int t0 = s - halo;  // >= -halo >= -(B-1)
#pragma unroll
                            for (int i = 0; i < B; ++i)
                                fma_v(acc, sw[(t0 + i + B) * B + i], ring[(q + k - i + 2 * R) % R]);
                            if (s >= s_begin && s < s_end) {
                                // slices >= acc_begin already hold what the successor rank owes them (received
                                // straight into this tensor): add instead of overwrite
                                if (s >= acc_begin) fma_v(acc, 1.f, ld_v(out + (int64_t)s * n_vec));
                                st_v(out + (int64_t)s * n_vec, acc);
                            }
                        }
                    }
                }
            }
        }
    }
    }
}

// ---- inverse M-transform  Y = inv(M) x_3 Z  as a banded substitution (ref: ehf:183-184, 223-224) -----------
// The reference multiplies by the dense T x T matrix inv(M).  M is banded lower triangular, so inv(M) x_3 Z is
// the solution of M x_3 Y = Z:  y[t] = (z[t] - sum_{i>=1} M[t,t-i] y[t-i]) / M[t,t]  -- the same time march
// as the forward stencil with the ring holding previous OUTPUTS (an IIR filter).  The recurrence runs in
// fp64 registers (loads / stores fp32) so the feedback does not amplify rounding over hundreds of slices.
// TRANSPOSED solves M^T x_3 g_in = g_out (the adjoint), marching downwards:
//   g_in[s] = (g_out[s] - sum_{i>=1} M[s+i, s] g_in[s+i]) / M[s,s].
// Time-sharded form: the rank's block continues a recurrence that started on another rank, so the ring starts
// from `h` halo rows -- the predecessor's last h OUTPUTS (forward) or the successor's first h outputs
// (transposed) -- instead of zeros, and the weight table of the transposed solve reaches h rows past the block
// (M[s+i, s] with s+i owned by the successor).  Rows are `ld` floats apart so a caller can pipeline the
// cross-rank scan over column chunks (sharding.solve_pipelined).
template <int B, bool TRANSPOSED>
__global__ void __launch_bounds__(256) solve_kernel(const float *__restrict__ src, float *__restrict__ dst,
                                                    const float *__restrict__ halo, int T, int h, int64_t n,
                                                    int64_t ld, int64_t ld_halo, const float *__restrict__ band_w,
                                                    int b) {
    extern __shared__ float sw[];
    load_weights<B>(sw, band_w, TRANSPOSED ? T + h : T, b);
    const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n) return;
    double ring[B];
#pragma unroll
    for (int k = 0; k < B; ++k) ring[k] = 0.0;
    // ring[j] holds the output of the step u with u = j (mod B); step -i (i = 1..h) is the i-th halo row
    // counted from the block: forward y[-i] = halo[h - i], transposed g[T - 1 + i] = halo[i - 1]
#pragma unroll
    for (int i = 1; i < B; ++i)
        if (i <= h) ring[B - i] = (double)__ldg(halo + (int64_t)(TRANSPOSED ? i - 1 : h - i) * ld_halo + pos);
    for (int u0 = 0; u0 < T; u0 += B) {
#pragma unroll
        for (int k = 0; k < B; ++k) {
            const int u = u0 + k;
            if (u < T) {
                const int t = TRANSPOSED ? T - 1 - u : u;
                double acc = (double)__ldg(src + (int64_t)t * ld + pos);
#pragma unroll
                for (int i = 1; i < B; ++i) {
                    // forward: M[t, t-i] = sw[(t+B)*B + i];  transposed: M[t+i, t] = sw[(t+i+B)*B + i]
                    const float w = TRANSPOSED ? sw[(t + i + B) * B + i] : sw[(t + B) * B + i];
                    acc -= (double)w * ring[(k - i + 2 * B) % B];
                }
                acc /= (double)sw[(t + B) * B];
                ring[k] = acc;
                dst[(int64_t)t * ld + pos] = (float)acc;
            }
        }
    }
}

template <bool TRANSPOSED>
static int solve_entry(const float *z, float *y, const float *halo, int T, int h, int64_t n, int64_t ld,
                       int64_t ld_halo, const float *band_w, int b, void *stream) {
    TMGCN_REQUIRE(T >= 0 && n >= 0 && h >= 0, "mtransform_dense_solve: negative size");
    TMGCN_REQUIRE(b >= 1 && b <= 32, "mtransform_dense_solve: band width b=%d outside [1, 32]", b);
    TMGCN_REQUIRE(h <= b - 1, "mtransform_dense_solve: halo=%d exceeds b-1=%d", h, b - 1);
    TMGCN_REQUIRE(ld >= n && (h == 0 || ld_halo >= n), "mtransform_dense_solve: row stride smaller than the row");
    if (n == 0 || T == 0) return 0;
    TMGCN_REQUIRE(z && y && band_w && (h == 0 || halo), "mtransform_dense_solve: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int T_tab = TRANSPOSED ? T + h : T;
#define TMGCN_SOLVE(BB)                                                                                     \
    if (b <= BB) {                                                                                          \
        const size_t smem = (size_t)(T_tab + 2 * BB) * BB * sizeof(float);                                  \
        TMGCN_REQUIRE(smem <= 200 * 1024, "mtransform_dense_solve: T=%d too large for the weight table", T); \
        auto kern = solve_kernel<BB, TRANSPOSED>;                                                           \
        if (smem > 48 * 1024)                                                                               \
            TMGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<(unsigned)ceil_div(n, 256), 256, smem, st>>>(z, y, halo, T, h, n, ld, ld_halo, band_w, b);   \
        return after_launch(TRANSPOSED ? "solve_bwd" : "solve_fwd");                                        \
    }
    TMGCN_SOLVE(2)
    TMGCN_SOLVE(4)
    TMGCN_SOLVE(8)
    TMGCN_SOLVE(12)
    TMGCN_SOLVE(16)
    TMGCN_SOLVE(20)
    TMGCN_SOLVE(32)
#undef TMGCN_SOLVE
    return 1;
}

template <int B, int V, bool REVERSE>
static int launch_stencil(const float *src_halo, const float *src, float *dst, int T_out, int halo, int64_t NF,
                          const float *band_w, int b, int s_begin, int s_end, int max_ctas, int acc_begin,
                          cudaStream_t st) {
    const int64_t n_vec = NF / V;
    const size_t smem = (size_t)(T_out + 2 * B) * B * sizeof(float);
    TMGCN_REQUIRE(smem <= 200 * 1024, "mtransform_dense: T_out=%d too large for the weight table (b=%d)", T_out, b);
    auto kern = stencil_kernel<B, V, REVERSE>;
    if (smem > 48 * 1024) TMGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // max_ctas > 0: a small persistent grid of 128-thread CTAs (the boundary launch that reads a peer GPU's
    // memory: it is NVLink-bound, and a full grid of stalled CTAs would hold the registers the concurrently
    // running interior stencil / SpMM need)
    const int threads = max_ctas > 0 ? 128 : 256;
    int64_t grid = ceil_div(n_vec, threads);
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    kern<<<(unsigned)grid, threads, smem, st>>>(src_halo, src, dst, T_out, halo, n_vec, band_w, b, s_begin, s_end,
                                                acc_begin);
    return after_launch(REVERSE ? "stencil_bwd" : "stencil_fwd");
}

template <int V, bool REVERSE>
static int dispatch_b(const float *src_halo, const float *src, float *dst, int T_out, int halo, int64_t NF,
                      const float *band_w, int b, int s_begin, int s_end, int max_ctas, int acc_begin,
                      cudaStream_t st) {
#define TMGCN_CASE(BB)                                                                                             \
    if (b <= BB)                                                                                                   \
        return launch_stencil<BB, V, REVERSE>(src_halo, src, dst, T_out, halo, NF, band_w, b, s_begin, s_end,     \
                                              max_ctas, acc_begin, st);
    TMGCN_CASE(1)
    TMGCN_CASE(2)
    TMGCN_CASE(4)
    TMGCN_CASE(6)
    TMGCN_CASE(8)
    TMGCN_CASE(10)
    TMGCN_CASE(12)
    TMGCN_CASE(16)
    TMGCN_CASE(20)
    TMGCN_CASE(24)
    TMGCN_CASE(32)
#undef TMGCN_CASE
    set_error("mtransform_dense: band width b=%d > 32 unsupported", b);
    return 1;
}

template <bool REVERSE>
static int stencil_entry(const float *src_halo, const float *src, float *dst, int T_out, int halo, int64_t NF,
                         const float *band_w, int b, int s_begin, int s_end, void *stream, int max_ctas = 0,
                         int acc_begin = 0x7fffffff) {
    TMGCN_REQUIRE(T_out >= 0 && NF >= 0 && halo >= 0, "mtransform_dense: negative size");
    TMGCN_REQUIRE(b >= 1 && b <= 32, "mtransform_dense: band width b=%d outside [1, 32]", b);
    TMGCN_REQUIRE(halo <= b - 1, "mtransform_dense: halo=%d exceeds b-1=%d", halo, b - 1);
    if (NF == 0 || halo + T_out == 0) return 0;
    TMGCN_REQUIRE(src && dst && band_w, "mtransform_dense: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec4 = (NF % 4 == 0) && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0) &&
                      ((uintptr_t)src_halo % 16 == 0);
    if (vec4)
        return dispatch_b<4, REVERSE>(src_halo, src, dst, T_out, halo, NF, band_w, b, s_begin, s_end, max_ctas,
                                      acc_begin, st);
    return dispatch_b<1, REVERSE>(src_halo, src, dst, T_out, halo, NF, band_w, b, s_begin, s_end, max_ctas, acc_begin,
                                  st);
}

}  // namespace tmgcn

extern "C" {
int tmgcn_mtransform_dense_fwd(const float *x_in, float *x_out, int T_out, int halo, int64_t NF,
                               const float *band_w, int b, void *stream) {
    // contiguous [halo | own] input: the own part starts `halo` slices in
    return tmgcn::stencil_entry<false>(x_in, x_in ? x_in + (int64_t)halo * NF : x_in, x_out, T_out, halo, NF, band_w, b,
                                       0, halo + T_out, stream);
}
int tmgcn_mtransform_dense_fwd_split(const float *x_halo, const float *x_own, float *x_out, int T_out, int halo,
                                     int64_t NF, const float *band_w, int b, int max_ctas, void *stream) {
    if (halo > 0 && !x_halo) {
        tmgcn::set_error("mtransform_dense_fwd_split: null halo pointer");
        return 1;
    }
    if (max_ctas < 0) {
        tmgcn::set_error("mtransform_dense_fwd_split: negative max_ctas");
        return 1;
    }
    return tmgcn::stencil_entry<false>(halo > 0 ? x_halo : x_own, x_own, x_out, T_out, halo, NF, band_w, b, 0,
                                       halo + T_out, stream, max_ctas);
}
int tmgcn_mtransform_dense_bwd(const float *g_out, float *g_in, int T_out, int halo, int64_t NF,
                               const float *band_w, int b, void *stream) {
    return tmgcn::stencil_entry<true>(g_out, g_out, g_in, T_out, halo, NF, band_w, b, 0, halo + T_out, stream);
}
int tmgcn_mtransform_dense_solve_fwd(const float *z, float *y, int T, int64_t NF, const float *band_w, int b,
                                     void *stream) {
    return tmgcn::solve_entry<false>(z, y, nullptr, T, 0, NF, NF, NF, band_w, b, stream);
}
int tmgcn_mtransform_dense_solve_bwd(const float *g_y, float *g_z, int T, int64_t NF, const float *band_w, int b,
                                     void *stream) {
    return tmgcn::solve_entry<true>(g_y, g_z, nullptr, T, 0, NF, NF, NF, band_w, b, stream);
}
int tmgcn_mtransform_dense_solve_part(const float *src, float *dst, const float *halo, int T, int h, int64_t n,
                                      int64_t ld, int64_t ld_halo, const float *band_w, int b, int transposed,
                                      void *stream) {
    if (transposed) return tmgcn::solve_entry<true>(src, dst, halo, T, h, n, ld, ld_halo, band_w, b, stream);
    return tmgcn::solve_entry<false>(src, dst, halo, T, h, n, ld, ld_halo, band_w, b, stream);
}
int tmgcn_mtransform_dense_bwd_range(const float *g_out, float *g_in, int T_out, int halo, int64_t NF,
                                     const float *band_w, int b, int s_begin, int s_end, int acc_begin, void *stream) {
    if (s_begin < 0 || s_end > halo + T_out || s_begin > s_end) {
        tmgcn::set_error("mtransform_dense_bwd_range: bad slice range [%d, %d)", s_begin, s_end);
        return 1;
    }
    if (acc_begin < 0) acc_begin = 0x7fffffff;
    return tmgcn::stencil_entry<true>(g_out, g_out, g_in, T_out, halo, NF, band_w, b, s_begin, s_end, stream, 0,
                                      acc_begin);
}
}
