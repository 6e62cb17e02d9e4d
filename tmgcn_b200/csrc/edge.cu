// (e) edge-endpoint gather readout and its backward.
//
// ref: flat ids ehf:196-198; gather + concat ehf:228-230 / 351-353 / 491-493;
// classifier ehf:232 / 355 / 495; backward = autograd index_put_(accumulate)
// (SURVEY.md section 8a rows a9-a11).
//
// HBM-bound row gathers: a group of G lanes (G*VEC >= F) moves one 4*F-byte
// endpoint row with vector loads.  The classifier (2F x C, C <= 8) is folded in
// so the (E, 2F) concat never round-trips HBM; the backward scatter-add is made
// deterministic by an incidence list (edges grouped by touched row, built once
// per edge set) instead of atomics.
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
template <int C>
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
        for (int f = lane; f < 2 * F; f += 32) {
            const float v = f < F ? __ldg(ys + f) : __ldg(yd + f - F);
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = fmaf(v, us[c * 2 * F + f], acc[c]);
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

// dy[row, f] = sum over incident (e, half): dz[e, half*F + f]      (CLASSIFY = false)
//            = sum over incident (e, half): sum_c dout[e, c] * u[half*F + f, c]   (CLASSIFY = true)
// one warp per touched row, incidences visited in list order (deterministic).
template <bool CLASSIFY>
__global__ void __launch_bounds__(256) scatter_rows_kernel(const float *__restrict__ g, const float *__restrict__ u,
                                                           const int64_t *__restrict__ row_ids,
                                                           const int64_t *__restrict__ seg_ptr,
                                                           const int64_t *__restrict__ perm, int64_t n_touched,
                                                           float *__restrict__ dy, int F, int C) {
    extern __shared__ float us[];  // [2F][C] as given
    if (CLASSIFY) {
        for (int i = threadIdx.x; i < 2 * F * C; i += blockDim.x) us[i] = u[i];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); k < n_touched; k += warps_total) {
        const int64_t row = row_ids[k];
        const int64_t s = seg_ptr[k], e = seg_ptr[k + 1];
        for (int f = lane; f < F; f += 32) {
            float acc = 0.f;
            for (int64_t q = s; q < e; ++q) {
                const int64_t code = perm[q];
                const int64_t edge = code >> 1;
                const int half = (int)(code & 1);
                if (CLASSIFY) {
                    const float *d = g + edge * C;
                    const float *uu = us + (half * F + f) * C;
                    float t = 0.f;
                    for (int c = 0; c < C; ++c) t = fmaf(__ldg(d + c), uu[c], t);
                    acc += t;
                } else {
                    acc += __ldg(g + edge * 2 * F + half * F + f);
                }
            }
            dy[row * F + f] = acc;
        }
    }
}

// du partial: block b handles edges [b*chunk, (b+1)*chunk); thread owns entries of (2F x C)
__global__ void __launch_bounds__(256) du_partial_kernel(const float *__restrict__ y, const int64_t *__restrict__ src,
                                                         const int64_t *__restrict__ dst,
                                                         const float *__restrict__ dout, float *__restrict__ partial,
                                                         int64_t E, int F, int C, int64_t chunk) {
    const int64_t e0 = (int64_t)blockIdx.x * chunk;
    const int64_t e1 = min(E, e0 + chunk);
    const int total = 2 * F * C;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int f2 = i / C, c = i % C;
        const int half = f2 >= F, f = f2 - half * F;
        float acc = 0.f;
        for (int64_t e = e0; e < e1; ++e) {
            const int64_t row = half ? dst[e] : src[e];
            acc = fmaf(__ldg(y + row * F + f), __ldg(dout + e * C + c), acc);
        }
        partial[(int64_t)blockIdx.x * total + i] = acc;
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

static int du_chunks() { return sm_count() * 4; }

static int warp_grid(int64_t n_warps) {
    int64_t blocks = ceil_div(n_warps, 8);
    const int64_t cap = (int64_t)sm_count() * 32;
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
#define TMGCN_RC(CC)                                                                   \
    case CC:                                                                           \
        readout_fwd_kernel<CC><<<grid, 256, smem, st>>>(y, src, dst, u, out, E, F);    \
        break;
    switch (C) {
        TMGCN_RC(1) TMGCN_RC(2) TMGCN_RC(3) TMGCN_RC(4) TMGCN_RC(5) TMGCN_RC(6) TMGCN_RC(7) TMGCN_RC(8)
    }
#undef TMGCN_RC
    return after_launch("readout_fwd");
}

int tmgcn_edge_gather_bwd(const float *dz, const int64_t *row_ids, const int64_t *seg_ptr, const int64_t *perm,
                          int64_t n_touched, float *dy, int64_t n_rows, int F, void *stream) {
    TMGCN_REQUIRE(n_touched >= 0 && n_rows >= 0 && F >= 1, "edge_gather_bwd: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rows > 0) TMGCN_CUDA(cudaMemsetAsync(dy, 0, (size_t)n_rows * F * sizeof(float), st));
    if (n_touched == 0) return 0;
    TMGCN_REQUIRE(dz && row_ids && seg_ptr && perm && dy, "edge_gather_bwd: null pointer");
    scatter_rows_kernel<false><<<warp_grid(n_touched), 256, 0, st>>>(dz, nullptr, row_ids, seg_ptr, perm, n_touched,
                                                                     dy, F, 0);
    return after_launch("scatter_rows");
}

size_t tmgcn_edge_du_ws_bytes(int F, int C) { return (size_t)du_chunks() * 2 * F * C * sizeof(float); }

int tmgcn_edge_readout_bwd(const float *y, const int64_t *src, const int64_t *dst, const float *u,
                           const float *dout, const int64_t *row_ids, const int64_t *seg_ptr, const int64_t *perm,
                           int64_t n_touched, float *dy, float *du, int64_t n_rows, int64_t E, int F, int C,
                           void *du_ws, void *stream) {
    TMGCN_REQUIRE(n_touched >= 0 && n_rows >= 0 && E >= 0 && F >= 1, "edge_readout_bwd: bad sizes");
    TMGCN_REQUIRE(C >= 1 && C <= MAXC, "edge_readout_bwd: C=%d outside [1, %d]", C, MAXC);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)2 * F * C * sizeof(float);
    TMGCN_REQUIRE(smem <= 48 * 1024, "edge_readout_bwd: 2*F*C too large");
    if (dy) {
        if (n_rows > 0) TMGCN_CUDA(cudaMemsetAsync(dy, 0, (size_t)n_rows * F * sizeof(float), st));
        if (n_touched > 0) {
            TMGCN_REQUIRE(dout && u && row_ids && seg_ptr && perm, "edge_readout_bwd: null pointer");
            scatter_rows_kernel<true><<<warp_grid(n_touched), 256, smem, st>>>(dout, u, row_ids, seg_ptr, perm,
                                                                               n_touched, dy, F, C);
            if (after_launch("scatter_rows<classify>")) return 1;
        }
    }
    if (du) {
        if (E == 0) {
            TMGCN_CUDA(cudaMemsetAsync(du, 0, smem, st));
            return 0;
        }
        TMGCN_REQUIRE(y && src && dst && dout && du_ws, "edge_readout_bwd: null pointer (du)");
        int64_t n_chunks = ceil_div(E, 64);
        if (n_chunks > du_chunks()) n_chunks = du_chunks();
        const int64_t chunk = ceil_div(E, n_chunks);
        n_chunks = ceil_div(E, chunk);
        du_partial_kernel<<<(unsigned)n_chunks, 256, 0, st>>>(y, src, dst, dout, (float *)du_ws, E, F, C, chunk);
        if (after_launch("du_partial")) return 1;
        reduce_partials_edge<<<(unsigned)ceil_div(2 * F * C, 256), 256, 0, st>>>((const float *)du_ws, du,
                                                                                 (int)n_chunks, 2 * F * C);
        if (after_launch("reduce_partials_edge")) return 1;
    }
    return 0;
}
}
