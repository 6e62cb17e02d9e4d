// (c) feature GEMM, SIMT fp32 path: any (K, Nf).  This is the path the
// reference's own shapes take (F = 2 -> 6 -> 2, C <= 3: no tensor-core tile fits)
// and the exact-fp32 checker for the tcgen05 path (gemm_tc.cu).
//
// ref: t.matmul(AtXt, W) ehf:222 / 330 / 344 / 486-489; nonlinearity ehf:332-335;
// backward = autograd of those:
//   dP = (dY * act'(Y)) . W^T ,  dW = P^T . (dY * act'(Y))  summed over all T*N rows.
#include "common.cuh"

namespace tmgcn {

constexpr int BM = 64, BN = 64, BK = 16;

// C[R, Nc] = A'[R, Kd] . B'[Kd, Nc]
//   A' = A                      (GRAD = false)
//   A' = A * act'(Yaux)         (GRAD = true; Yaux has A's shape)
//   B' = B (Kd x Nc row-major)  (TRANSB = false)   |  B'[k][n] = B[n][k], B is (Nc x Kd) (TRANSB = true)
// epilogue: C = act(C) (fwd only)
template <bool TRANSB, bool GRAD>
__global__ void __launch_bounds__(256) sgemm_rows(const float *__restrict__ A, const float *__restrict__ Yaux,
                                                  const float *__restrict__ B, float *__restrict__ C, int64_t R,
                                                  int Kd, int Nc, int act, int64_t strideA = 0, int64_t strideB = 0,
                                                  int64_t strideC = 0, const float *__restrict__ bias = nullptr) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    // grouped form (per-slice weights, ehf:188-191): blockIdx.z selects the slice, every operand advances by its stride
    A += (int64_t)blockIdx.z * strideA;
    if (GRAD) Yaux += (int64_t)blockIdx.z * strideA;
    B += (int64_t)blockIdx.z * strideB;
    C += (int64_t)blockIdx.z * strideC;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    const int col0 = blockIdx.y * BN;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < Kd; k0 += BK) {
        for (int idx = tid; idx < BM * BK; idx += 256) {
            const int r = idx / BK, k = idx % BK;
            const int64_t gr = row0 + r;
            float v = 0.f;
            if (gr < R && k0 + k < Kd) {
                v = A[gr * Kd + k0 + k];
                if (GRAD) v *= act_grad_rt(Yaux[gr * Kd + k0 + k], act);
            }
            As[k][r] = v;
        }
        for (int idx = tid; idx < BK * BN; idx += 256) {
            int k, n;
            if (TRANSB) {
                n = idx / BK;
                k = idx % BK;
            } else {
                k = idx / BN;
                n = idx % BN;
            }
            float v = 0.f;
            if (k0 + k < Kd && col0 + n < Nc)
                v = TRANSB ? B[(int64_t)(col0 + n) * Kd + k0 + k] : B[(int64_t)(k0 + k) * Nc + col0 + n];
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t gr = row0 + ty * 4 + i;
        if (gr >= R) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gc = col0 + tx * 4 + j;
            if (gc < Nc) C[gr * Nc + gc] = GRAD ? acc[i][j] : act_apply_rt(acc[i][j] + (bias ? bias[gc] : 0.f), act);
        }
    }
}

// partial[chunk][k][n] = sum over the chunk's rows of P[r][k] * dYeff[r][n]
__global__ void __launch_bounds__(256) dw_partial(const float *__restrict__ P, const float *__restrict__ Y,
                                                  const float *__restrict__ dY, float *__restrict__ partial,
                                                  int64_t R, int K, int Nf, int act, int64_t rows_per_chunk) {
    __shared__ float Ps[BK][BM + 4];
    __shared__ float Ds[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int k0 = blockIdx.y * BM, n0 = blockIdx.z * BN;
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_chunk;
    const int64_t r_end = min(R, r_begin + rows_per_chunk);
    float acc[4][4] = {};
    for (int64_t r0 = r_begin; r0 < r_end; r0 += BK) {
        for (int idx = tid; idx < BK * BM; idx += 256) {
            const int rr = idx / BM, k = idx % BM;
            float v = 0.f;
            if (r0 + rr < r_end && k0 + k < K) v = P[(r0 + rr) * K + k0 + k];
            Ps[rr][k] = v;
        }
        for (int idx = tid; idx < BK * BN; idx += 256) {
            const int rr = idx / BN, n = idx % BN;
            float v = 0.f;
            if (r0 + rr < r_end && n0 + n < Nf) {
                const int64_t o = (r0 + rr) * Nf + n0 + n;
                v = dY[o];
                if (act != TMGCN_ACT_NONE) v *= act_grad_rt(Y[o], act);
            }
            Ds[rr][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < BK; ++rr) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = Ps[rr][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Ds[rr][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *out = partial + (int64_t)blockIdx.x * K * Nf;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + ty * 4 + i;
        if (k >= K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < Nf) out[(int64_t)k * Nf + n] = acc[i][j];
        }
    }
}

// per-slice dW[t] = P[t]^T . (dY[t] * act'(Y[t])): one CTA per (slice, 64 x 64 tile of dW) walks all N rows of its
// slice in a fixed order (deterministic, no partials); used by the per-slice-weights path only
__global__ void __launch_bounds__(256) dw_grouped(const float *__restrict__ P, const float *__restrict__ Y,
                                                  const float *__restrict__ dY, float *__restrict__ dW, int64_t N,
                                                  int K, int Nf, int act) {
    __shared__ float Ps[BK][BM + 4];
    __shared__ float Ds[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t t = blockIdx.x;
    const int k0 = blockIdx.y * BM, n0 = blockIdx.z * BN;
    P += t * N * K;
    dY += t * N * Nf;
    if (Y) Y += t * N * Nf;
    float acc[4][4] = {};
    for (int64_t r0 = 0; r0 < N; r0 += BK) {
        for (int idx = tid; idx < BK * BM; idx += 256) {
            const int rr = idx / BM, k = idx % BM;
            Ps[rr][k] = (r0 + rr < N && k0 + k < K) ? P[(r0 + rr) * K + k0 + k] : 0.f;
        }
        for (int idx = tid; idx < BK * BN; idx += 256) {
            const int rr = idx / BN, n = idx % BN;
            float v = 0.f;
            if (r0 + rr < N && n0 + n < Nf) {
                const int64_t o = (r0 + rr) * Nf + n0 + n;
                v = dY[o];
                if (act != TMGCN_ACT_NONE) v *= act_grad_rt(Y[o], act);
            }
            Ds[rr][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < BK; ++rr) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = Ps[rr][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Ds[rr][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *out = dW + t * (int64_t)K * Nf;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + ty * 4 + i;
        if (k >= K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < Nf) out[(int64_t)k * Nf + n] = acc[i][j];
        }
    }
}

// fixed-order reduction of the per-chunk partials: deterministic dW
__global__ void reduce_partials(const float *__restrict__ partial, float *__restrict__ out, int n_chunks,
                                int64_t n_elem) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_elem) return;
    float s = 0.f;
    for (int c = 0; c < n_chunks; ++c) s += partial[(int64_t)c * n_elem + i];
    out[i] = s;
}

__global__ void act_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, int64_t n, int act) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = act_apply_rt(x[i], act);
}
__global__ void act_bwd_kernel(const float *__restrict__ y, const float *__restrict__ dy, float *__restrict__ dx,
                               int64_t n, int act) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dx[i] = dy[i] * act_grad_rt(y[i], act);
}

int dw_max_chunks() { return sm_count() * 4; }

int gemm_simt_fwd(const float *p, const float *w, float *y, int64_t R, int K, int Nf, int act, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(R, BM), (unsigned)ceil_div(Nf, BN));
    sgemm_rows<false, false><<<grid, 256, 0, st>>>(p, nullptr, w, y, R, K, Nf, act);
    return after_launch("sgemm_rows<fwd>");
}

// y = act(p . w + bias): the regression head's nn.Linear (ref: ehf:418-420)
int gemm_simt_bias_fwd(const float *p, const float *w, const float *bias, float *y, int64_t R, int K, int Nf, int act,
                       cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(R, BM), (unsigned)ceil_div(Nf, BN));
    sgemm_rows<false, false><<<grid, 256, 0, st>>>(p, nullptr, w, y, R, K, Nf, act, 0, 0, 0, bias);
    return after_launch("sgemm_rows<fwd+bias>");
}

int gemm_simt_dp(const float *w, const float *y, const float *dy, float *dp, int64_t R, int K, int Nf, int act,
                 cudaStream_t st) {
    // dP[R, K] = (dY * act'(Y))[R, Nf] . W^T ;  W is (K x Nf) row-major => TRANSB with Nc = K, Kd = Nf
    dim3 grid((unsigned)ceil_div(R, BM), (unsigned)ceil_div(K, BN));
    if (act == TMGCN_ACT_NONE)
        sgemm_rows<true, false><<<grid, 256, 0, st>>>(dy, nullptr, w, dp, R, Nf, K, TMGCN_ACT_NONE);
    else
        sgemm_rows<true, true><<<grid, 256, 0, st>>>(dy, y, w, dp, R, Nf, K, act);
    return after_launch("sgemm_rows<dP>");
}

int gemm_simt_dw(const float *p, const float *y, const float *dy, float *dw, int64_t R, int K, int Nf, int act,
                 float *ws, cudaStream_t st) {
    int64_t n_chunks = ceil_div(R, 512);
    if (n_chunks > dw_max_chunks()) n_chunks = dw_max_chunks();
    if (n_chunks < 1) n_chunks = 1;
    int64_t rows_per_chunk = ceil_div(R, n_chunks);
    rows_per_chunk = ceil_div(rows_per_chunk, BK) * BK;
    n_chunks = ceil_div(R, rows_per_chunk);
    dim3 grid((unsigned)n_chunks, (unsigned)ceil_div(K, BM), (unsigned)ceil_div(Nf, BN));
    dw_partial<<<grid, 256, 0, st>>>(p, y, dy, ws, R, K, Nf, act, rows_per_chunk);
    if (after_launch("dw_partial")) return 1;
    const int64_t n_elem = (int64_t)K * Nf;
    reduce_partials<<<(unsigned)ceil_div(n_elem, 256), 256, 0, st>>>(ws, dw, (int)n_chunks, n_elem);
    return after_launch("reduce_partials");
}

// ---- per-slice weights (condensed_W=False, ref: ehf:188-191, 222, 277-282, 330): all T slices in one launch ----
int gemm_simt_sliced_fwd(const float *p, const float *w, float *y, int T, int64_t N, int K, int Nf, int act,
                         cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(N, BM), (unsigned)ceil_div(Nf, BN), (unsigned)T);
    sgemm_rows<false, false><<<grid, 256, 0, st>>>(p, nullptr, w, y, N, K, Nf, act, N * K, (int64_t)K * Nf, N * Nf);
    return after_launch("sgemm_rows<sliced fwd>");
}
int gemm_simt_sliced_dp(const float *w, const float *y, const float *dy, float *dp, int T, int64_t N, int K, int Nf,
                        int act, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(N, BM), (unsigned)ceil_div(K, BN), (unsigned)T);
    if (act == TMGCN_ACT_NONE)
        sgemm_rows<true, false><<<grid, 256, 0, st>>>(dy, nullptr, w, dp, N, Nf, K, TMGCN_ACT_NONE, N * Nf,
                                                      (int64_t)K * Nf, N * K);
    else
        sgemm_rows<true, true><<<grid, 256, 0, st>>>(dy, y, w, dp, N, Nf, K, act, N * Nf, (int64_t)K * Nf, N * K);
    return after_launch("sgemm_rows<sliced dP>");
}
int gemm_simt_sliced_dw(const float *p, const float *y, const float *dy, float *dw, int T, int64_t N, int K, int Nf,
                        int act, cudaStream_t st) {
    dim3 grid((unsigned)T, (unsigned)ceil_div(K, BM), (unsigned)ceil_div(Nf, BN));
    dw_grouped<<<grid, 256, 0, st>>>(p, y, dy, dw, N, K, Nf, act);
    return after_launch("dw_grouped");
}

}  // namespace tmgcn

extern "C" {
int tmgcn_act_fwd(const float *x, float *y, int64_t n, int act, void *stream) {
    using namespace tmgcn;
    TMGCN_REQUIRE(n >= 0 && act >= 0 && act <= 3, "act_fwd: bad arguments");
    if (n == 0) return 0;
    int64_t blocks = ceil_div(n, 256);
    if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
    act_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, n, act);
    return after_launch("act_fwd");
}
int tmgcn_act_bwd(const float *y, const float *dy, float *dx, int64_t n, int act, void *stream) {
    using namespace tmgcn;
    TMGCN_REQUIRE(n >= 0 && act >= 0 && act <= 3, "act_bwd: bad arguments");
    if (n == 0) return 0;
    int64_t blocks = ceil_div(n, 256);
    if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
    act_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(y, dy, dx, n, act);
    return after_launch("act_bwd");
}
}
