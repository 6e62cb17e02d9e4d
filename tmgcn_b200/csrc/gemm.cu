// (c) feature GEMM entry points: dispatch between the tcgen05 3xTF32 kernel
// (gemm_tc.cu) and the SIMT fp32 kernel (gemm_simt.cu).
#include <stdlib.h>

#include "common.cuh"

namespace tmgcn {
int dw_max_chunks();
int gemm_simt_fwd(const float *p, const float *w, float *y, int64_t R, int K, int Nf, int act, cudaStream_t st);
int gemm_simt_dp(const float *w, const float *y, const float *dy, float *dp, int64_t R, int K, int Nf, int act,
                 cudaStream_t st);
int gemm_simt_dw(const float *p, const float *y, const float *dy, float *dw, int64_t R, int K, int Nf, int act,
                 float *ws, cudaStream_t st);

int gemm_simt_sliced_fwd(const float *p, const float *w, float *y, int T, int64_t N, int K, int Nf, int act,
                         cudaStream_t st);
int gemm_simt_sliced_dp(const float *w, const float *y, const float *dy, float *dp, int T, int64_t N, int K, int Nf,
                        int act, cudaStream_t st);
int gemm_simt_sliced_dw(const float *p, const float *y, const float *dy, float *dw, int T, int64_t N, int K, int Nf,
                        int act, cudaStream_t st);

int gemm_simt_bias_fwd(const float *p, const float *w, const float *bias, float *y, int64_t R, int K, int Nf, int act,
                       cudaStream_t st);

// tensor-core path (gemm_tc.cu)
bool gemm_tc_eligible(int64_t R, int K, int Nf);
int gemm_tc_fwd(const float *a, const float *w, float *c, int64_t R, int K, int Nf, int act, bool trans_w,
                const float *yaux, cudaStream_t st);
bool gemm_dw_tc_eligible(int64_t R, int KI, int NO);
int gemm_dw_tc(const float *p, const float *dy, float *dw, int64_t R, int KI, int NO, float *ws, cudaStream_t st);

static bool tc_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("TMGCN_DISABLE_TC");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}
}  // namespace tmgcn

using namespace tmgcn;

extern "C" {

int tmgcn_gemm_xw_fwd(const float *p, const float *w, float *y, int64_t R, int K, int Nf, int act, void *stream) {
    TMGCN_REQUIRE(R >= 0 && K >= 1 && Nf >= 1, "gemm_xw_fwd: bad sizes R=%lld K=%d Nf=%d", (long long)R, K, Nf);
    TMGCN_REQUIRE(act >= 0 && act <= 3, "gemm_xw_fwd: unknown activation %d", act);
    if (R == 0) return 0;
    TMGCN_REQUIRE(p && w && y, "gemm_xw_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (tc_enabled() && gemm_tc_eligible(R, K, Nf)) return gemm_tc_fwd(p, w, y, R, K, Nf, act, false, nullptr, st);
    return gemm_simt_fwd(p, w, y, R, K, Nf, act, st);
}

int tmgcn_gemm_xw_bias_fwd(const float *p, const float *w, const float *bias, float *y, int64_t R, int K, int Nf,
                           int act, void *stream) {
    TMGCN_REQUIRE(R >= 0 && K >= 1 && Nf >= 1, "gemm_xw_bias_fwd: bad sizes R=%lld K=%d Nf=%d", (long long)R, K, Nf);
    TMGCN_REQUIRE(act >= 0 && act <= 3, "gemm_xw_bias_fwd: unknown activation %d", act);
    if (R == 0) return 0;
    TMGCN_REQUIRE(p && w && bias && y, "gemm_xw_bias_fwd: null pointer");
    return gemm_simt_bias_fwd(p, w, bias, y, R, K, Nf, act, (cudaStream_t)stream);
}

size_t tmgcn_gemm_dw_ws_bytes(int K, int Nf) { return (size_t)dw_max_chunks() * (size_t)K * (size_t)Nf * sizeof(float); }

int tmgcn_gemm_dw_dx_bwd(const float *p, const float *w, const float *y, const float *dy, float *dp, float *dw,
                         int64_t R, int K, int Nf, int act, void *dw_ws, void *stream) {
    TMGCN_REQUIRE(R >= 0 && K >= 1 && Nf >= 1, "gemm_bwd: bad sizes R=%lld K=%d Nf=%d", (long long)R, K, Nf);
    TMGCN_REQUIRE(act >= 0 && act <= 3, "gemm_bwd: unknown activation %d", act);
    TMGCN_REQUIRE(act == TMGCN_ACT_NONE || y, "gemm_bwd: y is required when an activation is fused");
    cudaStream_t st = (cudaStream_t)stream;
    if (R == 0) {
        if (dw) TMGCN_CUDA(cudaMemsetAsync(dw, 0, (size_t)K * Nf * sizeof(float), st));
        return 0;
    }
    TMGCN_REQUIRE(w && dy, "gemm_bwd: null pointer");
    if (dp) {
        int rc;
        // the tensor-core kernel has no fused act'(y): callers wanting it pre-apply tmgcn_act_bwd (in place)
        if (tc_enabled() && act == TMGCN_ACT_NONE && gemm_tc_eligible(R, Nf, K))
            rc = gemm_tc_fwd(dy, w, dp, R, Nf, K, TMGCN_ACT_NONE, true, nullptr, st);
        else
            rc = gemm_simt_dp(w, y, dy, dp, R, K, Nf, act, st);
        if (rc) return rc;
    }
    if (dw) {
        TMGCN_REQUIRE(p && dw_ws, "gemm_bwd: p and dw_ws are required for dW");
        if (tc_enabled() && act == TMGCN_ACT_NONE && gemm_dw_tc_eligible(R, K, Nf)) {
            if (gemm_dw_tc(p, dy, dw, R, K, Nf, (float *)dw_ws, st)) return 1;
        } else if (gemm_simt_dw(p, y, dy, dw, R, K, Nf, act, (float *)dw_ws, st)) {
            return 1;
        }
    }
    return 0;
}

/* per-slice weights: y[t] = act(p[t] . w[t]) and its backward, all slices per call */
int tmgcn_gemm_xw_sliced_fwd(const float *p, const float *w, float *y, int T, int64_t N, int K, int Nf, int act,
                             void *stream) {
    TMGCN_REQUIRE(T >= 0 && N >= 0 && K >= 1 && Nf >= 1, "gemm_xw_sliced_fwd: bad sizes T=%d N=%lld K=%d Nf=%d", T,
                  (long long)N, K, Nf);
    TMGCN_REQUIRE(act >= 0 && act <= 3, "gemm_xw_sliced_fwd: unknown activation %d", act);
    if ((int64_t)T * N == 0) return 0;
    TMGCN_REQUIRE(p && w && y, "gemm_xw_sliced_fwd: null pointer");
    TMGCN_REQUIRE(T <= 65535, "gemm_xw_sliced_fwd: more than 65535 slices");
    cudaStream_t st = (cudaStream_t)stream;
    if (tc_enabled() && gemm_tc_eligible(N, K, Nf)) {      // wide features: the tensor-core kernel, slice by slice
        for (int t = 0; t < T; ++t)
            if (gemm_tc_fwd(p + (int64_t)t * N * K, w + (int64_t)t * K * Nf, y + (int64_t)t * N * Nf, N, K, Nf, act,
                            false, nullptr, st))
                return 1;
        return 0;
    }
    return gemm_simt_sliced_fwd(p, w, y, T, N, K, Nf, act, st);
}

int tmgcn_gemm_sliced_bwd(const float *p, const float *w, const float *y, const float *dy, float *dp, float *dw,
                          int T, int64_t N, int K, int Nf, int act, void *stream) {
    TMGCN_REQUIRE(T >= 0 && N >= 0 && K >= 1 && Nf >= 1, "gemm_sliced_bwd: bad sizes");
    TMGCN_REQUIRE(act >= 0 && act <= 3, "gemm_sliced_bwd: unknown activation %d", act);
    TMGCN_REQUIRE(act == TMGCN_ACT_NONE || y, "gemm_sliced_bwd: y is required when an activation is fused");
    TMGCN_REQUIRE(T <= 65535, "gemm_sliced_bwd: more than 65535 slices");
    cudaStream_t st = (cudaStream_t)stream;
    if (T == 0) return 0;
    if (N == 0) {
        if (dw) TMGCN_CUDA(cudaMemsetAsync(dw, 0, (size_t)T * K * Nf * sizeof(float), st));
        return 0;
    }
    TMGCN_REQUIRE(w && dy, "gemm_sliced_bwd: null pointer");
    if (dp && gemm_simt_sliced_dp(w, y, dy, dp, T, N, K, Nf, act, st)) return 1;
    if (dw) {
        TMGCN_REQUIRE(p, "gemm_sliced_bwd: p is required for dW");
        if (gemm_simt_sliced_dw(p, y, dy, dw, T, N, K, Nf, act, st)) return 1;
    }
    return 0;
}
}
