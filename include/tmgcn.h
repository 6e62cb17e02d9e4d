/*
 * tmgcn.h -- C ABI of libtmgcn_b200.so: the B200 (sm_100a) implementation of the
 * TM-GCN propagation hot path.
 *
 * The reference (IBM/TM-GCN, TensorGCN-master/) has no FFI: its boundary is the
 * Python call surface of func_MProduct / compute_AtXt / EmbeddingGCN*.  Each entry
 * point below names the reference lines it replaces ("ref:"; ehf =
 * embedding_help_functions.py).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with h_;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     every call is asynchronous on it and re-entrant across streams;
 *   - the library never allocates: outputs and workspaces are caller-owned,
 *     workspace sizes are queried first (the *_plan / *_ws_bytes calls);
 *   - return value 0 = ok, non-zero = error; tmgcn_last_error() gives the text
 *     (thread-local).  No C++ exception crosses this boundary.
 *
 * Data layout ("CSR-of-slices"): a T x N x N sparse tensor is stored as ONE CSR
 * over T*N rows: rowptr[T*N+1] int64 (row id = t*N + i), col[nnz] int32 (column
 * inside the slice, ascending inside a row), val[nnz] fp32 (or fp64 where said).
 * This is exactly the (t, i, j) order of a coalesced torch COO tensor
 * (invariant 3).  Dense tensors are (T, N, F) fp32, time-major,
 * contiguous -- the reference layout (ehf:204).
 *
 * M is banded lower-triangular (ref: SBM_our.py:88-96, read_data.py:56-62) and is
 * passed as its band: band_w[t*b + i] = M[t, t-i], 0 <= i < b; a zero weight
 * means "no entry" (the reference discovers the band with nonzero(M[:, j]),
 * read_data.py:216).  With time sharding a rank owns T_out consecutive output
 * slices and additionally holds `halo` (<= b-1) predecessor slices of every
 * INPUT tensor in front of its own: input slice index = halo + t - i.
 */
#ifndef TMGCN_H
#define TMGCN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMGCN_ABI_VERSION 2

/* epilogue / nonlinearity selector (ref: ehf:284-289) */
enum tmgcn_act { TMGCN_ACT_NONE = 0, TMGCN_ACT_RELU = 1, TMGCN_ACT_LEAKY = 2, TMGCN_ACT_SELU = 3 };

const char *tmgcn_last_error(void);
int tmgcn_abi_version(void);
/* number of kernels launched by this library in this process so far (bench.py's gpu_launches) */
int64_t tmgcn_launch_count(void);

/* ---- helpers --------------------------------------------------------- */
/* exclusive prefix sum of n int64 counts into out[n+1] (out[n] = total).
 * ws: tmgcn_scan_ws_bytes(n) bytes. */
size_t tmgcn_scan_ws_bytes(int64_t n);
int tmgcn_exclusive_scan_i64(const int64_t *counts, int64_t *out, int64_t n, void *ws, void *stream);

/* coalesced COO -> CSR-of-slices rowptr.  flat_row[nnz] int64 sorted ascending
 * (t*N + i of every stored entry); writes rowptr[n_rows+1].
 * Replaces the T boolean-mask passes of ref: ehf:561-572 / experiment_bitcoin_our.py:53-64. */
int tmgcn_rowptr_from_sorted_rows(const int64_t *flat_row, int64_t nnz, int64_t n_rows, int64_t *rowptr,
                                  void *stream);

/* ---- (a) sparse M-transform  A~ = A x_3 M ------------------------------
 * ref: func_MProduct, read_data.py:204-223 (bit-exact index output).
 * plan: writes out_counts[T_out*N] = size of the union pattern of each output row.
 * The caller scans them (tmgcn_exclusive_scan_i64) into out_rowptr and allocates
 * col/val of out_rowptr[T_out*N] entries.  run: fills out_col / out_val.
 * val_is_f64 selects fp64 (func_MProduct API parity) or fp32 values; sums are
 * always accumulated in fp64 in ascending source-slice order. */
int tmgcn_mtransform_sparse_plan(const int64_t *in_rowptr, const int32_t *in_col, int T_out, int halo, int64_t N,
                                 const double *band_w, int b, int64_t *out_counts, void *stream);
int tmgcn_mtransform_sparse_run(const int64_t *in_rowptr, const int32_t *in_col, const void *in_val, int T_out,
                                int halo, int64_t N, const double *band_w, int b, const int64_t *out_rowptr,
                                int32_t *out_col, void *out_val, int val_is_f64, void *stream);

/* The same two calls with a caller-owned workspace (tmgcn_mtransform_sparse_ws_bytes; 0 = not applicable, pass
 * NULL): the plan call records the union pattern of every (4 output slices x 32 rows) task in it and the run call
 * turns the record into values without merging again (about 2x faster; results are bit-identical).  The same
 * workspace, untouched in between, must be passed to both calls.  in_nnz = in_rowptr[(halo+T_out)*N].
 * fp32 values only: with val_is_f64 the run call ignores the record and merges as tmgcn_mtransform_sparse_run does. */
size_t tmgcn_mtransform_sparse_ws_bytes(int T_out, int halo, int64_t N, int b, int64_t in_nnz);
int tmgcn_mtransform_sparse_plan_ws(const int64_t *in_rowptr, const int32_t *in_col, int T_out, int halo, int64_t N,
                                    const double *band_w, int b, int64_t *out_counts, void *ws, size_t ws_bytes,
                                    void *stream);
int tmgcn_mtransform_sparse_run_ws(const int64_t *in_rowptr, const int32_t *in_col, const void *in_val, int T_out,
                                   int halo, int64_t N, const double *band_w, int b, const int64_t *out_rowptr,
                                   int32_t *out_col, void *out_val, int val_is_f64, void *ws, size_t ws_bytes,
                                   void *stream);

/* transpose every slice of a CSR-of-slices (for the backward SpMM).
 * plan: counts[T*N] (zero-initialised by the call) = entries per transposed row.
 * run: given the scanned t_rowptr, fills t_col / t_val with ascending columns.
 * ws: tmgcn_csr_transpose_ws_bytes(T*N, nnz, val_is_f64) bytes; values fp32 or fp64. */
size_t tmgcn_csr_transpose_ws_bytes(int64_t n_rows, int64_t nnz, int val_is_f64);
int tmgcn_csr_transpose_plan(const int64_t *rowptr, const int32_t *col, int T, int64_t N, int64_t *counts,
                             void *stream);
int tmgcn_csr_transpose_run(const int64_t *rowptr, const int32_t *col, const void *val, int T, int64_t N,
                            const int64_t *t_rowptr, int32_t *t_col, void *t_val, int val_is_f64, void *ws,
                            void *stream);

/* ---- graph preparation (the step before (a)) ---------------------------
 * ref: read_data.py:88-164 -- func_make_symmetric (A + A^T)/2, func_edge_life (a ones-band M-transform:
 * use tmgcn_mtransform_sparse_*), func_laplacian_transformation D^-1/2 (B + I) D^-1/2.
 * axpby     : C = alpha*A + beta*B by a sorted 2-way row merge (plan counts, caller scans, run fills);
 * row_sums  : out[row] = sum of the row's values (fp64);
 * scale_sym : val[k] *= deg[row]^-1/2 * deg[t*N + col[k]]^-1/2 in place. */
int tmgcn_csr_axpby_plan(const int64_t *a_rowptr, const int32_t *a_col, const int64_t *b_rowptr, const int32_t *b_col,
                         int64_t n_rows, int64_t *counts, void *stream);
int tmgcn_csr_axpby_run(const int64_t *a_rowptr, const int32_t *a_col, const void *a_val, const int64_t *b_rowptr,
                        const int32_t *b_col, const void *b_val, double alpha, double beta, int64_t n_rows,
                        const int64_t *c_rowptr, int32_t *c_col, void *c_val, int val_is_f64, void *stream);
int tmgcn_csr_row_sums(const int64_t *rowptr, const void *val, int64_t n_rows, double *out, int val_is_f64,
                       void *stream);
int tmgcn_csr_scale_sym(const int64_t *rowptr, const int32_t *col, void *val, int T, int64_t N, const double *deg,
                        int val_is_f64, void *stream);

/* ---- (b) dense M-transform  X~ = X x_3 M  (time stencil) ------------------
 * ref: ehf:204 / ehf:308 / ehf:346  (M @ X.reshape(T, N*F)).
 * fwd: x_in (halo+T_out, NF) -> x_out (T_out, NF).
 * bwd: g_out (T_out, NF) -> g_in (halo+T_out, NF) = M^T applied; the first `halo`
 *      slices of g_in are the partial sums owed to the predecessor rank.
 * band_w here is fp32 (device). NF = N*F. */
int tmgcn_mtransform_dense_fwd(const float *x_in, float *x_out, int T_out, int halo, int64_t NF,
                               const float *band_w, int b, void *stream);
int tmgcn_mtransform_dense_bwd(const float *g_out, float *g_in, int T_out, int halo, int64_t NF,
                               const float *band_w, int b, void *stream);
/* forward with the input in two pieces: the `halo` predecessor slices at x_halo, the own slices at x_own.
 * x_halo may point into a PEER GPU's memory (CUDA IPC / symmetric memory mapped over NVLink): the halo
 * exchange is then fused into the stencil -- the kernel loads the predecessor's slices straight from its
 * HBM, no staging copy and no halo region in the local tensor.  max_ctas > 0 runs it as a persistent grid
 * of at most that many 128-thread CTAs: the peer-reading launch is NVLink-bound, and a full grid of
 * stalled CTAs would take the registers the interior stencil / SpMM running beside it need (0 = full grid). */
int tmgcn_mtransform_dense_fwd_split(const float *x_halo, const float *x_own, float *x_out, int T_out, int halo,
                                     int64_t NF, const float *band_w, int b, int max_ctas, void *stream);
/* same, but only input slices s in [s_begin, s_end) of g_in are written (the others are left untouched):
 * lets a rank produce the `halo` slices it owes its predecessor first (and put them on the wire) and the
 * rest later, in place, without a staging copy.  Slices s >= acc_begin are ACCUMULATED into (g_in[s] += ...):
 * the partial sums the successor rank owes this rank's last slices are received straight into g_in while
 * the backward SpMM runs, and the stencil adds its own contribution on top (acc_begin < 0: overwrite all). */
int tmgcn_mtransform_dense_bwd_range(const float *g_out, float *g_in, int T_out, int halo, int64_t NF,
                                     const float *band_w, int b, int s_begin, int s_end, int acc_begin, void *stream);

/* inverse transform  Y = inv(M) x_3 Z  (ref: Minv = inv(M), ehf:183-184; applied at ehf:223-224, 331-332,
 * 338-341): inv(M) of a banded M is dense, so it is applied as the banded substitution M x_3 Y = Z marching
 * through time (fp64 recurrence, fp32 in/out); _bwd solves with M^T (the adjoint).
 * _part is the time-sharded, column-chunked form: it continues a recurrence started on another rank from `h`
 * halo rows (forward: the predecessor's last h OUTPUT slices; transposed: the successor's first h output
 * slices), works on n columns of rows that are `ld` floats apart (halo rows `ld_halo` apart), and -- transposed
 * -- needs band_w to hold T + h rows (M[s+i, s] with s+i in the successor's block).  A caller pipelines the
 * cross-rank scan over column chunks (tmgcn_b200/sharding.py: solve_pipelined). */
int tmgcn_mtransform_dense_solve_fwd(const float *z, float *y, int T, int64_t NF, const float *band_w, int b,
                                     void *stream);
int tmgcn_mtransform_dense_solve_bwd(const float *g_y, float *g_z, int T, int64_t NF, const float *band_w, int b,
                                     void *stream);
int tmgcn_mtransform_dense_solve_part(const float *src, float *dst, const float *halo, int T, int h, int64_t n,
                                      int64_t ld, int64_t ld_halo, const float *band_w, int b, int transposed,
                                      void *stream);

/* ---- (d) facewise SpMM  P_t = A~_t . X_t -----------------------------------
 * ref: the loop ehf:205-207 / ehf:309-311 and compute_AX ehf:301-305, 469-473.
 * y[t, i, :] = act( sum_k val[k] * x[t, col[k], :] ), all T slices in one launch.
 * The backward (dX_t = A~_t^T . dY_t) is the same call on the transposed CSR. */
int tmgcn_spmm_fwd(const int64_t *rowptr, const int32_t *col, const float *val, const float *x, float *y, int T,
                   int64_t N, int F, int act, void *stream);

/* ---- (c) feature GEMM  Y = act(P . W) --------------------------------------
 * ref: t.matmul(AtXt, W), ehf:222 / 330 / 344 / 486-489; nonlinearity ehf:332-335.
 * p (R, K) . w (K, Nf) -> y (R, Nf), R = T*N rows; fp32 in/out.  K = Nf = 128 (and
 * other multiples handled by the tensor-core path) run on tcgen05 with 3xTF32
 * error compensation; anything else takes the SIMT fp32 kernel.
 * bwd: given y (post-activation) and dy: dy <- dy * act'(y) implicitly, then
 *   dp (R, K) = dy . w^T   and   dw (K, Nf) = p^T . dy  (slice-summed, deterministic).
 * dw_ws: tmgcn_gemm_dw_ws_bytes(K, Nf) bytes of scratch for the per-CTA partials. */
int tmgcn_gemm_xw_fwd(const float *p, const float *w, float *y, int64_t R, int K, int Nf, int act, void *stream);
/* y = act(p . w + bias), bias[Nf]: the nn.Linear of the regression head (ref: ehf:418-420); fp32 SIMT kernel.
 * Its backward is tmgcn_gemm_dw_dx_bwd for dp / dw and the same call with p = ones(R, 1) for dbias = 1^T . dy. */
int tmgcn_gemm_xw_bias_fwd(const float *p, const float *w, const float *bias, float *y, int64_t R, int K, int Nf,
                           int act, void *stream);
size_t tmgcn_gemm_dw_ws_bytes(int K, int Nf);
int tmgcn_gemm_dw_dx_bwd(const float *p, const float *w, const float *y, const float *dy, float *dp, float *dw,
                         int64_t R, int K, int Nf, int act, void *dw_ws, void *stream);

/* per-slice weights (condensed_W=False; ref: ehf:188-191, 222, 277-282, 330): y[t] = act(p[t] . w[t]) with
 * p (T, N, K), w (T, K, Nf), y (T, N, Nf) -- all T slices per call (one grouped launch on the SIMT path).
 * bwd: dp[t] = (dy[t] * act'(y[t])) . w[t]^T and dw[t] = p[t]^T . (dy[t] * act'(y[t])); dp or dw may be NULL. */
int tmgcn_gemm_xw_sliced_fwd(const float *p, const float *w, float *y, int T, int64_t N, int K, int Nf, int act,
                             void *stream);
int tmgcn_gemm_sliced_bwd(const float *p, const float *w, const float *y, const float *dy, float *dp, float *dw,
                          int T, int64_t N, int K, int Nf, int act, void *stream);

/* ---- (e) edge-endpoint gather readout ------------------------------------
 * ref: flat ids ehf:196-198; gather + concat ehf:228-230 / 351-353 / 491-493;
 * classifier ehf:232 / 355 / 495.
 * gather_fwd : z[e, :] = [ y[src[e], :] || y[dst[e], :] ]          (E, 2F)
 * readout_fwd: out[e, :] = z[e, :] . u   without materialising z     (E, C), C <= 8
 * src/dst are flat row ids t*N + node (int64). */
int tmgcn_flat_edge_ids(const int64_t *edges /* (3,E) time,src,dst */, int64_t E, int64_t N, int64_t t_offset,
                        int64_t *src, int64_t *dst, void *stream);
int tmgcn_edge_gather_fwd(const float *y, const int64_t *src, const int64_t *dst, float *z, int64_t E, int F,
                          void *stream);
int tmgcn_edge_readout_fwd(const float *y, const int64_t *src, const int64_t *dst, const float *u, float *out,
                           int64_t E, int F, int C, void *stream);
/* backward.  The incidence CSR makes the scatter-add deterministic (no atomics):
 * inc_ptr[n_rows+1] over ALL T*N rows and perm[2E] = (e*2+half) grouped by endpoint row
 * (built once per edge set: stable sort of the 2E endpoint ids + tmgcn_rowptr_from_sorted_rows).
 * gather_bwd : dy[row, :] = sum over incident (e, half) of dz[e, half*F:(half+1)*F]
 * readout_bwd: dy[row, :] = sum dout[e, :] . u[half*F:(half+1)*F, :]^T ;  du = z^T . dout
 * dy (n_rows, F) is written exactly once, rows no edge touches become 0.  dy or du may be
 * NULL to skip that output.  ws: tmgcn_edge_readout_bwd_ws_bytes(n_rows, F, C) bytes of scratch
 * (per-row class sums + per-CTA dU partials). */
int tmgcn_edge_gather_bwd(const float *dz, const int64_t *inc_ptr, const int64_t *perm, float *dy, int64_t n_rows,
                          int F, void *stream);
size_t tmgcn_edge_readout_bwd_ws_bytes(int64_t n_rows, int F, int C);
/* The two passes of readout_bwd, exposed separately because the per-row class sums are a rank-2C
 * factorisation of the whole upstream gradient (dY = S . U~): when the layer between the propagation
 * and the readout is linear, every backward stage can run on the skinny factor (see layer_step.py).
 * class_sums  : S[row, h, c] = sum over incident (e, h) of dout[e, c]            S is (n_rows, 2, C)
 * factor_apply: dy[row, f] = sum_{h,c} S[row,h,c] * u[hF+f, c]     (if dy)      "expand"
 *               du[hF+f, c] = sum_rows y[row, f] * S[row,h,c]      (if du)      "reduce"
 * ws: tmgcn_edge_factor_ws_bytes(F, C) bytes, needed only when du is requested.
 * dout / perm may be null when there is no edge at all (inc_ptr all zero). */
size_t tmgcn_edge_factor_ws_bytes(int F, int C);
int tmgcn_edge_class_sums(const float *dout, const int64_t *inc_ptr, const int64_t *perm, float *S, int64_t n_rows,
                          int C, void *stream);
int tmgcn_edge_factor_apply(const float *y, const float *u, const float *S, float *dy, float *du, int64_t n_rows,
                            int F, int C, void *ws, void *stream);
/* act != TMGCN_ACT_NONE additionally folds the layer nonlinearity in front of the readout into the same pass:
 * dy <- dy * act'(y), y being the POST-activation embedding the readout gathered from (ref: ehf:332-335 followed by
 * ehf:351-355); needs both dy and du (y is in registers for du anyway). */
int tmgcn_edge_readout_bwd(const float *y, const float *u, const float *dout, const int64_t *inc_ptr,
                           const int64_t *perm, float *dy, float *du, int64_t n_rows, int F, int C, int act, void *ws,
                           void *stream);

/* ---- elementwise activation (layer boundary, ref: ehf:332-335) ----------- */
int tmgcn_act_fwd(const float *x, float *y, int64_t n, int act, void *stream);
int tmgcn_act_bwd(const float *y, const float *dy, float *dx, int64_t n, int act, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TMGCN_H */
