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
// A warp takes EB edges per iteration: the 2*EB endpoint ids are fetched with one coalesced load each and the
// 2*EB row gathers are all issued before any is consumed (EB KB in flight per warp at F = 128).
// u staged in shared memory transposed as us[c][2F].
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
    if (VEC == 4 && F == 128) {
        constexpr int EB = 4;
        const int64_t n_blk = (E + EB - 1) / EB;
        const float4 *y4 = reinterpret_cast<const float4 *>(y);
        for (int64_t blk = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); blk < n_blk; blk += warps_total) {
            const int64_t e0 = blk * EB;
            int64_t id = 0;                        // lanes 0..EB-1: src ids, lanes EB..2EB-1: dst ids
            if (lane < 2 * EB) {
                const int64_t e = e0 + (lane & (EB - 1));
                if (e < E) id = lane < EB ? src[e] : dst[e];
            }
            float4 v[2 * EB];
#pragma unroll
            for (int k = 0; k < 2 * EB; ++k) {
                const int64_t row = __shfl_sync(0xffffffffu, id, k);
                v[k] = ld_stream_f4(y4 + row * 32 + lane);
            }
            float acc[EB][C];
#pragma unroll
            for (int k = 0; k < EB; ++k)
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float4 ws = *reinterpret_cast<const float4 *>(us + c * 256 + 4 * lane);
                    const float4 wd = *reinterpret_cast<const float4 *>(us + c * 256 + 128 + 4 * lane);
                    const float4 a = v[k], bq = v[EB + k];
                    acc[k][c] = a.x * ws.x + a.y * ws.y + a.z * ws.z + a.w * ws.w + bq.x * wd.x + bq.y * wd.y +
                                bq.z * wd.z + bq.w * wd.w;
                }
#pragma unroll
            for (int k = 0; k < EB; ++k)
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc[k][c] += __shfl_xor_sync(0xffffffffu, acc[k][c], o);
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < EB; ++k)
                    if (e0 + k < E) {
#pragma unroll
                        for (int c = 0; c < C; ++c) out[(e0 + k) * C + c] = acc[k][c];
                    }
            }
        }
        return;
    }
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

// ---- classifier backward, two streaming passes (see the header comment) -------------------------
// pass 1: S[row][h][c] = sum over incident (e, h) of dOut[e, c].  One THREAD per row: the dependent chain
// inc_ptr -> perm -> dOut is hidden by tens of millions of independent threads; inc_ptr reads and S writes
// are coalesced, only the dOut reads are scattered (2*E small gathers).
template <int CM>
__global__ void __launch_bounds__(256) row_class_sums_kernel(const float *__restrict__ dout,
                                                             const int64_t *__restrict__ inc_ptr,
                                                             const int64_t *__restrict__ perm, int64_t n_rows,
                                                             float *__restrict__ S, int C) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const int64_t s = inc_ptr[row], e = inc_ptr[row + 1];
    float acc[2][CM];
#pragma unroll
    for (int c = 0; c < CM; ++c) acc[0][c] = acc[1][c] = 0.f;
    for (int64_t q = s; q < e; ++q) {
        const int64_t code = perm[q];
        const float *d = dout + (code >> 1) * C;
        const int h = (int)(code & 1);
#pragma unroll
        for (int c = 0; c < CM; ++c)
            if (c < C) {
                const float v = __ldg(d + c);
                if (h) acc[1][c] += v; else acc[0][c] += v;
            }
    }
    float *o = S + row * 2 * C;
#pragma unroll
    for (int c = 0; c < CM; ++c)
        if (c < C) {
            o[c] = acc[0][c];
            o[C + c] = acc[1][c];
        }
}

// pass 2: dY[row, :] = sum_h sum_c S[row][h][c] * U[hF + :, c]  (every row written once, zeros included) and
// dU[hF + f, c] += Y[row, f] * S[row][h][c] accumulated in registers (lane owns features VEC*(lane+32k)..).
// A warp takes RB consecutive rows per iteration and issues all of its loads (S and the RB rows of Y) up
// front; rows whose S is all zero (untouched, or sums that cancel) contribute nothing.
// HAS_DY / HAS_DU select the expand and reduce halves at compile time: the expand-only and reduce-only
// variants (low-rank backward) need far fewer registers and run at 5-6 CTAs per SM.
template <int VEC, int NCH, int CM, bool HAS_DY, bool HAS_DU>
__global__ void __launch_bounds__(256, (NCH == 1 && CM <= 4) ? ((HAS_DY && HAS_DU) ? 2 : (HAS_DU ? 3 : 4)) : 1)
    readout_bwd_kernel(const float *__restrict__ y, const float *__restrict__ u, const float *__restrict__ S, int64_t n_rows,
                                                          float *__restrict__ dy, float *__restrict__ du_partial,
                                                          int F, int C, int act) {
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

    // this lane's slice of U is loop-invariant: keep it in registers (re-reading it from shared memory for
    // every row cost an 8-way bank conflict per access and bounded the kernel at 2.2 TB/s)
    float ur[2][NCH][VEC][CM];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < NCH; ++k)
#pragma unroll
            for (int v = 0; v < VEC; ++v)
#pragma unroll
                for (int c = 0; c < CM; ++c) {
                    const int f = VEC * (lane + 32 * k) + v;
                    ur[h][k][v][c] = (f < F && c < C) ? us[(h * F + f) * C + c] : 0.f;
                }
    constexpr int RB = 4;
    const int sc = 2 * C;                                   // S floats per row (<= 16)
    const int64_t n_blocks = (n_rows + RB - 1) / RB;
    for (int64_t blk = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); blk < n_blocks; blk += warps_total) {
        const int64_t row0 = blk * RB;
        // Everything this block needs is requested up front and independently: its S values (RB * 2C <= 64
        // floats, two coalesced loads) and its RB rows of Y.  Y is read for every row -- gating the read on
        // "S != 0" would save the ~30 % untouched rows but chain a second DRAM latency behind the S load
        // (measured: 2.3 TB/s gated vs a plain stream).
        const int64_t sbase = row0 * sc;
        const int64_t slim = n_rows * sc;
        const float s_lo = (lane < RB * sc && sbase + lane < slim) ? __ldg(S + sbase + lane) : 0.f;
        const float s_hi = (lane + 32 < RB * sc && sbase + lane + 32 < slim) ? __ldg(S + sbase + lane + 32) : 0.f;
        float yv[RB][NCH][VEC];
#pragma unroll
        for (int r = 0; r < RB; ++r)
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int f0 = VEC * (lane + 32 * k);
#pragma unroll
                for (int v = 0; v < VEC; ++v) yv[r][k][v] = 0.f;
                if (HAS_DU && f0 < F && row0 + r < n_rows) {
                    if (VEC == 4) {
                        const float4 t4 = ld_stream_f4(reinterpret_cast<const float4 *>(y + (row0 + r) * F + f0));
                        yv[r][k][0] = t4.x; yv[r][k][1 % VEC] = t4.y; yv[r][k][2 % VEC] = t4.z; yv[r][k][3 % VEC] = t4.w;
                    } else {
                        yv[r][k][0] = __ldg(y + (row0 + r) * F + f0);
                    }
                }
            }
        float Sr[RB][2][CM];
        bool touched[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            touched[r] = false;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int c = 0; c < CM; ++c) {
                    float v = 0.f;
                    if (c < C) {
                        const int idx = r * sc + h * C + c;
                        const float a = __shfl_sync(0xffffffffu, s_lo, idx & 31);
                        const float bq = __shfl_sync(0xffffffffu, s_hi, idx & 31);
                        v = idx < 32 ? a : bq;
                    }
                    Sr[r][h][c] = v;
                    touched[r] = touched[r] || (v != 0.f);
                }
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const int64_t row = row0 + r;
            if (row >= n_rows) break;
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int f0 = VEC * (lane + 32 * k);
                if (f0 < F) {
                    float o[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; ++v) o[v] = 0.f;
                    if (touched[r]) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
#pragma unroll
                            for (int c = 0; c < CM; ++c)
                                if (c < C) {
                                    if (HAS_DY) {
                                        o[v] = fmaf(Sr[r][0][c], ur[0][k][v][c], o[v]);
                                        o[v] = fmaf(Sr[r][1][c], ur[1][k][v][c], o[v]);
                                    }
                                    if (HAS_DU) {
                                        acc[0][k][v][c] = fmaf(yv[r][k][v], Sr[r][0][c], acc[0][k][v][c]);
                                        acc[1][k][v][c] = fmaf(yv[r][k][v], Sr[r][1][c], acc[1][k][v][c]);
                                    }
                                }
                        }
                    }
                    if (HAS_DY) {
                        // layer nonlinearity folded in: dY <- dY * act'(Y), Y being in registers for dU anyway
                        // (saves the separate elementwise pass: one read of Y and a read + write of dY)
                        if (HAS_DU && act != TMGCN_ACT_NONE) {
#pragma unroll
                            for (int v = 0; v < VEC; ++v) o[v] *= act_grad_rt(yv[r][k][v], act);
                        }
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
    if (!HAS_DU) return;
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

// ---- single-output halves of the factor apply (the low-rank backward's expand and reduce) ----------------
// Pure streams: expand writes T*N*F floats from 2C per row, reduce reads T*N*F floats into a 2F x C sum.
// One row per warp step, lane owns float4 columns lane + 32k; the row's 2C factors are ONE broadcast load
// (all lanes, same address: a single 16/32 B request) instead of coalesced loads + 8 shuffles per row, the U
// slice / the accumulators stay in registers, UNR rows are in flight per warp, rows interleave across the
// grid's warps so neighbouring warps touch neighbouring rows.
template <int CM>
__device__ __forceinline__ void load_row_factors(const float *__restrict__ S, int64_t row, int C, bool vec,
                                                 float (&s)[2][CM]) {
    if (vec) {                                              // C == CM and 16-byte aligned rows
        const float4 *q = reinterpret_cast<const float4 *>(S + row * (2 * CM));
        float t[2 * CM];
#pragma unroll
        for (int i = 0; i < (2 * CM) / 4; ++i) {
            const float4 v = __ldg(q + i);
            t[4 * i] = v.x; t[4 * i + 1] = v.y; t[4 * i + 2] = v.z; t[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c = 0; c < CM; ++c) s[h][c] = t[h * CM + c];
    } else {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c = 0; c < CM; ++c) s[h][c] = c < C ? __ldg(S + row * (2 * C) + h * C + c) : 0.f;
    }
}

template <int CM, int NCH, int UNR>
__global__ void __launch_bounds__(256) factor_expand_kernel(const float *__restrict__ u, const float *__restrict__ S,
                                                            int64_t n_rows, float *__restrict__ dy, int F, int C,
                                                            int vec) {
    const int lane = threadIdx.x & 31;
    float ur[2][NCH][4][CM];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < NCH; ++k)
#pragma unroll
            for (int v = 0; v < 4; ++v)
#pragma unroll
                for (int c = 0; c < CM; ++c) {
                    const int f = 4 * (lane + 32 * k) + v;
                    ur[h][k][v][c] = (f < F && c < C) ? __ldg(u + ((int64_t)h * F + f) * C + c) : 0.f;
                }
    const int64_t W = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int64_t base = w0; base < n_rows; base += W * UNR) {
        float s[UNR][2][CM];
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
            const int64_t row = base + j * W;
            if (row < n_rows) load_row_factors<CM>(S, row, C, vec != 0, s[j]);
        }
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
            const int64_t row = base + j * W;
            if (row >= n_rows) break;
            bool touched = false;                           // untouched rows get exact zeros whatever U holds
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int c = 0; c < CM; ++c) touched = touched || (s[j][h][c] != 0.f);
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int f0 = 4 * (lane + 32 * k);
                if (f0 < F) {
                    float o[4] = {0.f, 0.f, 0.f, 0.f};
                    if (touched) {
#pragma unroll
                        for (int v = 0; v < 4; ++v)
#pragma unroll
                            for (int c = 0; c < CM; ++c) {
                                o[v] = fmaf(s[j][0][c], ur[0][k][v][c], o[v]);
                                o[v] = fmaf(s[j][1][c], ur[1][k][v][c], o[v]);
                            }
                    }
                    st_stream_f4(reinterpret_cast<float4 *>(dy + row * F + f0), make_float4(o[0], o[1], o[2], o[3]));
                }
            }
        }
    }
}

// expand, C == CM and aligned S: a warp takes 32 consecutive rows, lane j fetches row j's 2C factors (one
// coalesced request for the block) and the rows are then written one after another with the factors broadcast
// by shuffles -- 32 rows of stores ride on a single load latency, which is what a pure write stream needs
// (the per-row-load variant above reaches 4.5 TB/s, write-only peak is 7.5).
template <int CM, int NCH>
__global__ void __launch_bounds__(256) factor_expand32_kernel(const float *__restrict__ u, const float *__restrict__ S,
                                                              int64_t n_rows, float *__restrict__ dy, int F) {
    const int lane = threadIdx.x & 31;
    float ur[2][NCH][4][CM];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < NCH; ++k)
#pragma unroll
            for (int v = 0; v < 4; ++v)
#pragma unroll
                for (int c = 0; c < CM; ++c) {
                    const int f = 4 * (lane + 32 * k) + v;
                    ur[h][k][v][c] = f < F ? __ldg(u + ((int64_t)h * F + f) * CM + c) : 0.f;
                }
    const int64_t W = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_blk = (n_rows + 31) / 32;
    // the factors of the warp's NEXT block are requested before this block's 32 rows of stores go out
    auto fetch = [&](int64_t blk, float (&dst)[2 * CM]) {
#pragma unroll
        for (int i = 0; i < 2 * CM; ++i) dst[i] = 0.f;
        const int64_t row = blk * 32 + lane;
        if (blk < n_blk && row < n_rows) {
            const float4 *q = reinterpret_cast<const float4 *>(S + row * (2 * CM));
#pragma unroll
            for (int i = 0; i < (2 * CM) / 4; ++i) {
                const float4 t4 = __ldg(q + i);
                dst[4 * i] = t4.x; dst[4 * i + 1] = t4.y; dst[4 * i + 2] = t4.z; dst[4 * i + 3] = t4.w;
            }
        }
    };
    const int64_t blk0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float next[2 * CM];
    fetch(blk0, next);
    for (int64_t blk = blk0; blk < n_blk; blk += W) {
        const int64_t row0 = blk * 32;
        float mine[2 * CM];
#pragma unroll
        for (int i = 0; i < 2 * CM; ++i) mine[i] = next[i];
        fetch(blk + W, next);
        const int n_here = (int)min((int64_t)32, n_rows - row0);
#pragma unroll 4
        for (int j = 0; j < n_here; ++j) {
            float sj[2 * CM];
            bool touched = false;                           // untouched rows get exact zeros whatever U holds
#pragma unroll
            for (int i = 0; i < 2 * CM; ++i) {
                sj[i] = __shfl_sync(0xffffffffu, mine[i], j);
                touched = touched || (sj[i] != 0.f);
            }
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int f0 = 4 * (lane + 32 * k);
                if (f0 < F) {
                    float o[4] = {0.f, 0.f, 0.f, 0.f};
                    if (touched) {
#pragma unroll
                        for (int v = 0; v < 4; ++v)
#pragma unroll
                            for (int c = 0; c < CM; ++c) {
                                o[v] = fmaf(sj[c], ur[0][k][v][c], o[v]);
                                o[v] = fmaf(sj[CM + c], ur[1][k][v][c], o[v]);
                            }
                    }
                    st_stream_f4(reinterpret_cast<float4 *>(dy + (row0 + j) * F + f0),
                                 make_float4(o[0], o[1], o[2], o[3]));
                }
            }
        }
    }
}

template <int CM, int NCH, int UNR>
__global__ void __launch_bounds__(256) factor_reduce_kernel(const float *__restrict__ y, const float *__restrict__ S,
                                                            int64_t n_rows, float *__restrict__ du_partial, int F,
                                                            int C, int vec) {
    extern __shared__ float red[];                          // [8 warps][2F*C]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc[2][NCH][4][CM];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < NCH; ++k)
#pragma unroll
            for (int v = 0; v < 4; ++v)
#pragma unroll
                for (int c = 0; c < CM; ++c) acc[h][k][v][c] = 0.f;
    const int64_t W = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int64_t base = w0; base < n_rows; base += W * UNR) {
        float s[UNR][2][CM];
        float4 yv[UNR][NCH];
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
            const int64_t row = base + j * W;
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int f0 = 4 * (lane + 32 * k);
                yv[j][k] = (row < n_rows && f0 < F) ? ld_stream_f4(reinterpret_cast<const float4 *>(y + row * F + f0))
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (row < n_rows) {
                load_row_factors<CM>(S, row, C, vec != 0, s[j]);
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int c = 0; c < CM; ++c) s[j][h][c] = 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < UNR; ++j)
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const float yy[4] = {yv[j][k].x, yv[j][k].y, yv[j][k].z, yv[j][k].w};
#pragma unroll
                for (int v = 0; v < 4; ++v)
#pragma unroll
                    for (int c = 0; c < CM; ++c) {
                        acc[0][k][v][c] = fmaf(yy[v], s[j][0][c], acc[0][k][v][c]);
                        acc[1][k][v][c] = fmaf(yy[v], s[j][1][c], acc[1][k][v][c]);
                    }
            }
    }
    // block reduction in fixed warp order, then one partial per block
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < NCH; ++k)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int f = 4 * (lane + 32 * k) + v;
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

size_t tmgcn_edge_factor_ws_bytes(int F, int C) { return (size_t)du_blocks() * 2 * F * C * sizeof(float); }

size_t tmgcn_edge_readout_bwd_ws_bytes(int64_t n_rows, int F, int C) {
    // per-row class sums S (n_rows x 2 x C) followed by the per-CTA dU partials
    const size_t s_bytes = ((size_t)(n_rows > 0 ? n_rows : 0) * 2 * C * sizeof(float) + 255) & ~(size_t)255;
    return s_bytes + tmgcn_edge_factor_ws_bytes(F, C);
}

int tmgcn_edge_class_sums(const float *dout, const int64_t *inc_ptr, const int64_t *perm, float *S, int64_t n_rows,
                          int C, void *stream) {
    TMGCN_REQUIRE(n_rows >= 0, "edge_class_sums: bad sizes");
    TMGCN_REQUIRE(C >= 1 && C <= MAXC, "edge_class_sums: C=%d outside [1, %d]", C, MAXC);
    if (n_rows == 0) return 0;
    // dout / perm may be null when no edge touches any row (E = 0): the kernel then never dereferences them
    TMGCN_REQUIRE(inc_ptr && S, "edge_class_sums: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned g = (unsigned)ceil_div(n_rows, 256);
    if (C <= 2) row_class_sums_kernel<2><<<g, 256, 0, st>>>(dout, inc_ptr, perm, n_rows, S, C);
    else if (C <= 4) row_class_sums_kernel<4><<<g, 256, 0, st>>>(dout, inc_ptr, perm, n_rows, S, C);
    else row_class_sums_kernel<8><<<g, 256, 0, st>>>(dout, inc_ptr, perm, n_rows, S, C);
    return after_launch("row_class_sums");
}

static int factor_apply_impl(const float *y, const float *u, const float *S, float *dy, float *du, int64_t n_rows,
                             int F, int C, void *ws, int act, void *stream);

int tmgcn_edge_factor_apply(const float *y, const float *u, const float *S, float *dy, float *du, int64_t n_rows,
                            int F, int C, void *ws, void *stream) {
    return factor_apply_impl(y, u, S, dy, du, n_rows, F, C, ws, TMGCN_ACT_NONE, stream);
}

// act != none: dy <- dy * act'(y) (needs both outputs: y is only in registers on the fused dY + dU kernel)
static int factor_apply_impl(const float *y, const float *u, const float *S, float *dy, float *du, int64_t n_rows,
                             int F, int C, void *ws, int act, void *stream) {
    TMGCN_REQUIRE(n_rows >= 0 && F >= 1, "edge_factor_apply: bad sizes");
    TMGCN_REQUIRE(act >= 0 && act <= 3, "edge_readout_bwd: unknown activation %d", act);
    TMGCN_REQUIRE(act == TMGCN_ACT_NONE || (dy && du), "edge_readout_bwd: a fused activation needs both dy and du");
    TMGCN_REQUIRE(C >= 1 && C <= MAXC, "edge_factor_apply: C=%d outside [1, %d]", C, MAXC);
    cudaStream_t st = (cudaStream_t)stream;
    const int n_u = 2 * F * C;
    if (n_rows == 0) {
        if (du) TMGCN_CUDA(cudaMemsetAsync(du, 0, (size_t)n_u * sizeof(float), st));
        return 0;
    }
    TMGCN_REQUIRE(u && S && (dy || du), "edge_factor_apply: null pointer");
    TMGCN_REQUIRE(!du || (y && ws), "edge_factor_apply: y and ws are required for the reduction output");
    float *partial = du ? (float *)ws : nullptr;
    const bool v4 = F % 4 == 0 && (!dy || (uintptr_t)dy % 16 == 0) && (!y || (uintptr_t)y % 16 == 0);
    const int per_lane = v4 ? 4 : 1;
    const int nch = (F + 32 * per_lane - 1) / (32 * per_lane);
    TMGCN_REQUIRE(nch <= 4, "edge_factor_apply: F=%d too large", F);
    const int cm = C <= 2 ? 2 : (C <= 4 ? 4 : 8);
    if (v4 && (dy == nullptr) != (du == nullptr) && nch * cm <= 4) {
        // one output only: the streaming kernels
        const int vec = (C == cm && (uintptr_t)S % 16 == 0) ? 1 : 0;
        int grid = 0;
#define TMGCN_STREAM(CMX, K)                                                                                      \
    if (cm == CMX && nch == K) {                                                                                  \
        if (dy && vec) {                                                                                          \
            auto kern = factor_expand32_kernel<CMX, K>;                                                           \
            int per_sm = 1;                                                                                       \
            TMGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));                     \
            grid = sm_count() * (per_sm < 1 ? 1 : per_sm);                                                        \
            const int64_t want = ceil_div(n_rows, 8 * 32);                                                        \
            if (grid > want) grid = (int)want;                                                                    \
            kern<<<grid, 256, 0, st>>>(u, S, n_rows, dy, F);                                                      \
        } else if (dy) {                                                                                          \
            auto kern = factor_expand_kernel<CMX, K, 4>;                                                          \
            int per_sm = 1;                                                                                       \
            TMGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));                     \
            grid = sm_count() * (per_sm < 1 ? 1 : per_sm);                                                        \
            const int64_t want = ceil_div(n_rows, 8);                                                             \
            if (grid > want) grid = (int)want;                                                                    \
            kern<<<grid, 256, 0, st>>>(u, S, n_rows, dy, F, C, vec);                                              \
        } else {                                                                                                  \
            auto kern = factor_reduce_kernel<CMX, K, 4>;                                                          \
            const size_t rsm = (size_t)8 * n_u * sizeof(float);                                                   \
            if (rsm > 48 * 1024)                                                                                  \
                TMGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm));    \
            int per_sm = 1;                                                                                       \
            TMGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, rsm));                   \
            if (per_sm < 1) per_sm = 1;                                                                           \
            if (per_sm > 4) per_sm = 4; /* the partial workspace holds 4 CTAs per SM */                           \
            grid = sm_count() * per_sm;                                                                           \
            const int64_t want = ceil_div(n_rows, 8);                                                             \
            if (grid > want) grid = (int)want;                                                                    \
            kern<<<grid, 256, rsm, st>>>(y, S, n_rows, partial, F, C, vec);                                       \
        }                                                                                                         \
    }
        TMGCN_STREAM(2, 1) TMGCN_STREAM(2, 2) TMGCN_STREAM(4, 1)
#undef TMGCN_STREAM
        if (after_launch(dy ? "factor_expand" : "factor_reduce")) return 1;
        if (du) {
            reduce_partials_edge<<<(unsigned)ceil_div(n_u, 256), 256, 0, st>>>(partial, du, grid, n_u);
            if (after_launch("reduce_partials_edge")) return 1;
        }
        return 0;
    }
    const size_t smem = (size_t)n_u * sizeof(float) * (du ? 9 : 1);
    TMGCN_REQUIRE(smem <= 200 * 1024, "edge_factor_apply: 2*F*C too large");
    int grid = 0;
    // persistent grid = exactly one wave of resident CTAs (a partial second wave would run at low occupancy)
#define TMGCN_LAUNCH(V, K, CMX)                                                                                   \
    {                                                                                                             \
        auto kern = (dy && du) ? readout_bwd_kernel<V, K, CMX, true, true>                                        \
                               : (dy ? readout_bwd_kernel<V, K, CMX, true, false>                                 \
                                     : readout_bwd_kernel<V, K, CMX, false, true>);                               \
        if (smem > 48 * 1024)                                                                                     \
            TMGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
        int per_sm = 1;                                                                                           \
        TMGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));                      \
        if (per_sm < 1) per_sm = 1;                                                                               \
        if (du && per_sm > 4) per_sm = 4; /* the dU partial workspace holds 4 CTAs per SM */                      \
        grid = sm_count() * per_sm;                                                                               \
        const int64_t want = ceil_div(n_rows, 8 * 4);                                                             \
        if (grid > want) grid = (int)(want < 1 ? 1 : want);                                                       \
        kern<<<grid, 256, smem, st>>>(y, u, S, n_rows, dy, partial, F, C, act);                                   \
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

int tmgcn_edge_readout_bwd(const float *y, const float *u, const float *dout, const int64_t *inc_ptr,
                           const int64_t *perm, float *dy, float *du, int64_t n_rows, int F, int C, int act, void *ws,
                           void *stream) {
    TMGCN_REQUIRE(n_rows >= 0 && F >= 1, "edge_readout_bwd: bad sizes");
    TMGCN_REQUIRE(C >= 1 && C <= MAXC, "edge_readout_bwd: C=%d outside [1, %d]", C, MAXC);
    TMGCN_REQUIRE(n_rows == 0 || ws, "edge_readout_bwd: null workspace");
    float *S = (float *)ws;
    const size_t s_bytes = ((size_t)(n_rows > 0 ? n_rows : 0) * 2 * C * sizeof(float) + 255) & ~(size_t)255;
    if (tmgcn_edge_class_sums(dout, inc_ptr, perm, S, n_rows, C, stream)) return 1;
    return factor_apply_impl(y, u, S, dy, du, n_rows, F, C, ws ? (char *)ws + s_bytes : nullptr, act, stream);
}
}
