// (a) sparse M-transform: per-row sorted B-way merge of the band's CSR rows, plus the
// integer plumbing around it (exclusive scan, COO->CSR row pointers, per-slice
// CSR transpose for the backward SpMM).  All of it is HBM/latency-bound integer
// work: coalesced where the data allows, grids sized in multiples of the SM count.
//
// ref: func_MProduct, TensorGCN-master/read_data.py:204-223.
#include <limits.h>
#include <stdlib.h>

#include <utility>

#include "common.cuh"

namespace tmgcn {

// ------------------------------------------------------------------------
// exclusive scan (int64), three phases, deterministic
// ------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;  // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t *total) {
    // v: this thread's value; returns exclusive prefix inside the block
    __shared__ int64_t warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        int64_t s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t n = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += n;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    int64_t base = w > 0 ? warp_sums[w - 1] : 0;
    if (total) *total = warp_sums[SCAN_THREADS / 32 - 1];
    int64_t r = base + incl - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const int64_t *__restrict__ in, int64_t n,
                                                               int64_t *__restrict__ tile_sums) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int64_t s = 0;
    for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_THREADS) {
        int64_t k = base + i;
        if (k < n) s += in[k];
    }
    int64_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_offsets(int64_t *__restrict__ tile_sums, int64_t n_tiles,
                                                                  int64_t *__restrict__ grand_total) {
    // single block: in-place exclusive scan of tile_sums
    int64_t carry = 0;
    for (int64_t base = 0; base < n_tiles; base += SCAN_THREADS) {
        int64_t k = base + threadIdx.x;
        int64_t v = k < n_tiles ? tile_sums[k] : 0;
        int64_t total;
        int64_t ex = block_exclusive_scan(v, &total);
        if (k < n_tiles) tile_sums[k] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(const int64_t *__restrict__ in, int64_t n,
                                                           const int64_t *__restrict__ tile_offsets,
                                                           int64_t *__restrict__ out) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t v[SCAN_ITEMS];
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t k = base + i;
        v[i] = k < n ? in[k] : 0;
        s += v[i];
    }
    int64_t ex = block_exclusive_scan(s, nullptr) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t k = base + i;
        if (k < n) out[k] = ex;
        ex += v[i];
    }
}

// ------------------------------------------------------------------------
// sorted flat row ids -> rowptr
// ------------------------------------------------------------------------
__global__ void rowptr_from_rows(const int64_t *__restrict__ rows, int64_t nnz, int64_t n_rows,
                                 int64_t *__restrict__ rowptr) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > nnz) return;
    int64_t prev = k == 0 ? -1 : rows[k - 1];
    int64_t cur = k == nnz ? n_rows : rows[k];
    for (int64_t r = prev + 1; r <= cur; ++r) rowptr[r] = k;
}

// ------------------------------------------------------------------------
// merge: one THREAD per output row (t, i): a B-way sorted merge of row i of the source slices
// halo + t - l, l < b.  The B cursors live in registers (the loops over cursors are fully unrolled),
// each step emits the smallest pending column and every cursor sitting on it contributes w*val in fp64,
// in ascending source-slice order (the order coalesce() sums duplicates in) and advances.
// A warp-per-row version spent its time in per-output warp reductions (10 fp64 shuffles per emitted
// entry: 146 ms for 1.2 G outputs); here a warp advances 32 rows at once with ~2 instructions per cursor
// per output, the lists are read sequentially per thread (L1 sector reuse) and rows of neighbouring
// threads are neighbours in memory.
// ------------------------------------------------------------------------
template <int B, bool COUNT_ONLY, typename VT>
__global__ void __launch_bounds__(128) merge_rows(const int64_t *__restrict__ in_rowptr,
                                                  const int32_t *__restrict__ in_col, const VT *__restrict__ in_val,
                                                  int T_out, int halo, int64_t N, const double *__restrict__ band_w,
                                                  int b, int64_t *__restrict__ out_counts,
                                                  const int64_t *__restrict__ out_rowptr,
                                                  int32_t *__restrict__ out_col, VT *__restrict__ out_val) {
    const int64_t n_out_rows = (int64_t)T_out * N;
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_out_rows) return;
    const int t = (int)(row / N);
    const int64_t i = row - (int64_t)t * N;
    int64_t pos[B];      // cursor l walks source slice halo + t - l; slot B-1-l so that slot order = ascending slice
    int32_t left[B];     // entries left in the list
    int32_t cur[B];      // column under the cursor (INT_MAX when exhausted)
    double w[B];
#pragma unroll
    for (int k = 0; k < B; ++k) {
        const int l = B - 1 - k;
        pos[k] = 0;
        left[k] = 0;
        cur[k] = INT_MAX;
        w[k] = 0.0;
        if (l < b) {
            const int sl = halo + t - l;
            const double wl = band_w[(int64_t)t * b + l];
            if (sl >= 0 && wl != 0.0) {
                const int64_t p0 = in_rowptr[(int64_t)sl * N + i];
                const int64_t p1 = in_rowptr[(int64_t)sl * N + i + 1];
                pos[k] = p0;
                left[k] = (int32_t)(p1 - p0);
                w[k] = wl;
                if (p1 > p0) cur[k] = in_col[p0];
            }
        }
    }
    int64_t count = 0;
    const int64_t obase = COUNT_ONLY ? 0 : out_rowptr[row];
    while (true) {
        int m = cur[0];
#pragma unroll
        for (int k = 1; k < B; ++k) m = min(m, cur[k]);
        if (m == INT_MAX) break;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < B; ++k) {
            if (cur[k] == m) {
                if (!COUNT_ONLY) acc += w[k] * (double)in_val[pos[k]];
                ++pos[k];
                --left[k];
                cur[k] = left[k] > 0 ? in_col[pos[k]] : INT_MAX;
            }
        }
        if (!COUNT_ONLY) {
            out_col[obase + count] = m;
            out_val[obase + count] = (VT)acc;
        }
        ++count;
    }
    if (COUNT_ONLY) out_counts[row] = count;
}

// Staged variant (fill pass, see launch_merge).  A warp owns RW = 32/LPR consecutive rows of one output slice and
// LPR lanes share a row, each holding CB = ceil(B/LPR) of its B cursors.  For every source slice of the band
// the warp's rows are ONE contiguous range of the CSR, so the warp copies it into shared memory with coalesced
// cp.async (every fetched sector is fully used; the thread-per-row kernel above re-fetches a 32-byte sector for
// each 4-byte read: 195 GB of DRAM traffic for 16 GB of data) and the lanes then run the register-cursor merge
// out of shared memory, branch-free: a step is min over own cursors -> min over the row's lanes (shuffles) ->
// every cursor on the minimum contributes and reloads.  Splitting a row over LPR lanes divides both the staging
// footprint per warp (more resident warps: the loop is latency-bound) and the serial work per step.
// A block whose segments do not fit (hub rows) runs the same loop on global memory.
// Sum order: ONE fp64 FMA chain over the row's cursors in ascending source-slice order, handed from lane to
// lane of the row -- every variant produces the bits of the thread-per-row kernel.
template <int CB, int LPR, bool COUNT_ONLY, typename VT>
__device__ __forceinline__ int64_t merge_loop_global(const int32_t *__restrict__ colp, const VT *__restrict__ valp,
                                               const int64_t (&gb)[CB], int32_t (&p)[CB], const int32_t (&e)[CB],
                                               const double (&w)[CB], bool writer, int64_t obase,
                                               int32_t *__restrict__ out_col, VT *__restrict__ out_val) {
    int32_t cur[CB];
    VT v[CB];
#pragma unroll
    for (int k = 0; k < CB; ++k) {
        const bool in = p[k] < e[k];
        cur[k] = in ? colp[gb[k] + p[k]] : INT_MAX;
        v[k] = (!COUNT_ONLY && in) ? valp[gb[k] + p[k]] : (VT)0;
    }
    int64_t count = 0;
    while (true) {
        int m = cur[0];
#pragma unroll
        for (int k = 1; k < CB; ++k) m = min(m, cur[k]);
#pragma unroll
        for (int d = 1; d < LPR; d <<= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, d));
        const bool row_done = m == INT_MAX;
        if (LPR == 1) {
            if (row_done) break;
        } else {
            if (__all_sync(0xffffffffu, row_done)) break;
        }
        const int mm = (LPR > 1 && row_done) ? -1 : m;                      // finished rows match nothing
        double acc = 0.0;
        if (!COUNT_ONLY) {
            // ascending-slot chain across the row's lanes (see merge_loop_smem)
#pragma unroll
            for (int sI = 0; sI < LPR; ++sI) {
                double part = acc;
#pragma unroll
                for (int k = 0; k < CB; ++k)
                    if (cur[k] == mm) part = fma(w[k], (double)v[k], part);
                acc = LPR == 1 ? part : __shfl_sync(0xffffffffu, part, (threadIdx.x & ~(LPR - 1)) + sI);
            }
        }
#pragma unroll
        for (int k = 0; k < CB; ++k) p[k] += (cur[k] == mm) ? 1 : 0;
#pragma unroll
        for (int k = 0; k < CB; ++k) {
            const bool in = p[k] < e[k];
            cur[k] = in ? colp[gb[k] + p[k]] : INT_MAX;
            if (!COUNT_ONLY) v[k] = in ? valp[gb[k] + p[k]] : (VT)0;
        }
        if (!COUNT_ONLY) {
            if (writer && !row_done) {
                out_col[obase + count] = m;
                out_val[obase + count] = (VT)acc;
            }
        }
        count += row_done ? 0 : 1;
    }
    return count;
}

// Staged entries are interleaved {col, val}: 4 B (count pass), 8 B (fp32) or 16 B (fp64: col, pad, val), so a
// cursor reload is ONE shared load at one address; cursors are absolute shared byte addresses.
template <bool COUNT_ONLY, typename VT>
struct StageEntry {
    static constexpr int SIZE = COUNT_ONLY ? 4 : (sizeof(VT) == 4 ? 8 : 16);
    static constexpr int VAL_OFF = sizeof(VT) == 4 ? 4 : 8;
};

// Register image of a staged value: fp32 values travel as raw bits so one v2.b32 load fills {col, val}.
template <typename VT>
struct ValReg {
    using type = int32_t;
};
template <>
struct ValReg<double> {
    using type = double;
};

// One cursor, one step, as straight-line predicated PTX (the compiler's version of the same C spends ~14
// instructions per cursor on selects and register moves; this is 7):
//   hit = (col under the cursor == the step's minimum);  if hit: acc += w * val (fp64 FMA), cursor += 1 entry;
//   reload {col, val} at the cursor, or INT_MAX when the row's list is exhausted.
// With advance == false it is just the (re)load.
template <bool COUNT_ONLY, typename VT, bool FMA = true>
__device__ __forceinline__ void cursor_step(int32_t &c, typename ValReg<VT>::type &v, uint32_t &p, uint32_t e, int mm,
                                            double w, double &acc) {
    if (COUNT_ONLY) {
        asm volatile(
            "{\n .reg .pred h, q;\n"
            " setp.eq.s32 h, %0, %3;\n"
            " @h add.u32 %1, %1, 4;\n"
            " setp.lt.u32 q, %1, %2;\n"
            " mov.b32 %0, 0x7fffffff;\n"
            " @q ld.shared.b32 %0, [%1];\n}"
            : "+r"(c), "+r"(p)
            : "r"(e), "r"(mm));
    } else if (!FMA && sizeof(VT) == 4) {
        asm volatile(
            "{\n .reg .pred h, q;\n"
            " setp.eq.s32 h, %0, %4;\n"
            " @h add.u32 %2, %2, 8;\n"
            " setp.lt.u32 q, %2, %3;\n"
            " mov.b32 %0, 0x7fffffff;\n"
            " @q ld.shared.v2.b32 {%0, %1}, [%2];\n}"
            : "+r"(c), "+r"(*reinterpret_cast<int32_t *>(&v)), "+r"(p)
            : "r"(e), "r"(mm));
    } else if (!FMA) {
        asm volatile(
            "{\n .reg .pred h, q;\n"
            " setp.eq.s32 h, %0, %4;\n"
            " @h add.u32 %2, %2, 16;\n"
            " setp.lt.u32 q, %2, %3;\n"
            " mov.b32 %0, 0x7fffffff;\n"
            " @q ld.shared.b32 %0, [%2];\n"
            " @q ld.shared.b64 %1, [%2+8];\n}"
            : "+r"(c), "+d"(*reinterpret_cast<double *>(&v)), "+r"(p)
            : "r"(e), "r"(mm));
    } else if (sizeof(VT) == 4) {
        asm volatile(
            "{\n .reg .pred h, q;\n .reg .f32 vf;\n .reg .f64 vd;\n"
            " setp.eq.s32 h, %0, %5;\n"
            " mov.b32 vf, %1;\n"
            " selp.f32 vf, vf, 0f00000000, h;\n"          // w * (+0) leaves the fp64 chain unchanged
            " cvt.f64.f32 vd, vf;\n"
            " fma.rn.f64 %3, %6, vd, %3;\n"
            " @h add.u32 %2, %2, 8;\n"
            " setp.lt.u32 q, %2, %4;\n"
            " mov.b32 %0, 0x7fffffff;\n"
            " @q ld.shared.v2.b32 {%0, %1}, [%2];\n}"
            : "+r"(c), "+r"(*reinterpret_cast<int32_t *>(&v)), "+r"(p), "+d"(acc)
            : "r"(e), "r"(mm), "d"(w));
    } else {
        asm volatile(
            "{\n .reg .pred h, q;\n .reg .f64 vd;\n"
            " setp.eq.s32 h, %0, %5;\n"
            " selp.f64 vd, %1, 0d0000000000000000, h;\n"
            " fma.rn.f64 %3, %6, vd, %3;\n"
            " @h add.u32 %2, %2, 16;\n"
            " setp.lt.u32 q, %2, %4;\n"
            " mov.b32 %0, 0x7fffffff;\n"
            " @q ld.shared.b32 %0, [%2];\n"
            " @q ld.shared.b64 %1, [%2+8];\n}"
            : "+r"(c), "+d"(*reinterpret_cast<double *>(&v)), "+r"(p), "+d"(acc)
            : "r"(e), "r"(mm), "d"(w));
    }
}

// the accumulate half of cursor_step alone (split rows chain their lanes' partial sums one after another)
template <typename VT>
__device__ __forceinline__ void cursor_fma(int32_t c, typename ValReg<VT>::type v, int mm, double w, double &acc) {
    if (sizeof(VT) == 4) {
        asm volatile(
            "{\n .reg .pred h;\n .reg .f32 vf;\n .reg .f64 vd;\n"
            " setp.eq.s32 h, %1, %3;\n"
            " mov.b32 vf, %2;\n"
            " selp.f32 vf, vf, 0f00000000, h;\n"
            " cvt.f64.f32 vd, vf;\n"
            " fma.rn.f64 %0, %4, vd, %0;\n}"
            : "+d"(acc)
            : "r"(c), "r"(*reinterpret_cast<const int32_t *>(&v)), "r"(mm), "d"(w));
    } else {
        asm volatile(
            "{\n .reg .pred h;\n .reg .f64 vd;\n"
            " setp.eq.s32 h, %1, %3;\n"
            " selp.f64 vd, %2, 0d0000000000000000, h;\n"
            " fma.rn.f64 %0, %4, vd, %0;\n}"
            : "+d"(acc)
            : "r"(c), "d"(*reinterpret_cast<const double *>(&v)), "r"(mm), "d"(w));
    }
}

template <int CB, int LPR, bool COUNT_ONLY, typename VT>
__device__ __forceinline__ int64_t merge_loop_smem(uint32_t (&p)[CB], const uint32_t (&e)[CB], const double (&w)[CB],
                                                   bool writer, int64_t obase, int32_t *__restrict__ out_col,
                                                   VT *__restrict__ out_val) {
    int32_t cur[CB];
    typename ValReg<VT>::type v[CB];
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < CB; ++k) {
        v[k] = 0;
        cur[k] = -2;                                                        // matches nothing: plain load
        cursor_step<COUNT_ONLY, VT>(cur[k], v[k], p[k], e[k], -1, 0.0, acc);
    }
    int32_t *oc = out_col + obase;
    VT *ov = out_val + obase;
    int64_t count = 0;
    while (true) {
        int m = cur[0];
#pragma unroll
        for (int k = 1; k < CB; ++k) m = min(m, cur[k]);
#pragma unroll
        for (int d = 1; d < LPR; d <<= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, d));
        const bool row_done = m == INT_MAX;
        if (LPR == 1) {
            if (row_done) break;
        } else {
            if (__all_sync(0xffffffffu, row_done)) break;
        }
        const int mm = (LPR > 1 && row_done) ? -1 : m;                      // finished rows match nothing
        acc = 0.0;
        if (LPR == 1) {
#pragma unroll
            for (int k = 0; k < CB; ++k) cursor_step<COUNT_ONLY, VT>(cur[k], v[k], p[k], e[k], mm, w[k], acc);
        } else {
            // one chain over the row's B cursors in ascending slot order, exactly the sum the single-lane
            // variants form: lane `s` of the row continues from what lane s-1 handed over
            if (!COUNT_ONLY) {
#pragma unroll
                for (int sI = 0; sI < LPR; ++sI) {
                    double part = acc;
#pragma unroll
                    for (int k = 0; k < CB; ++k) cursor_fma<VT>(cur[k], v[k], mm, w[k], part);
                    acc = __shfl_sync(0xffffffffu, part, (threadIdx.x & ~(LPR - 1)) + sI);
                }
            }
            double unused = 0.0;
#pragma unroll
            for (int k = 0; k < CB; ++k)
                cursor_step<COUNT_ONLY, VT, false>(cur[k], v[k], p[k], e[k], mm, 0.0, unused);
        }
        if (!COUNT_ONLY) {
            if (writer && !row_done) {
                *oc = m;
                *ov = (VT)acc;
            }
            oc += row_done ? 0 : 1;
            ov += row_done ? 0 : 1;
        }
        count += row_done ? 0 : 1;
    }
    return count;
}

template <int B, int LPR, bool COUNT_ONLY, typename VT>
__global__ void __launch_bounds__(32) merge_rows_staged(const int64_t *__restrict__ in_rowptr,
                                                        const int32_t *__restrict__ in_col,
                                                        const VT *__restrict__ in_val, int T_out, int halo, int64_t N,
                                                        const double *__restrict__ band_w, int b, int cap,
                                                        int64_t *__restrict__ out_counts,
                                                        const int64_t *__restrict__ out_rowptr,
                                                        int32_t *__restrict__ out_col, VT *__restrict__ out_val) {
    constexpr int RW = 32 / LPR;                 // rows per warp
    constexpr int CB = (B + LPR - 1) / LPR;      // cursors per lane
    extern __shared__ __align__(16) uint8_t stage_raw[];
    constexpr int ES = StageEntry<COUNT_ONLY, VT>::SIZE;
    constexpr int VO = StageEntry<COUNT_ONLY, VT>::VAL_OFF;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stage_raw);   // [B][cap] entries of ES bytes
    const int lane = threadIdx.x;
    const int sub = lane % LPR;                  // which share of the row's cursors this lane holds
    const int rl = lane / LPR;                   // row of the block
    const int64_t nblk = (N + RW - 1) / RW;
    const int64_t n_tasks = (int64_t)T_out * nblk;
    for (int64_t task = blockIdx.x; task < n_tasks; task += gridDim.x) {
        const int t = (int)(task / nblk);
        const int64_t r0 = (task - (int64_t)t * nblk) * RW;
        const int64_t i = r0 + rl;
        const bool live = i < N;
        const int64_t ic = live ? i : N - 1;                                // clamp for the pointer loads
        // phase 1: row pointers of this lane's cursors -- all loads are issued before any is consumed.
        // Cursor slot kk = sub*CB + k walks source slice halo + t - (B-1-kk): ascending slot = ascending slice.
        int64_t p0s[CB], p1s[CB];
        double w[CB];
#pragma unroll
        for (int k = 0; k < CB; ++k) {
            const int kk = sub * CB + k;
            const int l = B - 1 - kk;
            p0s[k] = p1s[k] = 0;
            w[k] = 0.0;
            if (l >= 0 && l < b) {
                const int sl = halo + t - l;
                const double wl = band_w[(int64_t)t * b + l];
                if (sl >= 0 && wl != 0.0) {
                    w[k] = wl;
                    p0s[k] = in_rowptr[(int64_t)sl * N + ic];
                    p1s[k] = in_rowptr[(int64_t)sl * N + ic + 1];
                }
            }
            if (!live) p0s[k] = p1s[k];                                     // padding lanes own empty rows
        }
        // phase 2: segment bounds of every slot via shuffles (slot kk lives in sub-lane kk / CB), then the
        // staging copies as fire-and-forget cp.async: all B segments are in flight together, one wait for the lot
        bool fits = true;                                                   // warp-uniform
        int64_t seg0s[CB];
#pragma unroll
        for (int kk = 0; kk < CB * LPR; ++kk) {
            const int own = kk / CB, k = kk % CB;                           // compile-time after unrolling
            const int64_t seg0 = __shfl_sync(0xffffffffu, p0s[k], own);
            const int64_t seg1 = __shfl_sync(0xffffffffu, p1s[k], (RW - 1) * LPR + own);
            if (own == sub) seg0s[k] = seg0;
            fits = fits && (seg1 - seg0 <= cap);
        }
        if (fits) {
#pragma unroll
            for (int kk = 0; kk < CB * LPR; ++kk) {
                const int own = kk / CB, k = kk % CB;
                const int64_t seg0 = __shfl_sync(0xffffffffu, seg0s[k], own);
                const int64_t seg1 = __shfl_sync(0xffffffffu, p1s[k], (RW - 1) * LPR + own);
                const int64_t len = seg1 - seg0;
                uint32_t dst = sbase + (uint32_t)(kk * cap + lane) * ES;
                for (int64_t q = lane; q < len; q += 32, dst += 32 * ES) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(in_col + seg0 + q)
                                 : "memory");
                    if (!COUNT_ONLY) {
                        if (sizeof(VT) == 4)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + VO),
                                         "l"(in_val + seg0 + q)
                                         : "memory");
                        else
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + VO),
                                         "l"(in_val + seg0 + q)
                                         : "memory");
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        const bool writer = live && sub == 0;
        const int64_t obase = (COUNT_ONLY || !writer) ? 0 : out_rowptr[(int64_t)t * N + i];
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        int64_t count;
        if (fits) {
            uint32_t p[CB], e[CB];
#pragma unroll
            for (int k = 0; k < CB; ++k) {
                p[k] = sbase + (uint32_t)((sub * CB + k) * cap + (int32_t)(p0s[k] - seg0s[k])) * ES;
                e[k] = p[k] + (uint32_t)(p1s[k] - p0s[k]) * ES;
            }
            count = merge_loop_smem<CB, LPR, COUNT_ONLY, VT>(p, e, w, writer, obase, out_col, out_val);
        } else {
            int32_t p[CB], e[CB];
#pragma unroll
            for (int k = 0; k < CB; ++k) {
                p[k] = 0;
                e[k] = (int32_t)(p1s[k] - p0s[k]);
            }
            count = merge_loop_global<CB, LPR, COUNT_ONLY, VT>(in_col, in_val, p0s, p, e, w, writer, obase, out_col,
                                                               out_val);
        }
        if (COUNT_ONLY && writer) out_counts[(int64_t)t * N + i] = count;
        __syncwarp();                                                       // staging buffers are reused by the next task
    }
}

// ---- union-list variant of the tiled transform (default when the caller passes a workspace) ---------------------
// The count pass already walks the union pattern of every (group of TT output slices) x (32-row block) task; it
// records it -- column and hit mask of every union entry, iteration-major, plus the per-row union length -- in a
// fixed-stride workspace record, and the fill pass (fill_from_union) never merges again: it turns the record into
// values with one LANE per union entry.  Workspace: [header 256 B: int n_overflow_tasks][n_tasks records].
constexpr size_t UNION_HEADER = 256;
__host__ __device__ __forceinline__ size_t union_stride(int qc) { return 64 + (size_t)qc * 192; }
struct UnionTask {
    int g;
    int64_t blk;
};
__device__ __forceinline__ UnionTask union_task(int64_t task, int64_t nblk, int n_groups, bool block_major) {
    UnionTask u;
    if (block_major) {
        u.blk = task / n_groups;
        u.g = (int)(task - u.blk * n_groups);
    } else {
        u.g = (int)(task / nblk);
        u.blk = task - (int64_t)u.g * nblk;
    }
    return u;
}

// Time-tiled fill (default for fp32 values, b <= 12, short rows; TMGCN_MERGE_TT=0 disables).  TT = 4 consecutive output slices of the
// same 32 rows share B-1 of their B source slices, so ONE merge over the NS = B-1+TT sources feeds all four:
// a step takes the smallest pending column, every cursor sitting on it hands over its value, and output tt
// (whose window is slots tt .. tt+B-1) accumulates its own fp64 chain over its slots in ascending slice order
// -- bit for bit what the single-output kernels produce -- and is emitted iff one of its slots with a
// non-zero weight was hit.  Per output that is ~55 instructions instead of ~97, and the segments are staged
// once for four outputs.  Staging is a pool (segments packed back to back), not B fixed-size buffers.
template <int B, int TT, bool COUNT>
__global__ void __launch_bounds__(32) merge_rows_tiled(const int64_t *__restrict__ in_rowptr,
                                                       const int32_t *__restrict__ in_col,
                                                       const float *__restrict__ in_val, int T_out, int halo,
                                                       int64_t N, const double *__restrict__ band_w, int b,
                                                       int pool, int64_t *__restrict__ out_counts,
                                                       const int64_t *__restrict__ out_rowptr,
                                                       int32_t *__restrict__ out_col, float *__restrict__ out_val) {
    // COUNT: the plan pass on the same structure -- only the columns are staged (4-byte entries: half the
    // shared memory per warp, twice the resident warps), a step is the min tree + one predicated cursor step per
    // source, and output tt counts a column iff one of its non-zero-weight slots was hit.  ~7x fewer
    // instructions per output than the thread-per-row count pass (whose per-cursor branches diverge) and no
    // sector over-fetch (that pass read 29.6 GB from DRAM for 3.6 GB of columns).
    constexpr int NS = B - 1 + TT;
    constexpr int ES = COUNT ? 4 : 8;            // bytes per staged entry
    extern __shared__ __align__(16) uint8_t stage_raw[];
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stage_raw);   // `pool` entries of {col[, val]}
    const int lane = threadIdx.x;
    const int64_t nblk = (N + 31) / 32;
    const int n_groups = (T_out + TT - 1) / TT;
    const int64_t n_tasks = (int64_t)n_groups * nblk;
    for (int64_t task = blockIdx.x; task < n_tasks; task += gridDim.x) {
        const int g = (int)(task / nblk);
        const int t0 = g * TT;
        const int64_t i = (task - (int64_t)g * nblk) * 32 + lane;
        const bool live = i < N;
        const int64_t ic = live ? i : N - 1;
        // weights of output tt on slot k (lag l = B-1-k+tt), zero outside its window / the band / the tensor
        double w[COUNT ? 1 : TT][COUNT ? 1 : B];
        uint32_t nz[TT];                         // slots whose weight for output tt is non-zero
        bool used[NS];
#pragma unroll
        for (int k = 0; k < NS; ++k) used[k] = false;
#pragma unroll
        for (int tt = 0; tt < TT; ++tt) {
            nz[tt] = 0;
#pragma unroll
            for (int j = 0; j < B; ++j) {        // j-th slot of the window: slot k = tt + j, lag l = B-1-j
                const int l = B - 1 - j;
                const int sl = halo + t0 + tt - l;
                double wl = 0.0;
                if (l < b && t0 + tt < T_out && sl >= 0) wl = band_w[(int64_t)(t0 + tt) * b + l];
                if (!COUNT) w[COUNT ? 0 : tt][COUNT ? 0 : j] = wl;
                if (wl != 0.0) {
                    nz[tt] |= 1u << (tt + j);
                    used[tt + j] = true;
                }
            }
        }
        // row pointers of the used slots
        int64_t p0s[NS], p1s[NS];
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            p0s[k] = p1s[k] = 0;
            if (used[k]) {                       // warp-uniform
                const int sl = halo + t0 - (B - 1) + k;
                p0s[k] = in_rowptr[(int64_t)sl * N + ic];
                p1s[k] = in_rowptr[(int64_t)sl * N + ic + 1];
            }
            if (!live) p0s[k] = p1s[k];
        }
        // pack the segments back to back; all copies in flight together
        int64_t seg0s[NS];
        int32_t off[NS + 1];
        off[0] = 0;
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            seg0s[k] = __shfl_sync(0xffffffffu, p0s[k], 0);
            const int64_t seg1 = __shfl_sync(0xffffffffu, p1s[k], 31);
            const int64_t len = seg1 - seg0s[k];
            off[k + 1] = off[k] + (int32_t)(len > (int64_t)pool ? pool + 1 : len);
        }
        const bool fits = off[NS] <= pool;       // warp-uniform
        if (fits) {
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                const int len = off[k + 1] - off[k];
                uint32_t dst = sbase + (uint32_t)(off[k] + lane) * ES;
                for (int q = lane; q < len; q += 32, dst += 32 * ES) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(in_col + seg0s[k] + q)
                                 : "memory");
                    if (!COUNT)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4),
                                     "l"(in_val + seg0s[k] + q)
                                     : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        int32_t *oc[COUNT ? 1 : TT];
        float *ov[COUNT ? 1 : TT];
        int64_t cnt[TT];
#pragma unroll
        for (int tt = 0; tt < TT; ++tt) {
            cnt[tt] = 0;
            if (!COUNT) {
                const int64_t ob = (live && t0 + tt < T_out) ? out_rowptr[(int64_t)(t0 + tt) * N + i] : 0;
                oc[COUNT ? 0 : tt] = out_col + ob;
                ov[COUNT ? 0 : tt] = out_val + ob;
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (fits) {
            uint32_t p[NS], e[NS];
            int32_t cur[NS];
            int32_t v[NS];
            double unused = 0.0;
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                p[k] = sbase + (uint32_t)(off[k] + (int32_t)(p0s[k] - seg0s[k])) * ES;
                e[k] = p[k] + (uint32_t)(p1s[k] - p0s[k]) * ES;
                v[k] = 0;
                cur[k] = -2;
                if (COUNT)
                    cursor_step<true, float>(cur[k], v[k], p[k], e[k], -1, 0.0, unused);
                else
                    cursor_step<false, float, false>(cur[k], v[k], p[k], e[k], -1, 0.0, unused);
            }
            while (true) {
                int m = cur[0];
#pragma unroll
                for (int k = 1; k < NS; ++k) m = min(m, cur[k]);
                if (m == INT_MAX) break;
                uint32_t hits = 0;
                double vd[COUNT ? 1 : NS];
#pragma unroll
                for (int k = 0; k < NS; ++k) {
                    const bool hit = cur[k] == m;
                    hits |= hit ? (1u << k) : 0u;
                    if (!COUNT) vd[COUNT ? 0 : k] = hit ? (double)__int_as_float(v[k]) : 0.0;  // +0 leaves a chain unchanged
                }
#pragma unroll
                for (int tt = 0; tt < TT; ++tt) {
                    if (COUNT) {
                        cnt[tt] += (hits & nz[tt]) ? 1 : 0;
                    } else {
                        double acc = 0.0;
#pragma unroll
                        for (int j = 0; j < B; ++j) acc = fma(w[COUNT ? 0 : tt][COUNT ? 0 : j], vd[COUNT ? 0 : tt + j], acc);
                        if (hits & nz[tt]) {
                            *oc[COUNT ? 0 : tt] = m;
                            *ov[COUNT ? 0 : tt] = (float)acc;
                            ++oc[COUNT ? 0 : tt];
                            ++ov[COUNT ? 0 : tt];
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < NS; ++k) {
                    if (COUNT)
                        cursor_step<true, float>(cur[k], v[k], p[k], e[k], m, 0.0, unused);
                    else
                        cursor_step<false, float, false>(cur[k], v[k], p[k], e[k], m, 0.0, unused);
                }
            }
        } else {
            // hub blocks: the same merge on global memory
            int64_t gp[NS], ge[NS];
            int32_t cur[NS];
            float v[NS];
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                gp[k] = p0s[k];
                ge[k] = p1s[k];
                const bool in = gp[k] < ge[k];
                cur[k] = in ? in_col[gp[k]] : INT_MAX;
                v[k] = (!COUNT && in) ? in_val[gp[k]] : 0.f;
            }
            while (true) {
                int m = cur[0];
#pragma unroll
                for (int k = 1; k < NS; ++k) m = min(m, cur[k]);
                if (m == INT_MAX) break;
                uint32_t hits = 0;
                double vd[COUNT ? 1 : NS];
#pragma unroll
                for (int k = 0; k < NS; ++k) {
                    const bool hit = cur[k] == m;
                    hits |= hit ? (1u << k) : 0u;
                    if (!COUNT) vd[COUNT ? 0 : k] = hit ? (double)v[k] : 0.0;
                    gp[k] += hit ? 1 : 0;
                }
#pragma unroll
                for (int tt = 0; tt < TT; ++tt) {
                    if (COUNT) {
                        cnt[tt] += (hits & nz[tt]) ? 1 : 0;
                    } else {
                        double acc = 0.0;
#pragma unroll
                        for (int j = 0; j < B; ++j) acc = fma(w[COUNT ? 0 : tt][COUNT ? 0 : j], vd[COUNT ? 0 : tt + j], acc);
                        if (hits & nz[tt]) {
                            *oc[COUNT ? 0 : tt] = m;
                            *ov[COUNT ? 0 : tt] = (float)acc;
                            ++oc[COUNT ? 0 : tt];
                            ++ov[COUNT ? 0 : tt];
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < NS; ++k) {
                    const bool in = gp[k] < ge[k];
                    cur[k] = in ? in_col[gp[k]] : INT_MAX;
                    if (!COUNT) v[k] = in ? in_val[gp[k]] : 0.f;
                }
            }
        }
        if (COUNT) {
#pragma unroll
            for (int tt = 0; tt < TT; ++tt)
                if (live && t0 + tt < T_out) out_counts[(int64_t)(t0 + tt) * N + i] = cnt[tt];
        }
        __syncwarp();                            // the pool is reused by the next task
    }
}

// Count pass of the union-list variant: merge_rows_tiled<COUNT> plus the record, restructured around what its
// profile showed (ncu, benchmark shard: 6.2 K warp-instructions per task, of which the merge loop was 3.6 K):
//   * a CTA keeps ONE group of output slices and walks the row blocks (grid = n_groups x walkers), so the band
//     weights' non-zero pattern is derived once per CTA, not per task; the CTAs of the groups of one row block run
//     side by side and share its source segments through L2;
//   * the source segments are staged by the copy engine: one elected lane issues a 1-D bulk copy
//     (cp.async.bulk, mbarrier completion) per segment over the enclosing 16-byte-aligned range, instead of ~140
//     4-byte cp.async per lane;
//   * every merge iteration appends the union entry -- column and 16-bit hit mask -- to the task's record,
//     iteration-major ([q][lane]: the lanes still merging write one coalesced line each).
// Count-pass cursor step that also records the hit.  The merge loop is bound by the half-rate integer ALU pipe
// (ncu: alu 65 %, fma 13 %), so the two predicated updates -- cursor += 4, hits |= bit -- are written as
// multiply-adds with a run-time 1 (`unit`): ptxas cannot fold them into adds and issues them on the FMA pipe.
template <int K>
__device__ __forceinline__ void cursor_step_hits(int32_t &c, uint32_t &p, uint32_t e, int mm, uint32_t &hits,
                                                 uint32_t unit) {
    asm volatile(
        "{\n .reg .pred h, q;\n"
        " setp.eq.s32 h, %0, %4;\n"
        " @h mad.lo.u32 %1, %5, 4, %1;\n"
        " @h mad.lo.u32 %2, %5, %6, %2;\n"
        " setp.lt.u32 q, %1, %3;\n"
        " mov.b32 %0, 0x7fffffff;\n"
        " @q ld.shared.b32 %0, [%1];\n}"
        : "+r"(c), "+r"(p), "+r"(hits)
        : "r"(e), "r"(mm), "r"(unit), "n"(1u << K));
}
template <int... K>
__device__ __forceinline__ void cursor_steps_hits(std::integer_sequence<int, K...>, int32_t (&c)[sizeof...(K)],
                                                  uint32_t (&p)[sizeof...(K)], const uint32_t (&e)[sizeof...(K)],
                                                  int mm, uint32_t &hits, uint32_t unit) {
    (cursor_step_hits<K>(c[K], p[K], e[K], mm, hits, unit), ...);
}

__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    for (int spin = 0; spin < (1 << 26); ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();                                    // a lost copy must fail the launch, not hang the device
}

template <int B, int TT>
__global__ void __launch_bounds__(32, 10) count_union(const int64_t *__restrict__ in_rowptr,
                                                  const int32_t *__restrict__ in_col, int64_t in_nnz, int T_out,
                                                  int halo, int64_t N, const double *__restrict__ band_w, int b,
                                                  int pool, int64_t *__restrict__ out_counts,
                                                  uint8_t *__restrict__ utmp, int qc) {
    constexpr int NS = B - 1 + TT;
    static_assert(NS <= 16, "hit masks are 16 bits");
    extern __shared__ __align__(16) uint8_t stage_raw[];
    uint64_t *const bar = reinterpret_cast<uint64_t *>(stage_raw);          // 16 bytes reserved in front of the pool
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stage_raw) + 16;   // `pool` column entries
    const int lane = threadIdx.x;
    const int64_t nblk = (N + 31) / 32;
    const int n_groups = (T_out + TT - 1) / TT;
    const int g = (int)(blockIdx.x % (unsigned)n_groups);
    const int64_t walker = blockIdx.x / (unsigned)n_groups, n_walkers = gridDim.x / (unsigned)n_groups;
    const int t0 = g * TT;
    const size_t stride = union_stride(qc);
    const int64_t nnz4 = in_nnz & ~(int64_t)3;                              // bulk copies stop at the last whole quad
    const uint32_t unit = qc > 0 ? 1u : 0u;                                 // a 1 the compiler cannot see (cursor_step_hits)
    // slots with a non-zero weight for output tt (slot k = tt + j holds lag l = B-1-j), once per CTA
    uint32_t nz[TT];
    uint32_t used_bits = 0;
#pragma unroll
    for (int tt = 0; tt < TT; ++tt) {
        nz[tt] = 0;
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int l = B - 1 - j;
            const int sl = halo + t0 + tt - l;
            double wl = 0.0;
            if (l < b && t0 + tt < T_out && sl >= 0) wl = band_w[(int64_t)(t0 + tt) * b + l];
            if (wl != 0.0) nz[tt] |= 1u << (tt + j);
        }
        used_bits |= nz[tt];
    }
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0;
    for (int64_t blk = walker; blk < nblk; blk += n_walkers) {
        const int64_t task = blk * n_groups + g;
        const int64_t i = blk * 32 + lane;
        const bool live = i < N;
        const int64_t ic = live ? i : N - 1;
        uint8_t *const urec = utmp + UNION_HEADER + (size_t)task * stride;
        // row pointers of the used slots
        int64_t p0s[NS], p1s[NS];
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            p0s[k] = p1s[k] = 0;
            if (used_bits & (1u << k)) {         // warp-uniform
                const int sl = halo + t0 - (B - 1) + k;
                p0s[k] = in_rowptr[(int64_t)sl * N + ic];
                p1s[k] = in_rowptr[(int64_t)sl * N + ic + 1];
            }
            if (!live) p0s[k] = p1s[k];
        }
        // segment k of the block = [seg0, seg1) of in_col; it is staged as the enclosing aligned range [a0, a1)
        // at pool offset soff[k] (a multiple of 4 entries), so source and destination are 16-byte aligned
        int64_t a0s[NS];
        int32_t soff[NS + 1], nbulk[NS], tail0[NS], ntail[NS];
        soff[0] = 0;
        uint32_t total_bytes = 0;
        bool fits = true;
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            const int64_t seg0 = __shfl_sync(0xffffffffu, p0s[k], 0);
            const int64_t seg1 = __shfl_sync(0xffffffffu, p1s[k], 31);
            const int64_t a0 = seg0 & ~(int64_t)3;
            const int64_t a1 = (seg1 + 3) & ~(int64_t)3;
            const int64_t span = a1 - a0;
            a0s[k] = a0;
            fits = fits && span <= (int64_t)pool;
            const int64_t bulk_end = a1 < nnz4 ? a1 : nnz4;                 // entries [bulk_end, seg1) go by plain loads
            const int64_t nb = bulk_end > a0 ? bulk_end - a0 : 0;
            nbulk[k] = seg1 > seg0 ? (int32_t)(nb > (int64_t)pool ? pool : nb) : 0;
            const int64_t ts = seg0 > nnz4 ? seg0 : nnz4;
            tail0[k] = (int32_t)(ts - a0 > (int64_t)pool ? pool : ts - a0);
            ntail[k] = seg1 > ts ? (int32_t)(seg1 - ts) : 0;                // at most 3
            soff[k + 1] = soff[k] + (int32_t)(span > (int64_t)pool ? pool + 4 : span);
            total_bytes += (uint32_t)nbulk[k] * 4u;
        }
        fits = fits && soff[NS] <= pool;         // warp-uniform
        if (fits) {
            if (lane == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                                 (uint32_t)__cvta_generic_to_shared(bar)),
                             "r"(total_bytes)
                             : "memory");
#pragma unroll
                for (int k = 0; k < NS; ++k)
                    if (nbulk[k] > 0)
                        asm volatile(
                            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                sbase + (uint32_t)soff[k] * 4u),
                            "l"(in_col + a0s[k]), "r"((uint32_t)nbulk[k] * 4u),
                            "r"((uint32_t)__cvta_generic_to_shared(bar))
                            : "memory");
            }
            // the last partial quad of the whole column array (only the final segment can reach it)
#pragma unroll
            for (int k = 0; k < NS; ++k)
                if (lane < ntail[k])
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + (uint32_t)(soff[k] + tail0[k] + lane) * 4u),
                                 "r"(in_col[a0s[k] + tail0[k] + lane])
                                 : "memory");
            mbar_wait_bounded(bar, phase);
            phase ^= 1u;
        }
        __syncwarp();
        int64_t cnt[TT];
#pragma unroll
        for (int tt = 0; tt < TT; ++tt) cnt[tt] = 0;
        int32_t *const rcol = reinterpret_cast<int32_t *>(urec + 64) + lane;
        uint16_t *const rmask = reinterpret_cast<uint16_t *>(urec + 64 + (size_t)qc * 128) + lane;
        int q = 0;                               // union entries of this lane's row so far
        if (fits) {
            uint32_t p[NS], e[NS];
            int32_t cur[NS];
            int32_t v[NS];
            double unused = 0.0;
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                p[k] = sbase + (uint32_t)(soff[k] + (int32_t)(p0s[k] - a0s[k])) * 4u;
                e[k] = p[k] + (uint32_t)(p1s[k] - p0s[k]) * 4u;
                v[k] = 0;
                cur[k] = -2;
                cursor_step<true, float>(cur[k], v[k], p[k], e[k], -1, 0.0, unused);
            }
            while (true) {
                int m = cur[0];
#pragma unroll
                for (int k = 1; k < NS; ++k) m = min(m, cur[k]);
                if (m == INT_MAX) break;
                uint32_t hits = 0;
                cursor_steps_hits(std::make_integer_sequence<int, NS>{}, cur, p, e, m, hits, unit);
                if (q < qc) {
                    rcol[q * 32] = m;
                    rmask[q * 32] = (uint16_t)hits;
                }
                ++q;
#pragma unroll
                for (int tt = 0; tt < TT; ++tt) cnt[tt] += (hits & nz[tt]) ? 1 : 0;
            }
        } else {
            // hub blocks: the same merge on global memory; their record is marked overflowed below
            int64_t gp[NS], ge[NS];
            int32_t cur[NS];
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                gp[k] = p0s[k];
                ge[k] = p1s[k];
                cur[k] = gp[k] < ge[k] ? in_col[gp[k]] : INT_MAX;
            }
            while (true) {
                int m = cur[0];
#pragma unroll
                for (int k = 1; k < NS; ++k) m = min(m, cur[k]);
                if (m == INT_MAX) break;
                uint32_t hits = 0;
#pragma unroll
                for (int k = 0; k < NS; ++k) {
                    const bool hit = cur[k] == m;
                    hits |= hit ? (1u << k) : 0u;
                    gp[k] += hit ? 1 : 0;
                }
#pragma unroll
                for (int tt = 0; tt < TT; ++tt) cnt[tt] += (hits & nz[tt]) ? 1 : 0;
#pragma unroll
                for (int k = 0; k < NS; ++k) cur[k] = gp[k] < ge[k] ? in_col[gp[k]] : INT_MAX;
            }
            q = qc + 1;
        }
        {
            const bool over = q > qc;            // a row longer than the record (or a hub block): fill_union_overflow
            reinterpret_cast<uint16_t *>(urec)[lane] = over ? (uint16_t)0xFFFF : (uint16_t)q;
            if (__any_sync(0xffffffffu, over) && lane == 0) atomicAdd(reinterpret_cast<int *>(utmp), 1);
        }
#pragma unroll
        for (int tt = 0; tt < TT; ++tt)
            if (live && t0 + tt < T_out) out_counts[(int64_t)(t0 + tt) * N + i] = cnt[tt];
        __syncwarp();                            // the pool is reused by the next task
    }
}

// Fill pass of the union-list variant.  A warp owns one task (TT output slices x 32 rows).  The task's record
// gives the union pattern; its entries are walked in row-major order with one LANE per entry.  Because every
// source slice stores the block's rows back to back and in the same (row, column) order as the union list, the
// value of source slot k for an entry with hit bit k is at
//     first entry of the block in that slice + #(earlier union entries with bit k)
// -- a running count kept per slot plus a ballot/popcount rank inside the chunk: no merge, no compare, and the
// gathers of a warp are monotone (coalesced).  Output tt then forms ITS fp64 FMA chain over the slots of its window
// in ascending slice order (absent sources contribute w * (+0): the bits of the merging kernels) and the entries
// present in tt are written at base_tt + rank: again a ballot/popcount, fully coalesced stores.
// Tasks whose record overflowed (a row longer than qc, or a block that did not fit the count pass's staging pool)
// are re-merged from global memory (tiled_task_global): same sums, same order.
template <int B, int TT, typename VT>
__device__ __forceinline__ void tiled_task_global(const int64_t *__restrict__ in_rowptr,
                                                  const int32_t *__restrict__ in_col, const VT *__restrict__ in_val,
                                                  int T_out, int halo, int64_t N, const double *__restrict__ band_w,
                                                  int b, int t0, int64_t i, bool live,
                                                  const int64_t *__restrict__ out_rowptr,
                                                  int32_t *__restrict__ out_col, VT *__restrict__ out_val) {
    constexpr int NS = B - 1 + TT;
    const int64_t ic = live ? i : N - 1;
    double w[TT][B];
    uint32_t nz[TT];
    bool used[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) used[k] = false;
#pragma unroll
    for (int tt = 0; tt < TT; ++tt) {
        nz[tt] = 0;
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int l = B - 1 - j;
            const int sl = halo + t0 + tt - l;
            double wl = 0.0;
            if (l < b && t0 + tt < T_out && sl >= 0) wl = band_w[(int64_t)(t0 + tt) * b + l];
            w[tt][j] = wl;
            if (wl != 0.0) {
                nz[tt] |= 1u << (tt + j);
                used[tt + j] = true;
            }
        }
    }
    int64_t gp[NS], ge[NS];
    int32_t cur[NS];
    VT v[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        gp[k] = ge[k] = 0;
        if (used[k] && live) {
            const int sl = halo + t0 - (B - 1) + k;
            gp[k] = in_rowptr[(int64_t)sl * N + ic];
            ge[k] = in_rowptr[(int64_t)sl * N + ic + 1];
        }
        const bool in = gp[k] < ge[k];
        cur[k] = in ? in_col[gp[k]] : INT_MAX;
        v[k] = in ? in_val[gp[k]] : (VT)0;
    }
    int32_t *oc[TT];
    VT *ov[TT];
#pragma unroll
    for (int tt = 0; tt < TT; ++tt) {
        const int64_t ob = (live && t0 + tt < T_out) ? out_rowptr[(int64_t)(t0 + tt) * N + i] : 0;
        oc[tt] = out_col + ob;
        ov[tt] = out_val + ob;
    }
    while (true) {
        int m = cur[0];
#pragma unroll
        for (int k = 1; k < NS; ++k) m = min(m, cur[k]);
        if (m == INT_MAX) break;
        uint32_t hits = 0;
        double vd[NS];
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            const bool hit = cur[k] == m;
            hits |= hit ? (1u << k) : 0u;
            vd[k] = hit ? (double)v[k] : 0.0;
            gp[k] += hit ? 1 : 0;
        }
#pragma unroll
        for (int tt = 0; tt < TT; ++tt) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < B; ++j) acc = fma(w[tt][j], vd[tt + j], acc);
            if (hits & nz[tt]) {
                *oc[tt] = m;
                *ov[tt] = (VT)acc;
                ++oc[tt];
                ++ov[tt];
            }
        }
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            const bool in = gp[k] < ge[k];
            cur[k] = in ? in_col[gp[k]] : INT_MAX;
            v[k] = in ? in_val[gp[k]] : (VT)0;
        }
    }
}

// One source slot of one chunk of the lane-per-entry walk, as straight-line PTX (the compiler's version of the same
// C spends 14 instructions per slot re-deriving the predicate and re-loading the base pointer):
//   p = (mask has bit K);  bal = ballot(p);  v = p ? in_val[sb + popc(bal & lanes below)] : +0;
//   sb += total of the ballot, taken as (rank + own bit) of the last lane by a shuffle: the quarter-rate XU pipe
//   (popc, cvt) is the busiest pipe of this kernel (ncu: 57 %), so the second popcount is avoided.
// (A cp.async variant that parks the gathered values in shared memory -- commit groups instead of the register
// scoreboards, which alias between the two chunks in flight -- measured the same 7.7-7.9 ms and was dropped.)
template <int K>
__device__ __forceinline__ void slot_gather(uint32_t mask, uint32_t lt, uint32_t &sb, const float *base, float &v) {
    asm volatile(
        "{\n .reg .pred p;\n .reg .b32 t, bal, r, i;\n .reg .b64 a;\n"
        " and.b32 t, %2, %5;\n"
        " setp.ne.u32 p, t, 0;\n"
        " vote.sync.ballot.b32 bal, p, 0xffffffff;\n"
        " and.b32 r, bal, %3;\n"
        " popc.b32 r, r;\n"
        " add.u32 i, r, %1;\n"
        " mov.f32 %0, 0f00000000;\n"
        " mad.wide.u32 a, i, 4, %4;\n"
        " @p ld.global.nc.f32 %0, [a];\n"
        " selp.u32 t, 1, 0, p;\n"
        " add.u32 t, t, r;\n"
        " shfl.sync.idx.b32 t, t, 31, 31, 0xffffffff;\n"
        " add.u32 %1, %1, t;\n}"
        : "=f"(v), "+r"(sb)
        : "r"(mask), "r"(lt), "l"(base), "n"(1u << K)
        : "memory");
}
template <int K>
__device__ __forceinline__ void slot_gather(uint32_t mask, uint32_t lt, uint32_t &sb, const double *base, double &v) {
    asm volatile(
        "{\n .reg .pred p;\n .reg .b32 t, bal, r, i;\n .reg .b64 a;\n"
        " and.b32 t, %2, %5;\n"
        " setp.ne.u32 p, t, 0;\n"
        " vote.sync.ballot.b32 bal, p, 0xffffffff;\n"
        " and.b32 r, bal, %3;\n"
        " popc.b32 r, r;\n"
        " add.u32 i, r, %1;\n"
        " mov.f64 %0, 0d0000000000000000;\n"
        " mad.wide.u32 a, i, 8, %4;\n"
        " @p ld.global.nc.f64 %0, [a];\n"
        " selp.u32 t, 1, 0, p;\n"
        " add.u32 t, t, r;\n"
        " shfl.sync.idx.b32 t, t, 31, 31, 0xffffffff;\n"
        " add.u32 %1, %1, t;\n}"
        : "=d"(v), "+r"(sb)
        : "r"(mask), "r"(lt), "l"(base), "n"(1u << K)
        : "memory");
}
template <typename VT, typename IdxT, int... K>
__device__ __forceinline__ void gather_slots(std::integer_sequence<int, K...>, uint32_t mask, uint32_t lt,
                                             IdxT (&sb)[sizeof...(K)], const VT *base, VT (&v)[sizeof...(K)]) {
    (slot_gather<K>(mask, lt, sb[K], base, v[K]), ...);
}
// One output slice of one chunk: entries present in it (mask & nz) are written at ob + rank, ob advances.
__device__ __forceinline__ void slot_emit(uint32_t mask, uint32_t nz, uint32_t lt, uint32_t &ob, int32_t *out_col,
                                          float *out_val, int32_t col, float val) {
    asm volatile(
        "{\n .reg .pred p;\n .reg .b32 t, bal, r, i;\n .reg .b64 a;\n"
        " and.b32 t, %1, %2;\n"
        " setp.ne.u32 p, t, 0;\n"
        " vote.sync.ballot.b32 bal, p, 0xffffffff;\n"
        " and.b32 r, bal, %3;\n"
        " popc.b32 r, r;\n"
        " add.u32 i, r, %0;\n"
        " mad.wide.u32 a, i, 4, %4;\n"
        " @p st.global.b32 [a], %6;\n"
        " mad.wide.u32 a, i, 4, %5;\n"
        " @p st.global.f32 [a], %7;\n"
        " selp.u32 t, 1, 0, p;\n"
        " add.u32 t, t, r;\n"
        " shfl.sync.idx.b32 t, t, 31, 31, 0xffffffff;\n"
        " add.u32 %0, %0, t;\n}"
        : "+r"(ob)
        : "r"(mask), "r"(nz), "r"(lt), "l"(out_col), "l"(out_val), "r"(col), "f"(val)
        : "memory");
}
__device__ __forceinline__ void slot_emit(uint32_t mask, uint32_t nz, uint32_t lt, uint32_t &ob, int32_t *out_col,
                                          double *out_val, int32_t col, double val) {
    asm volatile(
        "{\n .reg .pred p;\n .reg .b32 t, bal, r, i;\n .reg .b64 a;\n"
        " and.b32 t, %1, %2;\n"
        " setp.ne.u32 p, t, 0;\n"
        " vote.sync.ballot.b32 bal, p, 0xffffffff;\n"
        " and.b32 r, bal, %3;\n"
        " popc.b32 r, r;\n"
        " add.u32 i, r, %0;\n"
        " mad.wide.u32 a, i, 4, %4;\n"
        " @p st.global.b32 [a], %6;\n"
        " mad.wide.u32 a, i, 8, %5;\n"
        " @p st.global.f64 [a], %7;\n"
        " selp.u32 t, 1, 0, p;\n"
        " add.u32 t, t, r;\n"
        " shfl.sync.idx.b32 t, t, 31, 31, 0xffffffff;\n"
        " add.u32 %0, %0, t;\n}"
        : "+r"(ob)
        : "r"(mask), "r"(nz), "r"(lt), "l"(out_col), "l"(out_val), "r"(col), "d"(val)
        : "memory");
}
template <int B, int TT, typename VT, typename IdxT>
__global__ void __launch_bounds__(32, 12) fill_from_union(const int64_t *__restrict__ in_rowptr,
                                                      const int32_t *__restrict__ in_col,
                                                      const VT *__restrict__ in_val, int T_out, int halo, int64_t N,
                                                      const double *__restrict__ band_w, int b,
                                                      const uint8_t *__restrict__ utmp, int qc,
                                                      const int64_t *__restrict__ out_rowptr,
                                                      int32_t *__restrict__ out_col, VT *__restrict__ out_val) {
    constexpr int NS = B - 1 + TT;
    static_assert(NS <= 16, "hit masks are 16 bits");
    extern __shared__ __align__(16) uint8_t stage_raw[];
    // [col: row-major packed, qc x 32 x i32][mask: row-major packed, qc x 32 x u16][raw masks of the record, qc x 32 x u16]
    const int32_t *const s_col = reinterpret_cast<const int32_t *>(stage_raw);
    uint16_t *const s_mask = reinterpret_cast<uint16_t *>(stage_raw + (size_t)qc * 128);
    const uint16_t *const s_raw = reinterpret_cast<const uint16_t *>(stage_raw + (size_t)qc * 192);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stage_raw);
    const int lane = threadIdx.x;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t nblk = (N + 31) / 32;
    const int n_groups = (T_out + TT - 1) / TT;
    const size_t stride = union_stride(qc);
    // a CTA keeps ONE group of output slices and walks the row blocks (grid = n_groups x walkers, like count_union):
    // the weights of the group's outputs and their non-zero patterns are set up once, not per task
    const int g = (int)(blockIdx.x % (unsigned)n_groups);
    const int64_t walker = blockIdx.x / (unsigned)n_groups, n_walkers = gridDim.x / (unsigned)n_groups;
    const int t0 = g * TT;
    // weights of output tt on the j-th slot of its window (slot k = tt + j, lag l = B-1-j) -- as merge_rows_tiled
    double w[TT][B];
    uint32_t nz[TT];
    uint32_t used_bits = 0;
#pragma unroll
    for (int tt = 0; tt < TT; ++tt) {
        nz[tt] = 0;
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int l = B - 1 - j;
            const int sl = halo + t0 + tt - l;
            double wl = 0.0;
            if (l < b && t0 + tt < T_out && sl >= 0) wl = band_w[(int64_t)(t0 + tt) * b + l];
            w[tt][j] = wl;
            if (wl != 0.0) nz[tt] |= 1u << (tt + j);
        }
        used_bits |= nz[tt];
    }
    for (int64_t blk = walker; blk < nblk; blk += n_walkers) {
        const int64_t task = blk * n_groups + g;
        const int64_t r0 = blk * 32;
        const uint8_t *const urec = utmp + UNION_HEADER + (size_t)task * stride;
        const uint32_t ul = reinterpret_cast<const uint16_t *>(urec)[lane];
        if (__any_sync(0xffffffffu, ul == 0xFFFFu)) continue;                 // overflowed record: fill_union_overflow
        // row offsets of the row-major walk (exclusive scan of the union lengths) and the longest row
        uint32_t incl = ul;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t nb = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += nb;
        }
        const int Ltot = (int)__shfl_sync(0xffffffffu, incl, 31);
        uint32_t qmax = ul;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) qmax = max(qmax, __shfl_xor_sync(0xffffffffu, qmax, o));
        // stage the record TRANSPOSED: the record is iteration-major ([q][row], what the merging lanes could write
        // coalesced), the walk is row-major.  Columns go straight to their packed place (4-byte copies: lane = row,
        // a coalesced line per iteration); the 2-byte hit masks are copied raw (16-byte copies) and transposed from
        // shared memory once they have landed.  Everything is in flight together.
        const uint32_t roff = incl - ul;
        {
            const int32_t *src = reinterpret_cast<const int32_t *>(urec + 64) + lane;
            uint32_t dst = sbase + roff * 4;
            for (uint32_t q = 0; q < ul; ++q, src += 32, dst += 4)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
            const uint8_t *msrc = urec + 64 + (size_t)qc * 128;
            const int nm = (int)qmax * 4;                                    // 64 B of hit masks per iteration
            for (int x = lane; x < nm; x += 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + (uint32_t)qc * 192 +
                                                                               (uint32_t)x * 16),
                             "l"(msrc + (size_t)x * 16)
                             : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        // running positions: first entry of the block in every source slot, first output entry of the block
        // (IdxT = uint32_t: the variant serves tensors with fewer than 2^32 entries; one IMAD.WIDE per address)
        IdxT sb[NS];
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            sb[k] = 0;
            if (used_bits & (1u << k)) sb[k] = (IdxT)in_rowptr[(int64_t)(halo + t0 - (B - 1) + k) * N + r0];
        }
        IdxT ob[TT];
#pragma unroll
        for (int tt = 0; tt < TT; ++tt) ob[tt] = t0 + tt < T_out ? (IdxT)out_rowptr[(int64_t)(t0 + tt) * N + r0] : 0;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        for (uint32_t q = 0; q < ul; ++q) s_mask[roff + q] = s_raw[q * 32 + lane];
        __syncwarp();
        // The walk is software-pipelined: the front half of chunk c+1 (entry lookup, ranks, the NS gathers) is
        // issued before the back half of chunk c (fp64 chains, stores), so the gathers' latency hides behind it.
        struct Front {
            int32_t col;
            uint32_t mask;
            VT v[NS];
        };
        auto front = [&](int e0, Front &f) {
            const int e = e0 + lane;
            f.col = 0;
            f.mask = 0;
            if (e < Ltot) {
                f.col = s_col[e];
                f.mask = s_mask[e];
            }
            gather_slots<VT, IdxT>(std::make_integer_sequence<int, NS>{}, f.mask, lt, sb, in_val, f.v);
        };
        auto back = [&](const Front &f) {
            double vd[NS];
#pragma unroll
            for (int k = 0; k < NS; ++k) vd[k] = (double)f.v[k];            // +0 leaves a chain unchanged
#pragma unroll
            for (int tt = 0; tt < TT; ++tt) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < B; ++j) acc = fma(w[tt][j], vd[tt + j], acc);
                slot_emit(f.mask, nz[tt], lt, ob[tt], out_col, out_val, f.col, (VT)acc);
            }
        };
        Front fa, fb;
        if (Ltot > 0) front(0, fa);
        for (int e0 = 0; e0 < Ltot; e0 += 64) {                             // all conditions are warp-uniform
            const bool more1 = e0 + 32 < Ltot;
            if (more1) front(e0 + 32, fb);
            back(fa);
            if (!more1) break;
            if (e0 + 64 < Ltot) front(e0 + 64, fa);
            back(fb);
        }
        __syncwarp();                            // the staging buffers are reused by the next task
    }
}

// second launch of the union-list fill: the tasks whose record overflowed are merged from global memory (kept out
// of fill_from_union so that its register allocation is that of the lane-per-entry loop alone)
template <int B, int TT, typename VT>
__global__ void __launch_bounds__(128) fill_union_overflow(const int64_t *__restrict__ in_rowptr,
                                                           const int32_t *__restrict__ in_col,
                                                           const VT *__restrict__ in_val, int T_out, int halo,
                                                           int64_t N, const double *__restrict__ band_w, int b,
                                                           const uint8_t *__restrict__ utmp, int qc,
                                                           const int64_t *__restrict__ out_rowptr,
                                                           int32_t *__restrict__ out_col, VT *__restrict__ out_val) {
    const int lane = threadIdx.x & 31;
    const int64_t nblk = (N + 31) / 32;
    const int n_groups = (T_out + TT - 1) / TT;
    const int64_t n_tasks = (int64_t)n_groups * nblk;
    const size_t stride = union_stride(qc);
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t task = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; task < n_tasks; task += warps_total) {
        const uint32_t ul = reinterpret_cast<const uint16_t *>(utmp + UNION_HEADER + (size_t)task * stride)[lane];
        if (!__any_sync(0xffffffffu, ul == 0xFFFFu)) continue;
        const UnionTask ut = union_task(task, nblk, n_groups, true);
        const int64_t i = ut.blk * 32 + lane;
        tiled_task_global<B, TT, VT>(in_rowptr, in_col, in_val, T_out, halo, N, band_w, b, ut.g * TT, i, i < N,
                                     out_rowptr, out_col, out_val);
        __syncwarp();
    }
}

// largest record depth (union entries per row) a workspace of ws_bytes gives the n_tasks tasks; 0 = no union variant
static int union_qc(const void *ws, size_t ws_bytes, int64_t n_tasks) {
    if (ws == nullptr || n_tasks <= 0 || ws_bytes <= UNION_HEADER) return 0;
    const size_t per_task = (ws_bytes - UNION_HEADER) / (size_t)n_tasks;
    if (per_task < union_stride(8)) return 0;
    size_t qc = (per_task - 64) / 192;
    if (qc > 240) qc = 240;
    return (int)qc;
}

static bool env_flag(const char *name, bool dflt) {
    const char *e = getenv(name);
    if (!e || !e[0]) return dflt;
    return e[0] != '0';
}

template <bool COUNT_ONLY, typename VT>
static int launch_merge(const int64_t *in_rowptr, const int32_t *in_col, const VT *in_val, int T_out, int halo,
                        int64_t N, const double *band_w, int b, int64_t *out_counts, const int64_t *out_rowptr,
                        int32_t *out_col, VT *out_val, cudaStream_t st, void *ws = nullptr, size_t ws_bytes = 0) {
    // Count pass: thread per row (cols only: issue-bound at full occupancy, faster than staging).
    // Fill pass: the staged kernel when the rows are short enough for a warp's block to fit its staging
    // buffers -- 1.9x faster on the 2 M-node shard (ncu: 18 GB of DRAM traffic instead of 205 GB for 390 M
    // outputs) -- else thread per row.  TMGCN_MERGE_STAGED = 0 | 1 | 2 | 4 forces the fill variant (lanes per
    // row; 0 = thread per row), TMGCN_MERGE_STAGE_KB the staging bytes per warp.
    static int forced = -2;
    static int stage_kb = 36;
    if (forced == -2) {
        const char *e = getenv("TMGCN_MERGE_STAGED");
        forced = e ? atoi(e) : -1;
        const char *k = getenv("TMGCN_MERGE_STAGE_KB");
        if (k && atoi(k) > 0) stage_kb = atoi(k);
    }
    const int threads = 128;
    const unsigned grid = (unsigned)ceil_div((int64_t)T_out * N, threads);
    // staged entry size of the FILL pass (the count pass takes the same decisions so both run the same variant)
    constexpr size_t entry = COUNT_ONLY ? 8 : StageEntry<COUNT_ONLY, VT>::SIZE;
    int sel = 0;
    int cap = 0;
    double mean_row = 0.0;                                                  // stored entries per source row
    int64_t nnz_in_host = -1;
    if (forced != 0) {
        int bb = 2;                                                         // the template width b rounds up to
        for (const int cand : {2, 4, 6, 8, 10, 12, 16, 20, 24, 32})
            if (b <= cand) {
                bb = cand;
                break;
            }
        cap = (int)(((size_t)stage_kb * 1024) / ((size_t)bb * entry)) & ~15;
        if (cap > 1024) cap = 1024;
        if (cap < 16) cap = 16;
        if (forced > 0) {
            sel = forced;
        } else {
            // mean stored entries per source row decide how many rows a warp can stage (mean + 4 sigma of a
            // Poisson block; longer blocks -- hub rows -- take the kernel's global-memory path)
            int64_t nnz_in = 0;
            const int64_t rows_in = (int64_t)(T_out + halo) * N;
            TMGCN_CUDA(cudaMemcpyAsync(&nnz_in, in_rowptr + rows_in, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
            TMGCN_CUDA(cudaStreamSynchronize(st));
            const double mean = rows_in > 0 ? (double)nnz_in / (double)rows_in : 0.0;
            mean_row = mean;
            nnz_in_host = nnz_in;
            for (const int ll : {1, 2}) {                                    // 4 lanes per row only pays when forced
                const double need = mean * (32 / ll);
                if (need + 4.0 * sqrt(need) + 16.0 <= (double)cap && ll <= bb) {
                    sel = ll;
                    break;
                }
            }
        }
    }
    {
        static int tiled = -1;
        if (tiled < 0) tiled = env_flag("TMGCN_MERGE_TT", true) ? 1 : 0;   // TMGCN_MERGE_TT=0: the per-slice kernels
        // the tiled kernels pay off when a warp's B-1+4 source segments fit its staging pool (sel >= 1 says the
        // per-slice staged kernel would fit too); forced variants (TMGCN_MERGE_STAGED) bypass them
        if (tiled && b <= 12 && forced < 0 && sel >= 1 && T_out >= 4) {
            const int64_t n_tasks = (int64_t)ceil_div(T_out, 4) * ceil_div(N, (int64_t)32);
            // union-list variant: the count pass records the union pattern in the workspace, the fill pass is
            // fill_from_union (TMGCN_MERGE_UNION=0 or no workspace: the merging fill kernels)
            static int union_on = -1;
            if (union_on < 0) union_on = env_flag("TMGCN_MERGE_UNION", true) ? 1 : 0;
            int qc = union_on ? union_qc(ws, ws_bytes, n_tasks) : 0;
            if (((uintptr_t)in_col & 15) != 0 || nnz_in_host < 0) qc = 0;     // the bulk copies need an aligned column array
            int pool = (38 * 1024) / 8;                                     // staged entries ({col, val}; count: col)
            if (COUNT_ONLY && qc) {
                // Only columns are staged here, so the pool can afford the spread of the block sums: a row's length
                // persists over the band's source slices, so the sum over a 32-row block of NS slices spreads like
                // NS * sqrt(32 * mean), not like the square root of the total (4 % of the benchmark's blocks
                // overflowed the fixed pool and were merged from global memory by both passes).
                int bb = 2;
                for (const int cand : {2, 4, 6, 8, 10, 12})
                    if (b <= cand) {
                        bb = cand;
                        break;
                    }
                const double ns = bb - 1 + 4, blk = 32.0 * mean_row;
                int want = ((int)(ns * blk + 3.5 * ns * sqrt(blk) + 64.0) + 15) & ~15;
                if (want > 12288) want = 12288;
                if (want > pool) pool = want;
            }
            const size_t smem = (size_t)pool * (COUNT_ONLY ? 4 : 8);
            int per_sm = (int)((220 * 1024) / (smem + 1024));
            if (per_sm > 16) per_sm = 16;
            int64_t g = (int64_t)sm_count() * per_sm;
            if (g > n_tasks) g = n_tasks;
#define TMGCN_UNION_COUNT(BB)                                                                                    \
    if (b <= BB) {                                                                                               \
        auto kern = count_union<BB, 4>;                                                                          \
        const size_t csmem = smem + 16;                                                                          \
        TMGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));         \
        const int64_t n_groups = ceil_div(T_out, 4), n_blk = ceil_div(N, (int64_t)32);                           \
        int occ = 0;                                                                                             \
        TMGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32, csmem));                        \
        int64_t walkers = ((int64_t)sm_count() * (occ > 0 ? occ : 1)) / n_groups;                                \
        if (walkers < 1) walkers = 1;                                                                            \
        if (walkers > n_blk) walkers = n_blk;                                                                    \
        kern<<<(unsigned)(n_groups * walkers), 32, csmem, st>>>(in_rowptr, in_col, nnz_in_host, T_out, halo, N,  \
                                                                band_w, b, pool, out_counts, (uint8_t *)ws, qc); \
        return after_launch("count_union");                                                                      \
    }
            if constexpr (COUNT_ONLY) {
                if (qc) {
                    TMGCN_UNION_COUNT(2) TMGCN_UNION_COUNT(4) TMGCN_UNION_COUNT(6) TMGCN_UNION_COUNT(8)
                    TMGCN_UNION_COUNT(10) TMGCN_UNION_COUNT(12)
                }
            }
#undef TMGCN_UNION_COUNT
            if (qc) TMGCN_REQUIRE(((uintptr_t)ws & 15) == 0, "mtransform_sparse: workspace must be 16-byte aligned");
            if (COUNT_ONLY && qc) TMGCN_CUDA(cudaMemsetAsync(ws, 0, UNION_HEADER, st));
            bool union_fill = false;
            int n_over = 0;
            bool idx32 = false;
            if (!COUNT_ONLY && qc) {
                int64_t nnz_io[2] = {0, 0};
                TMGCN_CUDA(cudaMemcpyAsync(&n_over, ws, sizeof(int), cudaMemcpyDeviceToHost, st));
                TMGCN_CUDA(cudaMemcpyAsync(&nnz_io[0], in_rowptr + (int64_t)(T_out + halo) * N, sizeof(int64_t),
                                           cudaMemcpyDeviceToHost, st));
                TMGCN_CUDA(cudaMemcpyAsync(&nnz_io[1], out_rowptr + (int64_t)T_out * N, sizeof(int64_t),
                                           cudaMemcpyDeviceToHost, st));
                TMGCN_CUDA(cudaStreamSynchronize(st));
                idx32 = nnz_io[0] < ((int64_t)1 << 32) && nnz_io[1] < ((int64_t)1 << 32);
                // mostly overflowed records: merge again.  Tensors with 2^32 entries or more: the merging kernels
                // (the lane-per-entry walk keeps its running positions in 32 bits).
                union_fill = (int64_t)n_over * 20 <= n_tasks && idx32;
            }
#define TMGCN_UNION_FILL(BB)                                                                                     \
    if (b <= BB) {                                                                                               \
        auto kern = fill_from_union<BB, 4, VT, uint32_t>;                                                        \
        const size_t usmem = (size_t)qc * 256;                                                                   \
        TMGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)usmem));         \
        int occ = 0;                                                                                             \
        TMGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32, usmem));                        \
        const int64_t n_groups = ceil_div(T_out, 4), n_blk = ceil_div(N, (int64_t)32);                           \
        int64_t walkers = ((int64_t)sm_count() * (occ > 0 ? occ : 1)) / n_groups;                                \
        if (walkers < 1) walkers = 1;                                                                            \
        if (walkers > n_blk) walkers = n_blk;                                                                    \
        const int64_t ug = n_groups * walkers;                                                                   \
        kern<<<(unsigned)ug, 32, usmem, st>>>(in_rowptr, in_col, in_val, T_out, halo, N, band_w, b,              \
                                              (const uint8_t *)ws, qc, out_rowptr, out_col, out_val);            \
        if (after_launch("fill_from_union")) return 1;                                                           \
        if (n_over == 0) return 0;                                                                               \
        int64_t og = ceil_div(n_tasks, (int64_t)4);                                                              \
        if (og > (int64_t)sm_count() * 8) og = (int64_t)sm_count() * 8;                                          \
        fill_union_overflow<BB, 4, VT><<<(unsigned)og, 128, 0, st>>>(in_rowptr, in_col, in_val, T_out, halo, N,  \
                                                                     band_w, b, (const uint8_t *)ws, qc,         \
                                                                     out_rowptr, out_col, out_val);              \
        return after_launch("fill_union_overflow");                                                              \
    }
            // fp64 values (the func_MProduct API-parity path) keep the merging fill: the union fill is validated
            // and sanitizer-clean for the fp32 layout only.  In the one fp64 run with a partly overflowed,
            // hand-sized workspace (tests: workspace_paths) the results were bit-identical but the workspace
            // header read back changed; until a compute-sanitizer pass on a GPU box clears that, fp64 stays on
            // the round-2 kernels.
            if constexpr (!COUNT_ONLY && sizeof(VT) == 4) {
                if (union_fill) {
                    TMGCN_UNION_FILL(2) TMGCN_UNION_FILL(4) TMGCN_UNION_FILL(6) TMGCN_UNION_FILL(8) TMGCN_UNION_FILL(10)
                    TMGCN_UNION_FILL(12)
                }
            }
#undef TMGCN_UNION_FILL
            if constexpr (sizeof(VT) == 4) {
#define TMGCN_TILED(BB)                                                                                          \
    if (b <= BB) {                                                                                               \
        auto kern = merge_rows_tiled<BB, 4, COUNT_ONLY>;                                                         \
        TMGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
        kern<<<(unsigned)g, 32, smem, st>>>(in_rowptr, in_col, (const float *)in_val, T_out, halo, N, band_w, b, \
                                            pool, out_counts, out_rowptr, out_col, (float *)out_val);            \
        return after_launch(COUNT_ONLY ? "merge_rows_tiled<count>" : "merge_rows_tiled<fill>");                  \
    }
                TMGCN_TILED(2) TMGCN_TILED(4) TMGCN_TILED(6) TMGCN_TILED(8) TMGCN_TILED(10) TMGCN_TILED(12)
#undef TMGCN_TILED
            }
        }
    }
#define TMGCN_MERGE_STAGED(BB, LL)                                                                               \
    if constexpr (!COUNT_ONLY && BB >= LL) {                                                                     \
        const size_t smem = (size_t)BB * cap * entry;                                                            \
        const int64_t n_tasks = (int64_t)T_out * ceil_div(N, (int64_t)(32 / LL));                                \
        auto kern = merge_rows_staged<BB, LL, COUNT_ONLY, VT>;                                                   \
        TMGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
        int per_sm = (int)((220 * 1024) / (smem + 1024));                                                        \
        if (per_sm > 32) per_sm = 32;                                                                            \
        int64_t g = (int64_t)sm_count() * per_sm;                                                                \
        if (g > n_tasks) g = n_tasks;                                                                            \
        kern<<<(unsigned)g, 32, smem, st>>>(in_rowptr, in_col, in_val, T_out, halo, N, band_w, b, cap,           \
                                            out_counts, out_rowptr, out_col, out_val);                           \
        return after_launch("merge_rows_staged<fill>");                                                          \
    }
#define TMGCN_MERGE(BB)                                                                                          \
    if (b <= BB) {                                                                                               \
        if (sel == 1) TMGCN_MERGE_STAGED(BB, 1)                                                                  \
        if (sel == 2) TMGCN_MERGE_STAGED(BB, 2)                                                                  \
        if (sel == 4) TMGCN_MERGE_STAGED(BB, 4)                                                                  \
        merge_rows<BB, COUNT_ONLY, VT><<<grid, threads, 0, st>>>(in_rowptr, in_col, in_val, T_out, halo, N, band_w, \
                                                                 b, out_counts, out_rowptr, out_col, out_val);   \
        return after_launch(COUNT_ONLY ? "merge_rows<count>" : "merge_rows<fill>");                              \
    }
    TMGCN_MERGE(2)
    TMGCN_MERGE(4)
    TMGCN_MERGE(6)
    TMGCN_MERGE(8)
    TMGCN_MERGE(10)
    TMGCN_MERGE(12)
    TMGCN_MERGE(16)
    TMGCN_MERGE(20)
    TMGCN_MERGE(24)
    TMGCN_MERGE(32)
#undef TMGCN_MERGE
#undef TMGCN_MERGE_STAGED
    set_error("mtransform_sparse: band width b=%d > 32 unsupported", b);
    return 1;
}

// ------------------------------------------------------------------------
// per-slice transpose
// ------------------------------------------------------------------------
__global__ void transpose_count(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, int64_t N,
                                int64_t n_rows, unsigned long long *__restrict__ counts) {
    // warp per row so the col reads are coalesced
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < n_rows; row += warps_total) {
        const int64_t tbase = (row / N) * N;
        const int64_t s = rowptr[row], e = rowptr[row + 1];
        for (int64_t k = s + lane; k < e; k += 32) atomicAdd(&counts[tbase + col[k]], 1ULL);
    }
}

template <typename VT>
struct RowVal {
    VT val;
    int32_t row;
};

template <typename VT>
__global__ void transpose_fill(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                               const VT *__restrict__ val, int64_t N, int64_t n_rows,
                               unsigned long long *__restrict__ cursor, RowVal<VT> *__restrict__ tmp) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < n_rows; row += warps_total) {
        const int64_t t = row / N;
        const int64_t tbase = t * N;
        const int32_t i = (int32_t)(row - tbase);
        const int64_t s = rowptr[row], e = rowptr[row + 1];
        for (int64_t k = s + lane; k < e; k += 32) {
            unsigned long long pos = atomicAdd(&cursor[tbase + col[k]], 1ULL);
            RowVal<VT> rv;
            rv.row = i;
            rv.val = val[k];
            tmp[pos] = rv;
        }
    }
}

// rank sort of every transposed row (keys are unique inside a row): restores
// ascending order so the backward SpMM sums in a reproducible order.
template <typename VT>
__global__ void transpose_rank_sort(const int64_t *__restrict__ t_rowptr, const RowVal<VT> *__restrict__ tmp,
                                    int64_t n_rows, int32_t *__restrict__ t_col, VT *__restrict__ t_val) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < n_rows; row += warps_total) {
        const int64_t s = t_rowptr[row], e = t_rowptr[row + 1];
        const int64_t len = e - s;
        if (len <= 32) {
            RowVal<VT> mine;
            mine.row = INT_MAX;
            mine.val = 0;
            if (lane < len) mine = tmp[s + lane];
            int rank = 0;
            for (int j = 0; j < (int)len; ++j) {
                int kj = __shfl_sync(0xffffffffu, mine.row, j);
                rank += kj < mine.row;
            }
            if (lane < len) {
                t_col[s + rank] = mine.row;
                t_val[s + rank] = mine.val;
            }
        } else {
            for (int64_t a = lane; a < len; a += 32) {
                RowVal<VT> mine = tmp[s + a];
                int64_t rank = 0;
                for (int64_t j = 0; j < len; ++j) rank += tmp[s + j].row < mine.row;
                t_col[s + rank] = mine.row;
                t_val[s + rank] = mine.val;
            }
        }
    }
}

// ------------------------------------------------------------------------
// graph preparation (ref: read_data.py:88-164): C = alpha*A + beta*B by a 2-way sorted row merge,
// row sums and the symmetric degree scaling D^-1/2 (.) D^-1/2.  Thread per row.
// ------------------------------------------------------------------------
template <bool COUNT_ONLY, typename VT>
__global__ void csr_axpby_kernel(const int64_t *__restrict__ arow, const int32_t *__restrict__ acol,
                                 const VT *__restrict__ aval, const int64_t *__restrict__ brow,
                                 const int32_t *__restrict__ bcol, const VT *__restrict__ bval, double alpha,
                                 double beta, int64_t n_rows, int64_t *__restrict__ counts,
                                 const int64_t *__restrict__ crow, int32_t *__restrict__ ccol,
                                 VT *__restrict__ cval) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    int64_t pa = arow[row], ea = arow[row + 1], pb = brow[row], eb = brow[row + 1];
    int64_t o = COUNT_ONLY ? 0 : crow[row], n = 0;
    while (pa < ea || pb < eb) {
        const int ca = pa < ea ? acol[pa] : INT_MAX, cb = pb < eb ? bcol[pb] : INT_MAX;
        const int m = min(ca, cb);
        if (!COUNT_ONLY) {
            double v = 0.0;
            if (ca == m) v += alpha * (double)aval[pa];
            if (cb == m) v += beta * (double)bval[pb];
            ccol[o + n] = m;
            cval[o + n] = (VT)v;
        }
        pa += ca == m;
        pb += cb == m;
        ++n;
    }
    if (COUNT_ONLY) counts[row] = n;
}

template <typename VT>
__global__ void csr_row_sums_kernel(const int64_t *__restrict__ rowptr, const VT *__restrict__ val, int64_t n_rows,
                                    double *__restrict__ out) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    double s = 0.0;
    for (int64_t k = rowptr[row]; k < rowptr[row + 1]; ++k) s += (double)val[k];
    out[row] = s;
}

// val[k] <- val[k] * deg[row]^-1/2 * deg[t*N + col[k]]^-1/2   (ref: read_data.py:143-159)
template <typename VT>
__global__ void csr_scale_sym_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                     VT *__restrict__ val, int64_t N, int64_t n_rows,
                                     const double *__restrict__ deg) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const int64_t tbase = (row / N) * N;
    const double di = 1.0 / sqrt(deg[row]);
    for (int64_t k = rowptr[row]; k < rowptr[row + 1]; ++k)
        val[k] = (VT)(((double)val[k] * di) * (1.0 / sqrt(deg[tbase + col[k]])));
}

static int grid_for_warps(int64_t n_warps, int threads) {
    int64_t blocks = ceil_div(n_warps * 32, threads);
    int64_t cap = (int64_t)sm_count() * 32;  // persistent-ish: grid-stride beyond this
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace tmgcn

using namespace tmgcn;

extern "C" {

size_t tmgcn_scan_ws_bytes(int64_t n) { return (size_t)(ceil_div(n > 0 ? n : 1, SCAN_TILE) + 1) * sizeof(int64_t); }

int tmgcn_exclusive_scan_i64(const int64_t *counts, int64_t *out, int64_t n, void *ws, void *stream) {
    TMGCN_REQUIRE(n >= 0, "scan: n < 0");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        TMGCN_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t), st));
        return 0;
    }
    TMGCN_REQUIRE(counts && out && ws, "scan: null pointer");
    int64_t *tile_sums = (int64_t *)ws;
    const int64_t n_tiles = ceil_div(n, SCAN_TILE);
    scan_tile_sums<<<(unsigned)n_tiles, SCAN_THREADS, 0, st>>>(counts, n, tile_sums);
    if (after_launch("scan_tile_sums")) return 1;
    scan_tile_offsets<<<1, SCAN_THREADS, 0, st>>>(tile_sums, n_tiles, out + n);
    if (after_launch("scan_tile_offsets")) return 1;
    scan_apply<<<(unsigned)n_tiles, SCAN_THREADS, 0, st>>>(counts, n, tile_sums, out);
    return after_launch("scan_apply");
}

int tmgcn_rowptr_from_sorted_rows(const int64_t *flat_row, int64_t nnz, int64_t n_rows, int64_t *rowptr,
                                  void *stream) {
    TMGCN_REQUIRE(nnz >= 0 && n_rows >= 0 && rowptr, "rowptr_from_sorted_rows: bad arguments");
    const int threads = 256;
    rowptr_from_rows<<<(unsigned)ceil_div(nnz + 1, threads), threads, 0, (cudaStream_t)stream>>>(flat_row, nnz,
                                                                                                 n_rows, rowptr);
    return after_launch("rowptr_from_rows");
}

static int check_band(int T_out, int halo, int64_t N, int b) {
    TMGCN_REQUIRE(T_out >= 0 && N >= 0, "mtransform: negative size");
    TMGCN_REQUIRE(b >= 1 && b <= 32, "mtransform: band width b=%d outside [1, 32]", b);
    TMGCN_REQUIRE(halo >= 0, "mtransform: negative halo");
    return 0;
}

size_t tmgcn_mtransform_sparse_ws_bytes(int T_out, int halo, int64_t N, int b, int64_t in_nnz) {
    // record depth: union rows of persistent dynamic graphs run to ~2-3x the mean source row (the band's sources
    // overlap); deeper rows overflow their record and are re-merged by the fill pass.  0 = variant not applicable.
    if (T_out < 4 || N <= 0 || b < 1 || b > 12 || halo < 0 || in_nnz <= 0) return 0;
    if (!env_flag("TMGCN_MERGE_UNION", true) || !env_flag("TMGCN_MERGE_TT", true)) return 0;
    const double rows_in = (double)(T_out + halo) * (double)N;
    const double mean = (double)in_nnz / rows_in;
    int qc = ((int)(mean * 4.4) + 8 + 3) & ~3;
    if (qc < 16) qc = 16;
    if (qc > 96) qc = 96;
    const int64_t n_tasks = ceil_div(T_out, 4) * ceil_div(N, (int64_t)32);
    const size_t budget = (size_t)16 << 30;
    while (qc >= 16 && UNION_HEADER + (size_t)n_tasks * union_stride(qc) > budget) qc -= 4;
    if (qc < 16 || (double)qc < 1.5 * mean) return 0;
    return UNION_HEADER + (size_t)n_tasks * union_stride(qc);
}

int tmgcn_mtransform_sparse_plan_ws(const int64_t *in_rowptr, const int32_t *in_col, int T_out, int halo, int64_t N,
                                    const double *band_w, int b, int64_t *out_counts, void *ws, size_t ws_bytes,
                                    void *stream) {
    if (check_band(T_out, halo, N, b)) return 1;
    if ((int64_t)T_out * N == 0) return 0;
    TMGCN_REQUIRE(in_rowptr && band_w && out_counts, "mtransform_sparse_plan: null pointer");
    return launch_merge<true, float>(in_rowptr, in_col, nullptr, T_out, halo, N, band_w, b, out_counts, nullptr,
                                     nullptr, nullptr, (cudaStream_t)stream, ws, ws_bytes);
}

int tmgcn_mtransform_sparse_run_ws(const int64_t *in_rowptr, const int32_t *in_col, const void *in_val, int T_out,
                                   int halo, int64_t N, const double *band_w, int b, const int64_t *out_rowptr,
                                   int32_t *out_col, void *out_val, int val_is_f64, void *ws, size_t ws_bytes,
                                   void *stream) {
    if (check_band(T_out, halo, N, b)) return 1;
    if ((int64_t)T_out * N == 0) return 0;
    TMGCN_REQUIRE(in_rowptr && band_w && out_rowptr, "mtransform_sparse_run: null pointer");
    if (val_is_f64)
        return launch_merge<false, double>(in_rowptr, in_col, (const double *)in_val, T_out, halo, N, band_w, b,
                                           nullptr, out_rowptr, out_col, (double *)out_val, (cudaStream_t)stream, ws,
                                           ws_bytes);
    return launch_merge<false, float>(in_rowptr, in_col, (const float *)in_val, T_out, halo, N, band_w, b, nullptr,
                                      out_rowptr, out_col, (float *)out_val, (cudaStream_t)stream, ws, ws_bytes);
}

int tmgcn_mtransform_sparse_plan(const int64_t *in_rowptr, const int32_t *in_col, int T_out, int halo, int64_t N,
                                 const double *band_w, int b, int64_t *out_counts, void *stream) {
    return tmgcn_mtransform_sparse_plan_ws(in_rowptr, in_col, T_out, halo, N, band_w, b, out_counts, nullptr, 0, stream);
}

int tmgcn_mtransform_sparse_run(const int64_t *in_rowptr, const int32_t *in_col, const void *in_val, int T_out,
                                int halo, int64_t N, const double *band_w, int b, const int64_t *out_rowptr,
                                int32_t *out_col, void *out_val, int val_is_f64, void *stream) {
    return tmgcn_mtransform_sparse_run_ws(in_rowptr, in_col, in_val, T_out, halo, N, band_w, b, out_rowptr, out_col,
                                          out_val, val_is_f64, nullptr, 0, stream);
}

int tmgcn_csr_axpby_plan(const int64_t *a_rowptr, const int32_t *a_col, const int64_t *b_rowptr, const int32_t *b_col,
                         int64_t n_rows, int64_t *counts, void *stream) {
    TMGCN_REQUIRE(n_rows >= 0, "csr_axpby_plan: bad sizes");
    if (n_rows == 0) return 0;
    TMGCN_REQUIRE(a_rowptr && b_rowptr && counts, "csr_axpby_plan: null pointer");
    csr_axpby_kernel<true, float><<<(unsigned)ceil_div(n_rows, 128), 128, 0, (cudaStream_t)stream>>>(
        a_rowptr, a_col, nullptr, b_rowptr, b_col, nullptr, 0.0, 0.0, n_rows, counts, nullptr, nullptr, nullptr);
    return after_launch("csr_axpby<count>");
}

int tmgcn_csr_axpby_run(const int64_t *a_rowptr, const int32_t *a_col, const void *a_val, const int64_t *b_rowptr,
                        const int32_t *b_col, const void *b_val, double alpha, double beta, int64_t n_rows,
                        const int64_t *c_rowptr, int32_t *c_col, void *c_val, int val_is_f64, void *stream) {
    TMGCN_REQUIRE(n_rows >= 0, "csr_axpby_run: bad sizes");
    if (n_rows == 0) return 0;
    TMGCN_REQUIRE(a_rowptr && b_rowptr && c_rowptr, "csr_axpby_run: null pointer");
    const unsigned grid = (unsigned)ceil_div(n_rows, 128);
    if (val_is_f64)
        csr_axpby_kernel<false, double><<<grid, 128, 0, (cudaStream_t)stream>>>(
            a_rowptr, a_col, (const double *)a_val, b_rowptr, b_col, (const double *)b_val, alpha, beta, n_rows,
            nullptr, c_rowptr, c_col, (double *)c_val);
    else
        csr_axpby_kernel<false, float><<<grid, 128, 0, (cudaStream_t)stream>>>(
            a_rowptr, a_col, (const float *)a_val, b_rowptr, b_col, (const float *)b_val, alpha, beta, n_rows, nullptr,
            c_rowptr, c_col, (float *)c_val);
    return after_launch("csr_axpby<fill>");
}

int tmgcn_csr_row_sums(const int64_t *rowptr, const void *val, int64_t n_rows, double *out, int val_is_f64,
                       void *stream) {
    TMGCN_REQUIRE(n_rows >= 0, "csr_row_sums: bad sizes");
    if (n_rows == 0) return 0;
    TMGCN_REQUIRE(rowptr && out, "csr_row_sums: null pointer");
    const unsigned grid = (unsigned)ceil_div(n_rows, 256);
    if (val_is_f64)
        csr_row_sums_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(rowptr, (const double *)val, n_rows, out);
    else
        csr_row_sums_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(rowptr, (const float *)val, n_rows, out);
    return after_launch("csr_row_sums");
}

int tmgcn_csr_scale_sym(const int64_t *rowptr, const int32_t *col, void *val, int T, int64_t N, const double *deg,
                        int val_is_f64, void *stream) {
    const int64_t n_rows = (int64_t)T * N;
    TMGCN_REQUIRE(T >= 0 && N >= 0, "csr_scale_sym: bad sizes");
    if (n_rows == 0) return 0;
    TMGCN_REQUIRE(rowptr && deg, "csr_scale_sym: null pointer");
    const unsigned grid = (unsigned)ceil_div(n_rows, 256);
    if (val_is_f64)
        csr_scale_sym_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(rowptr, col, (double *)val, N, n_rows, deg);
    else
        csr_scale_sym_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(rowptr, col, (float *)val, N, n_rows, deg);
    return after_launch("csr_scale_sym");
}

size_t tmgcn_csr_transpose_ws_bytes(int64_t n_rows, int64_t nnz, int val_is_f64) {
    return (size_t)nnz * (val_is_f64 ? sizeof(RowVal<double>) : sizeof(RowVal<float>)) +
           (size_t)n_rows * sizeof(int64_t);
}

int tmgcn_csr_transpose_plan(const int64_t *rowptr, const int32_t *col, int T, int64_t N, int64_t *counts,
                             void *stream) {
    const int64_t n_rows = (int64_t)T * N;
    if (n_rows == 0) return 0;
    TMGCN_REQUIRE(rowptr && counts, "csr_transpose_plan: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    TMGCN_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_rows * sizeof(int64_t), st));
    const int threads = 256;
    transpose_count<<<grid_for_warps(n_rows, threads), threads, 0, st>>>(rowptr, col, N, n_rows,
                                                                         (unsigned long long *)counts);
    return after_launch("transpose_count");
}

int tmgcn_csr_transpose_run(const int64_t *rowptr, const int32_t *col, const void *val, int T, int64_t N,
                            const int64_t *t_rowptr, int32_t *t_col, void *t_val, int val_is_f64, void *ws,
                            void *stream) {
    const int64_t n_rows = (int64_t)T * N;
    if (n_rows == 0) return 0;
    TMGCN_REQUIRE(rowptr && t_rowptr && ws, "csr_transpose_run: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *cursor = (unsigned long long *)ws;
    void *tmp = (char *)ws + (size_t)n_rows * sizeof(int64_t);
    TMGCN_CUDA(cudaMemcpyAsync(cursor, t_rowptr, (size_t)n_rows * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    const int threads = 256;
    const int grid = grid_for_warps(n_rows, threads);
    if (val_is_f64) {
        transpose_fill<double><<<grid, threads, 0, st>>>(rowptr, col, (const double *)val, N, n_rows, cursor,
                                                         (RowVal<double> *)tmp);
        if (after_launch("transpose_fill")) return 1;
        transpose_rank_sort<double><<<grid, threads, 0, st>>>(t_rowptr, (const RowVal<double> *)tmp, n_rows, t_col,
                                                              (double *)t_val);
    } else {
        transpose_fill<float><<<grid, threads, 0, st>>>(rowptr, col, (const float *)val, N, n_rows, cursor,
                                                        (RowVal<float> *)tmp);
        if (after_launch("transpose_fill")) return 1;
        transpose_rank_sort<float><<<grid, threads, 0, st>>>(t_rowptr, (const RowVal<float> *)tmp, n_rows, t_col,
                                                             (float *)t_val);
    }
    return after_launch("transpose_rank_sort");
}

}  // extern "C"
