// (a) sparse M-transform: per-row sorted B-way merge of the band's CSR rows, plus the
// integer plumbing around it (exclusive scan, COO->CSR row pointers, per-slice
// CSR transpose for the backward SpMM).  All of it is HBM/latency-bound integer
// work: coalesced where the data allows, grids sized in multiples of the SM count.
//
// ref: func_MProduct, TensorGCN-master/read_data.py:204-223.
#include <limits.h>
#include <stdlib.h>

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

// Staged variant (opt-in, see launch_merge): a warp owns 32 consecutive rows of one output slice.  For every source
// slice of the band those rows' entries are ONE contiguous range of the CSR, so the warp copies it into
// shared memory with coalesced loads (every fetched sector is fully used; the thread-per-row kernel above
// re-fetched a 32-byte sector for each 4-byte read: 195 GB of DRAM traffic for 16 GB of data) and the lanes
// then run the same B-way register-cursor merge out of shared memory.  A block whose segment exceeds the
// staging capacity (hub rows) falls back, warp-uniformly, to reading that source from global memory.
template <int B, bool COUNT_ONLY, typename VT>
__global__ void __launch_bounds__(32) merge_rows_staged(const int64_t *__restrict__ in_rowptr,
                                                        const int32_t *__restrict__ in_col,
                                                        const VT *__restrict__ in_val, int T_out, int halo, int64_t N,
                                                        const double *__restrict__ band_w, int b, int cap,
                                                        int64_t *__restrict__ out_counts,
                                                        const int64_t *__restrict__ out_rowptr,
                                                        int32_t *__restrict__ out_col, VT *__restrict__ out_val) {
    extern __shared__ __align__(16) uint8_t stage_raw[];
    int32_t *scol = reinterpret_cast<int32_t *>(stage_raw);                 // [B][cap]
    VT *sval = reinterpret_cast<VT *>(stage_raw + (size_t)B * cap * 4);     // [B][cap]   (fill pass only)
    const int lane = threadIdx.x;
    const int64_t nblk = (N + 31) / 32;
    const int64_t n_tasks = (int64_t)T_out * nblk;
    for (int64_t task = blockIdx.x; task < n_tasks; task += gridDim.x) {
        const int t = (int)(task / nblk);
        const int64_t r0 = (task - (int64_t)t * nblk) * 32;
        const int64_t i = r0 + lane;
        const bool live = i < N;
        const int64_t ic = live ? i : N - 1;                                // clamp for the pointer loads
        int32_t pos[B], left[B], cur[B];
        int64_t gbase[B];        // global position of staged element 0 (or of the lane's cursor when not staged)
        bool staged[B];
        double w[B];
        // phase 1: row pointers of every source slice -- all loads are issued before any is consumed
        int64_t p0s[B], p1s[B];
#pragma unroll
        for (int k = 0; k < B; ++k) {
            const int l = B - 1 - k;
            p0s[k] = p1s[k] = 0;
            w[k] = 0.0;
            staged[k] = false;                                              // "valid" until the bounds are known
            if (l < b) {
                const int sl = halo + t - l;
                const double wl = band_w[(int64_t)t * b + l];
                if (sl >= 0 && wl != 0.0) {
                    staged[k] = true;
                    w[k] = wl;
                    p0s[k] = in_rowptr[(int64_t)sl * N + ic];
                    p1s[k] = in_rowptr[(int64_t)sl * N + ic + 1];
                }
            }
        }
        // phase 2: segment bounds via shuffles, then the staging copies as fire-and-forget cp.async (4 B each):
        // all B segments are in flight together, one wait for the lot
#pragma unroll
        for (int k = 0; k < B; ++k) {
            int64_t p0 = p0s[k];
            const int64_t p1 = p1s[k];
            if (!live) p0 = p1;                                             // padding lanes own an empty row
            const int64_t seg0 = __shfl_sync(0xffffffffu, p0, 0);
            const int64_t seg1 = __shfl_sync(0xffffffffu, p1, 31);
            const int64_t len = seg1 - seg0;
            staged[k] = staged[k] && len <= cap;                            // warp-uniform
            left[k] = (int32_t)(p1 - p0);
            if (staged[k]) {
                gbase[k] = seg0;
                pos[k] = (int32_t)(p0 - seg0);
                for (int64_t q = lane; q < len; q += 32) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(
                                     (uint32_t)__cvta_generic_to_shared(scol + k * cap + q)),
                                 "l"(in_col + seg0 + q)
                                 : "memory");
                    if (!COUNT_ONLY) {
                        if (sizeof(VT) == 4)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(
                                             (uint32_t)__cvta_generic_to_shared(sval + k * cap + q)),
                                         "l"(in_val + seg0 + q)
                                         : "memory");
                        else
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(
                                             (uint32_t)__cvta_generic_to_shared(sval + k * cap + q)),
                                         "l"(in_val + seg0 + q)
                                         : "memory");
                    }
                }
            } else {
                gbase[k] = p0;
                pos[k] = 0;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int k = 0; k < B; ++k)
            cur[k] = left[k] > 0 ? (staged[k] ? scol[k * cap + pos[k]] : in_col[gbase[k]]) : INT_MAX;
        // phase 3: B-way merge, one row per lane
        int64_t count = 0;
        const int64_t obase = (COUNT_ONLY || !live) ? 0 : out_rowptr[(int64_t)t * N + i];
        while (true) {
            int m = cur[0];
#pragma unroll
            for (int k = 1; k < B; ++k) m = min(m, cur[k]);
            if (m == INT_MAX) break;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < B; ++k) {
                if (cur[k] == m) {
                    if (!COUNT_ONLY)
                        acc += w[k] * (double)(staged[k] ? sval[k * cap + pos[k]] : in_val[gbase[k] + pos[k]]);
                    ++pos[k];
                    --left[k];
                    cur[k] = left[k] > 0 ? (staged[k] ? scol[k * cap + pos[k]] : in_col[gbase[k] + pos[k]]) : INT_MAX;
                }
            }
            if (!COUNT_ONLY) {
                out_col[obase + count] = m;
                out_val[obase + count] = (VT)acc;
            }
            ++count;
        }
        if (COUNT_ONLY && live) out_counts[(int64_t)t * N + i] = count;
        __syncwarp();                                                       // staging buffers are reused by the next task
    }
}

template <bool COUNT_ONLY, typename VT>
static int launch_merge(const int64_t *in_rowptr, const int32_t *in_col, const VT *in_val, int T_out, int halo,
                        int64_t N, const double *band_w, int b, int64_t *out_counts, const int64_t *out_rowptr,
                        int32_t *out_col, VT *out_val, cudaStream_t st) {
    // The staged kernel cuts DRAM traffic 10x (ncu: 18 GB vs 205 GB for the 12-slice, 390 M-output case) but
    // its 36 KB of staging per warp leaves 6 warps per SM and the ~300-instruction merge step then runs
    // latency-bound (25 ms vs 20 ms for the thread-per-row kernel at full occupancy), so it is opt-in
    // (TMGCN_MERGE_STAGED=1) until the merge step itself is cheaper.
    static int use_staged = -1;
    if (use_staged < 0) {
        const char *e = getenv("TMGCN_MERGE_STAGED");
        use_staged = (e && e[0] == '1') ? 1 : 0;
    }
    const int threads = 128;
    const unsigned grid = (unsigned)ceil_div((int64_t)T_out * N, threads);
    const size_t entry = 4 + (COUNT_ONLY ? 0 : sizeof(VT));
    const int64_t n_tasks = (int64_t)T_out * ceil_div(N, 32);
#define TMGCN_MERGE(BB)                                                                                          \
    if (b <= BB) {                                                                                               \
        if (use_staged) {                                                                                        \
            /* ~36 KB of staging per warp: 6 resident warps per SM; capacity per source slice in entries */      \
            int cap = (int)((36 * 1024) / (BB * entry)) & ~15;                                                   \
            if (cap > 1024) cap = 1024;                                                                          \
            const size_t smem = (size_t)BB * cap * entry;                                                        \
            auto kern = merge_rows_staged<BB, COUNT_ONLY, VT>;                                                   \
            TMGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
            int64_t g = (int64_t)sm_count() * (int64_t)((220 * 1024) / (smem + 1024));                            \
            if (g > n_tasks) g = n_tasks;                                                                        \
            kern<<<(unsigned)g, 32, smem, st>>>(in_rowptr, in_col, in_val, T_out, halo, N, band_w, b, cap,       \
                                                out_counts, out_rowptr, out_col, out_val);                       \
            return after_launch(COUNT_ONLY ? "merge_rows_staged<count>" : "merge_rows_staged<fill>");            \
        }                                                                                                        \
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

int tmgcn_mtransform_sparse_plan(const int64_t *in_rowptr, const int32_t *in_col, int T_out, int halo, int64_t N,
                                 const double *band_w, int b, int64_t *out_counts, void *stream) {
    if (check_band(T_out, halo, N, b)) return 1;
    if ((int64_t)T_out * N == 0) return 0;
    TMGCN_REQUIRE(in_rowptr && band_w && out_counts, "mtransform_sparse_plan: null pointer");
    return launch_merge<true, float>(in_rowptr, in_col, nullptr, T_out, halo, N, band_w, b, out_counts, nullptr,
                                     nullptr, nullptr, (cudaStream_t)stream);
}

int tmgcn_mtransform_sparse_run(const int64_t *in_rowptr, const int32_t *in_col, const void *in_val, int T_out,
                                int halo, int64_t N, const double *band_w, int b, const int64_t *out_rowptr,
                                int32_t *out_col, void *out_val, int val_is_f64, void *stream) {
    if (check_band(T_out, halo, N, b)) return 1;
    if ((int64_t)T_out * N == 0) return 0;
    TMGCN_REQUIRE(in_rowptr && band_w && out_rowptr, "mtransform_sparse_run: null pointer");
    if (val_is_f64)
        return launch_merge<false, double>(in_rowptr, in_col, (const double *)in_val, T_out, halo, N, band_w, b,
                                           nullptr, out_rowptr, out_col, (double *)out_val, (cudaStream_t)stream);
    return launch_merge<false, float>(in_rowptr, in_col, (const float *)in_val, T_out, halo, N, band_w, b, nullptr,
                                      out_rowptr, out_col, (float *)out_val, (cudaStream_t)stream);
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
