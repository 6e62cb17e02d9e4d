"""Synthetic dynamic graphs of the benchmark shapes.

Slice 0 holds ~m distinct uniform-random directed pairs (i != j); slice t keeps
each pair of slice t-1 with probability rho and refills with fresh uniform pairs.
Every slice is then prepared the way the reference prepares its data:
symmetrise (A + A^T)/2 (ref: read_data.py:98-99), add I and scale
D^-1/2 (A + I) D^-1/2 (ref: read_data.py:130-164).

Pure torch tensor ops, device-agnostic (CPU for the small parity cases, the GPU
for the 2M-node benchmark slices): this is input generation, not the hot path.
"""
from __future__ import annotations

from typing import Iterator, Tuple

import torch


def _fresh_pairs(n: int, N: int, gen: torch.Generator, device) -> torch.Tensor:
    i = torch.randint(0, N, (n,), generator=gen, device=device, dtype=torch.int64)
    j = torch.randint(0, N - 1, (n,), generator=gen, device=device, dtype=torch.int64)
    j = j + (j >= i).to(torch.int64)  # j != i, uniform over the other N-1 nodes
    return i * N + j


def synth_slices(N: int, T: int, m: int, rho: float, seed: int = 20261017,
                 device="cpu") -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
    """Yield per slice (row int64, col int64, val fp64), sorted by (row, col)."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    m = min(m, N * (N - 1))
    keys = torch.unique(_fresh_pairs(m, N, gen, device))
    for t in range(T):
        if t > 0:
            keep = torch.rand(keys.numel(), generator=gen, device=device) < rho
            keys = keys[keep]
            need = max(m - keys.numel(), 0)
            keys = torch.unique(torch.cat([keys, _fresh_pairs(need, N, gen, device)]))
        yield _prepare(keys, N, device)


def _prepare(keys: torch.Tensor, N: int, device):
    """pairs (i*N + j) of one slice -> (row, col, val fp64) of D^-1/2 ((A + A^T)/2 + I) D^-1/2, sorted."""
    i, j = keys // N, keys % N
    diag = torch.arange(N, device=device, dtype=torch.int64)
    k_all = torch.cat([keys, j * N + i, diag * N + diag])
    v_all = torch.cat([torch.full((2 * keys.numel(),), 0.5, dtype=torch.float64, device=device),
                       torch.ones(N, dtype=torch.float64, device=device)])
    uk, inv = torch.unique(k_all, return_inverse=True)
    v = torch.zeros(uk.numel(), dtype=torch.float64, device=device).index_add_(0, inv, v_all)
    r, c = uk // N, uk % N
    deg = torch.zeros(N, dtype=torch.float64, device=device).index_add_(0, r, v)
    dinv = 1.0 / torch.sqrt(deg)
    return r, c, v * dinv[r] * dinv[c]


def life_cap(rho: float) -> int:
    """Lifetimes of the global process are capped where rho^L < 1e-3, so a window [t_lo, t_hi) only needs the
    births of the L slices before it -- every rank sees the same graph whatever the partition is."""
    if rho <= 0.0:
        return 1
    if rho >= 1.0:
        return 4096
    import math
    return max(1, min(4096, int(math.ceil(math.log(1e-3) / math.log(rho)))))


def _births(N: int, m: int, rho: float, tau: int, seed: int, device):
    """pairs born at time tau and the first slice they are absent from (deterministic in (seed, tau))."""
    n = m if tau == 0 else int(round(m * (1.0 - rho)))
    gen = torch.Generator(device=device)
    gen.manual_seed((seed * 1_000_003 + tau * 7_919 + 12_345) % (2 ** 62))
    keys = _fresh_pairs(n, N, gen, device)
    if rho <= 0.0:
        life = torch.ones(n, dtype=torch.int64, device=device)
    elif rho >= 1.0:
        life = torch.full((n,), life_cap(rho), dtype=torch.int64, device=device)
    else:
        life = torch.empty(n, dtype=torch.float32, device=device).geometric_(1.0 - rho, generator=gen)
        life = life.clamp_(max=float(life_cap(rho))).to(torch.int64)
    return keys, tau + life


def synth_slices_global(N: int, t_lo: int, t_hi: int, m: int, rho: float, seed: int = 20261017,
                        device="cpu") -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
    """Slices [t_lo, t_hi) of ONE dynamic graph defined for all t >= 0, so that time-sharded ranks hold
    consecutive windows of the same graph (the legacy `synth_slices` chain restarts at every call).
    Birth/lifetime form of the keep-with-probability-rho process: m pairs are born at t = 0 and m(1 - rho)
    at every later t, each alive for a Geometric(1 - rho) number of slices (capped at `life_cap`); slice t is
    the set of pairs alive at t (expected size m).  Yields (row, col, val fp64) prepared like the reference."""
    device = torch.device(device)
    m = min(m, N * (N - 1))
    L = life_cap(rho)
    keys = torch.empty(0, dtype=torch.int64, device=device)
    death = torch.empty(0, dtype=torch.int64, device=device)
    for tau in range(max(0, t_lo - L + 1), t_hi):
        k, d = _births(N, m, rho, tau, seed, device)
        alive = death > tau
        keys, death = torch.cat([keys[alive], k]), torch.cat([death[alive], d])
        if tau >= t_lo:
            yield _prepare(torch.unique(keys), N, device)


def synth_coo(N, T, m, rho, seed=20261017, device="cpu", t_start=None):
    """Coalesced (t, i, j)-ordered COO: idx (3, nnz) int64, val fp64.  t_start = None: the legacy chain;
    an integer: slices [t_start, t_start + T) of the global process (synth_slices_global)."""
    ts, rs, cs, vs = [], [], [], []
    it = (synth_slices(N, T, m, rho, seed, device) if t_start is None else
          synth_slices_global(N, t_start, t_start + T, m, rho, seed, device))
    for t, (r, c, v) in enumerate(it):
        ts.append(torch.full_like(r, t))
        rs.append(r)
        cs.append(c)
        vs.append(v)
    return torch.stack([torch.cat(ts), torch.cat(rs), torch.cat(cs)]), torch.cat(vs)


def synth_csr(N, T, m, rho, seed=20261017, t_start=None):
    """Same graph straight into a device CSR-of-slices (fp32 values, int32 columns)
    without ever holding the int64 COO of all slices (t_start: see synth_coo)."""
    from . import _lib, ops
    lib = _lib.load()
    dev = ops._dev()
    cols, vals, rps = [], [], []
    base = 0
    it = (synth_slices(N, T, m, rho, seed, dev) if t_start is None else
          synth_slices_global(N, t_start, t_start + T, m, rho, seed, dev))
    for r, c, v in it:
        rp = torch.empty(N + 1, dtype=torch.int64, device=dev)
        r = r.contiguous()
        _lib.check(lib.tmgcn_rowptr_from_sorted_rows(ops._p(r), r.numel(), N, ops._p(rp), ops._stream()))
        rps.append(rp[:-1] + base)
        base += r.numel()
        cols.append(c.to(torch.int32))
        vals.append(v.to(torch.float32))
    rps.append(torch.tensor([base], dtype=torch.int64, device=dev))
    return ops.SliceCSR(T, N, torch.cat(rps), torch.cat(cols), torch.cat(vals))


def synth_edges(csr, E: int, seed: int = 20261017) -> torch.Tensor:
    """E readout edges sampled from the stored entries, time-sorted: (3, E) int64."""
    gen = torch.Generator(device=csr.col.device)
    gen.manual_seed(seed + 1)
    pick = torch.randint(0, csr.nnz, (E,), generator=gen, device=csr.col.device, dtype=torch.int64)
    pick, _ = torch.sort(pick)
    row = torch.searchsorted(csr.rowptr, pick, right=True) - 1
    return torch.stack([row // csr.N, row % csr.N, csr.col[pick].to(torch.int64)])
