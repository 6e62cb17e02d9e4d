"""Time sharding across ranks: rank r owns a contiguous block of
time slices; the banded M needs only a (b-1)-slice halo from the predecessor rank.

  forward  : rank r sends its last b-1 INPUT slices to rank r+1 (dense H every step,
             sparse A once per dataset);
  backward : rank r sends the b-1 partial dH slices it computed for its predecessor's
             block back to rank r-1, which adds them; dW / dU are all-reduced.

One process per GPU, `torch.distributed` point-to-point + all-reduce (NCCL on the
GPUs, gloo in the CPU tests).  No wrap-around: rank 0 has a truncated window.
The functions are device-agnostic (they only slice, send, receive and add).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist

from .ops import SliceCSR


def shard_bounds(T: int, world: int, rank: int):
    """Contiguous block [t0, t1) of rank `rank`; blocks differ by at most one slice."""
    base, rem = divmod(T, world)
    t0 = rank * base + min(rank, rem)
    return t0, t0 + base + (1 if rank < rem else 0)


def balanced_bounds(weights, world: int):
    """Contiguous blocks [(t0, t1)] * world over len(weights) slices whose summed weights are as even as
    the slice granularity allows (greedy on the prefix sums: block r ends where the running weight is
    closest to (r + 1) / world of the total).  Used to balance nnz(A~_t): the first b-1 windows of the
    tensor are truncated, so equal slice counts leave rank 0 with less work."""
    w = [float(x) for x in weights]
    T = len(w)
    if world < 1 or T < world:
        raise ValueError(f"cannot split {T} slices over {world} ranks")
    pre = [0.0]
    for x in w:
        pre.append(pre[-1] + x)
    cuts = [0]
    for r in range(1, world):
        target = pre[-1] * r / world
        lo, hi = cuts[-1] + 1, T - (world - r)          # leave at least one slice per remaining rank
        best = min(range(lo, hi + 1), key=lambda t: (abs(pre[t] - target), t))
        cuts.append(best)
    cuts.append(T)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def slice_weight_estimate(T: int, b: int, N: int, m: int, rho: float):
    """Expected nnz(A~_t) of the synthetic graphs (tmgcn_b200.synth): the diagonal plus 2m stored pairs, plus the
    (1 - rho) share of each of the window's older slices that is no longer alive."""
    return [N + 2.0 * m * (1.0 + (min(t + 1, b) - 1) * (1.0 - rho)) for t in range(T)]


def assert_single_hop(T_own: int, h: int, world: int, device=None):
    """Every rank's block must hold at least h = b-1 slices: a halo then comes from the predecessor alone and
    the send / receive sizes of neighbouring ranks agree (ranks with fewer slices would need a multi-hop halo,
    which is not implemented).  Collective: raises on every rank if any rank fails."""
    if world <= 1:
        return
    t = torch.tensor([int(T_own)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if int(t.item()) < h:
        raise ValueError(f"time sharding needs at least b-1 = {h} slices per rank (smallest block: {int(t.item())})")


def _chain(send: Optional[torch.Tensor], dst: int, recv: Optional[torch.Tensor], src: int):
    ops = []
    if send is not None:
        ops.append(dist.P2POp(dist.isend, send, dst))
    if recv is not None:
        ops.append(dist.P2POp(dist.irecv, recv, src))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def exchange_sparse_halo(A_own: SliceCSR, halo_out: int, rank: int, world: int) -> SliceCSR:
    """Return [last `halo_out` slices of rank-1's block | own slices] as one CSR-of-slices
    (rank 0: own slices unchanged).  Done once per dataset, before the sparse M-transform."""
    T, N = A_own.T, A_own.N
    dev = A_own.rowptr.device
    assert_single_hop(T, halo_out, world, dev)
    h = halo_out
    send_meta = recv_meta = None
    if rank < world - 1:
        r0 = (T - h) * N
        base = A_own.rowptr[r0]
        s_rowptr = (A_own.rowptr[r0:] - base).contiguous()
        lo = int(base.item())
        s_col = A_own.col[lo:].contiguous()
        s_val = A_own.val[lo:].contiguous()
        send_meta = torch.tensor([s_col.numel()], dtype=torch.int64, device=dev)
    if rank > 0:
        recv_meta = torch.zeros(1, dtype=torch.int64, device=dev)
    _chain(send_meta, rank + 1, recv_meta, rank - 1)
    r_rowptr = r_col = r_val = None
    if rank > 0:
        n = int(recv_meta.item())
        r_rowptr = torch.empty(h * N + 1, dtype=torch.int64, device=dev)
        r_col = torch.empty(n, dtype=torch.int32, device=dev)
        r_val = torch.empty(n, dtype=A_own.val.dtype, device=dev)
    for s, r in ((s_rowptr if rank < world - 1 else None, r_rowptr), (s_col if rank < world - 1 else None, r_col),
                 (s_val if rank < world - 1 else None, r_val)):
        _chain(s, rank + 1, r, rank - 1)
    if rank == 0:
        return A_own
    rowptr = torch.cat([r_rowptr[:-1], A_own.rowptr + r_rowptr[-1]])
    return SliceCSR(T + h, N, rowptr, torch.cat([r_col, A_own.col]), torch.cat([r_val, A_own.val]))


class DenseHalo:
    """Per-step halo exchange of the dense layer input / its gradient."""

    def __init__(self, NF: int, h: int, rank: int, world: int, T_own: Optional[int] = None, device=None):
        self.NF, self.h, self.rank, self.world = NF, h, rank, world
        if T_own is not None:
            assert_single_hop(T_own, h, world, device)

    def forward(self, H: torch.Tensor, T_own: int, halo: int):
        """H = [halo | T_own] slices.  Fill H[:halo] from the predecessor's last slices."""
        h = min(self.h, T_own)
        send = H[halo + T_own - h:] if self.rank < self.world - 1 else None
        recv = H[:halo] if self.rank > 0 and halo > 0 else None
        _chain(send, self.rank + 1, recv, self.rank - 1)

    def backward(self, dH: torch.Tensor, T_own: int, halo: int, scratch: Optional[torch.Tensor] = None):
        """dH = [halo | T_own] slices.  Ship dH[:halo] to the predecessor, add what the
        successor computed for our last slices."""
        h = min(self.h, T_own)
        send = dH[:halo] if self.rank > 0 and halo > 0 else None
        recv = None
        if self.rank < self.world - 1:
            n = h * dH[0].numel()
            recv = (scratch.reshape(-1)[:n] if scratch is not None else
                    torch.empty(n, dtype=dH.dtype, device=dH.device)).view((h,) + tuple(dH.shape[1:]))
        _chain(send, self.rank - 1, recv, self.rank + 1)
        if recv is not None:
            dH[halo + T_own - h:].add_(recv)


_NEIGHBOUR_GROUPS = {}


def neighbour_groups(world: int):
    """one process group per pair of neighbouring ranks (i, i+1), created once per process.  Collective: every
    rank of the default group must call it (the first ShardComm does)."""
    if world not in _NEIGHBOUR_GROUPS:
        _NEIGHBOUR_GROUPS[world] = [dist.new_group([i, i + 1]) for i in range(world - 1)]
    return _NEIGHBOUR_GROUPS[world]


class ShardComm:
    """Overlapped halo exchange for `LayerStep`: the NCCL point-to-point traffic runs on its own CUDA
    streams while the main stream works on the slices that do not need the halo.

    forward : `start_forward(H)` ships our last b-1 input slices to rank+1 and receives the
              predecessor's into H[:halo]; `main.wait_event(fwd_done)` before the first b-1 outputs.
    backward: `start_backward(send, recv)` ships the partial dH owed to rank-1 and receives what
              rank+1 owes us; the caller waits on `bwd_recv` (and adds it, or lets the stencil accumulate).

    Every pair of neighbours talks over its own process group (its own NCCL communicator and stream), and a
    rank's sends and receives are issued on different CUDA streams: with one communicator all point-to-point
    operations of a rank are serialised in issue order, and "send to r-1, then receive from r+1" on every
    rank turns the G-1 independent neighbour transfers into one chain that runs from the last rank down
    (measured: up to 33 ms of waiting per step for the 9.2 GB gradient halo on 8 GPUs)."""

    def __init__(self, h: int, rank: int, world: int, device, T_own: Optional[int] = None, pair_groups: bool = True):
        self.h, self.rank, self.world = h, rank, world
        if T_own is not None:
            assert_single_hop(T_own, h, world, device)
        self.stream = torch.cuda.Stream(device=device, priority=-1)          # receives, all-reduce
        self.send_stream = torch.cuda.Stream(device=device, priority=-1)     # sends
        self.fwd_done = torch.cuda.Event()
        self.fwd_sent = torch.cuda.Event()
        self.bwd_sent = torch.cuda.Event()
        self.bwd_recv = torch.cuda.Event()
        self.grads_done = torch.cuda.Event()
        self.send_pending = False
        self.fwd_send_pending = False
        self._ready = torch.cuda.Event()
        self.pg_prev = self.pg_next = None
        if pair_groups and world > 1:
            groups = neighbour_groups(world)
            self.pg_prev = groups[rank - 1] if rank > 0 else None
            self.pg_next = groups[rank] if rank < world - 1 else None

    def _group_for(self, peer: int):
        return self.pg_prev if peer < self.rank else self.pg_next

    def _on_stream(self, stream, fn, done_event):
        main = torch.cuda.current_stream()
        self._ready.record(main)
        with torch.cuda.stream(stream):
            stream.wait_event(self._ready)
            fn()
            done_event.record(stream)

    def _send(self, t: torch.Tensor, peer: int):
        dist.isend(t, peer, group=self._group_for(peer)).wait()

    def _recv(self, t: torch.Tensor, peer: int):
        dist.irecv(t, peer, group=self._group_for(peer)).wait()

    def start_forward(self, H: torch.Tensor, T_own: int, halo: int):
        h = min(self.h, T_own)
        if self.rank > 0 and halo > 0:
            recv = H[:halo]
            self._on_stream(self.stream, lambda: self._recv(recv, self.rank - 1), self.fwd_done)
        else:
            self.fwd_done.record(torch.cuda.current_stream())
        if self.rank < self.world - 1:
            send = H[halo + T_own - h:]
            self._on_stream(self.send_stream, lambda: self._send(send, self.rank + 1), self.fwd_sent)
            self.fwd_send_pending = True

    def start_backward(self, send: Optional[torch.Tensor], recv: Optional[torch.Tensor]):
        """The receive and the send complete independently: the step only needs `bwd_recv`; `bwd_sent`
        guards the re-use of the send staging buffer (checked lazily at the next forward)."""
        if recv is not None:
            self._on_stream(self.stream, lambda: self._recv(recv, self.rank + 1), self.bwd_recv)
        if send is not None:
            self._on_stream(self.send_stream, lambda: self._send(send, self.rank - 1), self.bwd_sent)
            self.send_pending = True

    def wait_send_buffer_free(self):
        if self.send_pending:
            self.wait(self.bwd_sent)
            self.send_pending = False

    def wait_forward_sent(self):
        """H's tail was read by the forward send: order whatever rewrites H after it"""
        if self.fwd_send_pending:
            self.wait(self.fwd_sent)
            self.fwd_send_pending = False

    def start_allreduce(self, grads: List[torch.Tensor]):
        self._on_stream(self.stream, lambda: allreduce_grads(grads), self.grads_done)

    def wait(self, event):
        torch.cuda.current_stream().wait_event(event)


class PeerHalo:
    """Forward halo fused into the stencil over NVLink peer memory.

    Every rank keeps its layer input H (T_own, N, F) in symmetric memory; rank r maps rank r-1's block and
    the boundary stencil kernel (`tmgcn_mtransform_dense_fwd_split`) loads the predecessor's last b-1 slices
    straight from its HBM: no NCCL copy, no staging, no halo region in the local tensor.  Blocks may differ
    in length (nnz-balanced shards): the block lengths are all-gathered at construction and the tail is
    indexed from the predecessor's real length.  Two stream-ordered cross-rank barriers per step fence the
    reads: "every H is ready" before, "all reads are done" after; `LayerStep.forward` makes the caller's
    stream wait on `reads_done` before it returns, so whatever rewrites H next is ordered after the
    successor's NVLink reads.

    `storage` = (buffer, handle) of an existing symmetric allocation of at least T_own*N*F floats lets
    several workloads share one allocation (symmetric memory is not returned to the caching allocator)."""

    def __init__(self, T_own: int, N: int, F: int, h: int, rank: int, world: int, device, group=None,
                 storage=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.T, self.N, self.F, self.rank, self.world = T_own, N, F, rank, world
        lens = torch.zeros(world, dtype=torch.int64, device=device)
        lens[rank] = T_own
        dist.all_reduce(lens, group=group)
        self.T_all = [int(x) for x in lens.tolist()]
        if min(self.T_all) < h:
            raise ValueError(f"PeerHalo needs at least b-1 = {h} slices per rank (blocks: {self.T_all})")
        self.h = h
        n_max = max(self.T_all) * N * F
        if storage is None:
            buf = symm_mem.empty(n_max, dtype=torch.float32, device=device)
            hdl = symm_mem.rendezvous(buf, (group or dist.group.WORLD).group_name)
        else:
            buf, hdl = storage
            if buf.numel() < n_max:
                raise ValueError("PeerHalo: the shared symmetric buffer is too small for this workload")
        self.buf, self.hdl = buf, hdl
        self.H = buf[: T_own * N * F].view(T_own, N, F)
        self.prev = None
        if rank > 0:
            Tp = self.T_all[rank - 1]
            self.prev = hdl.get_buffer(rank - 1, (Tp, N, F), torch.float32)
        # high priority: the boundary kernel must not queue behind the main stream's persistent SpMM grid
        self.stream = torch.cuda.Stream(device=device, priority=-1)
        self.boundary_done = torch.cuda.Event()
        self.reads_done = torch.cuda.Event()
        self._ready = torch.cuda.Event()
        # opt-in (TMGCN_PEER_SIGNALS=1): neighbour-to-neighbour signals instead of the two all-rank barriers -- a
        # rank only has to agree with its predecessor ("your H is complete") and its successor ("I have finished
        # reading you").  Measured no faster than the barriers (C4 on 8 GPUs: 2.386 vs 2.393 ms/step), and one
        # full 8-GPU bench run with it died with a device-side trap that was not reproduced: barriers stay the default.
        self.signals = os.environ.get("TMGCN_PEER_SIGNALS", "0") == "1" and hasattr(hdl, "put_signal")

    @staticmethod
    def allocate(numel: int, device, group=None):
        """one symmetric allocation (buffer, handle) to be shared by several PeerHalo objects"""
        import torch.distributed._symmetric_memory as symm_mem
        buf = symm_mem.empty(int(numel), dtype=torch.float32, device=device)
        return buf, symm_mem.rendezvous(buf, (group or dist.group.WORLD).group_name)

    def tail(self) -> torch.Tensor:
        """the predecessor's last h slices -- a view of PEER memory"""
        return self.prev[self.T_all[self.rank - 1] - self.h:]

    def run_boundary(self, fn):
        """fn() launches the peer-reading boundary stencil; runs on the side stream between the two barriers."""
        main = torch.cuda.current_stream()
        self._ready.record(main)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self._ready)
            if self.signals:
                T_MS = 20000                     # a lost signal traps after 20 s instead of hanging the GPU
                if self.rank < self.world - 1:
                    self.hdl.put_signal(self.rank + 1, channel=0, timeout_ms=T_MS)   # my H is complete
                if self.rank > 0:
                    self.hdl.wait_signal(self.rank - 1, channel=0, timeout_ms=T_MS)  # ... and so is my predecessor's
                    fn()
                self.boundary_done.record(self.stream)
                if self.rank > 0:
                    self.hdl.put_signal(self.rank - 1, channel=1, timeout_ms=T_MS)   # finished reading rank-1
                if self.rank < self.world - 1:
                    self.hdl.wait_signal(self.rank + 1, channel=1, timeout_ms=T_MS)  # rank+1 finished reading me
                self.reads_done.record(self.stream)
                return
            self.hdl.barrier(channel=0)          # every rank's H is complete
            if self.rank > 0:
                fn()
            self.boundary_done.record(self.stream)
            self.hdl.barrier(channel=1)          # every rank has finished reading its predecessor
            self.reads_done.record(self.stream)

    def wait_reads_done(self):
        """order the current stream after every rank's NVLink reads of this step (call before H is rewritten)"""
        torch.cuda.current_stream().wait_event(self.reads_done)


def _device_local_solve(src, dst, halo, T, h, n, ld, ld_halo, w, b, transposed):
    """one column chunk of the banded substitution on the device (C ABI: tmgcn_mtransform_dense_solve_part)"""
    import ctypes
    from . import _lib
    from .ops import _p, _stream

    def view_ptr(v, row_stride):      # a column window of a row-major matrix: rows `row_stride` floats apart
        assert v.is_cuda and v.dtype == torch.float32 and v.stride(-1) == 1 and (v.shape[0] <= 1 or v.stride(0) == row_stride)
        return ctypes.c_void_p(v.data_ptr())
    lib = _lib.load()
    _lib.check(lib.tmgcn_mtransform_dense_solve_part(view_ptr(src, ld), view_ptr(dst, ld),
                                                     view_ptr(halo, ld_halo) if h > 0 else None, T, h, n, ld,
                                                     ld_halo, _p(w, torch.float32), b, 1 if transposed else 0,
                                                     _stream()))


def solve_pipelined(z: torch.Tensor, band, t0: int, t1: int, rank: int, world: int, transposed: bool = False,
                    chunks: int = 8, local_solve=None) -> torch.Tensor:
    """Y = inv(M) x_3 Z (ref: ehf:223-224) -- or the adjoint solve with M^T -- on a time-sharded tensor.

    The substitution y[t] = (z[t] - sum_i M[t,t-i] y[t-i]) / M[t,t] is a recurrence through time, so rank r can
    only start once rank r-1 has produced its last b-1 output slices: a cross-rank scan.  It is pipelined over
    `chunks` column ranges of the N*F columns: rank r solves chunk c as soon as the halo of chunk c has
    arrived and sends its own last b-1 rows of chunk c on while it works on chunk c+1, so G ranks take
    (G + chunks - 1) / chunks block-times instead of G.  The transposed solve runs the same way from the last
    rank down.  z: this rank's (T_own, ...) block [t0, t1); returns y of the same shape.  `local_solve` is the
    per-chunk kernel (default: the device kernel through the C ABI; the CPU tests inject a torch one)."""
    T_own = t1 - t0
    assert z.shape[0] == T_own and z.is_contiguous()
    b = band.b
    h = b - 1
    NF = z[0].numel() if T_own else 0
    y = torch.empty_like(z)
    cuda = z.is_cuda
    solve = local_solve or _device_local_solve
    up, down = (rank + 1, rank - 1) if transposed else (rank - 1, rank + 1)     # halo comes from `up`, goes to `down`
    has_up = 0 <= up < world and h > 0
    has_down = 0 <= down < world and h > 0
    if world > 1:
        assert_single_hop(T_own, h, world, z.device)
    # the transposed solve reaches b-1 weight rows into the successor's block (M[s+i, s])
    rows_w = min(band.T, t1 + h) if transposed else t1
    if cuda:
        w = band.device_weights(t0, rows_w, torch.float32)
    else:
        w = band.w[t0:rows_w].to(torch.float32).contiguous()
    h_w = rows_w - t1 if transposed else 0
    bounds = [NF * c // chunks for c in range(chunks + 1)]
    cols = [(bounds[c], bounds[c + 1]) for c in range(chunks) if bounds[c + 1] > bounds[c]]
    z2, y2 = z.view(T_own, NF), y.view(T_own, NF)
    halo_in = [torch.empty(h, c1 - c0, dtype=z.dtype, device=z.device) for c0, c1 in cols] if has_up else None
    stage = [torch.empty(h, c1 - c0, dtype=z.dtype, device=z.device) for c0, c1 in cols] if has_down else None
    if cuda:
        main = torch.cuda.current_stream()
        rs, ss = torch.cuda.Stream(device=z.device), torch.cuda.Stream(device=z.device)
        ev_recv = [torch.cuda.Event() for _ in cols]
        ev_done = [torch.cuda.Event() for _ in cols]
        ev_start = torch.cuda.Event()
        ev_start.record(main)
        if has_up:                                  # every receive is posted up front, in chunk order
            with torch.cuda.stream(rs):
                rs.wait_event(ev_start)
                for k in range(len(cols)):
                    dist.recv(halo_in[k], up)
                    ev_recv[k].record(rs)
    for k, (c0, c1) in enumerate(cols):
        if has_up:
            if cuda:
                main.wait_event(ev_recv[k])
            else:
                dist.recv(halo_in[k], up)
        hk = h if has_up else 0
        # transposed on the last rank / forward on rank 0: no halo, the recurrence starts from zeros
        solve(z2[:, c0:], y2[:, c0:], halo_in[k] if has_up else None, T_own, hk, c1 - c0, NF, c1 - c0, w, b,
              transposed)
        if has_down:
            if cuda:
                ev_done[k].record(main)
                with torch.cuda.stream(ss):
                    ss.wait_event(ev_done[k])
                    stage[k].copy_(y2[:h, c0:c1] if transposed else y2[T_own - h:, c0:c1])
                    dist.send(stage[k], down)
            else:
                stage[k].copy_(y2[:h, c0:c1] if transposed else y2[T_own - h:, c0:c1])
                dist.send(stage[k], down)
    if cuda:
        main.wait_stream(ss)                        # the staging buffers die with this call
        main.wait_stream(rs)
    del h_w
    return y


def allreduce_grads(grads: List[torch.Tensor]):
    """Sum the shared-parameter gradients (dW, dU) over ranks."""
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()
