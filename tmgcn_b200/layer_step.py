"""Workspace-planned forward+backward of one TM-GCN layer (the unit bench.py times).

Same arithmetic as `TMGCNLayer` + autograd (H -> M x_3 H -> A~_t . H~_t -> act(. W) ->
edge readout . U, then the full backward; ref: ehf:342-344, 351-355 and autograd of
those), but every (T, N, F) intermediate lives in one of three caller-visible work
buffers that are reused as tensors die, so a 2M-node, 32-slice, F=128 shard
(32.8 GB per tensor) fits one 180 GB B200:

    fwd   B1 = stencil(H)         H~
          B2 = spmm(A~, B1)       P          (B1 dead)
          B1 = act(B2 . W)        Y
          out = readout(B1, U)
    bwd   dU  = Z^T dOut          (reads Y rows)
          B3 = scatter(dOut, U)   dY         (then dY *= act'(Y) in place; Y dead)
          dW  = B2^T . B3 ,  B1 = B3 . W^T   dP   (P, dY dead)
          B3 = spmm(A~^T, B1)     dH~
          B2 = stencil^T(B3)      dH  (returned as a view of B2, halo slices included)

With time sharding (`halo` > 0) the caller owns H as [halo | T_own] slices and is
responsible for the halo exchange (see sharding.py); dH then carries the `halo`
partial slices owed to the predecessor rank in front.

Low-rank backward (`bwd_mode="lowrank"`, chosen automatically when the layer is linear, i.e.
act = none as in layer 2 of the reference model, ehf:342-355): the readout hands back
dY = S . U~ with S = per-row class sums, (T*N) x 2C -- a rank-2C factorisation of the whole
upstream gradient.  By associativity every stage then runs on the skinny factor and only the
last one expands to F columns:

    S   = class_sums(dOut)                      (T*N, 2C)
    G   = P^T S                                 one pass over P  -> dW = G^T U~,  dU_h = W^T G_h
    Q   = A~^T S      (SpMM^T on 2C columns)    16 B gathers instead of 512 B
    Q'  = M^T x_3 Q   (transposed stencil)      halo exchange = (b-1) * N * 2C floats
    dH  = Q' . V,  V = U~ W^T                   the only F-wide write of the backward

Same gradients to rounding (parity-tested against the CPU restatement of the reference and against the dense path); it replaces
~200 GB of HBM traffic per 32-slice shard by ~70 GB.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import torch

from . import _lib, ops
from .ops import ACT, Band, EdgePlan, SliceCSR, _p, _stream


class LayerStep:
    def __init__(self, At: SliceCSR, band: Band, plan: EdgePlan, F_in: int, F_out: int, C: int, act=None,
                 t0: int = 0, t1: Optional[int] = None, halo: int = 0, bwd_mode: str = "auto",
                 boundary_ctas: Optional[int] = None):
        self.lib = _lib.load()
        if boundary_ctas is None:     # persistent grid of the peer-reading boundary stencil (0 = full grid)
            boundary_ctas = int(os.environ.get("TMGCN_BOUNDARY_CTAS", 2 * torch.cuda.get_device_properties(
                At.rowptr.device).multi_processor_count))
        self.boundary_ctas = boundary_ctas
        self.At, self.AtT = At, At.transpose()
        if band.b > Band.MAX_KERNEL_BAND:
            raise NotImplementedError(f"LayerStep: band width {band.b} > {Band.MAX_KERNEL_BAND} (use the autograd ops)")
        if C > ops.MAX_FUSED_CLASSES:
            raise NotImplementedError(f"LayerStep: {C} classes > {ops.MAX_FUSED_CLASSES} (use the autograd ops)")
        self.band, self.plan = band, plan
        self.t0, self.t1, self.halo = t0, (band.T if t1 is None else t1), halo
        self.T, self.N = At.T, At.N
        assert self.T == self.t1 - self.t0
        self.F_in, self.F_out, self.C = F_in, F_out, C
        self.act = ACT[act] if not isinstance(act, int) else act
        dev = At.rowptr.device
        self.Fmax = max(F_in, F_out)
        n = self.T * self.N * self.Fmax
        self.B1 = torch.empty(n, dtype=torch.float32, device=dev)
        self.B2 = torch.empty((self.T + halo) * self.N * self.Fmax, dtype=torch.float32, device=dev)  # also holds dH
        self.B3 = torch.empty(n, dtype=torch.float32, device=dev)
        self.w_f32 = band.device_weights(self.t0, self.t1, torch.float32)
        self.out = torch.empty(plan.E, C, dtype=torch.float32, device=dev)
        self.dW = torch.empty(F_in, F_out, dtype=torch.float32, device=dev)
        self.dU = torch.empty(2 * F_out, C, dtype=torch.float32, device=dev)
        self.dw_ws = ops._ws(self.lib.tmgcn_gemm_dw_ws_bytes(F_in, F_out))
        self.du_ws = ops._ws(self.lib.tmgcn_edge_readout_bwd_ws_bytes(self.T * self.N, F_out, C))
        self.inc = plan.incidence(self.T * self.N)
        if bwd_mode == "auto":
            bwd_mode = "lowrank" if (self.act == 0 and 4 * C <= F_in) else "dense"
        if bwd_mode == "lowrank" and self.act != 0:
            raise ValueError("the low-rank backward needs a linear layer (act = none)")
        self.bwd_mode = bwd_mode
        if bwd_mode == "lowrank":
            J = 2 * C
            self.S = torch.empty(self.T * self.N * J, dtype=torch.float32, device=dev)
            self.Q = torch.empty(self.T * self.N * J, dtype=torch.float32, device=dev)
            self.Qp = torch.empty((self.T + halo) * self.N * J, dtype=torch.float32, device=dev)
            self.Qrecv = torch.empty(min(band.b - 1, self.T) * self.N * J, dtype=torch.float32, device=dev)
            self.G = torch.empty(2 * F_in, C, dtype=torch.float32, device=dev)
            self.fac_ws = ops._ws(self.lib.tmgcn_edge_factor_ws_bytes(max(F_in, F_out), C))
        self.hook: Optional[Callable[[str], None]] = None   # called before each stage (bench timing)

    @staticmethod
    def _check_f32(**tensors):
        # raw pointers go straight to fp32 kernels: another dtype would be silently reinterpreted
        for name, t in tensors.items():
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise TypeError(f"LayerStep: {name} must be a contiguous float32 CUDA tensor (got {t.dtype}, "
                                f"{'cuda' if t.is_cuda else 'cpu'})")

    def _view(self, buf, T, F):
        return buf[: T * self.N * F].view(T, self.N, F)

    def _mark(self, name):
        if self.hook is not None:
            self.hook(name)

    # ---- kernels on slice sub-ranges (pointer offsets; rowptr entries are absolute) -----------------
    def _stencil_fwd(self, H, Ht, t_lo, t_hi, h_in=None):
        """outputs [t_lo, t_hi) of this shard; H physically starts with `h_in` halo slices (default: all of
        them; 0 when the halo lives in a peer GPU's memory)."""
        h_in = self.halo if h_in is None else h_in
        hh = min(self.band.b - 1, h_in + t_lo)
        NF = self.N * self.F_in
        _lib.check(self.lib.tmgcn_mtransform_dense_fwd(
            _p(H[h_in + t_lo - hh:]), _p(Ht[t_lo:]), t_hi - t_lo, hh, NF,
            _p(self.w_f32[t_lo:]), self.band.b, _stream()))

    def _spmm(self, csr, x, y, t_lo, t_hi, F):
        _lib.check(self.lib.tmgcn_spmm_fwd(_p(csr.rowptr[t_lo * self.N:]), _p(csr.col), _p(csr.val), _p(x[t_lo:]),
                                           _p(y[t_lo:]), t_hi - t_lo, self.N, F, 0, _stream()))

    def forward(self, H: torch.Tensor, W: torch.Tensor, U: torch.Tensor, comm=None, peer=None) -> torch.Tensor:
        """comm: a sharding.ShardComm -> the forward halo exchange (NCCL) runs on its stream while the slices
        that do not depend on the halo are transformed and propagated.
        peer: a sharding.PeerHalo -> H is the rank's own (T, N, F) block in symmetric memory and the boundary
        stencil reads the predecessor's slices from peer memory over NVLink (exchange fused into the kernel)."""
        lib, T, N, st = self.lib, self.T, self.N, _stream()
        h_in = 0 if peer is not None else self.halo
        assert H.shape == (T + h_in, N, self.F_in) and H.is_contiguous()
        self._check_f32(H=H, W=W, U=U)
        assert W.shape == (self.F_in, self.F_out) and U.shape == (2 * self.F_out, self.C)
        Ht = self._view(self.B1, T, self.F_in)
        P = self._view(self.B2, T, self.F_in)
        Y = self._view(self.B1, T, self.F_out)
        overlap = (comm is not None or peer is not None) and self.halo > 0
        hb = min(self.band.b - 1, T) if overlap else 0
        if comm is not None:
            comm.wait_send_buffer_free()      # last step's gradient halo is sent from B2, which SpMM overwrites below
        if peer is not None:
            self._mark("halo_fwd_start")
            NF = N * self.F_in

            def boundary():     # outputs [0, hb): halo slices from the predecessor's HBM, the rest from ours
                _lib.check(lib.tmgcn_mtransform_dense_fwd_split(_p(peer.tail()), _p(H), _p(Ht), hb, self.halo, NF,
                                                                _p(self.w_f32), self.band.b, self.boundary_ctas,
                                                                _stream()))
            peer.run_boundary(boundary)
        elif comm is not None:
            self._mark("halo_fwd_start")
            comm.start_forward(H, T, self.halo)
        if hb < T:
            self._mark("stencil_fwd")
            self._stencil_fwd(H, Ht, hb, T, h_in)
            self._mark("spmm_fwd")
            self._spmm(self.At, Ht, P, hb, T, self.F_in)
        if hb > 0:
            self._mark("halo_fwd_wait")
            if peer is not None:
                torch.cuda.current_stream().wait_event(peer.boundary_done)
            else:
                comm.wait(comm.fwd_done)
                self._mark("stencil_fwd")
                self._stencil_fwd(H, Ht, 0, hb, h_in)
            self._mark("spmm_fwd")
            self._spmm(self.At, Ht, P, 0, hb, self.F_in)
        self._mark("gemm_fwd")
        _lib.check(lib.tmgcn_gemm_xw_fwd(_p(P), _p(W), _p(Y), T * N, self.F_in, self.F_out, self.act, st))
        self._mark("readout_fwd")
        _lib.check(lib.tmgcn_edge_readout_fwd(_p(Y), _p(self.plan.src), _p(self.plan.dst), _p(U), _p(self.out),
                                              self.plan.E, self.F_out, self.C, st))
        if peer is not None:
            # whatever rewrites H after this call is ordered after every rank's NVLink reads of this step
            peer.wait_reads_done()
        elif comm is not None:
            comm.wait_forward_sent()          # ... or after the NCCL send of our last b-1 slices
        self._mark("end")
        return self.out

    def _backward_lowrank(self, dOut, W, U, comm=None):
        """see the module docstring; -> dH (halo + T, N, F_in) view of B2, dW, dU."""
        lib, T, N, st, C = self.lib, self.T, self.N, _stream(), self.C
        Fi, Fo, J = self.F_in, self.F_out, 2 * self.C
        inc_ptr, perm = self.inc
        P = self._view(self.B2, T, Fi)
        dH = self._view(self.B2, T + self.halo, Fi)
        self._mark("readout_bwd")
        _lib.check(lib.tmgcn_edge_class_sums(_p(dOut), _p(inc_ptr), _p(perm), _p(self.S), T * N, C, st))
        # G[(h, k), c] = sum_rows P[row, k] S[row, h, c]   (the "reduce" half of factor_apply, y := P)
        self._mark("gemm_bwd")
        _lib.check(lib.tmgcn_edge_factor_apply(_p(P), _p(self.G), _p(self.S), None, _p(self.G), T * N, Fi, C,
                                               _p(self.fac_ws), st))
        # tiny (2C x F) algebra: dW = G~^T U~, dU_h = W^T G_h, V = U~ W^T
        dW_, dU_, Vu = ops.lowrank_small(W, U, self.G)
        self.dW.copy_(dW_)
        self.dU.copy_(dU_)
        if comm is not None:
            comm.start_allreduce([self.dW, self.dU])
        self._mark("spmm_bwd")
        Sv, Qv = self.S.view(T, N, J), self.Q.view(T, N, J)
        _lib.check(lib.tmgcn_spmm_fwd(_p(self.AtT.rowptr), _p(self.AtT.col), _p(self.AtT.val), _p(Sv), _p(Qv), T, N,
                                      J, 0, st))
        self._mark("stencil_bwd")
        _lib.check(lib.tmgcn_mtransform_dense_bwd(_p(self.Q), _p(self.Qp), T, self.halo, N * J, _p(self.w_f32),
                                                  self.band.b, st))
        Qp = self.Qp.view(T + self.halo, N, J)
        lo, hi, recv = 0, T + self.halo, None
        if comm is not None:
            h = min(self.band.b - 1, T)
            send = Qp[: self.halo] if self.halo > 0 else None
            recv = self.Qrecv.view(h, N, J) if comm.rank < comm.world - 1 else None
            self._mark("halo_bwd_start")
            comm.start_backward(send, recv)
            lo = self.halo                       # the halo slices belong to the predecessor: not expanded here
            if recv is not None:
                hi = self.halo + T - h           # the last h slices still miss what the successor owes them

        def expand(a, b_):
            if b_ > a:
                _lib.check(lib.tmgcn_edge_factor_apply(None, _p(Vu), _p(Qp[a:b_]), _p(dH[a:b_]), None, (b_ - a) * N,
                                                       Fi, C, None, st))
        self._mark("expand_dH")
        expand(lo, hi)                           # everything that does not wait for the neighbour first
        if recv is not None:
            self._mark("halo_bwd_wait")
            comm.wait(comm.bwd_recv)
            self._mark("expand_dH")
            Qp[hi:].add_(recv)
            expand(hi, T + self.halo)
        if comm is not None:
            self._mark("grads_wait")
            comm.wait(comm.grads_done)
        self._mark("end")
        return dH, self.dW, self.dU

    def backward(self, dOut: torch.Tensor, W: torch.Tensor, U: torch.Tensor, comm=None):
        """-> dH (halo + T, N, F_in) view of a work buffer, dW, dU.  With `comm` the partial-gradient halo
        and the dW/dU all-reduce overlap the main backward stencil; dH[halo:] is then complete (and dH[:halo]
        is what was sent to the predecessor / unspecified)."""
        self._check_f32(dOut=dOut, W=W, U=U)
        assert dOut.shape == (self.plan.E, self.C)
        if self.bwd_mode == "lowrank":
            return self._backward_lowrank(dOut, W, U, comm)
        lib, T, N, st = self.lib, self.T, self.N, _stream()
        inc_ptr, perm = self.inc
        P = self._view(self.B2, T, self.F_in)
        Y = self._view(self.B1, T, self.F_out)
        dY = self._view(self.B3, T, self.F_out)
        dP = self._view(self.B1, T, self.F_in)
        dHt = self._view(self.B3, T, self.F_in)
        dH = self._view(self.B2, T + self.halo, self.F_in)
        NF = N * self.F_in
        self._mark("readout_bwd")
        # the layer nonlinearity's derivative is folded into the same pass (dY <- dY * act'(Y))
        _lib.check(lib.tmgcn_edge_readout_bwd(_p(Y), _p(U), _p(dOut), _p(inc_ptr), _p(perm), _p(dY), _p(self.dU),
                                              T * N, self.F_out, self.C, self.act, _p(self.du_ws), st))
        self._mark("gemm_bwd")
        _lib.check(lib.tmgcn_gemm_dw_dx_bwd(_p(P), _p(W), None, _p(dY), _p(dP), _p(self.dW), T * N, self.F_in,
                                            self.F_out, 0, _p(self.dw_ws), st))
        recv = None
        if comm is None:
            self._mark("spmm_bwd")
            self._spmm(self.AtT, dP, dHt, 0, T, self.F_in)
        else:
            comm.start_allreduce([self.dW, self.dU])
            h = min(self.band.b - 1, T)
            if comm.rank < comm.world - 1:
                # P (B2) is dead since dW: what the successor owes our last h slices is received straight into
                # dH while the whole backward SpMM runs; the main stencil below adds its own part on top
                recv = dH[self.halo + T - h:]
                self._mark("halo_bwd_start")
                comm.start_backward(None, recv)
            # Gradient halo first: the partial sums owed to the predecessor depend only on our first h
            # outputs, so propagate those slices, run the small transposed stencil and put the result on
            # the wire while the remaining slices and the main stencil run.
            self._mark("spmm_bwd")
            self._spmm(self.AtT, dP, dHt, 0, h, self.F_in)
            if self.halo > 0:
                # the halo slices of dH are produced in place and sent from there
                self._mark("stencil_bwd")
                _lib.check(lib.tmgcn_mtransform_dense_bwd_range(_p(dHt), _p(dH), h, self.halo, NF, _p(self.w_f32),
                                                                self.band.b, 0, self.halo, -1, st))
                self._mark("halo_bwd_start")
                comm.start_backward(dH[: self.halo], None)
            if h < T:
                self._mark("spmm_bwd")
                self._spmm(self.AtT, dP, dHt, h, T, self.F_in)
            if recv is not None:
                self._mark("halo_bwd_wait")
                comm.wait(comm.bwd_recv)
        self._mark("stencil_bwd")
        _lib.check(lib.tmgcn_mtransform_dense_bwd_range(_p(dHt), _p(dH), T, self.halo, NF, _p(self.w_f32),
                                                        self.band.b, self.halo if comm is not None else 0,
                                                        self.halo + T,
                                                        self.halo + T - recv.shape[0] if recv is not None else -1, st))
        if comm is not None:
            self._mark("grads_wait")
            comm.wait(comm.grads_done)
        self._mark("end")
        return dH, self.dW, self.dU

    # ---- algorithmic bytes per stage, whole shard --------
    def algorithmic_bytes(self, l2_bytes: int) -> dict:
        T, N, Fi, Fo, E, C = self.T, self.N, self.F_in, self.F_out, self.plan.E, self.C
        nnz = self.At.nnz
        n_t = self.At.slice_nnz().double()

        def spmm(F):
            if 4 * N * F > l2_bytes:
                gather = 4.0 * F * nnz
            else:
                gather = 4.0 * F * float(torch.clamp(n_t, max=N).sum())
            return 8.0 * nnz + 4.0 * (N + 1) * T + gather + 4.0 * N * F * T
        # classifier folded in: the (E, 2F) concat is never written (SURVEY's unfused figure adds 8*F*E)
        edge_fwd = 16.0 * E + 8.0 * Fo * E + 4.0 * C * E
        # dY written once (4NF per slice), Y read once per touched row (<= 4NF), perm + dOut + inc_ptr
        touched = min(2.0 * E, float(N) * T)
        edge_bwd = 16.0 * E + 8.0 * C * E + 8.0 * N * T + 4.0 * N * Fo * T + 4.0 * Fo * touched
        out = {
            "stencil_fwd": 8.0 * N * Fi * T, "stencil_bwd": 8.0 * N * Fi * T,
            "spmm_fwd": spmm(Fi), "spmm_bwd": spmm(Fi),
            "gemm_fwd": 4.0 * N * (Fi + Fo) * T + 4.0 * Fi * Fo,
            "gemm_bwd": 2 * (4.0 * N * (Fi + Fo) * T + 4.0 * Fi * Fo),
            "readout_fwd": edge_fwd, "readout_bwd": edge_bwd,
        }
        if self.bwd_mode == "lowrank":
            J = 2 * C   # every backward stage works on the (T*N, 2C) factor; only expand_dH is F wide
            out.update({
                "readout_bwd": 16.0 * E + 4.0 * C * E + 8.0 * N * T + 4.0 * J * N * T,        # class sums
                "gemm_bwd": 4.0 * N * Fi * T + 4.0 * J * N * T,                             # G = P^T S (one pass over P)
                "spmm_bwd": spmm(J),
                "stencil_bwd": 8.0 * N * J * T,
                "expand_dH": 4.0 * N * Fi * T + 4.0 * J * N * T,
            })
        return out
