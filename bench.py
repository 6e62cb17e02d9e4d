#!/usr/bin/env python
"""Benchmark of the TM-GCN propagation hot path on B200 (bench contract: see README / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = one TM-GCN layer forward + backward (dense M-transform stencil, facewise
SpMM, feature GEMM, edge readout + classifier, and the backward of all of them) over
one shard of the synthetic dynamic graph.  Metric: slice-edges/s = sum_t nnz(A~_t)
processed per second, whole job.  Workload = BASELINE.json configs[4] cut to what one
GPU holds: N = 2M nodes, ~22M stored entries per input slice, b = 10, F = 128 -> 128,
T = 32 slices per GPU (T = 256 over 8 GPUs, weak scaling), rho = 0.9.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line (rank 0): library banners (e.g. NCCL's version line) are sent to stderr
# by pointing fd 1 at fd 2 and keeping a private handle on the real stdout for the result.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

import torch  # noqa: E402

SEED = 20261017


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nodes", type=int, default=2_000_000)
    ap.add_argument("--slices", type=int, default=32, help="time slices per GPU")
    ap.add_argument("--pairs", type=int, default=10_000_000, help="undirected pairs per slice (m)")
    ap.add_argument("--rho", type=float, default=0.9)
    ap.add_argument("--band", type=int, default=10)
    ap.add_argument("--feat", type=int, default=128)
    ap.add_argument("--classes", type=int, default=2)
    ap.add_argument("--act", default="none")
    ap.add_argument("--bwd", default="auto", choices=["auto", "dense", "lowrank"],
                    help="backward formulation (auto = low-rank when the layer is linear, see layer_step.py)")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU forward halo: fused into the stencil over NVLink peer memory, or NCCL send/recv")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-nodes", type=int, default=100_000, help="bounded CPU sample: nodes")
    ap.add_argument("--cpu-slices", type=int, default=8)
    return ap.parse_args()


def measured_traffic(args, T_local):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this very config
    (profiles/r01_traffic.json), or None when the config differs."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        d = json.load(open(p))
        c = d["config"]
        same = (c["nodes"] == args.nodes and c["slices"] == T_local and c["pairs"] == args.pairs
                and abs(c["rho"] - args.rho) < 1e-12 and c["band"] == args.band and c["feat"] == args.feat)
        for k, v in d["kernels"].items():
            if same and k.startswith("void spmm_rows<4, 32, 0"):
                return v["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        # median over the samples taken under load (upper half: idle samples sit at the low end)
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (reference-as-is structure: fp64
# M-transform + per-slice sparse.mm into an fp32 buffer, autograd backward)
# ----------------------------------------------------------------------------------
def cpu_sample(args, steps=1, warmup=0):
    import oracle
    from tmgcn_b200 import synth
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    N, T, b, F, C = args.cpu_nodes, args.cpu_slices, args.band, args.feat, args.classes
    m = max(int(args.pairs * (N / args.nodes)), 1)
    idx, val = synth.synth_coo(N, T, m, args.rho, seed=SEED, device="cpu")
    M = oracle.create_matrix_M(T, b)
    t_mp = time.perf_counter()
    ai, av = oracle.func_MProduct(idx.numpy(), val.numpy(), (T, N, N), M.numpy())
    t_mp = time.perf_counter() - t_mp
    At = oracle.split_slices(ai, av, T, N)
    nnz = int(ai.shape[1])
    g = torch.Generator().manual_seed(SEED)
    H = torch.rand(T, N, F, generator=g)
    W = torch.randn(F, F, generator=g) / F ** 0.5
    U = torch.randn(2 * F, C, generator=g)
    E = m * T // 8
    pick = torch.sort(torch.randint(0, nnz, (E,), generator=g)).values
    edges = torch.from_numpy(ai[:, pick.numpy()])
    dOut = torch.randn(E, C, generator=g)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        oracle.layer_fwd_bwd(At, H, M, W, U, edges, dOut, args.act, as_reference=True)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return {"value": nnz / sec, "unit": "slice-edges/s", "cores": cores, "kind": "port",
            "sample": f"oracle port of ehf:307-312,342-355 + autograd backward (reference-as-is slice-assign loop), "
                      f"N={N} T={T} m={m} rho={args.rho} b={b} F={F}: {nnz} slice-edges, {sec:.2f} s/step",
            "seconds_per_step": sec, "slice_edges": nnz, "mtransform_sparse_s": t_mp}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    r = cpu_sample(args, steps=max(1, min(args.steps, 3)), warmup=1 if args.warmup else 0)
    line = {
        "impl": "reference", "metric": "TM-GCN layer fwd+bwd slice-edges/s", "value": r["value"],
        "unit": "slice-edges/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64/f32 (reference dtypes)", "data": "synthetic",
        "config": workload_name(args),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "slice-edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


def workload_name(args, T_local=None):
    T_local = args.slices if T_local is None else T_local
    return {"workload": f"configs[4] shard: synthetic dynamic graph N={args.nodes}, m={args.pairs} pairs/slice "
                        f"(~{2 * args.pairs + args.nodes} stored entries/input slice), rho={args.rho}, b={args.band}, "
                        f"F={args.feat}->{args.feat}, C={args.classes}, T={T_local}/GPU "
                        f"(T={T_local * args.gpus} total, time-sharded), E=m*T/8 readout edges, act={args.act}",
            "l2_policy": "inputs exceed L2 (each stage streams >= 1 GB per slice; 126 MB L2)",
            "seed": SEED}


# ----------------------------------------------------------------------------------
# CUDA arm
# ----------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import tmgcn_b200 as tg
    from tmgcn_b200 import _lib, ops, sharding, synth
    from tmgcn_b200.layer_step import LayerStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load(build_if_missing=False)

    N, b, F, C = args.nodes, args.band, args.feat, args.classes
    T_local = args.slices
    halo = (b - 1) if rank > 0 else 0

    # ---- memory plan: 3 work buffers + H (+halo) + A~ and its transpose + inputs ----
    free, total = torch.cuda.mem_get_info()
    est_nnz_t = (2 * args.pairs + N) * (1 + (b - 1) * (1 - args.rho) * 1.05)

    def need(Tl):
        # steady state: H + three work buffers, A~ and its transpose, edge ids + incidence list
        # (the input tensor and the transpose scratch are freed before the dense buffers exist)
        dense = 4.0 * N * F * (4 * Tl + 2 * halo)
        sparse = 2 * (est_nnz_t * Tl * 8 + 8.0 * N * Tl)
        edges = args.pairs * Tl / 8 * (16 + 16 + 24)
        return dense + sparse + edges + 4e9
    while T_local > b and need(T_local) > 0.97 * free:
        T_local //= 2
    if world > 1:
        tl = torch.tensor([T_local], device=dev)
        dist.all_reduce(tl, op=dist.ReduceOp.MIN)
        T_local = int(tl.item())
    T_total = T_local * world
    t0, t1 = rank * T_local, (rank + 1) * T_local

    # ---- inputs (setup, untimed) ----
    M = tg.create_matrix_M(T_total, b)
    band = tg.Band(M)
    A_own = synth.synth_csr(N, T_local, args.pairs, args.rho, seed=SEED + rank)
    A_in = sharding.exchange_sparse_halo(A_own, halo_out=b - 1, rank=rank, world=world) if world > 1 else A_own
    At = ops.mtransform_sparse(A_in, band, t0, t1, halo)          # cold run (also the one the bench uses)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_tr = float("inf")
    for _ in range(2):   # warm runs, timed on the device: plan, scan, fill.  The first one may pay a cudaMalloc of
        ev0.record()     # the 10 GB output inside the region; the second reuses the block the first one freed.
        At2 = ops.mtransform_sparse(A_in, band, t0, t1, halo)
        ev1.record()
        torch.cuda.synchronize()
        t_tr = min(t_tr, ev0.elapsed_time(ev1) * 1e-3)
        assert torch.equal(At2.rowptr, At.rowptr) and torch.equal(At2.col, At.col) and torch.equal(At2.val, At.val)
        del At2
    nnz_in = A_in.nnz
    tr_bytes = 8.0 * nnz_in + 4.0 * (N + 1) * (T_local + halo) + 8.0 * At.nnz + 4.0 * (N + 1) * T_local
    del A_own, A_in
    torch.cuda.empty_cache()
    E = args.pairs * T_local // 8
    edges = synth.synth_edges(At, E, seed=SEED + rank)
    plan = tg.EdgePlan(edges, N)
    del edges
    step = LayerStep(At, band, plan, F, F, C, args.act, t0, t1, halo, bwd_mode=args.bwd)
    torch.cuda.empty_cache()
    gen = torch.Generator(device=dev).manual_seed(SEED + 100 + rank)
    peer, halo_mode = None, ("none" if world == 1 else args.halo)
    if world > 1 and args.halo == "peer":
        try:        # layer input in symmetric memory: the boundary stencil reads the predecessor's HBM over NVLink
            peer = sharding.PeerHalo(T_local, N, F, b - 1, rank, world, dev)
        except Exception as ex:  # pragma: no cover - depends on the box
            print(f"[bench] symmetric memory unavailable ({ex}); falling back to the NCCL halo", file=sys.stderr)
            halo_mode = "nccl"
        ok = torch.tensor([1 if peer is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            peer, halo_mode = None, "nccl"
    H = peer.H if peer is not None else torch.empty(T_local + halo, N, F, device=dev)
    for t in range(H.shape[0]):
        H[t].copy_(torch.rand(N, F, generator=gen, device=dev))
    gw = torch.Generator().manual_seed(SEED)
    W = (torch.randn(F, F, generator=gw) / F ** 0.5).to(dev)
    U = torch.randn(2 * F, C, generator=gw).to(dev)
    dOut_host = torch.randn(E, C, generator=torch.Generator().manual_seed(SEED + 7 + rank)).pin_memory()
    dOut = dOut_host.to(dev)
    out_host = torch.empty(E, C).pin_memory()
    dW_host = torch.empty(F, F).pin_memory()
    dU_host = torch.empty(2 * F, C).pin_memory()
    slice_edges_local = At.nnz
    comm = sharding.ShardComm(b - 1, rank, world, dev) if world > 1 else None
    copy_stream = torch.cuda.Stream(device=dev)
    ev_h2d, ev_fwd, ev_prev = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()

    def one_step(e2e):
        main = torch.cuda.current_stream()
        if e2e:
            # host -> device copy of this step's loss gradient on the copy stream: it is only needed by
            # the backward, so it overlaps the forward (after the previous step has released the buffer)
            ev_prev.record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_prev)
                dOut.copy_(dOut_host, non_blocking=True)
                ev_h2d.record(copy_stream)
        # with several ranks the halo exchange (fwd and bwd) and the dW/dU all-reduce run on the
        # communication stream inside forward()/backward(), overlapped with interior slices
        step.forward(H, W, U, comm, peer)
        if e2e:
            ev_fwd.record(main)
            with torch.cuda.stream(copy_stream):        # logits go back while the backward runs
                copy_stream.wait_event(ev_fwd)
                out_host.copy_(step.out, non_blocking=True)
            main.wait_event(ev_h2d)
        dH, dW, dU = step.backward(dOut, W, U, comm)
        if e2e:
            dW_host.copy_(dW, non_blocking=True)
            dU_host.copy_(dU, non_blocking=True)
            main.synchronize()
            copy_stream.synchronize()                   # the step's results are on the host
        return dH

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(K, e2e, stage_times=None):
        events = []

        def hook(name):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            events.append((name, ev))
        step.hook = hook if stage_times is not None else None
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        s.record()
        for _ in range(K):
            one_step(e2e)
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        step.hook = None
        if stage_times is not None:
            per_step = {}
            for (n1, e1), (n2, e2) in zip(events[:-1], events[1:]):
                if n1 != "end":
                    per_step[n1] = per_step.get(n1, 0.0) + e1.elapsed_time(e2)
            for k, v in per_step.items():      # a stage may run in several launches per step: sum them
                stage_times.setdefault(k, []).append(v / K)
        if world > 1:
            tms = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms, _lib.launch_count() - n0

    for _ in range(max(args.warmup, 3)):
        one_step(False)
    sampler = ClockSampler(local) if rank == 0 else None
    stage_times = {}
    ms, launches = timed(args.steps, False, stage_times)
    clocks = sampler.stop() if sampler else None
    one_step(True)
    ms_e2e, _ = timed(args.steps, True)
    ms_dense = None
    if step.bwd_mode == "lowrank":     # for transparency also time the general (dense-gradient) backward
        step.bwd_mode = "dense"
        one_step(False)
        ms_dense, _ = timed(args.steps, False)
        step.bwd_mode = "lowrank"

    stages_all = None
    if world > 1:
        mine = {k: round(sum(v) / len(v), 3) for k, v in stage_times.items()}
        stages_all = [None] * world
        dist.all_gather_object(stages_all, mine)
    tot = torch.tensor([slice_edges_local, nnz_in], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot)
    slice_edges, nnz_in_total = float(tot[0].item()), float(tot[1].item())

    if rank == 0:
        peak, peak_src = peaks()
        l2 = torch.cuda.get_device_properties(dev).L2_cache_size
        alg = step.algorithmic_bytes(l2)
        stages = {k: sum(v) / len(v) for k, v in stage_times.items()}
        for k in ("stencil_fwd", "spmm_fwd", "gemm_fwd", "readout_fwd", "readout_bwd", "gemm_bwd", "spmm_bwd",
                  "stencil_bwd"):
            stages.setdefault(k, float("nan"))
        dom = "spmm_fwd"
        achieved = alg[dom] / (stages[dom] * 1e-3) / 1e9
        per_stage = {k: {"ms": round(stages[k], 3),
                         "GB/s": round(alg[k] / (stages[k] * 1e-3) / 1e9, 1) if k in alg else None}
                     for k in stages}
        sec = ms * 1e-3 / args.steps
        sec_e2e = ms_e2e * 1e-3 / args.steps
        layer_bytes = sum(alg.values())
        line = {
            "metric": "TM-GCN layer fwd+bwd slice-edges/s", "value": slice_edges / sec, "unit": "slice-edges/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_name(args, T_local),
            "slice_edges_per_step": slice_edges, "input_edges_per_step": nnz_in_total,
            "clocks": clocks,
            "e2e": {"value": slice_edges / sec_e2e, "unit": "slice-edges/s",
                    "h2d_bytes_per_step": dOut_host.numel() * 4,
                    "d2h_bytes_per_step": (out_host.numel() + dW_host.numel() + dU_host.numel()) * 4,
                    "ms_per_step": sec_e2e * 1e3,
                    "what": "LayerStep.forward+backward through the C ABI; per step dOut (E x C, the host-side loss "
                            "gradient) comes from pinned host memory and logits + dW + dU go back to the host; "
                            "H, A~ and the edge list stay device-resident as the reference's ctor caches them "
                            "(ehf:195-198)"},
            "gpu_launches": launches,
            "halo": {"forward": halo_mode,
                     "note": "peer = boundary stencil loads the predecessor's b-1 slices from its HBM over NVLink "
                             "(symmetric memory), fused into the kernel; nccl = send/recv on a side stream"},
            "backward": {"mode": step.bwd_mode,
                         "note": "lowrank = exact re-association of the backward through the rank-2C factor the "
                                 "C-class readout hands back (linear layer, act=none); dense = general path",
                         "dense_ms_per_step": None if ms_dense is None else ms_dense / args.steps,
                         "dense_value": None if ms_dense is None else slice_edges / (ms_dense * 1e-3 / args.steps)},
            "roofline": {"bound": "hbm", "kernel": "spmm_rows (forward SpMM, all slices in one launch)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(args, T_local) if world == 1 else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg[dom]},
            "layer_hbm_frac": layer_bytes / sec / 1e9 / peak,
            "layer_algorithmic_bytes": layer_bytes,
            "stages": per_stage,
            "stages_ms_per_rank": stages_all,
            "mtransform_sparse": {"seconds": t_tr, "transform_edges_per_s": slice_edges_local / t_tr,
                                  "algorithmic_bytes": tr_bytes, "GB/s": tr_bytes / t_tr / 1e9,
                                  "hbm_frac": tr_bytes / t_tr / 1e9 / peak,
                                  "note": "stage (a), rank-0 shard: count pass + scan + fill pass, best of two warm runs, CUDA events "
                                          "(includes the host read of the output size); bit-identical to the cold run"},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_sample(args)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as ex:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": "slice-edges/s", "cores": None, "kind": "port",
                                        "sample": f"failed: {ex}"}
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
