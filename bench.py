#!/usr/bin/env python
"""Benchmark of the TM-GCN propagation hot path on B200 (bench contract: see README / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--preset NAME] [--no-extras]

A "step" = one TM-GCN layer forward + backward (dense M-transform stencil, facewise SpMM, feature GEMM,
edge readout + classifier, and the backward of all of them) over one shard of the synthetic dynamic
graph.  Metric: slice-edges/s = sum_t nnz(A~_t) processed per second, whole job.

Workloads (`--preset`; every field can be overridden on the command line):
  c5shard (default)  BASELINE.json configs[4] cut to what one GPU holds: N = 2M nodes, ~22M stored entries per
                     input slice, b = 10, F = 128 -> 128, T = 32 slices per GPU (T = 256 over 8 GPUs): WEAK scaling
  c5cut              a fits-one-GPU cut of configs[4] with the total work fixed: N = 500k, m = 2.5M, T = 128,
                     b = 10 (T/G = 16 >= b-1 at 8 GPUs): STRONG scaling
  c4                 configs[3], Reddit shape: N = 55 863, T = 178, b = 20, F = 128: STRONG scaling
  c1f128             configs[0] shape at F = 128 (N = 5 881, T = 95, b = 20): small enough for the reference's
                     CPU path to run the WHOLE config, so both arms measure the same thing
The default invocation times c5shard (the headline line) and then, with fewer steps, c5cut and c4
(`strong_scaling`), c1f128 against the CPU port on the same data (`same_config`, N = 1 only) and, at N > 1,
a sharded-vs-unsharded numerical check (`parity_multi_gpu`) before any timing.
"""
import argparse
import gc
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
METRIC = "TM-GCN layer fwd+bwd slice-edges/s"

PRESETS = {
    "c5shard": dict(nodes=2_000_000, pairs=10_000_000, rho=0.9, band=10, feat=128, classes=2, slices=32,
                    scaling="weak", label="configs[4] shard"),
    "c5cut": dict(nodes=500_000, pairs=2_500_000, rho=0.9, band=10, feat=128, classes=2, total_slices=128,
                  scaling="strong", label="configs[4] cut to fit one GPU (fixed total work)"),
    "c4": dict(nodes=55_863, pairs=32_000, rho=0.9, band=20, feat=128, classes=2, total_slices=178,
               scaling="strong", label="configs[3] Reddit shape"),
    "c1f128": dict(nodes=5_881, pairs=2_580, rho=0.9, band=20, feat=128, classes=2, total_slices=95,
                   scaling="strong", label="configs[0] Bitcoin-OTC shape at F=128", host_data=True),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default="c5shard", choices=sorted(PRESETS))
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"])
    ap.add_argument("--nodes", type=int, default=None)
    ap.add_argument("--slices", type=int, default=None, help="time slices per GPU (weak scaling)")
    ap.add_argument("--total-slices", type=int, default=None, help="time slices of the whole tensor (strong scaling)")
    ap.add_argument("--pairs", type=int, default=None, help="undirected pairs per slice (m)")
    ap.add_argument("--rho", type=float, default=None)
    ap.add_argument("--band", type=int, default=None)
    ap.add_argument("--feat", type=int, default=None)
    ap.add_argument("--classes", type=int, default=None)
    ap.add_argument("--act", default="none")
    ap.add_argument("--bwd", default="auto", choices=["auto", "dense", "lowrank"],
                    help="backward formulation (auto = low-rank when the layer is linear, see layer_step.py)")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU forward halo: fused into the stencil over NVLink peer memory, or NCCL send/recv")
    ap.add_argument("--no-extras", action="store_true",
                    help="time only the selected workload (no strong_scaling / same_config / parity legs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lean", action="store_true",
                    help="profiling runs: only the warm-up and the timed main steps of the selected workload "
                         "(no e2e / dense / stage-(a) timing, no extras, no CPU baseline)")
    ap.add_argument("--cpu-nodes", type=int, default=50_000, help="bounded CPU sample: nodes")
    ap.add_argument("--cpu-slices", type=int, default=8, help="bounded CPU sample: slices")
    return ap.parse_args()


def resolve_workload(args, preset=None, cli=True):
    """preset (+ explicit command-line overrides) -> workload dict"""
    wl = dict(PRESETS[preset or args.preset])
    wl["name"] = preset or args.preset
    if cli:
        for k in ("nodes", "pairs", "rho", "band", "feat", "classes", "slices", "scaling"):
            v = getattr(args, k)
            if v is not None:
                wl[k] = v
        if args.total_slices is not None:
            wl["total_slices"] = args.total_slices
    wl["act"] = args.act
    if wl["scaling"] == "weak" and "slices" not in wl:
        wl["slices"] = wl.pop("total_slices")
    if wl["scaling"] == "strong" and "total_slices" not in wl:
        wl["total_slices"] = wl.pop("slices")
    return wl


def describe(wl, T_total, world, blocks=None, flush=False):
    N, m = wl["nodes"], wl["pairs"]
    shard = (f"T={wl['slices']}/GPU (T={T_total} total, weak scaling)" if wl["scaling"] == "weak" else
             f"T={T_total} total over {world} GPU(s) (strong scaling: fixed total work)")
    d = {"workload": f"{wl['label']} [{wl['name']}]: synthetic dynamic graph N={N}, m={m} pairs/slice "
                     f"(~{2 * m + N} stored entries/input slice), rho={wl['rho']}, b={wl['band']}, "
                     f"F={wl['feat']}->{wl['feat']}, C={wl['classes']}, {shard}, time-sharded in nnz-balanced "
                     f"contiguous blocks, E=m*T/8 readout edges, act={wl['act']}",
         "l2_policy": "L2 flushed between timed steps (a 256 MB buffer is rewritten before every step)" if flush else
                      "inputs exceed L2 (every stage streams several times the 126 MB L2 per step)",
         "seed": SEED}
    if blocks is not None:
        d["time_blocks"] = [list(x) for x in blocks]
    return d


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def tf32_peak():
    """TF32 tensor-core peak measured on this pool's B200 by scripts/measure_tf32_peak.py (cuBLAS, 8192^3),
    committed as profiles/r02_tf32_peak.json -- MEASURED_PEAKS.json holds only the bf16 figure."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_tf32_peak.json")))
        return float(d["tf32_tflops"]), float(d["tf32_tflops_sustained"])
    except Exception:
        return None, None


def measured_traffic(wl, T_own, world):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this very config
    (profiles/r02_traffic.json, else r01), or None when the config differs."""
    if world != 1:
        return None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
            c = d["config"]
            same = (c["nodes"] == wl["nodes"] and c["slices"] == T_own and c["pairs"] == wl["pairs"]
                    and abs(c["rho"] - wl["rho"]) < 1e-12 and c["band"] == wl["band"] and c["feat"] == wl["feat"]
                    and c.get("generator", "chain") == "global")
            for k, v in d["kernels"].items():
                if same and k.startswith("void spmm_rows<4, 32, 0"):
                    return v["dram_bytes_per_launch"]
        except Exception:
            pass
    return None


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
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (reference-as-is structure: fp64
# M-transform + per-slice sparse.mm into an fp32 buffer, autograd backward)
# ----------------------------------------------------------------------------------
def cpu_run(wl, N, T, m, steps, warmup, what):
    """`steps` timed + `warmup` untimed layer fwd+bwd steps of the CPU port on an (N, T, m) graph with the
    workload's density / band / widths.  Returns the measurement with the configuration it REALLY ran."""
    import oracle
    from tmgcn_b200 import synth
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    b, F, C = wl["band"], wl["feat"], wl["classes"]
    idx, val = synth.synth_coo(N, T, m, wl["rho"], seed=SEED, device="cpu")
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
    E = max(m * T // 8, 1)
    pick = torch.sort(torch.randint(0, nnz, (E,), generator=g)).values
    edges = torch.from_numpy(ai[:, pick.numpy()])
    dOut = torch.randn(E, C, generator=g)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        oracle.layer_fwd_bwd(At, H, M, W, U, edges, dOut, wl["act"], as_reference=True)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return {"value": nnz / sec, "unit": "slice-edges/s", "cores": cores, "kind": "port",
            "sample": f"{what}: oracle port of ehf:307-312,342-355 + autograd backward (reference-as-is slice-assign "
                      f"loop), N={N} T={T} m={m} rho={wl['rho']} b={b} F={F} C={C}: {nnz} slice-edges, "
                      f"{sec:.2f} s/step over {len(times)} timed step(s)",
            "seconds_per_step": sec, "steps_run": len(times), "warmup_run": warmup, "slice_edges": nnz,
            "mtransform_sparse_s": t_mp, "ran": {"nodes": N, "slices": T, "pairs": m, "edges": E}}


def cpu_sample_shape(args, wl):
    """what the CPU arm runs for this workload: the whole thing when it is small enough (host_data presets),
    else a bounded sample of the same density / band / widths"""
    T_total = wl.get("total_slices", wl.get("slices"))
    if wl.get("host_data"):
        return wl["nodes"], T_total, wl["pairs"], True
    N = min(args.cpu_nodes, wl["nodes"])
    T = min(args.cpu_slices, T_total)
    return N, T, max(int(wl["pairs"] * (N / wl["nodes"])), 1), False


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    wl = resolve_workload(args)
    N, T, m, whole = cpu_sample_shape(args, wl)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    r = cpu_run(wl, N, T, m, steps, warmup,
                "the whole workload" if whole else "bounded sample of the workload (same density, band and widths)")
    T_total = wl.get("total_slices", wl.get("slices", 0) * max(1, args.gpus))
    cfg = {"workload": f"CPU arm ran N={N}, T={T}, m={m} pairs/slice, rho={wl['rho']}, b={wl['band']}, "
                       f"F={wl['feat']}->{wl['feat']}, C={wl['classes']}, E={r['ran']['edges']} readout edges, "
                       f"act={wl['act']} on {r['cores']} host cores"
                       + ("" if whole else f" -- a bounded sample of [{wl['name']}] (the GPU arm's N={wl['nodes']}, "
                                           f"m={wl['pairs']}, T={T_total}); its rate is an extrapolation, not a "
                                           f"same-config measurement"),
           "sample_of": describe(wl, T_total, max(1, args.gpus))["workload"], "same_config_as_gpu_arm": whole,
           "extrapolated": not whole, "seed": SEED}
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "slice-edges/s", "n_gpus": args.gpus,
        "steps": r["steps_run"], "warmup": r["warmup_run"], "ms_per_step": r["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
        "dtype": "f64/f32 (reference dtypes)", "data": "synthetic", "config": cfg,
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "slice-edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


# ----------------------------------------------------------------------------------
# CUDA arm
# ----------------------------------------------------------------------------------
class Ctx:
    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.args = args
        self.peer_storage = None          # one symmetric allocation shared by every workload of the run
        self.halo_mode = "none" if self.world == 1 else args.halo
        self.l2 = torch.cuda.get_device_properties(self.dev).L2_cache_size
        self.flush_buf = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def all_min(self, v):
        if self.world == 1:
            return v
        t = torch.tensor([v], device=self.dev, dtype=torch.int64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return int(t.item())

    def ensure_peer_storage(self, numel):
        """symmetric memory for the layer input of every workload (allocated once, sized for the largest)"""
        from tmgcn_b200 import sharding
        if self.world == 1 or self.halo_mode != "peer":
            return
        if self.peer_storage is not None and self.peer_storage[0].numel() >= numel:
            return
        ok = 1
        try:
            self.peer_storage = sharding.PeerHalo.allocate(numel, self.dev)
        except Exception as ex:  # pragma: no cover - depends on the box
            print(f"[bench] symmetric memory unavailable ({ex}); falling back to the NCCL halo", file=sys.stderr)
            ok = 0
        if self.all_min(ok) == 0:
            self.peer_storage, self.halo_mode = None, "nccl"

    def flush_l2(self):
        if self.flush_buf is None:
            self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.flush_buf.fill_(1)


def plan_blocks(wl, ctx):
    """-> (T_total, blocks) with nnz-balanced contiguous time blocks; shrinks a weak-scaling shard that
    does not fit this GPU (all ranks agree)."""
    from tmgcn_b200 import sharding
    N, b, F, m, rho = wl["nodes"], wl["band"], wl["feat"], wl["pairs"], wl["rho"]
    world, rank = ctx.world, ctx.rank
    free, _ = torch.cuda.mem_get_info()
    est_nnz_t = (2 * m + N) * (1 + (b - 1) * (1 - rho) * 1.05)

    def blocks_for(T_total):
        w = sharding.slice_weight_estimate(T_total, b, N, m, rho)
        return sharding.balanced_bounds(w, world) if world > 1 else [(0, T_total)]

    def need(Tl, halo):
        # steady state: H + three work buffers, A~ and its transpose, edge ids + incidence list
        dense = 4.0 * N * F * (4 * Tl + 2 * halo)
        sparse = 2 * (est_nnz_t * Tl * 8 + 8.0 * N * Tl)
        edges = m * Tl / 8 * (16 + 16 + 24)
        return dense + sparse + edges + 4e9

    if wl["scaling"] == "weak":
        per = wl["slices"]
        while True:
            T_total = per * world
            blocks = blocks_for(T_total)
            t0, t1 = blocks[rank]
            fits = 1 if (need(t1 - t0, (b - 1) if rank > 0 else 0) <= 0.97 * free or per <= b) else 0
            if ctx.all_min(fits):
                break
            per //= 2
        wl["slices"] = per
    else:
        T_total = wl["total_slices"]
        blocks = blocks_for(T_total)
        t0, t1 = blocks[rank]
        if not ctx.all_min(1 if need(t1 - t0, (b - 1) if rank > 0 else 0) <= 0.97 * free else 0):
            raise RuntimeError(f"workload [{wl['name']}] does not fit {world} GPU(s)")
    return T_total, blocks


def run_workload(wl, ctx, steps, warmup, full):
    """Build the shard of workload `wl` this rank owns, time `steps` layer steps (max over ranks) and return the
    measurements.  full = True adds the e2e timing, the clock sampler and the stage-(a) timing (headline line)."""
    import tmgcn_b200 as tg
    from tmgcn_b200 import _lib, ops, sharding, synth
    from tmgcn_b200.layer_step import LayerStep
    dist, world, rank, dev = ctx.dist, ctx.world, ctx.rank, ctx.dev
    N, b, F, C, m, rho = wl["nodes"], wl["band"], wl["feat"], wl["classes"], wl["pairs"], wl["rho"]
    T_total, blocks = plan_blocks(wl, ctx)
    t0, t1 = blocks[rank]
    T_own = t1 - t0
    halo = (b - 1) if rank > 0 else 0
    if world > 1:
        sharding.assert_single_hop(T_own, b - 1, world, dev)
    flush = 4.0 * N * F * T_own <= 4 * ctx.l2          # small workloads would otherwise be timed out of L2

    # ---- inputs (setup, untimed): this rank's window of ONE global dynamic graph -------------------------
    M = tg.create_matrix_M(T_total, b)
    band = tg.Band(M)
    if wl.get("host_data"):      # small config: the legacy seeded chain on the CPU, the same data the CPU arm uses
        idx, val = synth.synth_coo(N, T_total, m, rho, seed=SEED, device="cpu")
        sel = (idx[0] >= t0) & (idx[0] < t1)
        own = idx[:, sel].clone()
        own[0] -= t0
        A_own = tg.SliceCSR.from_coo(own, val[sel], T_own, N)
        del idx, val, own, sel
    else:
        A_own = synth.synth_csr(N, T_own, m, rho, seed=SEED, t_start=t0)
    A_in = sharding.exchange_sparse_halo(A_own, halo_out=b - 1, rank=rank, world=world) if world > 1 else A_own
    At = ops.mtransform_sparse(A_in, band, t0, t1, halo)          # cold run (also the one the bench uses)
    torch.cuda.synchronize()
    t_tr, tr_bytes = None, None
    nnz_in = A_in.nnz
    if full:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_tr = float("inf")
        for _ in range(3):   # warm runs, timed on the device: plan, scan, fill.  The first ones may pay a cudaMalloc
            ev0.record()     # of the outputs / the workspace inside the region; later ones reuse the freed blocks.
            At2 = ops.mtransform_sparse(A_in, band, t0, t1, halo)
            ev1.record()
            torch.cuda.synchronize()
            t_tr = min(t_tr, ev0.elapsed_time(ev1) * 1e-3)
            assert torch.equal(At2.rowptr, At.rowptr) and torch.equal(At2.col, At.col) and torch.equal(At2.val, At.val)
            del At2
        tr_bytes = 8.0 * nnz_in + 4.0 * (N + 1) * (T_own + halo) + 8.0 * At.nnz + 4.0 * (N + 1) * T_own
    del A_own, A_in
    torch.cuda.empty_cache()
    E = max(m * T_own // 8, 1)
    edges = synth.synth_edges(At, E, seed=SEED + rank)
    plan = tg.EdgePlan(edges, N, T=T_own)
    del edges
    step = LayerStep(At, band, plan, F, F, C, wl["act"], t0, t1, halo, bwd_mode=ctx.args.bwd)
    torch.cuda.empty_cache()
    gen = torch.Generator(device=dev).manual_seed(SEED + 100 + rank)
    peer, halo_mode = None, ctx.halo_mode
    if world > 1 and halo_mode == "peer":
        tmax = max(x1 - x0 for x0, x1 in blocks)
        ctx.ensure_peer_storage(tmax * N * F)
        halo_mode = ctx.halo_mode
        if halo_mode == "peer":
            peer = sharding.PeerHalo(T_own, N, F, b - 1, rank, world, dev, storage=ctx.peer_storage)
    h_in = 0 if peer is not None else halo
    H = peer.H if peer is not None else torch.empty(T_own + h_in, N, F, device=dev)
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
    comm = sharding.ShardComm(b - 1, rank, world, dev, T_own=T_own) if world > 1 else None
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

    host_ms = [0.0]

    def timed(K, e2e, stage_times=None):
        events = []

        def hook(name):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            events.append((name, ev))
        step.hook = hook if stage_times is not None else None
        ctx.barrier()
        n0 = _lib.launch_count()
        if flush:
            # small workload: rewrite a buffer larger than L2 before every step; each step is timed on its own
            # pair of events so the flush stays outside the timed region
            ms, host = 0.0, 0.0
            for _ in range(K):
                ctx.flush_l2()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                h0 = time.perf_counter()
                one_step(e2e)
                host += time.perf_counter() - h0
                e.record()
                torch.cuda.synchronize()
                ms += s.elapsed_time(e)
            ctx.barrier()
        else:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            h0 = time.perf_counter()
            for _ in range(K):
                one_step(e2e)
            host = time.perf_counter() - h0          # how long the host needed to ENQUEUE the K steps
            e.record()
            ctx.barrier()
            ms = s.elapsed_time(e)
        host_ms[0] = host * 1e3 / K
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

    def note(msg):           # progress on stderr (rank 0): which leg a failure belongs to
        if rank == 0:
            print(f"[bench] {wl['name']} x{world}: {msg}", file=sys.stderr, flush=True)
    note(f"setup done (T_own={T_own}, nnz={At.nnz}, halo={halo_mode}, bwd={step.bwd_mode})")
    W_ = max(warmup, 3)
    for _ in range(W_):
        one_step(False)
    torch.cuda.synchronize()
    note("warm-up done")
    sampler = ClockSampler(ctx.local) if (rank == 0 and full) else None
    stage_times = {}
    ms, launches = timed(steps, False, stage_times)
    host_enqueue_ms = host_ms[0]
    note(f"main timed: {ms / steps:.3f} ms/step")
    clocks = sampler.stop() if sampler else None
    ms_e2e = None
    if full:
        one_step(True)
        ms_e2e, _ = timed(steps, True)
        note(f"e2e timed: {ms_e2e / steps:.3f} ms/step")
    alg = step.algorithmic_bytes(ctx.l2)
    ms_dense, dense_stage_times, alg_dense = None, {}, None
    if step.bwd_mode == "lowrank" and not ctx.args.lean:     # the general (dense-gradient) layer of SURVEY 8(d)
        step.bwd_mode = "dense"
        one_step(False)
        torch.cuda.synchronize()
        note("dense warm-up done")
        ms_dense, _ = timed(steps, False, dense_stage_times)
        note(f"dense timed: {ms_dense / steps:.3f} ms/step")
        alg_dense = step.algorithmic_bytes(ctx.l2)
        step.bwd_mode = "lowrank"

    def gather_stages(st):
        mine = {k: round(sum(v) / len(v), 3) for k, v in st.items()}
        if world == 1:
            return [mine]
        allr = [None] * world
        dist.all_gather_object(allr, mine)
        return allr
    stages_all = gather_stages(stage_times)
    dense_stages_all = gather_stages(dense_stage_times) if ms_dense is not None else None
    tot = torch.tensor([At.nnz, nnz_in], device=dev, dtype=torch.float64)
    per_rank_nnz = [float(At.nnz)]
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, float(At.nnz))
        per_rank_nnz = allr
        dist.all_reduce(tot)
    slice_edges, nnz_in_total = float(tot[0].item()), float(tot[1].item())

    def stage_table(stages, algb):
        out = {}
        for k, v in stages.items():
            out[k] = {"ms": round(v, 3), "GB/s": round(algb[k] / (v * 1e-3) / 1e9, 1) if (k in algb and v > 0) else None}
        return out

    sec = ms * 1e-3 / steps
    res = {
        "config": describe(wl, T_total, world, blocks, flush),
        "scaling": wl["scaling"], "T_total": T_total, "ms_per_step": sec * 1e3, "value": slice_edges / sec,
        "steps": steps, "warmup": W_, "slice_edges_per_step": slice_edges, "input_edges_per_step": nnz_in_total,
        "slice_edges_per_rank": per_rank_nnz, "gpu_launches": launches, "halo_mode": halo_mode,
        "backward_mode": step.bwd_mode, "clocks": clocks, "host_enqueue_ms_per_step": host_enqueue_ms,
        "stages": stage_table(stages_all[0], alg), "stages_ms_per_rank": stages_all if world > 1 else None,
        "alg": alg, "layer_algorithmic_bytes": sum(alg.values()),
        "E": E, "F": F, "C": C,
    }
    if ms_e2e is not None:
        sec_e2e = ms_e2e * 1e-3 / steps
        res["e2e"] = {"value": slice_edges / sec_e2e, "unit": "slice-edges/s",
                      "h2d_bytes_per_step": dOut_host.numel() * 4,
                      "d2h_bytes_per_step": (out_host.numel() + dW_host.numel() + dU_host.numel()) * 4,
                      "ms_per_step": sec_e2e * 1e3,
                      "what": "LayerStep.forward+backward through the C ABI; per step dOut (E x C, the host-side loss "
                              "gradient) comes from pinned host memory and logits + dW + dU go back to the host; "
                              "H, A~ and the edge list stay device-resident as the reference's ctor caches them "
                              "(ehf:195-198)"}
    if ms_dense is not None:
        sd = ms_dense * 1e-3 / steps
        res["dense"] = {"ms_per_step": sd * 1e3, "value": slice_edges / sd,
                        "stages": stage_table(dense_stages_all[0], alg_dense),
                        "stages_ms_per_rank": dense_stages_all if world > 1 else None,
                        "layer_algorithmic_bytes": sum(alg_dense.values()), "alg": alg_dense}
    if t_tr is not None:
        res["mtransform_sparse"] = {"seconds": t_tr, "transform_edges_per_s": At.nnz / t_tr,
                                    "algorithmic_bytes": tr_bytes, "GB/s": tr_bytes / t_tr / 1e9,
                                    "note": "stage (a), rank-0 shard: count pass (records the union pattern) + scan + "
                                            "union-list fill pass, best of three warm runs, CUDA events (includes the "
                                            "workspace allocation and the host reads of the output size and of the "
                                            "overflow counter); bit-identical to the cold run"}
    res["T_own"] = T_own
    # release everything before the next workload
    del step, At, plan, H, peer, comm, dOut, W, U
    gc.collect()
    torch.cuda.empty_cache()
    return res


def module_fresh_input_leg(args, ctx, with_cpu=True, calls=15):
    """The reference's OTHER call form, `gcn(At, X, edges)` with fresh inputs (ehf:212-215: every evaluation,
    experiment_bitcoin_our.py:134,145), end to end through the module API with HOST inputs: a new Python list of
    CPU sparse slices, a CPU fp64 X and CPU edges go in, CPU logits come out, every call (list -> device CSR,
    M-transform of X, facewise SpMM, GEMM, readout; no_grad as in the reference's evaluation).  Config:
    configs[0] shape with the experiment's own widths (N = 5 881, T = 95, b = 20, F 2 -> 6 -> 2, 1 layer,
    experiment_bitcoin_our.py:109).  The CPU arm runs the oracle port of the same call on the same data."""
    import oracle
    import tmgcn_b200 as tg
    from tmgcn_b200 import synth
    N, T, m, b, F0, F1, C = 5_881, 95, 2_580, 20, 2, 6, 2
    idx, val = synth.synth_coo(N, T, m, 0.9, seed=SEED, device="cpu")
    M = oracle.create_matrix_M(T, b)
    ai, av = oracle.func_MProduct(idx.numpy(), val.numpy(), (T, N, N), M.numpy())
    nnz = int(ai.shape[1])
    At = [a.coalesce() for a in oracle.split_slices(ai, av, T, N)]
    g = torch.Generator().manual_seed(SEED)
    X = torch.rand(T, N, F0, generator=g, dtype=torch.float64)
    E = m * T // 8
    pick = torch.sort(torch.randint(0, nnz, (E,), generator=g)).values
    edges = torch.from_numpy(ai[:, pick.numpy()])
    torch.manual_seed(SEED)
    gcn = tg.EmbeddingGCN(At, X, edges, M, hidden_feat=[F1, C], condensed_W=True, use_Minv=False)
    h2d = nnz * (4 + 4 + 4) + T * 8 + X.numel() * 8 + edges.numel() * 8     # (row, col) int32 + fp32 value per entry
    d2h = E * C * 4
    with torch.no_grad():
        for _ in range(2):
            out = gcn(list(At), X, edges).cpu()
        torch.cuda.synchronize()
        per_call = []
        for _ in range(calls):
            t0 = time.perf_counter()
            out = gcn(list(At), X, edges).cpu()          # a NEW list object: the module's CSR cache cannot hit
            torch.cuda.synchronize()
            per_call.append(time.perf_counter() - t0)
        # host-side wall clock on a shared box: the median call (the mean of 5 calls moved 8 -> 90 ms between boxes)
        sec = sorted(per_call)[len(per_call) // 2]
    res = {"what": "module call gcn(At, X, edges) with fresh HOST inputs (ehf:212-215), no_grad, logits back on the "
                   "host; wall clock per call including list -> CSR conversion and all copies",
           "config": f"configs[0] shape, reference widths: N={N}, T={T}, b={b}, F {F0}->{F1}->{C}, 1 layer "
                     f"(EmbeddingGCN), E={E}, {nnz} slice-edges",
           "value": nnz / sec, "unit": "slice-edges/s", "ms_per_call": sec * 1e3,
           "ms_per_call_min_mean_max": [min(per_call) * 1e3, sum(per_call) / len(per_call) * 1e3, max(per_call) * 1e3],
           "calls": calls, "statistic": "median", "h2d_bytes_per_call": h2d, "d2h_bytes_per_call": d2h}
    if with_cpu:
        ref = oracle.OracleGCN(At, X, edges, M, gcn.W.detach().cpu(), gcn.U.detach().cpu(), as_reference=True)
        with torch.no_grad():
            ref(At, X, edges)
            per_r = []
            for _ in range(min(calls, 7)):
                t0 = time.perf_counter()
                out_r = ref(At, X, edges)
                per_r.append(time.perf_counter() - t0)
            sec_r = sorted(per_r)[len(per_r) // 2]
        err = ((out.double() - out_r.double()).abs().max() / out_r.double().abs().max()).item()
        res["cpu"] = {"value": nnz / sec_r, "unit": "slice-edges/s", "ms_per_call": sec_r * 1e3,
                      "cores": len(os.sched_getaffinity(0)), "kind": "port"}
        res["ratio"] = sec_r / sec
        res["rel_err_vs_cpu_port"] = err
    del gcn
    gc.collect()
    torch.cuda.empty_cache()
    return res


def roofline_of(res, stages_key, dom, peak, peak_src, traffic=None):
    st = res[stages_key] if stages_key == "stages" else res["dense"]["stages"]
    algb = res["alg"] if stages_key == "stages" else res["dense"]["alg"]
    achieved = algb[dom] / (st[dom]["ms"] * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "spmm_rows (forward SpMM, all slices in one launch)", "achieved": achieved,
            "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": algb[dom]}


def slim(res, peak):
    """the part of a workload's measurements that goes into the strong_scaling / same_config sub-objects"""
    out = {k: res[k] for k in ("config", "scaling", "T_total", "ms_per_step", "value", "steps", "warmup",
                               "slice_edges_per_step", "slice_edges_per_rank", "halo_mode", "backward_mode",
                               "host_enqueue_ms_per_step", "stages", "stages_ms_per_rank")}
    out["layer_hbm_frac"] = res["layer_algorithmic_bytes"] / (res["ms_per_step"] * 1e-3) / 1e9 / peak
    if "dense" in res:
        out["dense"] = {k: res["dense"][k] for k in ("ms_per_step", "value", "stages", "stages_ms_per_rank")}
    return out


def run_ours(args):
    from tmgcn_b200 import _lib
    if args.lean:
        args.no_extras = args.no_cpu_baseline = True
    ctx = Ctx(args)
    _lib.load(build_if_missing=False)
    world, rank = ctx.world, ctx.rank
    peak, peak_src = peaks()
    extras = not args.no_extras and args.preset == "c5shard"

    parity = None
    if world > 1 and not args.no_extras:
        from tmgcn_b200 import selfcheck
        try:
            parity = selfcheck.multi_gpu_parity(rank, world, ctx.dev, try_peer=(ctx.halo_mode == "peer"))
        except Exception as ex:  # the check must never hide a crash: report it as a failed check
            parity = {"ok": False, "error": f"{type(ex).__name__}: {ex}"}
        gc.collect()
        torch.cuda.empty_cache()

    wl = resolve_workload(args)
    res = run_workload(wl, ctx, args.steps, args.warmup, full=not args.lean)

    strong, same_cfg, fresh = {}, None, None
    if extras:
        k_x = max(3, min(args.steps, 10))
        for name in ("c5cut", "c4"):
            try:
                r = run_workload(resolve_workload(args, name, cli=False), ctx, k_x, 3, full=False)
                strong[name] = slim(r, peak)
            except Exception as ex:
                strong[name] = {"error": f"{type(ex).__name__}: {ex}"}
        if world == 1:
            try:
                wl1 = resolve_workload(args, "c1f128", cli=False)
                r = run_workload(wl1, ctx, k_x, 3, full=False)
                same_cfg = {"gpu": slim(r, peak)}
                if not args.no_cpu_baseline:
                    cb = cpu_run(wl1, wl1["nodes"], wl1["total_slices"], wl1["pairs"], 1, 0, "the whole workload")
                    same_cfg["cpu"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
                    same_cfg["ratio"] = r["value"] / cb["value"]
                    same_cfg["ratio_dense_backward"] = r["dense"]["value"] / cb["value"] if "dense" in r else None
                    same_cfg["note"] = ("both arms ran this whole workload on the same seeded graph (generated on the "
                                        "host); the CPU arm is one timed step without warm-up")
            except Exception as ex:
                same_cfg = {"error": f"{type(ex).__name__}: {ex}"}
            try:
                fresh = module_fresh_input_leg(args, ctx, with_cpu=not args.no_cpu_baseline)
            except Exception as ex:
                fresh = {"error": f"{type(ex).__name__}: {ex}"}

    if rank == 0:
        dom = "spmm_fwd"
        line = {
            "metric": METRIC, "value": res["value"], "unit": "slice-edges/s", "n_gpus": world, "steps": res["steps"],
            "warmup": res["warmup"], "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": res["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": res["config"],
            "slice_edges_per_step": res["slice_edges_per_step"], "input_edges_per_step": res["input_edges_per_step"],
            "slice_edges_per_rank": res["slice_edges_per_rank"],
            "clocks": res["clocks"], "e2e": res.get("e2e"), "gpu_launches": res["gpu_launches"],
            "halo": {"forward": res["halo_mode"],
                     "note": "peer = boundary stencil loads the predecessor's b-1 slices from its HBM over NVLink "
                             "(symmetric memory), fused into the kernel; nccl = send/recv on a side stream"},
            "backward": {"mode": res["backward_mode"],
                         "note": "lowrank = exact re-association of the backward through the rank-2C factor the "
                                 "C-class readout hands back (linear layer, act=none); dense = the general layer of "
                                 "SURVEY 8(d) (any activation), timed in the same run: see `dense`",
                         "dense_ms_per_step": res["dense"]["ms_per_step"] if "dense" in res else None,
                         "dense_value": res["dense"]["value"] if "dense" in res else None},
            "roofline": roofline_of(res, "stages", dom, peak, peak_src, measured_traffic(wl, res["T_own"], world)),
            "layer_hbm_frac": res["layer_algorithmic_bytes"] / (res["ms_per_step"] * 1e-3) / 1e9 / peak,
            "layer_algorithmic_bytes": res["layer_algorithmic_bytes"],
            "stages": res["stages"], "stages_ms_per_rank": res["stages_ms_per_rank"],
            "host_enqueue_ms_per_step": res["host_enqueue_ms_per_step"],
        }
        if "dense" in res:
            d = res["dense"]
            line["dense"] = {"what": "the same step with the general dense-gradient backward (what any activation "
                                     "takes): readout bwd -> act' -> dP, dW -> SpMM^T -> stencil^T",
                             "ms_per_step": d["ms_per_step"], "value": d["value"], "stages": d["stages"],
                             "stages_ms_per_rank": d["stages_ms_per_rank"],
                             "layer_hbm_frac": d["layer_algorithmic_bytes"] / (d["ms_per_step"] * 1e-3) / 1e9 / peak,
                             "roofline": roofline_of(res, "dense", "spmm_bwd", peak, peak_src)}
            line["dense"]["roofline"]["kernel"] = "spmm_rows on the transposed CSR (backward SpMM)"
        burst, sustained = tf32_peak()
        g_ms = res["stages"].get("gemm_fwd", {}).get("ms")
        if burst and g_ms:
            rows = res["T_own"] * wl["nodes"] if world == 1 else None
            if rows:
                useful = 2.0 * rows * wl["feat"] * wl["feat"] / (g_ms * 1e-3) / 1e12
                line["gemm_tensor_pipe"] = {
                    "kernel": "gemm_tf32x3_kernel (tcgen05.mma kind::tf32, 3 MMAs per useful product)",
                    "useful_tflops": useful, "executed_tflops": 3.0 * useful, "tf32_peak_burst": burst,
                    "tf32_peak_sustained": sustained, "executed_frac_of_sustained_peak": 3.0 * useful / sustained,
                    "executed_frac_of_burst_peak": 3.0 * useful / burst,
                    "note": "the kernel is HBM-bound (AI = 32 flop/B): see stages.gemm_fwd GB/s; peaks from "
                            "profiles/r02_tf32_peak.json (cuBLAS TF32 8192^3 on this pool's B200)"}
        if "mtransform_sparse" in res:
            ms_ = res["mtransform_sparse"]
            ms_["hbm_frac"] = ms_["GB/s"] / peak
            line["mtransform_sparse"] = ms_
        if parity is not None:
            line["parity_multi_gpu"] = parity
        if strong:
            line["strong_scaling"] = strong
        if same_cfg is not None:
            line["same_config"] = same_cfg
            line["extra"] = {"same_config_ratio": same_cfg.get("ratio")}
        if fresh is not None:
            line["e2e_fresh_inputs"] = fresh
        if world == 1 and not args.no_cpu_baseline:
            try:
                N, T, m, whole = cpu_sample_shape(args, wl)
                cb = cpu_run(wl, N, T, m, 1, 0, "the whole workload" if whole else
                             "bounded sample of the workload (same density, band and widths)")
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as ex:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": "slice-edges/s", "cores": None, "kind": "port",
                                        "sample": f"failed: {ex}"}
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
