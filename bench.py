#!/usr/bin/env python
"""Benchmark of the Chebyshev graph-wavelet feature path (BASELINE.json metric:
"Chebyshev-wavelet nnz*K*F/s & HBM GB/s at 1/2/4/8 B200 vs ref CPU").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload reddit|arxiv|physics|pubmed|cora] [--f F] [--order K] [--scales S]

One "step" = one pass of the hot path over one synthetic graph of the named
shape: K fused Chebyshev orders (CSR SpMM over the implicit scaled Laplacian +
recurrence + scale accumulation + L1 normalisation).  Default workload: the
Reddit shape (232,965 nodes, 114.6 M stored entries) with the reference's
defaults K=3, S=1 (s=0.8), F=1 (X0 = log1p(degree)) - the largest named
configuration; it fits one B200 and is the only one whose CSR streams from HBM.

Prints ONE JSON line (see README/DESIGN.md for the keys).  `value` is timed with
the graph resident in HBM; `e2e` is the same metric through the host-buffer
entry (pinned host CSR -> H2D -> degree pass -> orders -> D2H of the features).
`--impl reference` times the CPU oracle port of the reference's scipy path
(the reference itself is pure Python + scipy and is not present on the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "chebyshev_wavelet_nnz_k_f_per_s"
UNIT = "nnz*K*F/s"
DEFAULT_F = {"reddit": 1, "arxiv": 128, "physics": 1, "pubmed": 1, "cora": 1}
CPU_RATE_GUESS = 2.0e7      # nnz*K*F/s of the oracle port on one host core (sizes the bounded CPU sample)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="reddit", choices=sorted(DEFAULT_F))
    ap.add_argument("--f", type=int, default=None, help="feature columns of X0 (default: per workload)")
    ap.add_argument("--order", type=int, default=3, help="Chebyshev order K")
    ap.add_argument("--scales", type=int, default=1, help="number of wavelet scales S")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="do not replay the step as a CUDA graph")
    ap.add_argument("--entry", default="csr", choices=["csr", "dense"],
                    help="e2e entry: pinned host CSR (default) or the reference's own boundary, a dense [N,N] float32 "
                         "adjacency (calibration/WATS.py:99; small shapes only)")
    ap.add_argument("--replicas", action="store_true",
                    help="N>1: UGCA replica mode - every rank holds the whole graph and recomputes the features of its "
                         "own perturbation candidates (--flips, default 5); no exchange, weak scaling")
    ap.add_argument("--no-sell", action="store_true", help="F=1: use the generic CSR kernel instead of the SELL plan")
    ap.add_argument("--flips", type=int, default=0,
                    help="UGCA mode (N=1): every step recomputes the features of the graph with this many symmetric "
                         "edge flips around a random target node applied on top of the resident graph")
    ap.add_argument("--check", action="store_true", help="N>1: compare every rank's rows with the single-GPU path")
    return ap.parse_args()


def scale_list(n_scales):
    base = [0.8, 0.4, 1.6, 3.2, 0.2, 6.4, 0.1, 12.8]
    return base[0] if n_scales == 1 else base[:n_scales]


def algorithmic_bytes(n, nnz, f, k_max, n_scales):
    """Compulsory-traffic model B_k of SURVEY 8d (int32 indices, fp32 data,
    binary adjacency) for each order k = 1..K."""
    out = []
    for k in range(1, k_max + 1):
        t_terms = 1 + (1 if k >= 2 else 0) + (1 if k < k_max else 0)
        acc = 1 if k == 1 else 2
        out.append(4 * nnz + 4 * (n + 1) + 4 * n + 4 * n * f * t_terms + 4 * n_scales * n * f * acc)
    return out


class ClockSampler(threading.Thread):
    """Polls NVML for SM clock / throttle reasons while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def poll(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "hw_power_brake_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        for name, bit in names.items():
            if r & bit:
                self.reasons.add(name)

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag:
            try:
                self.poll()
            except Exception:
                break
            time.sleep(0.002)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# --------------------------------------------------------------------------- #
# CPU arm: the oracle port of the reference's scipy path                        #
# --------------------------------------------------------------------------- #
def cpu_sample_graph(workload, target_seconds, k_max, f, gen_device):
    """A bounded sample of the workload: the same generator at a reduced scale
    (nodes and entries shrunk together, same mean degree)."""
    import scipy.sparse as sp
    from efficient_gnn_b200 import synth
    sh = synth.SHAPES[workload]
    want_nnz = CPU_RATE_GUESS * target_seconds / max(1, k_max)      # F enters through the recurrence only
    scale = float(min(1.0, max(2e-3, want_nnz / sh.nnz)))
    rp, ci, n = synth.synth_csr(workload, self_loops=True, device=gen_device, scale=scale)
    rp, ci = rp.cpu().numpy(), ci.cpu().numpy()
    adj = sp.csr_matrix((np.ones(ci.size, np.float32), ci, rp), shape=(n, n))
    return adj, scale


def time_oracle(adj, k_max, scales, f, repeats=1):
    from oracle import wats_oracle as orc
    x0 = None
    if f > 1:
        x0 = np.random.default_rng(3).standard_normal((adj.shape[0], f)).astype(np.float32)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.wavelet_features(adj, k=k_max, s=scales, x0=x0)
        best = min(best, time.perf_counter() - t0)
    return best


def run_reference(args, rank, world):
    """`--impl reference`: the CPU implementation of the path (oracle port of
    calibration/WATS.py:39-74 over scipy, single-threaded like the reference)
    on a bounded sample per step."""
    if rank != 0:
        return
    f = args.f or DEFAULT_F[args.workload]
    scales = scale_list(args.scales)
    total = max(1, args.steps + args.warmup)
    per_step = min(10.0, max(0.5, 150.0 / total))
    adj, scale = cpu_sample_graph(args.workload, per_step, args.order, f, "cpu")
    for _ in range(args.warmup):
        time_oracle(adj, args.order, scales, f)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        time_oracle(adj, args.order, scales, f)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = adj.nnz * args.order * f / dt
    sample = (f"{args.workload}-shape generator at scale {scale:.4f} (N={adj.shape[0]}, nnz={adj.nnz}), "
              f"K={args.order}, S={args.scales}, F={f}; one full path per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}-shape", "k": args.order, "scales": args.scales, "f": f,
                   "sample_scale": scale, "n": int(adj.shape[0]), "nnz": int(adj.nnz)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- #
# GPU arm                                                                       #
# --------------------------------------------------------------------------- #
def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist
    import efficient_gnn_b200 as egnn
    from efficient_gnn_b200 import synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    f = args.f or DEFAULT_F[args.workload]
    k_max, n_scales = args.order, args.scales
    scales = scale_list(n_scales)
    sh = synth.SHAPES[args.workload]

    if world > 1 and args.replicas:
        return run_replicas(args, rank, local_rank, world)
    if world > 1:
        from efficient_gnn_b200 import sharded
        return sharded.bench_entry(args, rank, local_rank, world, METRIC, UNIT, algorithmic_bytes,
                                   ClockSampler, physical_gpu_index, scale_list)

    # synthetic graph of the named shape, generated directly in HBM
    rp, ci, n = synth.synth_csr(args.workload, self_loops=True, device=dev)
    graph = egnn.CsrGraph(rp, ci, None, n)
    nnz = graph.nnz
    if f == 1:
        x0 = None                                  # reference default: log1p(degree), produced by the degree pass
    else:
        x0 = torch.randn(n, f, device=dev, generator=torch.Generator(device=dev).manual_seed(sh.seed))
    work = float(nnz) * k_max * f

    use_sell = False if args.no_sell else None
    flips = None
    if args.flips > 0:               # calib_fga.py:897-904: budget symmetric flips incident to one target node
        gen = torch.Generator().manual_seed(7)
        picks = torch.randint(0, n, (args.flips + 1,), generator=gen).tolist()
        target, others = picks[0], [j for j in picks[1:] if j != picks[0]]
        flips = ([target] * len(others) + others, others + [target] * len(others), [1.0] * (2 * len(others)))

    def step(events=None):
        return egnn.graph_wavelet_features(graph, k=k_max, s=scales, X0=x0, deltas=flips, _order_events=events,
                                           _use_sell=use_sell)

    # per-order CUDA events (recorded by the library on the launching stream) on every 8th step of the
    # timed region: recording them on every step costs ~9 % of the step (6 records + a ctypes array)
    EV_STRIDE = 8
    ev_steps = list(range(0, args.steps, EV_STRIDE))
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2 * k_max)] for _ in ev_steps]
    for row in ev:                                 # materialise the handles
        for e in row:
            e.record()
    torch.cuda.synchronize()
    import ctypes as C
    ev_arrays = [(C.c_void_p * (2 * k_max))(*[e.cuda_event for e in row]) for row in ev]

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    # steps without events replay the same pass as one CUDA graph (public API: WaveletSession)
    session = None
    if flips is None and not args.no_graph and not args.no_sell:
        session = egnn.WaveletSession(graph, k=k_max, s=scales, f=f)
        if x0 is not None:
            session.x0.copy_(x0)
        for _ in range(3):
            session()
        torch.cuda.synchronize()

    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for i in range(args.steps):
        if i % EV_STRIDE == 0:
            step(ev_arrays[i // EV_STRIDE])
        elif session is not None:
            session()
        else:
            step()
    stop.record()
    torch.cuda.synchronize()
    sampler.stop_flag = True
    sampler.join()
    ms_per_step = start.elapsed_time(stop) / args.steps
    value = work / (ms_per_step * 1e-3)

    order_ms = np.array([[row[2 * j].elapsed_time(row[2 * j + 1]) for j in range(k_max)] for row in ev])
    avg_launch_ms = float(order_ms.mean())
    b_k = algorithmic_bytes(n, nnz, f, k_max, n_scales)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = (sum(b_k) / k_max) / (avg_launch_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.isfile(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"{args.workload}_f{f}_k{k_max}_s{n_scales}")
        except Exception:
            traffic = None
    if f == 1 and k_max >= 1 and not args.no_sell and graph.sell_plan() is not None:
        kernel_name = "sell_spmv_kernel"
    else:
        kernel_name = "cheb_wide_kernel" if f >= 8 else "cheb_order_kernel"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel_name, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": sum(b_k) / k_max, "avg_launch_ms": avg_launch_ms,
                "per_order_ms": [float(v) for v in order_ms.mean(axis=0)],
                "order_kernel_share_of_step": float(order_ms.sum(axis=1).mean() / ms_per_step)}

    # end to end through the host-buffer entry: pinned CSR -> device -> features -> host
    e2e = None
    if not args.no_e2e:
        rp_h = graph.rowptr.cpu().pin_memory()
        ci_h = graph.colidx.cpu().pin_memory()
        x0_h = None if x0 is None else x0.cpu().pin_memory()
        out_h = torch.empty((n, n_scales * f), dtype=torch.float32).pin_memory()
        h2d = rp_h.numel() * 4 + ci_h.numel() * 4 + (0 if x0_h is None else x0_h.numel() * 4)
        d2h = out_h.numel() * 4

        entry = "CsrGraph.from_host_csr + graph_wavelet_features (pinned host CSR in, host features out)"
        dense_h = None
        if args.entry == "dense":
            if 4 * n * n > 8e9:
                raise SystemExit("--entry dense needs the [N,N] float32 adjacency to fit comfortably; use cora/pubmed/physics")
            import scipy.sparse as sp
            adj = sp.csr_matrix((np.ones(nnz, np.float32), graph.colidx.cpu().numpy(), graph.rowptr.cpu().numpy()),
                                shape=(n, n))
            dense_h = torch.from_numpy(adj.toarray()).pin_memory()
            h2d = dense_h.numel() * 4 + (0 if x0_h is None else x0_h.numel() * 4)
            entry = ("graph_wavelet_features(dense [N,N] float32 adjacency): pinned host dense in -> device "
                     "dense->CSR kernels -> features -> host (the reference boundary, calibration/WATS.py:99-100)")

        def e2e_step():
            if dense_h is not None:
                g = egnn.CsrGraph.from_dense(dense_h.to(dev, non_blocking=True))
            else:
                g = egnn.CsrGraph.from_host_csr(rp_h, ci_h, None, n, device=dev)
            xx = None if x0_h is None else x0_h.to(dev, non_blocking=True)
            feats = egnn.graph_wavelet_features(g, k=k_max, s=scales, X0=xx, _use_sell=use_sell)
            out_h.copy_(feats, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        e_steps = max(3, min(args.steps, 20))
        for _ in range(3):
            e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e_dt = (time.perf_counter() - t0) / e_steps
        e2e = {"value": work / e_dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e_dt * 1e3, "steps": e_steps, "entry": entry}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        adj, scale = cpu_sample_graph(args.workload, 12.0, k_max, f, dev)
        t_cpu = time_oracle(adj, k_max, scales, f)
        cpu_baseline = {"value": adj.nnz * k_max * f / t_cpu, "unit": UNIT, "cores": 1, "kind": "port",
                        "host_cores": os.cpu_count(), "seconds": t_cpu,
                        "sample": (f"{args.workload}-shape generator at scale {scale:.4f} (N={adj.shape[0]}, "
                                   f"nnz={adj.nnz}), same K/S/F; oracle port of calibration/WATS.py:39-74 "
                                   "(scipy, single-threaded like the reference), one run")}

    # kernels of ours per step: SELL path = prescale + K x (SpMV + epilogue); wide path = padded prescale +
    # K orders (+ L1 normalisation when the row spans several feature tiles); narrow generic = prescale + K
    if kernel_name == "sell_spmv_kernel":
        launches_per_step = 1 + 2 * k_max
    elif f >= 8:
        launches_per_step = 1 + k_max + (1 if (f + 3) // 4 * 4 > 128 else 0)
    else:
        launches_per_step = k_max + (1 if f <= 4 else 0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}-shape", "n": n, "nnz": nnz, "k": k_max, "scales": n_scales, "f": f,
                   "self_loops": True, "parallelism": "1 GPU", "ugca_flips": int(args.flips),
                   "cuda_graph": session is not None, "event_stride": EV_STRIDE, "l2_policy": (
                       "inputs larger than L2 (CSR %.0f MB vs 126 MB L2), no flush" % (4 * nnz / 1e6)
                       if 4 * nnz > 126e6 else "inputs fit in L2: latency-bound configuration, no flush")},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
        "gpu_launches": int(launches_per_step * args.steps),
        "clocks": sampler.summary(),
    }
    print(json.dumps(line), flush=True)


def run_replicas(args, rank, local_rank, world):
    """UGCA replica mode: candidates are independent, so each rank keeps the whole
    graph resident and recomputes the features of its own candidates (edge flips around a target
    node applied on top of the resident CSR / SELL plan).  No data-path collective."""
    import torch.distributed as dist
    import efficient_gnn_b200 as egnn
    from efficient_gnn_b200 import synth
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    f = args.f or DEFAULT_F[args.workload]
    k_max, n_scales = args.order, args.scales
    scales = scale_list(n_scales)
    rp, ci, n = synth.synth_csr(args.workload, self_loops=True, device=dev)
    graph = egnn.CsrGraph(rp, ci, None, n)
    nnz = graph.nnz
    x0 = None if f == 1 else torch.randn(n, f, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    budget = args.flips or 5
    gen = torch.Generator().manual_seed(1000 + rank)
    cands = []
    for _ in range(args.steps + max(3, args.warmup)):
        picks = torch.randint(0, n, (budget + 1,), generator=gen).tolist()
        t, others = picks[0], [j for j in picks[1:] if j != picks[0]]
        cands.append(([t] * len(others) + others, others + [t] * len(others), [1.0] * (2 * len(others))))
    it = iter(cands)
    for _ in range(max(3, args.warmup)):
        egnn.graph_wavelet_features(graph, k=k_max, s=scales, X0=x0, deltas=next(it))
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        egnn.graph_wavelet_features(graph, k=k_max, s=scales, X0=x0, deltas=next(it))
    stop.record()
    torch.cuda.synchronize()
    sampler.stop_flag = True
    sampler.join()
    ms = torch.tensor([start.elapsed_time(stop) / args.steps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.barrier()
    ms_per_step = float(ms.item())
    if rank == 0:
        work = float(nnz) * k_max * f * world            # every rank finishes one candidate per step
        line = {
            "metric": METRIC, "value": work / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}-shape", "n": n, "nnz": nnz, "k": k_max, "scales": n_scales, "f": f,
                       "self_loops": True, "ugca_flips": budget,
                       "parallelism": f"{world} replicas: whole graph per GPU, one perturbation candidate per rank "
                                      "per step, no exchange"},
            "roofline": None, "cpu_baseline": None, "e2e": None,
            "gpu_launches": int((1 + 2 * k_max + 1) * args.steps * world), "clocks": sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
