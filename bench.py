#!/usr/bin/env python
"""bench.py -- headline benchmark of the MCMC balanced-colouring sweep (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c5|small]

metric  : vertex-updates / second of one synchronous MCMC sweep (edges/s alongside), whole job over all N GPUs
workload: BASELINE config 3 -- Erdos-Renyi n = 100 M, mean degree 16, nCol = maxDeg (numColRatio 1.0) -- which fits
          one B200 (6.8 GB of CSR); synthetic, generated on the device(s), never crossing PCIe.
step    : one sweep over all n vertices starting from the uniform random colouring (every vertex active, ~31 % of
          them conflicting -- the most expensive sweep of a chain).  The colouring is reset between steps outside
          the timed region; the kernel is timed with CUDA events on the stream it is launched on.
e2e     : the same sweep through the reference-shaped host API with HOST buffers: colouring H2D from pinned memory
          (mcmcb200_init_colors), sweep, counters (mcmcb200_status) and colouring D2H (mcmcb200_get_colors).
roofline: algorithmic bytes 8*nnz + 12*n + 4 per sweep (SURVEY 8d / DESIGN.md) over the live kernel time, against the
          measured HBM copy bandwidth in MEASURED_PEAKS.json.
cpu_baseline / --impl reference: the UNMODIFIED reference CPU sampler (oracle/_ref, else the C port) on a bounded
          sample of the same workload, single thread (the reference has no threading).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n, mean degree, description)
    "c3": (100_000_000, 16, "BASELINE config 3: Erdos-Renyi n=100M, mean degree 16, nCol=maxDeg"),
    "c2": (1_000_000, 32, "BASELINE config 2: Erdos-Renyi n=1M, mean degree 32, nCol=maxDeg"),
    "c5": (10_000_000, 16, "BASELINE config 5: Erdos-Renyi n=10M, mean degree 16, nCol=maxDeg"),
    # config 4 = "R-MAT / power-law synthetic graph with 50M vertices and a Reddit-like degree skew".  Reddit (233 K vertices, mean degree
    # 492, maximum 21 657) has max/mean degree 44; R-MAT (0.45,0.15,0.15,0.25) at scale 26 gives max/mean ~ 100 (natural vertex ids: the
    # hubs sit together at the low ids, which is what stresses the degree binning), and the reference's own palette rule nCol = maxDeg
    # (numColRatio 1.0, main.cu:162) stays usable -- a few thousand colours, the wide-palette kernel -- and reaches a proper colouring.
    "c4": (50_000_000, 16, "BASELINE config 4: R-MAT (0.45,0.15,0.15,0.25) scale 26 trimmed to 50M vertices, edge factor 16 (max/mean degree ~100, "
                           "Reddit: 44), nCol=maxDeg"),
    # the round-1 graph: Graph500 parameters, max/mean degree 25 000 (hub rows of ~10^6 neighbours).  No palette a sampler of this kind can
    # use gives a proper colouring here (two adjacent hubs that start with the same colour both see every colour taken and never move), so
    # this one is a load-balance stress test of the kernels only: nCol = 1024.
    "c4heavy": (50_000_000, 16, "R-MAT (0.57,0.19,0.19,0.05) [Graph500] scale 26 trimmed to 50M vertices, edge factor 16, max/mean degree 25 000, nCol=1024 "
                                "(kernel stress test; no proper colouring exists for this sampler)"),
    "c4small": (1_500_000, 16, "R-MAT (0.57,0.19,0.19,0.05) scale 21 trimmed to 1.5M vertices, edge factor 16, nCol=1024"),
    "small": (200_000, 16, "smoke-sized Erdos-Renyi n=200k, mean degree 16"),
}
GRAPH_SEED = 42
# measured DRAM traffic of one sweep (ncu dram__bytes_read.sum + dram__bytes_write.sum, both launches), bytes
TRAFFIC = {"c3": (14.93e9, "profiles/r02_ncu_blocked_kernels_c3.md (ncu dram__bytes of both kernels of one sweep: 7.48 + 1.73 + 5.62 + 0.10 GB)"),
           "c4": (11.59e9, "profiles/r02_config4_wide_palette.md (ncu dram__bytes of the four kernels of one sweep)")}
CHAIN_SEED = 1


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.index, self.proc = [], index, None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def __enter__(self):
        self.t.start()
        time.sleep(0.3)                                    # let the first samples arrive before the timed region
        return self

    def __exit__(self, *a):
        time.sleep(0.15)
        if self.proc is not None:
            self.proc.terminate()
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


RMAT = {"c4": (0.45, 0.15, 0.15), "c4heavy": (0.57, 0.19, 0.19), "c4small": (0.57, 0.19, 0.19)}


def palette_for(workload, max_deg):
    """numColRatio 1.0: nCol = maxDeg (main.cu:162); the R-MAT hubs have degree ~1e6, there the ratio is raised so that
    the palette stays tractable."""
    return min(max_deg, 1024) if workload in ("c4heavy", "c4small") else max_deg


def sample_start(workload, n):
    return (n // 2) if workload.startswith("c4") else 0


def gen_graph_device(n, deg, device, workload="c3"):
    import torch
    from mcmc_colorer_b200.graphgen import er_graph_torch, rmat_graph_torch
    if workload.startswith("c4"):
        scale = max(1, (n - 1).bit_length())
        a, b, c = RMAT[workload]
        rowptr64, neighs, nnz, max_deg = rmat_graph_torch(scale, deg, GRAPH_SEED, n_keep=n, a=a, b=b, c=c, device=device)
    else:
        rowptr64, neighs, nnz, max_deg = er_graph_torch(n, deg, GRAPH_SEED, device=device)
    assert nnz < 2 ** 31, "this bench keeps CSR offsets in int32 tensors"
    rowptr = rowptr64.to(torch.int32)
    del rowptr64
    torch.cuda.synchronize()
    return rowptr, neighs, nnz, max_deg


def cpu_reference_rate(rowptr, neighs, n, nnz, nCol, target_seconds=12.0, steps=1, warmup=0, v0=0):
    """Times the reference CPU sampler (oracle/_ref: the unmodified ColoringMCMC_CPU methods; else the C port) on the
    m vertices [v0, v0+m) of the same graph (v0 = 0 on the Erdos-Renyi workloads; the middle of the id range on R-MAT, whose
    lowest ids are the hubs), single thread (the reference has no threading at all).  m is sized by a short
    calibration so that the whole call costs about target_seconds of CPU work.
    Returns (vertex-updates/s, kind, sample description, per-step seconds, m)."""
    import numpy as np
    from oracle.pyoracle import Port, Ref
    P = Port()
    use_ref = Ref.available()
    colors = P.init_colors(CHAIN_SEED, n, nCol)
    u = P.tape(CHAIN_SEED, 1, n)

    def run_sample(m, steps, warmup):
        e_0, e_m = int(rowptr[v0].item()), int(rowptr[v0 + m].item())
        cumul = np.zeros(n + 1, np.uint32)                # vertices outside the sample: empty rows, never visited
        cumul[v0:v0 + m + 1] = (rowptr[v0:v0 + m + 1].cpu().numpy().astype(np.int64) - e_0).astype(np.uint32)
        cumul[v0 + m + 1:] = e_m - e_0
        nb = neighs[e_0:max(e_m, e_0 + 1)].cpu().numpy().astype(np.uint32)[:e_m - e_0]
        out = []
        if use_ref:
            R = Ref()
            g = R.graph_from_csr(cumul, nb, float(nnz) / n / n)
            h = R.mcmc(g, nCol, CHAIN_SEED)
            for i in range(warmup + steps):
                R.set_colors(h, colors)
                sec = R.sweep_range_timed(h, u, v0, v0 + m)     # the reference's own per-vertex loop, timed inside the harness
                if i >= warmup:
                    out.append(sec)
            R.L.ref_mcmc_free(h)
            R.L.ref_graph_free(g)
        else:
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                P.sweep(cumul, nb, nCol, 1e-8, colors, u, 0, vb=v0, ve=v0 + m)
                if i >= warmup:
                    out.append(time.perf_counter() - t0)
        return out

    m0 = min(n - v0, 500_000)
    rate0 = m0 / run_sample(m0, 1, 0)[0]
    m = int(min(n - v0, max(100_000, rate0 * target_seconds / max(1, steps + warmup))))
    secs = run_sample(m, steps, warmup)
    rate = float(np.mean([m / s for s in secs]))
    kind = "reference" if use_ref else "port"
    sample = (f"vertices [{v0}, {v0 + m}) of the {n} of the same graph ({int(rowptr[v0 + m].item()) - int(rowptr[v0].item())} directed edges), "
              f"full {n}-entry colour array, first sweep from the uniform random colouring")
    return rate, kind, sample, secs, m


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on host cores."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n, deg, desc = WORKLOADS[args.workload]
    if args.n:
        n = args.n
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    if dev == "cpu":
        import numpy as np
        from mcmc_colorer_b200.graphgen import er_graph_numpy
        n = min(n, 2_000_000)
        cumul, nb = er_graph_numpy(n, deg, GRAPH_SEED)
        rowptr, neighs = torch.from_numpy(cumul.astype(np.int64)), torch.from_numpy(nb.astype(np.int64))
        nnz, max_deg = len(nb), int(np.diff(cumul.astype(np.int64)).max())
    else:
        rowptr, neighs, nnz, max_deg = gen_graph_device(n, deg, dev, args.workload)
    nCol = args.ncol if args.ncol else palette_for(args.workload, max_deg)
    rate, kind, sample, secs, m = cpu_reference_rate(rowptr, neighs, n, nnz, nCol, target_seconds=max(20.0, 6.0 * (args.steps + args.warmup)),
                                                     steps=args.steps, warmup=args.warmup, v0=sample_start(args.workload, n))
    ms = 1e3 * sum(secs) / len(secs)
    line = {"impl": "reference", "metric": "vertex_updates_per_sec", "value": rate, "unit": "vertex-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32 colours / f32 CDF", "data": "synthetic",
            "config": {"workload": desc, "n": n, "nnz_directed": nnz, "nCol": nCol, "proposal": "uniform",
                       "step": "bounded sample of one sweep: " + sample},
            "cpu_baseline": {"value": rate, "unit": "vertex-updates/s", "cores": 1, "kind": kind, "sample": sample,
                             "sample_vertices": int(m), "sample_fraction": m / n,
                             "note": "the reference is single-threaded; the rate of the sampled vertex range stands for the whole sweep (per-vertex cost is uniform on this graph)"},
            "e2e": {"value": rate, "unit": "vertex-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "edges_per_sec": rate * nnz / n, "host_cores_available": os.cpu_count()}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="override the vertex count (debugging)")
    ap.add_argument("--proposal", default="uniform", choices=["uniform", "dynamic"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stage-cap-bytes", type=int, default=0, help="tuning experiments: mcmcb200_params.stageCapBytes (0 = automatic)")
    ap.add_argument("--item-bits", type=int, default=0, help="tuning experiments: mcmcb200_params.itemBits (0 = automatic)")
    ap.add_argument("--stage-buffers", type=int, default=0, help="tuning experiments: mcmcb200_params.stageBuffers (0 = automatic)")
    ap.add_argument("--no-overlap", action="store_true", help="tuning experiments: the two passes of the blocked sweep back to back")
    ap.add_argument("--expected-sweeps", type=int, default=0, help="mcmcb200_params.expectedSweeps of the measured handle (0 = many: blocked layout on large graphs)")
    ap.add_argument("--ncol", type=int, default=0, help="override the palette size (default: palette_for(workload, maxDeg))")
    ap.add_argument("--traj", type=int, default=0, help="with --quick: also run a chain of this many sweeps and report its violation trajectory")
    ap.add_argument("--quick", action="store_true", help="tuning experiments: kernel timing only (no e2e, no time-to-colouring, no CPU baseline)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import mcmc_colorer_b200 as mc
    from mcmc_colorer_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: libmcmcb200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        from mcmc_colorer_b200 import multigpu
        return multigpu.bench_main(args, WORKLOADS, GRAPH_SEED, CHAIN_SEED, measured_peak, ClockSampler, palette_for, RMAT)

    n, deg, desc = WORKLOADS[args.workload]
    if args.n:
        n = args.n
    t_gen = time.perf_counter()
    rowptr, neighs, nnz, max_deg = gen_graph_device(n, deg, dev, args.workload)
    t_gen = time.perf_counter() - t_gen
    nCol = args.ncol if args.ncol else palette_for(args.workload, max_deg)
    proposal = mc.PROPOSAL_UNIFORM if args.proposal == "uniform" else mc.PROPOSAL_DYNAMIC
    prm = mc.ColoringMCMCParams(nCol=nCol, proposal=proposal,
                                convergence=mc.CONVERGE_VERTICES if proposal == mc.PROPOSAL_UNIFORM else mc.CONVERGE_EDGES,
                                seed=CHAIN_SEED)
    torch.cuda.synchronize()
    t_create = time.perf_counter()
    ch = mc.Chain(params=prm, device=local_rank, flags=mc.FLAG_NO_EARLY_STOP | (mc.FLAG_NO_OVERLAP if args.no_overlap else 0),
                  n_global=n, v_begin=0, v_end=n, device_csr=(rowptr.data_ptr(), neighs.data_ptr(), nnz),
                  stage_cap_bytes=args.stage_cap_bytes, item_bits=args.item_bits, stage_buffers=args.stage_buffers, expected_sweeps=args.expected_sweeps)
    ch.synchronize()
    create_ms = 1e3 * (time.perf_counter() - t_create)   # mcmcb200_create: allocations + (blocked path) the layout build on the device
    layout_bytes = ch.layout_bytes()
    if args.quick:
        for _ in range(args.warmup):
            ch.init_colors(None); ch.sweep(1); ch.synchronize()
        ms = []
        for _ in range(args.steps):
            ch.init_colors(None); torch.cuda.synchronize(); ch.sweep(1); ms.append(ch.last_sweep_ms())
        ch.init_colors(None); ch.sweep(10); chain_ms = ch.last_sweep_ms() / 10.0
        st = ch.status()
        traj = []
        if args.traj:
            ch.init_colors(None)
            done = 0
            for k in [1, 2, 3, 5, 10, 20, 40, 80, 160, 250]:
                if k > args.traj:
                    break
                ch.sweep(k - done); done = k
                s_ = ch.status()
                traj.append([k, int(s_.conflictEdges), int(s_.violatingVertices), int(s_.usedColors)])
        t = float(np.mean(ms))
        print(json.dumps({"quick": True, "traj": traj, "maxDeg": int(max_deg), "nnz": int(nnz), "workload": args.workload, "ms_per_step": t, "min_ms": float(np.min(ms)), "frac": (8 * nnz + 12 * n + 4) / (t * 1e-3) / 1e9 / measured_peak()[0],
                          "chain_ms_per_sweep": chain_ms, "kernel_mode": ch.kernel_mode(), "create_ms": create_ms, "nCol": nCol,
                          "after_10": [int(st.conflictEdges), int(st.violatingVertices)],
                          "tuning": [args.stage_cap_bytes, args.item_bits, args.stage_buffers, bool(args.no_overlap)], "lib": os.environ.get("MCMCB200_LIB", "")}))
        ch.close()
        return 0

    # ---- device-resident timing: `value` ----
    for _ in range(args.warmup):
        ch.init_colors(None)
        ch.sweep(1)
        ch.synchronize()
    launches0 = ch.launch_count()
    kernel_ms = []
    with ClockSampler(local_rank) as clocks:
        t_wall = time.perf_counter()
        for _ in range(args.steps):
            ch.init_colors(None)                          # untimed: reset to the uniform random colouring
            torch.cuda.synchronize()
            l0 = ch.launch_count()
            ch.sweep(1)                                   # one launch of sweep_kernel, CUDA events on its stream
            kernel_ms.append(ch.last_sweep_ms())
            launches_per_step = ch.launch_count() - l0
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall
        # a chain of consecutive sweeps (no reset), for the time-to-colouring side of the metric
        ch.init_colors(None)
        ch.sweep(10)
        chain_ms = ch.last_sweep_ms() / 10.0
    st = ch.status()
    # time to a proper colouring (the other half of BASELINE's metric), the reference's --tailcut protocol: sweep until at most
    # z = max(50, n/2000) vertices are violated (coloringMCMC_main.cu:150-170), then the greedy tail-cutting repair (:271-290).
    # (At n = 1e8 the epsilon tails alone re-colour ~n*nCol*eps = 45 vertices per sweep, so the plain chain idles at a few
    # dozen violations: the threshold is part of the algorithm, not a shortcut.)  A second handle with params.tailcut = 1 (its
    # mcmcb200_create is timed too: `setup_ms` is the layout build a user pays once per graph); the chain stops ON THE DEVICE at
    # the threshold -- the host only polls every 4 sweeps -- and the repair works from the violator list the last sweep emitted.
    prm_tc = mc.ColoringMCMCParams(nCol=nCol, proposal=proposal, convergence=prm.convergence, seed=CHAIN_SEED, tailcut=True)
    z_tail = max(50, n // 2000)

    def time_to_colouring(expected_sweeps):
        torch.cuda.synchronize()
        t_setup = time.perf_counter()
        c = mc.Chain(params=prm_tc, device=local_rank, flags=(mc.FLAG_NO_OVERLAP if args.no_overlap else 0), n_global=n, v_begin=0, v_end=n,
                     device_csr=(rowptr.data_ptr(), neighs.data_ptr(), nnz), stage_cap_bytes=args.stage_cap_bytes, item_bits=args.item_bits,
                     stage_buffers=args.stage_buffers, expected_sweeps=expected_sweeps)
        c.synchronize()
        setup = 1e3 * (time.perf_counter() - t_setup)
        mode = c.kernel_mode()
        c.init_colors(None)
        c.synchronize()
        t0 = time.perf_counter()
        while True:
            c.sweep(4)
            s_ = c.status()
            if s_.converged or s_.sweep >= 250 or time.perf_counter() - t0 > 5.0:
                break
        sweeps = int(s_.sweep)
        t_sw = time.perf_counter() - t0
        # the repair pass only once the chain is below the threshold (a palette that cannot get there is reported as not proper)
        reached = bool(s_.converged)
        viol_at_z = int(s_.violatingVertices)
        # (c4heavy: hub rows of ~1e6 neighbours see every colour taken -- no repair exists for this sampler, see WORKLOADS; skipped)
        rounds = c.tailcut(64) if (reached and s_.conflictEdges > 0 and args.workload != "c4heavy") else 0
        s_ = c.status()
        t_all = time.perf_counter() - t0
        c.close()
        return {"sweeps": sweeps, "tailcut_rounds": int(rounds), "ms": 1e3 * t_all, "ms_sweeps": 1e3 * t_sw, "setup_ms": setup,
                "total_ms_with_setup": setup + 1e3 * t_all, "kernel_mode": mode, "z": z_tail, "reached_z": reached, "violating_at_z": viol_at_z,
                "proper": bool(s_.conflictEdges == 0 and s_.violatingVertices == 0), "conflictEdges_left": int(s_.conflictEdges),
                "usedColors": int(s_.usedColors), "nCol": nCol}

    ttc = time_to_colouring(args.expected_sweeps)           # layout amortised over many chains (setup reported beside it)
    # one chain on a new graph: expectedSweeps = 8 keeps the layout-free kernels (mcmcb200.h), so the total includes a cheap create
    ttc_one_shot = time_to_colouring(8)
    ms_per_step = float(np.mean(kernel_ms))
    value = n / (ms_per_step * 1e-3)
    alg_bytes = 8 * nnz + 12 * n + 4
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the host API with HOST buffers (pinned), every step: colouring H2D, sweep, counters D2H, colouring D2H ----
    # (a) the library's narrow interface (mcmcb200_{init,get}_colors_narrow: the device's own u8/u16 colour format -- what a caller that
    #     only wants the sweep moves); (b) the reference-shaped uint32 calls (mcmcb200_init_colors / _get_colors: 4 bytes per vertex each way)
    eb = ch.color_bytes()
    nd = torch.uint8 if eb == 1 else torch.int16
    pin_n_in = torch.empty(n, dtype=nd).pin_memory()
    pin_n_out = torch.empty(n, dtype=nd).pin_memory()
    ch.init_colors(None)
    ch.get_colors_narrow_ptr(pin_n_in.data_ptr(), eb)
    e2e_t = []
    for i in range(max(2, args.warmup) + min(args.steps, 5)):
        t0 = time.perf_counter()
        ch.init_colors_narrow_ptr(pin_n_in.data_ptr(), eb)   # H2D n * eb bytes from pinned memory
        ch.sweep(1)
        s2 = ch.status()                                      # counters D2H
        ch.get_colors_narrow_ptr(pin_n_out.data_ptr(), eb)    # D2H n * eb bytes
        dt = time.perf_counter() - t0
        if i >= max(2, args.warmup):
            e2e_t.append(dt)
    e2e_value = n / float(np.mean(e2e_t))
    pinned_in = torch.empty(n, dtype=torch.int32).pin_memory()
    pinned_out = torch.empty(n, dtype=torch.int32).pin_memory()
    ch.init_colors(None)
    ch.get_colors_ptr(pinned_in.data_ptr())
    e2e32_t = []
    for i in range(2 + min(args.steps, 3)):
        t0 = time.perf_counter()
        ch.init_colors_ptr(pinned_in.data_ptr())          # H2D 4n bytes from pinned memory
        ch.sweep(1)
        s2 = ch.status()                                  # counters D2H
        ch.get_colors_ptr(pinned_out.data_ptr())          # D2H 4n bytes
        dt = time.perf_counter() - t0
        if i >= 2:
            e2e32_t.append(dt)
    e2e32_value = n / float(np.mean(e2e32_t))

    line = {
        "metric": "vertex_updates_per_sec", "value": value, "unit": "vertex-updates/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8 colours / u32 ids / f32 CDF", "data": "synthetic",
        "config": {"workload": desc, "n": n, "nnz_directed": nnz, "nCol": nCol, "maxDeg": max_deg, "proposal": args.proposal,
                   "step": "one sweep from the uniform random colouring (all vertices active)",
                   "l2": "inputs (CSR %.1f GB) larger than L2; no flush needed" % ((4 * nnz + 4 * n) / 1e9),
                   "graph_gen_s": round(t_gen, 2), "create_ms": round(create_ms, 1), "layout_bytes": int(layout_bytes),
                   "csr_bytes": int(4 * (n + 1) + 4 * nnz)},
        "edges_per_sec": value * nnz / n,
        "chain_ms_per_sweep": chain_ms,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": TRAFFIC[args.workload][0] if args.workload in TRAFFIC and args.proposal == "uniform" and not args.ncol else None,
                     "traffic_source": TRAFFIC[args.workload][1] if args.workload in TRAFFIC and args.proposal == "uniform" and not args.ncol else None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                     "kernel": {"direct": "sweep_kernel (one launch per sweep)",
                                "direct-binned": "binned_sweep_kernel (one launch per sweep; thread / warp / CTA rows by degree)",
                                "wide-binned": "wide_tables_kernel + wide_sweep_kernel (palettes above 512 colours: thread / warp / CTA rows by degree, colour "
                                               "lists and shared-memory bitmaps)",
                                "blocked": "blocked_gather_kernel then blocked_sweep_kernel (two launches per sweep)",
                                "blocked-overlapped": "blocked_gather_kernel || blocked_sweep_kernel (two launches per sweep, concurrent on two "
                                                      "streams; launch_ms = CUDA events around the pair)"}[ch.kernel_mode()],
                     "kernel_mode": ch.kernel_mode(),
                     "launch_ms": ms_per_step, "launches_per_sweep": int(launches_per_step)},
        "e2e": {"value": e2e_value, "unit": "vertex-updates/s", "h2d_bytes_per_step": eb * n,
                "d2h_bytes_per_step": eb * n + 40 + 8 * nCol, "ms_per_step": 1e3 * float(np.mean(e2e_t)),
                "api": "mcmcb200_init_colors_narrow + mcmcb200_sweep + mcmcb200_status + mcmcb200_get_colors_narrow (u%d colours)" % (8 * eb),
                "u32_interface": {"value": e2e32_value, "ms_per_step": 1e3 * float(np.mean(e2e32_t)), "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": 4 * n + 40 + 8 * nCol,
                                  "api": "mcmcb200_init_colors + mcmcb200_sweep + mcmcb200_status + mcmcb200_get_colors (the reference's uint32 colour layout)"}},
        "gpu_launches": int(launches_per_step * args.steps),
        "clocks": clocks.summary(),
        "after_10_chain_sweeps": {"conflictEdges": int(st.conflictEdges), "violatingVertices": int(st.violatingVertices)},
        "time_to_proper_coloring": dict(ttc, one_shot=ttc_one_shot,
                                        setup_note="setup_ms = wall time of mcmcb200_create: layout-construction kernels (75 ms on config 3) + ~30 cudaMalloc/"
                                                   "cudaFree of GB-sized buffers; 90-140 ms on an idle box (profiles/r02_create_times.jsonl), more on a shared one"),
    }
    if not args.no_cpu_baseline:
        rate, kind, sample, secs, m = cpu_reference_rate(rowptr, neighs, n, nnz, nCol, target_seconds=12.0, v0=sample_start(args.workload, n))
        line["cpu_baseline"] = {"value": rate, "unit": "vertex-updates/s", "cores": 1, "kind": kind, "sample": sample,
                                "sample_vertices": int(m), "sample_fraction": m / n, "host_cores_available": os.cpu_count()}
    ch.close()
    print(json.dumps(line))
    return 0


class _QuietStdout:
    """fd 1 is pointed at stderr while the benchmark runs (NCCL prints its version banner to stdout), so that the ONE
    JSON line is the only thing the caller finds on stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        self.py_stdout = sys.stdout
        sys.stdout = os.fdopen(os.dup(self.saved), "w")
        return self

    def __exit__(self, *a):
        sys.stdout.flush()
        sys.stdout = self.py_stdout
        os.dup2(self.saved, 1)
        os.close(self.saved)


if __name__ == "__main__":
    with _QuietStdout():
        rc = main()
    sys.exit(rc)
