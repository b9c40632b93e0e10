"""Vertex-partitioned multi-GPU sweeps: one process per GPU, torch.distributed (NCCL over NVLink) for the exchange.

The reference is single-GPU; this is new work defined by the north star.  The path shards by vertex:

  * contiguous vertex ranges [vBegin, vEnd), one per rank, each rank holding only its CSR rows (global neighbour ids);
  * the narrow (u8/u16) colour array is replicated and double buffered; a sweep writes the owned slice of C_{t+1};
  * per sweep ONE real exchange step:  all-gather of the owned colour slices (n bytes in total for u8 colours) and
    one all-reduce of int64[2 + nCol] = {directed conflicts, violating vertices, class-size deltas};
  * draws are Philox keyed by the GLOBAL vertex id, so trajectories are bit-identical for 1, 2, 4 and 8 GPUs.

The driver below is engine-agnostic: `engine` is anything with the small interface of GpuEngine (the B200 engine over
libmcmcb200).  The CPU tests drive the same driver over gloo with a test double of that interface (defined in
tests/, backed by the CPU checker), which covers the partition arithmetic, the exchange layout and the convergence protocol without a GPU.
"""
import os
import time

import numpy as np

from . import capi
from .colorer import Chain, ColoringMCMCParams


def partition(n, world, align=256):
    """Equal chunks of `chunk` vertices (a multiple of `align`, the sweep kernel's tile) -> [(vBegin, vEnd)], chunk.
    Equal chunks keep the all-gather a single contiguous in-place collective; the colour buffers are padded for it."""
    chunk = -(-n // world)
    chunk = -(-chunk // align) * align
    return [(min(n, r * chunk), min(n, (r + 1) * chunk)) for r in range(world)], chunk


class _DevPtr:
    """zero-copy torch view of raw device memory (no ownership)"""

    def __init__(self, ptr, nelem, typestr):
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _view(ptr, nelem, typestr, device):
    import torch
    return torch.as_tensor(_DevPtr(ptr, nelem, typestr), device=device)


class GpuEngine:
    """The B200 engine of one rank: a partitioned mcmcb200 handle plus torch views of its exchange buffers."""

    def __init__(self, d_rowptr_local, d_neighs_local, nnz_local, n_global, v_begin, v_end, params, device_index):
        import torch
        self.device = f"cuda:{device_index}"
        flags = capi.FLAG_NO_FUSED_FINALIZE | capi.FLAG_NO_EARLY_STOP
        self.keep = (d_rowptr_local, d_neighs_local)
        self.chain = Chain(params=params, device=device_index, flags=flags, n_global=n_global, v_begin=v_begin, v_end=v_end,
                           device_csr=(d_rowptr_local.data_ptr(), d_neighs_local.data_ptr(), nnz_local))
        self.n, self.nCol = n_global, params.nCol
        self.stream = torch.cuda.ExternalStream(self.chain.stream(), device=self.device)
        self.t = 0
        self._bufs = None
        self.p2p = False

    def enable_p2p(self, rank, world, group=None):
        """Fused exchange: map every rank's colour replicas (CUDA IPC) so that the sweep kernel stores each finished tile's
        new colours straight into all of them over NVLink; the NCCL all-gather disappears, the counter all-reduce stays
        (it is also the inter-rank barrier).  Returns False (and keeps the all-gather) where it does not apply."""
        import torch.distributed as dist
        if world < 2 or world > 8:
            return False
        try:
            mine = self.chain.ipc_export()
        except Exception:
            mine = None
        table = [None] * world
        dist.all_gather_object(table, mine, group=group)
        if any(t is None for t in table):
            return False
        ok = True
        try:
            self.chain.ipc_attach(world, rank, b"".join(table))
        except Exception:
            ok = False
        flags = [None] * world
        dist.all_gather_object(flags, ok, group=group)
        self.p2p = all(flags)
        assert self.p2p or not ok, "fused exchange attached on some ranks only"
        return self.p2p

    def _views(self):
        if self._bufs is None:
            p0, nbytes, eb = self.chain.device_view(capi.VIEW_COLORS_CUR)
            p1, _, _ = self.chain.device_view(capi.VIEW_COLORS_NEXT)
            self.elem_bytes = eb                      # colours are exchanged as raw bytes (u8, or u16 little endian)
            self._bufs = [_view(p0, nbytes, "|u1", self.device), _view(p1, nbytes, "|u1", self.device)]
            pc, cbytes, _ = self.chain.device_view(capi.VIEW_COUNTERS)
            self._counters = _view(pc, cbytes // 8, "<i8", self.device)
        return self._bufs

    def init_colors(self, colors=None):
        self.chain.init_colors(colors)
        self.t = 0
        self._bufs = None
        self._views()

    def init_colors_slice(self, host_ptr, sweeper):
        """Host interface at N GPUs: every rank uploads only the colours it owns; the slices are all-gathered on the device."""
        self.chain.init_colors_slice_ptr(host_ptr)
        self.t = 0
        self._views()
        sweeper.gather_current()
        self.chain.init_colors_finish()

    def local_sweep(self):
        self.chain.sweep(1)

    def local_count(self):
        return self.chain.status()          # launches the local counting pass when the counters are stale

    def next_colors(self):
        return self._views()[(self.t + 1) & 1]

    def cur_colors(self):
        return self._views()[self.t & 1]

    def counters(self):
        self._views()
        return self._counters

    def finalize(self, advanced):
        self.chain.finalize_sweep()
        if advanced:
            self.t += 1

    def status(self):
        return self.chain.status()

    def colors_host(self, which="cur"):
        self.chain.synchronize()
        buf = self._views()[(self.t if which == "cur" else self.t + 1) & 1]
        raw = buf[: self.n * self.elem_bytes].cpu().numpy()
        return (raw if self.elem_bytes == 1 else raw.view("<u2")).astype(np.uint32)


class DistributedSweeper:
    """Sweep driver of one rank.  All collectives are issued on the engine's stream so they order with its kernels."""

    def __init__(self, engine, rank, world, chunk, group=None):
        self.e, self.rank, self.world, self.chunk, self.group = engine, rank, world, chunk, group

    def _on_stream(self):
        import contextlib
        import torch
        s = getattr(self.e, "stream", None)
        return torch.cuda.stream(s) if s is not None else contextlib.nullcontext()

    def _exchange(self, colours):
        import torch.distributed as dist
        with self._on_stream():
            if colours:
                cb = self.chunk * self.e.elem_bytes
                full = self.e.next_colors()[: cb * self.world]
                mine = full[self.rank * cb:(self.rank + 1) * cb]
                dist.all_gather_into_tensor(full, mine, group=self.group)            # owned slices of C_{t+1}
            dist.all_reduce(self.e.counters(), op=dist.ReduceOp.SUM, group=self.group)  # conflicts, violations, class deltas

    def gather_current(self):
        """all-gather the owned slices of the CURRENT colouring (sliced init)"""
        import torch.distributed as dist
        with self._on_stream():
            cb = self.chunk * self.e.elem_bytes
            full = self.e.cur_colors()[: cb * self.world]
            dist.all_gather_into_tensor(full, full[self.rank * cb:(self.rank + 1) * cb], group=self.group)

    def sweep(self, k=1):
        for _ in range(k):
            self.e.local_sweep()
            self._exchange(colours=not getattr(self.e, "p2p", False))   # fused exchange: the kernel already stored into the peers
            self.e.finalize(advanced=True)

    def status(self):
        """Global counters of the current colouring (runs a distributed counting pass if they are stale)."""
        st = self.e.status()
        if st.countsSweep != st.sweep:
            self._exchange(colours=False)
            self.e.finalize(advanced=False)
            st = self.e.status()
        return st


def slice_csr_torch(rowptr, neighs, vb, ve):
    """Owned rows of a device-resident CSR as a standalone (rebased, 32-byte aligned, padded) local CSR."""
    import torch
    e0, e1 = int(rowptr[vb].item()), int(rowptr[ve].item())
    rp = (rowptr[vb:ve + 1] - rowptr[vb]).to(torch.int32).contiguous()
    nb = torch.zeros(e1 - e0 + 16, dtype=torch.int32, device=neighs.device)
    nb[: e1 - e0] = neighs[e0:e1]
    return rp, nb, e1 - e0


def bench_main(args, WORKLOADS, GRAPH_SEED, CHAIN_SEED, measured_peak, ClockSampler):
    """bench.py under torchrun (N > 1): strong scaling of the same workload, vertex-partitioned over N GPUs."""
    import json
    import torch
    import torch.distributed as dist
    from .graphgen import er_graph_torch
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = f"cuda:{local_rank}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    n, deg, desc = WORKLOADS[args.workload]
    if args.n:
        n = args.n
    # every rank generates the same graph on its own GPU (deterministic), keeps its rows and frees the rest
    rowptr64, neighs, nnz, max_deg = er_graph_torch(n, deg, GRAPH_SEED, device=dev)
    parts, chunk = partition(n, world)
    vb, ve = parts[rank]
    rp, nb, nnz_local = slice_csr_torch(rowptr64, neighs, vb, ve)
    del rowptr64, neighs
    torch.cuda.empty_cache()
    nCol = max_deg
    proposal = capi.PROPOSAL_UNIFORM if args.proposal == "uniform" else capi.PROPOSAL_DYNAMIC
    prm = ColoringMCMCParams(nCol=nCol, proposal=proposal, seed=CHAIN_SEED,
                             convergence=capi.CONVERGE_VERTICES if proposal == capi.PROPOSAL_UNIFORM else capi.CONVERGE_EDGES)
    eng = GpuEngine(rp, nb, nnz_local, n, vb, ve, prm, local_rank)
    fused = eng.enable_p2p(rank, world) if os.environ.get("MCMCB200_NO_P2P") is None else False
    sw = DistributedSweeper(eng, rank, world, chunk)

    def timed_step():
        eng.init_colors(None)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(eng.stream):
            e0.record()
        sw.sweep(1)
        with torch.cuda.stream(eng.stream):
            e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)                                      # max over ranks
        return float(ms.item())

    for _ in range(max(args.warmup, 1)):
        timed_step()
    l0 = eng.chain.launch_count()
    with ClockSampler(local_rank) as clocks:
        steps_ms = [timed_step() for _ in range(args.steps)]
    launches = eng.chain.launch_count() - l0
    ms_per_step = float(np.mean(steps_ms))
    # end to end with host buffers: every rank uploads the colours of the vertices it owns from pinned memory (H2D), the slices
    # are exchanged on the device, then sweep + exchange, global counters (D2H) and the owned slice of the result (D2H)
    n_own = ve - vb
    pin_in = torch.empty(max(n_own, 1), dtype=torch.int32).pin_memory()
    pin_out = torch.empty(max(n_own, 1), dtype=torch.int32).pin_memory()
    eng.init_colors(None)
    eng.chain.get_colors_slice_ptr(pin_in.data_ptr())
    e2e = []
    for i in range(5):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        eng.init_colors_slice(pin_in.data_ptr(), sw)
        sw.sweep(1)
        st = sw.status()
        eng.chain.get_colors_slice_ptr(pin_out.data_ptr())
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if i >= 2:
            e2e.append(float(dt.item()))
    # ten chained sweeps, then the global counters (also checks the N-GPU trajectory against the 1-GPU invariants)
    eng.init_colors(None)
    sw.sweep(10)
    st = sw.status()
    if rank == 0:
        alg_bytes = 8 * nnz + 12 * n + 4
        peak, peak_src = measured_peak()
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
        value = n / (ms_per_step * 1e-3)
        line = {
            "metric": "vertex_updates_per_sec", "value": value, "unit": "vertex-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8 colours / u32 ids / f32 CDF", "data": "synthetic",
            "config": {"workload": desc, "n": n, "nnz_directed": nnz, "nCol": nCol, "proposal": args.proposal,
                       "parallelism": f"vertex partition x{world}; colour exchange: " + ("fused into the sweep kernel (peer stores over NVLink)" if fused else "NCCL all-gather of u8 slices") + "; counters: one NCCL all-reduce per sweep",
                       "step": "one sweep from the uniform random colouring incl. the colour exchange (max over ranks)",
                       "l2": "inputs larger than L2; no flush needed"},
            "edges_per_sec": value * nnz / n,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s",
                         "frac": achieved / (peak * world), "traffic": None, "peak_source": peak_src + f" x {world} GPUs",
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel": "per rank: " + eng.chain.kernel_mode() + " sweep + colour exchange + counter all-reduce"},
            "e2e": {"value": n / float(np.mean(e2e)), "unit": "vertex-updates/s", "h2d_bytes_per_step": 4 * n,
                    "d2h_bytes_per_step": 4 * n + 40 * world, "ms_per_step": 1e3 * float(np.mean(e2e)),
                    "note": "each rank moves only the colours of the vertices it owns"},
            "gpu_launches": int(launches), "clocks": clocks.summary(),
            "after_10_chain_sweeps": {"conflictEdges": int(st.conflictEdges), "violatingVertices": int(st.violatingVertices)},
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
    return 0
