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


def partition_by_nnz(rowptr, world, align=256):
    """Contiguous vertex ranges balanced by the number of directed edges (prefix of cumulDegs): cut r is the first
    vertex whose offset reaches r/world of nnz, rounded to a multiple of `align` (the kernels move owned colours as 16-byte
    vectors / bulk copies; mcmcb200_create_partition insists on 256).  rowptr: numpy array or torch tensor of n+1 offsets.
    Returns [(vBegin, vEnd)] -- ranges may be empty at the end of a very skewed graph."""
    n = len(rowptr) - 1
    nnz = int(rowptr[n])
    cuts = [0]
    for r in range(1, world):
        target = nnz * r // world
        if hasattr(rowptr, "cpu"):                       # torch tensor (device or host)
            import torch
            v = int(torch.searchsorted(rowptr, torch.tensor([target], dtype=rowptr.dtype, device=rowptr.device)).item())
        else:
            v = int(np.searchsorted(rowptr, target))
        v = min(n, (v + align // 2) // align * align)
        cuts.append(max(v, cuts[-1]))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


class _DevPtr:
    """zero-copy torch view of raw device memory (no ownership)"""

    def __init__(self, ptr, nelem, typestr):
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _view(ptr, nelem, typestr, device):
    import torch
    return torch.as_tensor(_DevPtr(ptr, nelem, typestr), device=device)


class GpuEngine:
    """The B200 engine of one rank: a partitioned mcmcb200 handle plus torch views of its exchange buffers."""

    def __init__(self, d_rowptr_local, d_neighs_local, nnz_local, n_global, v_begin, v_end, params, device_index, extra_flags=0, early_stop=False, **tuning):
        """early_stop: let the chain stop on the device at its threshold (the --tailcut protocol: params.tailcut, then DistributedSweeper.tailcut);
        by default the sweeps keep advancing (benchmarks, trajectory checks)."""
        import torch
        self.device = f"cuda:{device_index}"
        flags = capi.FLAG_NO_FUSED_FINALIZE | (0 if early_stop else capi.FLAG_NO_EARLY_STOP) | extra_flags
        self.keep = (d_rowptr_local, d_neighs_local)
        self.chain = Chain(params=params, device=device_index, flags=flags, n_global=n_global, v_begin=v_begin, v_end=v_end,
                           device_csr=(d_rowptr_local.data_ptr(), d_neighs_local.data_ptr(), nnz_local), **tuning)
        self.n, self.nCol = n_global, params.nCol
        self.stream = torch.cuda.ExternalStream(self.chain.stream(), device=self.device)
        self.t = 0
        self._bufs = None
        self.p2p = False

    def enable_p2p(self, rank, world, group=None):
        """Fused exchange: map every rank's colour replicas and counter-exchange block (CUDA IPC) so that the sweep kernel stores
        each finished tile's new colours straight into all replicas over NVLink and all-reduces the counters itself -- no NCCL
        call and no host in the sweep loop.  Returns False (and keeps all-gather + all-reduce) where it does not apply.
        Protocol: (1) every rank says whether it is eligible (source-blocked sweep, export works) and NOBODY attaches unless all
        are; (2) attach; (3) if any rank failed to map a peer, everybody detaches again."""
        import torch.distributed as dist
        if world < 2 or world > 8:
            return False
        mine = None
        if self.chain.kernel_mode() in ("blocked", "blocked-overlapped"):
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
        if not self.p2p and ok:
            self.chain.ipc_detach()
        return self.p2p

    def _views(self):
        if self._bufs is None:
            p0, nbytes, eb = self.chain.device_view(capi.VIEW_COLORS_CUR)
            p1, _, _ = self.chain.device_view(capi.VIEW_COLORS_NEXT)
            self.elem_bytes = eb                      # colours are exchanged as raw bytes (u8, or u16 little endian)
            self._bufs = [_view(p0, nbytes, "|u1", self.device), _view(p1, nbytes, "|u1", self.device)]
            self._ptrs = (p0, p1) if (self.t & 1) == 0 else (p1, p0)        # physical buffers 0 / 1 (colouring t lives in buffer t & 1)
            if self.t & 1:
                self._bufs.reverse()
            pc, cbytes, _ = self.chain.device_view(capi.VIEW_COUNTERS)
            self._counters = _view(pc, cbytes // 8, "<i8", self.device)
        return self._bufs

    def init_colors(self, colors=None):
        self.chain.init_colors(colors)
        self.t = 0
        self._bufs = None
        self._views()
        self._fence_ranks()

    def _fence_ranks(self):
        """fused exchange: (re)initialisation clears this rank's exchange block; no peer may start the next collective (and add
        into that block) before every rank is through it"""
        if self.p2p:
            import torch.distributed as dist
            self.chain.synchronize()
            dist.barrier()

    def init_colors_slice(self, host_ptr, sweeper, elem_bytes=4):
        """Host interface at N GPUs: every rank uploads only the colours it owns (uint32, or the device's narrow format when elem_bytes
        is 1 / 2); the slices are all-gathered on the device."""
        if elem_bytes == 4:
            self.chain.init_colors_slice_ptr(host_ptr)
        else:
            self.chain.init_colors_slice_narrow_ptr(host_ptr, elem_bytes)
        self.t = 0
        self._views()
        sweeper.gather_current()
        self.chain.init_colors_finish()
        self._fence_ranks()

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

    # distributed tail cutting: this rank's part (Chain.tc_*  ->  mcmcb200_tailcut_dist_*)
    def class_sizes(self):
        return self.chain.class_sizes()

    def tc_begin(self, order):
        return self.chain.tc_begin(order)

    def tc_mark(self, ids):
        self.chain.tc_mark(ids)

    def tc_round(self):
        return self.chain.tc_round()

    def tc_apply(self, ids, cols):
        self.chain.tc_apply(ids, cols)

    def tc_recount(self):
        return self.chain.tc_recount()

    def tc_end(self, directed, viol, exact):
        self.chain.tc_end(directed, viol, exact)

    def sync_t(self):
        """A chain created without FLAG_NO_EARLY_STOP stops advancing on the device once it is at its threshold, while this object
        counts every sweep it launched: re-align the buffer parity with the device (mcmcb200_device_view only reads the state;
        mcmcb200_status would start a counting pass, which is a collective in this driver)."""
        self._views()
        cur, _, _ = self.chain.device_view(capi.VIEW_COLORS_CUR)
        par = 0 if cur == self._ptrs[0] else 1
        if (self.t & 1) != par:
            self.t -= 1

    def colors_host(self, which="cur"):
        self.chain.synchronize()
        self.sync_t()
        buf = self._views()[(self.t if which == "cur" else self.t + 1) & 1]
        raw = buf[: self.n * self.elem_bytes].cpu().numpy()
        return (raw if self.elem_bytes == 1 else raw.view("<u2")).astype(np.uint32)


class DistributedSweeper:
    """Sweep driver of one rank.  All collectives are issued on the engine's stream so they order with its kernels.
    parts: [(vBegin, vEnd)] of every rank (equal chunks -> one in-place all-gather; nnz-balanced ranges -> one broadcast per rank)."""

    def __init__(self, engine, rank, world, chunk, group=None, parts=None):
        self.e, self.rank, self.world, self.chunk, self.group = engine, rank, world, chunk, group
        self.parts = parts
        self.uniform = parts is None or all(vb == r * chunk for r, (vb, ve) in enumerate(parts))

    def _on_stream(self):
        import contextlib
        import torch
        s = getattr(self.e, "stream", None)
        return torch.cuda.stream(s) if s is not None else contextlib.nullcontext()

    def _gather(self, buf):
        import torch.distributed as dist
        eb = self.e.elem_bytes
        if self.uniform:
            cb = self.chunk * eb
            full = buf[: cb * self.world]
            dist.all_gather_into_tensor(full, full[self.rank * cb:(self.rank + 1) * cb], group=self.group)
        else:
            for r, (vb, ve) in enumerate(self.parts):
                if ve > vb:
                    dist.broadcast(buf[vb * eb: ve * eb], src=r, group=self.group)

    def _exchange(self, colours):
        import torch.distributed as dist
        with self._on_stream():
            if colours:
                self._gather(self.e.next_colors())                                   # owned slices of C_{t+1}
            dist.all_reduce(self.e.counters(), op=dist.ReduceOp.SUM, group=self.group)  # conflicts, violations, class deltas

    def gather_current(self):
        """all-gather the owned slices of the CURRENT colouring (sliced init)"""
        with self._on_stream():
            self._gather(self.e.cur_colors())

    def sweep(self, k=1):
        if getattr(self.e, "p2p", False):
            # fused exchange: colours, counters and the inter-rank barrier all happen inside the sweep kernels -- k sweeps are
            # 2k launches per rank, nothing else
            self.e.chain.sweep(k)
            self.e.t += k
            return
        for _ in range(k):
            self.e.local_sweep()
            self._exchange(colours=True)
            self.e.finalize(advanced=True)

    def tailcut(self, max_passes=64):
        """Tail cutting of the N-GPU chain (the reference's --tailcut repair, coloringMCMC_main.cu:271-290, is single-GPU): the same
        sequential-greedy result, computed by the ranks together.  Per pass: every rank flags the violating vertices it owns that
        the reference would visit; the flagged ids are exchanged (they are few: the chain stopped at <= z = max(50, n/2000)
        violators); then rounds -- a flagged vertex is ready when no flagged neighbour with a smaller id, on any rank, is still
        pending; ready vertices are pairwise non-adjacent, so all ranks repair theirs at once and tell each other the (vertex,
        colour) pairs.  Returns the number of passes.  All small exchanges go through all_gather_object (host), nothing is large."""
        import torch.distributed as dist
        if hasattr(self.e, "sync_t"):
            self.e.sync_t()
        # colours in ascending class size, computed ONCE before the passes (coloringMCMC_main.cu:272-277); ties by colour index --
        # identical on every rank (the class sizes are global)
        order = np.argsort(np.asarray(self.e.class_sizes(), dtype=np.int64), kind="stable").astype(np.uint32)
        passes = 0
        while passes < max_passes:
            mine = np.asarray(self.e.tc_begin(order), dtype=np.uint32)
            table = [None] * self.world
            dist.all_gather_object(table, mine, group=self.group)
            total = int(sum(len(t) for t in table))
            if total == 0:
                d, v, _ = self.e.tc_recount()
                tot = [None] * self.world
                dist.all_gather_object(tot, (d, v), group=self.group)
                self.e.tc_end(sum(t[0] for t in tot), sum(t[1] for t in tot), True)
                return passes
            passes += 1
            self.e.tc_mark(np.concatenate(table))
            inexact = False
            while True:
                ids, cols, left, inx = self.e.tc_round()
                got = [None] * self.world
                dist.all_gather_object(got, (ids, cols, left, inx), group=self.group)
                for r, (ri, rc, rl, rx) in enumerate(got):
                    inexact = inexact or rx
                    if r != self.rank and len(ri):
                        self.e.tc_apply(ri, rc)
                if sum(g[2] for g in got) == 0:
                    break
            d, v, nf = self.e.tc_recount()
            tot = [None] * self.world
            dist.all_gather_object(tot, (d, v, nf), group=self.group)
            gd, gv, gnf = (sum(t[i] for t in tot) for i in range(3))
            # stop: nothing left to flag, a repaired vertex found every colour taken (vertices outside the lists may violate now:
            # full recount at the next status), or no progress is possible
            if inexact or gnf == 0 or gnf >= total:
                self.e.tc_end(gd, gv, not inexact)
                return passes
        d, v, _ = self.e.tc_recount()
        tot = [None] * self.world
        dist.all_gather_object(tot, (d, v), group=self.group)
        self.e.tc_end(sum(t[0] for t in tot), sum(t[1] for t in tot), True)
        return passes

    def status(self):
        """Global counters of the current colouring (runs a distributed counting pass if they are stale)."""
        st = self.e.status()
        if getattr(self.e, "p2p", False):
            return st                                   # the counting pass (if one was needed) reduced across the ranks on the device
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


def bench_main(args, WORKLOADS, GRAPH_SEED, CHAIN_SEED, measured_peak, ClockSampler, palette_for, RMAT):
    """bench.py under torchrun (N > 1): strong scaling of the same workload, vertex-partitioned over N GPUs."""
    import json
    import torch
    import torch.distributed as dist
    from .graphgen import er_graph_torch, rmat_graph_torch
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = f"cuda:{local_rank}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    n, deg, desc = WORKLOADS[args.workload]
    if args.n:
        n = args.n
    # every rank generates the same graph on its own GPU (deterministic), keeps its rows and frees the rest
    skewed = args.workload.startswith("c4")
    if skewed:
        ra, rb, rc = RMAT[args.workload]
        rowptr64, neighs, nnz, max_deg = rmat_graph_torch(max(1, (n - 1).bit_length()), deg, GRAPH_SEED, n_keep=n, a=ra, b=rb, c=rc, device=dev)
    else:
        rowptr64, neighs, nnz, max_deg = er_graph_torch(n, deg, GRAPH_SEED, device=dev)
    # contiguous vertex ranges: equal vertex counts on Erdos-Renyi (they are nnz-balanced to a fraction of a percent and keep the
    # NCCL fallback a single in-place all-gather), balanced by directed edges on skewed graphs
    parts, chunk = partition(n, world)
    if skewed:
        parts = partition_by_nnz(rowptr64, world)
    vb, ve = parts[rank]
    rp, nb, nnz_local = slice_csr_torch(rowptr64, neighs, vb, ve)
    del rowptr64, neighs
    torch.cuda.empty_cache()
    nCol = getattr(args, "ncol", 0) or palette_for(args.workload, max_deg)      # (bench.py: R-MAT hubs have degree ~1e6)
    proposal = capi.PROPOSAL_UNIFORM if args.proposal == "uniform" else capi.PROPOSAL_DYNAMIC
    prm = ColoringMCMCParams(nCol=nCol, proposal=proposal, seed=CHAIN_SEED,
                             convergence=capi.CONVERGE_VERTICES if proposal == capi.PROPOSAL_UNIFORM else capi.CONVERGE_EDGES)
    eng = GpuEngine(rp, nb, nnz_local, n, vb, ve, prm, local_rank)
    fused = eng.enable_p2p(rank, world) if os.environ.get("MCMCB200_NO_P2P") is None else False
    sw = DistributedSweeper(eng, rank, world, chunk, parts=parts)

    sync_word = torch.zeros(1, device=dev)

    def timed_step(k=1):
        """k sweeps timed on the device, max over ranks.  Host barrier + synchronize on both sides; in addition the ranks meet ON THE
        DEVICE right before the start event (a one-word all-reduce on the sweep stream), so that the start skew of the host processes
        after dist.barrier() (tens of microseconds, against a 0.5 ms step at 8 GPUs) is not charged to the sweep."""
        eng.init_colors(None)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(eng.stream):
            dist.all_reduce(sync_word)
            e0.record()
        sw.sweep(k)
        with torch.cuda.stream(eng.stream):
            e1.record()
        torch.cuda.synchronize()
        dist.barrier()
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
    chain_ms = timed_step(10) / 10.0                      # ten consecutive sweeps in one timed region (steady state of a chain)
    # end to end with host buffers: every rank uploads the colours of the vertices it owns from pinned memory (H2D), the slices
    # are exchanged on the device, then sweep + exchange, global counters (D2H) and the owned slice of the result (D2H)
    # (the device's narrow colour format, mcmcb200_{init,get}_colors_slice_narrow; the reference-shaped uint32 slices are timed beside it)
    n_own = ve - vb
    eb = eng.chain.color_bytes()

    def e2e_loop(elem_bytes):
        dt_t = torch.uint8 if elem_bytes == 1 else torch.int16 if elem_bytes == 2 else torch.int32
        pin_in = torch.empty(max(n_own, 1), dtype=dt_t).pin_memory()
        pin_out = torch.empty(max(n_own, 1), dtype=dt_t).pin_memory()
        eng.init_colors(None)
        if elem_bytes == 4:
            eng.chain.get_colors_slice_ptr(pin_in.data_ptr())
        else:
            eng.chain.get_colors_slice_narrow_ptr(pin_in.data_ptr(), elem_bytes)
        out = []
        for i in range(5):
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            eng.init_colors_slice(pin_in.data_ptr(), sw, elem_bytes)
            sw.sweep(1)
            sw.status()
            if elem_bytes == 4:
                eng.chain.get_colors_slice_ptr(pin_out.data_ptr())
            else:
                eng.chain.get_colors_slice_narrow_ptr(pin_out.data_ptr(), elem_bytes)
            dt = torch.tensor([time.perf_counter() - t0], device=dev)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if i >= 2:
                out.append(float(dt.item()))
        return out

    e2e = e2e_loop(eb)
    e2e32 = e2e_loop(4)
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
                       "parallelism": f"vertex partition x{world} ({'nnz-balanced' if skewed else 'equal vertex counts'}); " + ("colour exchange AND counter all-reduce + inter-rank barrier fused into the sweep kernels (peer stores / system-scope reductions over NVLink): no NCCL call, no host in the sweep loop" if fused else "NCCL all-gather of the narrow colour slices + one NCCL all-reduce of the counters per sweep"),
                       "step": "one sweep from the uniform random colouring incl. the colour exchange (max over ranks)",
                       "l2": "inputs larger than L2; no flush needed"},
            "edges_per_sec": value * nnz / n, "chain_ms_per_sweep": chain_ms,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s",
                         "frac": achieved / (peak * world), "traffic": None, "peak_source": peak_src + f" x {world} GPUs",
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel": "per rank: " + eng.chain.kernel_mode() + " sweep + colour exchange + counter all-reduce"},
            "e2e": {"value": n / float(np.mean(e2e)), "unit": "vertex-updates/s", "h2d_bytes_per_step": eb * n,
                    "d2h_bytes_per_step": eb * n + 40 * world, "ms_per_step": 1e3 * float(np.mean(e2e)),
                    "note": "each rank moves only the colours of the vertices it owns, in the device's narrow format (u%d)" % (8 * eb),
                    "u32_interface": {"value": n / float(np.mean(e2e32)), "ms_per_step": 1e3 * float(np.mean(e2e32)),
                                      "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": 4 * n + 40 * world}},
            "gpu_launches": int(launches), "clocks": clocks.summary(),
            "after_10_chain_sweeps": {"conflictEdges": int(st.conflictEdges), "violatingVertices": int(st.violatingVertices)},
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
    return 0
