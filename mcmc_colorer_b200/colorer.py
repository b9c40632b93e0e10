"""Host-side mirror of the reference's colourer interface over the C ABI.

  ColoringMCMCParams  <-  struct ColoringMCMCParams            graph_coloring/coloring.h:65-74 (defaults main.cu:160-168)
  Graph               <-  Graph<float,float> / GraphStruct     graph/graph.h:37-129 (CSR: cumulDegs, neighs; nEdges counts both directions)
  ColoringMCMC        <-  ColoringMCMC<float,float>            graph_coloring/coloringMCMC.h:43-140
                          (ctor(graph, randStates, params), setDirectoryPath, run(iteration), log + colours files
                           in the formats of coloringMCMC_prints.cu:27-230)
  Chain               the handle itself (mcmcb200_*), used by the parity tests and bench.py

The production C++ twin of this file is mcmc_colorer_b200/host/ (same ABI); this Python layer exists because the
tests, bench.py and the torch.distributed multi-GPU driver are Python.
"""
import ctypes as C
import math
import time
from dataclasses import dataclass

import numpy as np

from . import capi


@dataclass
class ColoringMCMCParams:
    """Field names and defaults of the reference struct (coloring.h:65-74, main.cu:160-168)."""
    maxRip: int = 250
    nCol: int = 0
    numColorRatio: float = 1.0
    lambda_: float = 1.0
    epsilon: float = 1e-8
    ratioFreezed: float = 1e-2
    tabooIteration: int = 0
    tailcut: bool = False
    # additions of the B200 build
    proposal: int = capi.PROPOSAL_DYNAMIC      # shipped GPU default: COLOR_BALANCE_DYNAMIC_DISTR (coloringMCMC.h:39)
    convergence: int = capi.CONVERGE_EDGES     # GPU loop test (coloringMCMC_main.cu:169)
    seed: int = 0


class Graph:
    """CSR graph with the reference's accessors (graph/graph.h:119-129, doStats graphCPU.cpp:432-450)."""

    def __init__(self, cumulDegs, neighs, prob=None):
        self.cumulDegs = np.ascontiguousarray(cumulDegs, np.uint32)
        self.neighs = np.ascontiguousarray(neighs, np.uint32)
        self.nNodes = len(self.cumulDegs) - 1
        self.nEdges = int(len(self.neighs))
        deg = np.diff(self.cumulDegs.astype(np.int64)) if self.nNodes else np.zeros(0, np.int64)
        self.maxDeg = int(deg.max()) if self.nNodes else 0
        self.minDeg = int(deg.min()) if self.nNodes else 0
        self.meanDeg = float(np.float32(self.nEdges) / np.float32(self.nNodes)) if self.nNodes else 0.0
        # main.cu:68: file graphs get prob = nEdges / n^2
        self.prob = float(prob) if prob is not None else (self.nEdges / float(self.nNodes * self.nNodes) if self.nNodes else 0.0)

    def getMaxNodeDeg(self):
        return self.maxDeg

    def getMinNodeDeg(self):
        return self.minDeg

    def getMeanNodeDeg(self):
        return self.meanDeg

    @staticmethod
    def default_ncol(maxDeg, numColRatio):
        """nCol = maxDeg * (1/ratio) with the reference's float arithmetic and truncation (main.cu:53,162)."""
        return int(np.float32(maxDeg) * (np.float32(1.0) / np.float32(numColRatio)))


class DeviceCsr:
    """CSR built on the GPU from an edge list (mcmcb200_csr_from_edges): the device twin of Graph::setupImporterNew
    (graphCPU.cpp:112-170) -- self-loops dropped, back-edges added, duplicates kept, rows in file order.
    src / dst: integer arrays on the host (numpy) or raw device pointers (ints) of m entries.
    Use as Chain(params=..., n_global=n, v_begin=0, v_end=n, device_csr=csr.as_tuple())."""

    def __init__(self, n, src, dst, m=None, device=0):
        self.L = capi.lib()
        self.n, self.device = int(n), device
        keep = []
        def ptr(x):
            if isinstance(x, int):
                return C.c_void_p(x)
            a = np.ascontiguousarray(x, np.uint32)
            keep.append(a)
            return C.c_void_p(a.ctypes.data)
        m = int(m if m is not None else len(src))
        self.rowptr, self.neighs, nnz = C.c_void_p(), C.c_void_p(), C.c_uint64()
        capi.check(self.L.mcmcb200_csr_from_edges(self.n, m, ptr(src), ptr(dst), device, C.byref(self.rowptr), C.byref(self.neighs),
                                                  C.byref(nnz)), "mcmcb200_csr_from_edges")
        self.nnz = int(nnz.value)

    def as_tuple(self):
        return (self.rowptr.value, self.neighs.value, self.nnz)

    def close(self):
        if self.rowptr is not None:
            self.L.mcmcb200_csr_free(self.rowptr, self.neighs)
            self.rowptr = self.neighs = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Chain:
    """One Markov chain on one GPU: a mcmcb200_handle."""

    def __init__(self, cumulDegs=None, neighs=None, params: ColoringMCMCParams = None, device=-1, flags=0,
                 n_global=None, v_begin=0, v_end=None, device_csr=None, stage_cap_bytes=0, item_bits=0, stage_buffers=0, expected_sweeps=0):
        """stage_cap_bytes / item_bits / stage_buffers: tuning of the source-blocked sweep (mcmcb200_params), 0 = automatic.
        expected_sweeps: mcmcb200_params.expectedSweeps (0 = many: build the blocked layout on large graphs)."""
        self.L = capi.lib()
        self.params = params
        p = capi.Params(nCol=params.nCol, epsilon=params.epsilon, lambda_=params.lambda_,
                        numColorRatio=params.numColorRatio, ratioFreezed=params.ratioFreezed,
                        tabooIteration=params.tabooIteration, maxRip=params.maxRip, tailcut=int(params.tailcut),
                        proposal=params.proposal, convergence=params.convergence, seed=params.seed, device=device,
                        flags=flags, stageCapBytes=stage_cap_bytes, itemBits=item_bits, stageBuffers=stage_buffers, expectedSweeps=expected_sweeps)
        self.h = C.c_void_p()
        if device_csr is not None:
            d_rowptr, d_neighs, nnz_local = device_csr
            self.n = int(n_global)
            self.v_begin, self.v_end = v_begin, self.n if v_end is None else v_end
            self._keep = device_csr
            capi.check(self.L.mcmcb200_create_device_csr(C.byref(self.h), self.n, self.v_begin, self.v_end, nnz_local,
                                                         C.c_void_p(d_rowptr), C.c_void_p(d_neighs), C.byref(p)),
                       "mcmcb200_create_device_csr")
        else:
            cumul, cp = capi._u32(cumulDegs)
            nb, nbp = capi._u32(neighs if len(neighs) else np.zeros(1, np.uint32))
            if n_global is None:
                self.n = len(cumul) - 1
                self.v_begin, self.v_end = 0, self.n
                capi.check(self.L.mcmcb200_create(C.byref(self.h), self.n, len(neighs), cp, nbp, C.byref(p)),
                           "mcmcb200_create")
            else:
                self.n = int(n_global)
                self.v_begin, self.v_end = v_begin, v_end
                capi.check(self.L.mcmcb200_create_partition(C.byref(self.h), self.n, v_begin, v_end, cp, nbp, C.byref(p)),
                           "mcmcb200_create_partition")
        self.nCol = params.nCol

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.mcmcb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init_colors(self, colors=None):
        if colors is None:
            capi.check(self.L.mcmcb200_init_colors(self.h, None), "mcmcb200_init_colors")
        else:
            a, ap = capi._u32(colors)
            assert len(a) == self.n
            capi.check(self.L.mcmcb200_init_colors(self.h, ap), "mcmcb200_init_colors")

    def init_colors_ptr(self, host_ptr):
        capi.check(self.L.mcmcb200_init_colors(self.h, C.c_void_p(host_ptr)), "mcmcb200_init_colors")

    def color_bytes(self):
        eb = C.c_uint32()
        capi.check(self.L.mcmcb200_color_bytes(self.h, C.byref(eb)), "mcmcb200_color_bytes")
        return eb.value

    def init_colors_narrow_ptr(self, host_ptr, elem_bytes):
        capi.check(self.L.mcmcb200_init_colors_narrow(self.h, C.c_void_p(host_ptr), elem_bytes), "mcmcb200_init_colors_narrow")

    def get_colors_narrow_ptr(self, host_ptr, elem_bytes):
        capi.check(self.L.mcmcb200_get_colors_narrow(self.h, C.c_void_p(host_ptr), elem_bytes), "mcmcb200_get_colors_narrow")

    def init_colors_narrow(self, colors):
        a = np.ascontiguousarray(colors, np.uint8 if self.color_bytes() == 1 else np.uint16)
        assert len(a) == self.n
        self.init_colors_narrow_ptr(a.ctypes.data, a.itemsize)

    def get_colors_narrow(self):
        out = np.empty(self.n, np.uint8 if self.color_bytes() == 1 else np.uint16)
        self.get_colors_narrow_ptr(out.ctypes.data, out.itemsize)
        return out

    def set_tape(self, u):
        if u is None:
            capi.check(self.L.mcmcb200_set_tape(self.h, None, 0), "mcmcb200_set_tape")
            return
        u = np.ascontiguousarray(u, np.float32).reshape(-1, self.n)
        capi.check(self.L.mcmcb200_set_tape(self.h, u.ctypes.data_as(C.c_void_p), u.shape[0]), "mcmcb200_set_tape")

    def sweep(self, k=1):
        capi.check(self.L.mcmcb200_sweep(self.h, k), "mcmcb200_sweep")

    def finalize_sweep(self):
        capi.check(self.L.mcmcb200_finalize_sweep(self.h), "mcmcb200_finalize_sweep")

    def status(self):
        st = capi.Status()
        capi.check(self.L.mcmcb200_status(self.h, C.byref(st)), "mcmcb200_status")
        return st

    def get_colors(self, out=None):
        if out is None:
            out = np.empty(self.n, np.uint32)
        capi.check(self.L.mcmcb200_get_colors(self.h, out.ctypes.data_as(C.c_void_p)), "mcmcb200_get_colors")
        return out

    def get_colors_ptr(self, host_ptr):
        capi.check(self.L.mcmcb200_get_colors(self.h, C.c_void_p(host_ptr)), "mcmcb200_get_colors")

    def class_sizes(self):
        out = np.empty(self.nCol, np.uint64)
        capi.check(self.L.mcmcb200_get_class_sizes(self.h, out.ctypes.data_as(C.c_void_p)), "mcmcb200_get_class_sizes")
        return out

    def history(self, cap=4096):
        out = np.zeros(2 * cap, np.uint64)
        cnt = C.c_uint32()
        capi.check(self.L.mcmcb200_get_history(self.h, out.ctypes.data_as(C.c_void_p), cap, C.byref(cnt)),
                   "mcmcb200_get_history")
        return out[:2 * cnt.value].reshape(-1, 2)

    def tailcut(self, max_rounds=64):
        r = C.c_uint32()
        capi.check(self.L.mcmcb200_tailcut(self.h, max_rounds, C.byref(r)), "mcmcb200_tailcut")
        return r.value

    # ---- distributed tail cutting (mcmcb200_tailcut_dist_*): one rank's part; multigpu.DistributedSweeper.tailcut drives it ----
    def tc_begin(self, order, cap=None):
        cap = max(1, min(self.n, 1 << 26)) if cap is None else cap
        o = np.ascontiguousarray(order, np.uint32)
        out = np.empty(cap, np.uint32)
        cnt = C.c_uint32()
        capi.check(self.L.mcmcb200_tailcut_dist_begin(self.h, o.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), cap, C.byref(cnt)),
                   "mcmcb200_tailcut_dist_begin")
        return out[:cnt.value].copy()

    def tc_mark(self, ids):
        a = np.ascontiguousarray(ids, np.uint32)
        capi.check(self.L.mcmcb200_tailcut_dist_mark(self.h, a.ctypes.data_as(C.c_void_p), len(a)), "mcmcb200_tailcut_dist_mark")

    def tc_round(self, cap=None):
        cap = max(1, min(self.n, 1 << 26)) if cap is None else cap
        ids, cols = np.empty(cap, np.uint32), np.empty(cap, np.uint32)
        done, left, inexact = C.c_uint32(), C.c_uint32(), C.c_uint32()
        capi.check(self.L.mcmcb200_tailcut_dist_round(self.h, ids.ctypes.data_as(C.c_void_p), cols.ctypes.data_as(C.c_void_p), cap,
                                                      C.byref(done), C.byref(left), C.byref(inexact)), "mcmcb200_tailcut_dist_round")
        return ids[:done.value].copy(), cols[:done.value].copy(), left.value, bool(inexact.value)

    def tc_apply(self, ids, cols):
        a, b = np.ascontiguousarray(ids, np.uint32), np.ascontiguousarray(cols, np.uint32)
        capi.check(self.L.mcmcb200_tailcut_dist_apply(self.h, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), len(a)), "mcmcb200_tailcut_dist_apply")

    def tc_recount(self):
        d, v, nf = C.c_uint64(), C.c_uint64(), C.c_uint32()
        capi.check(self.L.mcmcb200_tailcut_dist_recount(self.h, C.byref(d), C.byref(v), C.byref(nf)), "mcmcb200_tailcut_dist_recount")
        return d.value, v.value, nf.value

    def tc_end(self, directed, viol, exact=True):
        capi.check(self.L.mcmcb200_tailcut_dist_end(self.h, int(directed), int(viol), 1 if exact else 0), "mcmcb200_tailcut_dist_end")

    def conflicts_of(self, colors):
        a, ap = capi._u32(colors)
        e, v = C.c_uint64(), C.c_uint64()
        capi.check(self.L.mcmcb200_conflicts_of(self.h, ap, C.byref(e), C.byref(v)), "mcmcb200_conflicts_of")
        return e.value, v.value

    def debug_occupancy(self, v):
        words = (self.nCol + 31) // 32
        out = np.zeros(words, np.uint32)
        capi.check(self.L.mcmcb200_debug_occupancy(self.h, v, out.ctypes.data_as(C.c_void_p)), "mcmcb200_debug_occupancy")
        return out

    def debug_all_occupancy(self):
        nloc = self.v_end - self.v_begin
        w64 = (self.nCol + 63) // 64
        w64 = 1 if w64 <= 1 else 2 if w64 <= 2 else 4 if w64 <= 4 else 8 if w64 <= 8 else w64
        masks = np.zeros((max(nloc, 1), w64), np.uint64)
        same = np.zeros(max(nloc, 1), np.uint32)
        capi.check(self.L.mcmcb200_debug_all_occupancy(self.h, masks.ctypes.data_as(C.c_void_p),
                                                       same.ctypes.data_as(C.c_void_p)), "mcmcb200_debug_all_occupancy")
        return masks[:nloc], same[:nloc]

    def device_view(self, which):
        ptr, nbytes, eb = C.c_void_p(), C.c_uint64(), C.c_uint32()
        capi.check(self.L.mcmcb200_device_view(self.h, which, C.byref(ptr), C.byref(nbytes), C.byref(eb)),
                   "mcmcb200_device_view")
        return ptr.value, nbytes.value, eb.value

    def init_colors_slice_narrow_ptr(self, host_ptr, elem_bytes):
        capi.check(self.L.mcmcb200_init_colors_slice_narrow(self.h, C.c_void_p(host_ptr), elem_bytes), "mcmcb200_init_colors_slice_narrow")

    def get_colors_slice_narrow_ptr(self, host_ptr, elem_bytes):
        capi.check(self.L.mcmcb200_get_colors_slice_narrow(self.h, C.c_void_p(host_ptr), elem_bytes), "mcmcb200_get_colors_slice_narrow")

    def init_colors_slice_ptr(self, host_ptr):
        capi.check(self.L.mcmcb200_init_colors_slice(self.h, C.c_void_p(host_ptr)), "mcmcb200_init_colors_slice")

    def init_colors_finish(self):
        capi.check(self.L.mcmcb200_init_colors_finish(self.h), "mcmcb200_init_colors_finish")

    def get_colors_slice_ptr(self, host_ptr):
        capi.check(self.L.mcmcb200_get_colors_slice(self.h, C.c_void_p(host_ptr)), "mcmcb200_get_colors_slice")

    def ipc_detach(self):
        capi.check(self.L.mcmcb200_ipc_detach(self.h), "mcmcb200_ipc_detach")

    def ipc_export(self):
        buf = (C.c_ubyte * 192)()
        capi.check(self.L.mcmcb200_ipc_export(self.h, C.cast(buf, C.c_void_p)), "mcmcb200_ipc_export")
        return bytes(buf)

    def ipc_attach(self, n_ranks, my_rank, handles: bytes):
        buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        capi.check(self.L.mcmcb200_ipc_attach(self.h, n_ranks, my_rank, C.cast(buf, C.c_void_p)), "mcmcb200_ipc_attach")

    def stream(self):
        s = C.c_void_p()
        capi.check(self.L.mcmcb200_stream(self.h, C.byref(s)), "mcmcb200_stream")
        return s.value or 0

    def synchronize(self):
        capi.check(self.L.mcmcb200_synchronize(self.h), "mcmcb200_synchronize")

    def last_sweep_ms(self):
        ms = C.c_float()
        capi.check(self.L.mcmcb200_last_sweep_ms(self.h, C.byref(ms)), "mcmcb200_last_sweep_ms")
        return ms.value

    def layout_bytes(self):
        b = C.c_uint64()
        capi.check(self.L.mcmcb200_layout_bytes(self.h, C.byref(b)), "mcmcb200_layout_bytes")
        return b.value

    KERNEL_MODES = ("direct", "blocked", "blocked-overlapped", "direct-binned", "wide-binned")

    def kernel_mode(self):
        """which sweep implementation the handle runs: 'direct', 'blocked' or 'blocked-overlapped' (mcmcb200_kernel_mode)"""
        m = C.c_int(0)
        capi.check(self.L.mcmcb200_kernel_mode(self.h, C.byref(m)), "mcmcb200_kernel_mode")
        return self.KERNEL_MODES[m.value]

    def launch_count(self):
        k = C.c_uint64()
        capi.check(self.L.mcmcb200_launch_count(self.h, C.byref(k)), "mcmcb200_launch_count")
        return k.value


def luby_color(cumulDegs, neighs, seed=0, device=-1):
    """ColoringLuby cross-check (graph_coloring/coloringLuby.cu:364-501): returns (colors uint32[n], 1-based; nCol; rounds)."""
    L = capi.lib()
    cumul, cp = capi._u32(cumulDegs)
    nb, nbp = capi._u32(neighs if len(neighs) else np.zeros(1, np.uint32))
    n = len(cumul) - 1
    out = np.zeros(n, np.uint32)
    ncol, rounds = C.c_uint32(), C.c_uint32()
    capi.check(L.mcmcb200_luby_color(n, len(neighs), cp, nbp, seed, device, out.ctypes.data_as(C.c_void_p), C.byref(ncol),
                                     C.byref(rounds)), "mcmcb200_luby_color")
    return out, ncol.value, rounds.value


def occupancy_bits(mask_words32, nCol):
    """uint32 mask words -> uint8[nCol] occupancy row (1 = a neighbour has that colour)."""
    bits = np.unpackbits(np.asarray(mask_words32, np.uint32).view(np.uint8), bitorder="little")
    return bits[:nCol].astype(np.uint8)


def color_stats(hist, n, prob):
    """Final statistics block of getStatsNumColors (coloringMCMC_prints.cu:140-174), float32 arithmetic in the
    reference's order."""
    f32 = np.float32
    nCol = len(hist)
    average = f32(n) / f32(nCol)
    counter, max_i, min_i, max_c, min_c = 0, 0, n, 0, n
    bal = f32(0)
    for i, hcount in enumerate(hist):
        hcount = int(hcount)
        if hcount > 0:
            counter += 1
            if hcount > max_c:
                max_i, max_c = i, hcount
            if hcount < min_c:
                min_i, min_c = i, hcount
            d = f32(f32(hcount) - average)
            bal = f32(bal + f32(d * d))
    bal = f32(bal / f32(f32(n) * f32(prob))) if prob > 0 else f32(0)
    bal = f32(math.sqrt(bal))
    var = f32(0)
    for hcount in hist:
        d = f32(f32(int(hcount)) - average)
        var = f32(var + f32(d * d))
    var = f32(var / f32(nCol))
    return dict(used=counter, most=(max_i, max_c), least=(min_i, min_c), average=float(average), variance=float(var),
                std=float(f32(math.sqrt(var))), balancingIndex=float(bal))


def _g(x):
    """ostream << float formatting (6 significant digits, %g)."""
    return "%g" % x


class ColoringMCMC:
    """ColoringMCMC<float,float> (graph_coloring/coloringMCMC.h:43-140) over libmcmcb200.

    `randStates` is accepted for signature compatibility and ignored: the RNG is stateless Philox keyed by
    params.seed (include/mcmcb200.h, RNG contract)."""

    def __init__(self, graph: Graph, randStates, params: ColoringMCMCParams, device=-1, sweeps_per_check=1):
        self.graph = graph
        self.param = params
        self.directory = None
        self.rip = 0
        self.maxIterReached = False
        self.duration = 0.0
        self.sweeps_per_check = sweeps_per_check
        self.chain = Chain(graph.cumulDegs, graph.neighs, params, device=device)

    def setDirectoryPath(self, directory):
        self.directory = directory

    def run(self, iteration=0):
        """coloringMCMC_main.cu:100-298: init, sweep until the conflict count is <= z or maxRip, optional tail cut,
        final statistics + colours file."""
        p, ch = self.param, self.chain
        log = open(self.directory + ".log", "w") if self.directory else None
        colf = open(self.directory + "-colors.txt", "w") if self.directory else None

        def w(s):
            if log:
                log.write(s + "\n")

        # __customPrintRun0_start (coloringMCMC_prints.cu:37-48)
        w("numCol: %d" % p.nCol)
        w("epsilon: " + _g(p.epsilon))
        w("lambda: " + _g(p.lambda_))
        w("ratioFreezed: " + _g(p.ratioFreezed))
        w("maxRip: %d" % p.maxRip)
        w("")
        w("numColorRatio: " + _g(p.numColorRatio))
        ch.init_colors(None)
        t0 = time.perf_counter()
        self.rip = 0
        st = ch.status()
        use_edges = p.convergence == capi.CONVERGE_EDGES
        while True:                                           # do { rip++; ... } while (rip < maxRip), _main.cu:160-269
            self.rip += 1
            count = st.conflictEdges if use_edges else st.violatingVertices
            if st.converged:
                break
            w("***** Tentativo numero: %d" % self.rip)         # __customPrintRun2_conflicts
            w("conflitti rilevati: %d" % count)
            ch.sweep(self.sweeps_per_check)
            st = ch.status()
            w("nuovi conflitti rilevati: %d" % (st.conflictEdges if use_edges else st.violatingVertices))
            if self.rip >= p.maxRip:
                break
        if p.tailcut and (st.conflictEdges > 0):              # _main.cu:271-290
            w("***** Tentativo numero: %d" % self.rip)
            w("---> TailCutting")
            w("conflitti rilevati: %d" % st.conflictEdges)
            ch.tailcut()
            st = ch.status()
            w("nuovi conflitti rilevati: %d" % st.conflictEdges)
        self.duration = time.perf_counter() - t0
        self.maxIterReached = self.rip == p.maxRip            # _main.cu:294-295
        self.status = st
        # __customPrintRun7_end + getStatsNumColors("end_") (coloringMCMC_prints.cu:96-230)
        hist = ch.class_sizes()
        colors = ch.get_colors()
        s = color_stats(hist, self.graph.nNodes, self.graph.prob)
        self.stats = s
        w("COLORAZIONE FINALE")
        w("Time " + _g(self.duration))
        w("Max iteration reached " + ("yes" if self.rip >= p.maxRip else "no"))
        if colf:
            colf.write("".join("%d %d\n" % (i, c) for i, c in enumerate(colors)))
        w("Number of used colors is %d on %d available" % (s["used"], p.nCol))
        w("Most used colors is %d used %d times" % s["most"])
        w("Least used colors is %d used %d times" % s["least"])
        w("")
        w("Average " + _g(s["average"]))
        w("Variance " + _g(s["variance"]))
        w("StandardDeviation " + _g(s["std"]))
        w("BalancingIndex " + _g(s["balancingIndex"]))
        w("")
        w("")
        w("end colorazione finale -------------------------------------------------------------------")
        w("")
        if log:
            log.close()
        if colf:
            colf.close()
        return colors
