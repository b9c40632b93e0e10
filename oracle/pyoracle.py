"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the two CPU checkers.

  Port : oracle/liboracle_port.so   (plain-C restatement, oracle/mcmc_oracle.c)
  Ref  : oracle/_ref/libmcmc_ref.so (UNMODIFIED reference CPU colourer + oracle/ref_harness.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module, and only as the checker.  Nothing under mcmc_colorer_b200/ may import it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle_port.so")
REF_SO = os.path.join(HERE, "_ref", "libmcmc_ref.so")

UNIFORM, DYNAMIC = 0, 1

_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build(which=("port", "ref")):
    """Compile the checkers (oracle/Makefile).  `ref` is a no-op where /root/reference is absent."""
    for w in which:
        subprocess.run(["make", "-s", "-C", HERE, w], check=True)


class ColorStats(C.Structure):
    _fields_ = [("usedColors", C.c_uint32), ("mostUsed", C.c_uint32), ("mostUsedCount", C.c_uint32),
                ("leastUsed", C.c_uint32), ("leastUsedCount", C.c_uint32),
                ("meanCPU", C.c_float), ("varianceCPU", C.c_float), ("stdCPU", C.c_float),
                ("averageGPU", C.c_float), ("varianceGPU", C.c_float), ("stdGPU", C.c_float),
                ("balancingIndex", C.c_float)]


class Port:
    """oracle/mcmc_oracle.h"""

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build(("port",))
        L = self.L = C.CDLL(PORT_SO)
        L.orc_philox4x32_10.argtypes = [_u32p, _u32p, _u32p]
        L.orc_draw_bits.restype = C.c_uint32
        L.orc_draw_bits.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_draw_uniform.restype = C.c_float
        L.orc_draw_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int]
        L.orc_init_color.restype = C.c_uint32
        L.orc_init_color.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_init_colors.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _u32p]
        L.orc_fill_bits.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _u32p]
        L.orc_fill_tape.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, _f32p]
        L.orc_setup_rnd2.restype = C.c_int
        L.orc_setup_rnd2.argtypes = [C.c_uint32, C.c_float, _u32p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.orc_violation_count.restype = C.c_uint64
        L.orc_violation_count.argtypes = [C.c_uint32, _u32p, _u32p, _u32p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_conflict_edges.restype = C.c_uint64
        L.orc_conflict_edges.argtypes = [C.c_uint32, _u32p, _u32p, _u32p, C.c_uint32, C.c_uint32]
        L.orc_occupancy.restype = C.c_uint32
        L.orc_occupancy.argtypes = [C.c_uint32, _u32p, _u32p, _u32p, C.c_uint32, _u8p]
        L.orc_fill_p_uniform.argtypes = [C.c_uint32, C.c_float, _u8p, C.c_uint32, _f32p]
        L.orc_sweep.restype = C.c_uint64
        L.orc_sweep.argtypes = [C.c_uint32, _u32p, _u32p, C.c_uint32, C.c_float, C.c_uint32, C.c_int, _u32p, _u32p,
                                C.c_void_p, _f32p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_class_sizes.argtypes = [C.c_uint32, _u32p, C.c_uint32, _u32p]
        L.orc_color_stats.argtypes = [C.c_uint32, C.c_uint32, _u32p, C.c_float, C.POINTER(ColorStats)]
        L.orc_run.restype = C.c_uint32
        L.orc_run.argtypes = [C.c_uint32, _u32p, _u32p, C.c_uint32, C.c_float, C.c_uint32, C.c_int, C.c_uint64,
                              C.c_uint32, C.c_uint64, _u32p, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        L.orc_tailcut.restype = C.c_uint32
        L.orc_tailcut.argtypes = [C.c_uint32, _u32p, _u32p, C.c_uint32, _u32p, C.c_uint32, C.POINTER(C.c_uint64)]
        self.libc = C.CDLL(None)

    # -- RNG ------------------------------------------------------------------------------------
    def philox(self, ctr, key):
        out = np.zeros(4, np.uint32)
        self.L.orc_philox4x32_10(np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), out)
        return out

    def tape(self, seed, sweep, n, proposal=UNIFORM, vb=0):
        u = np.empty(n - vb, np.float32)
        self.L.orc_fill_tape(seed, sweep, vb, n, proposal, u)
        return u

    def draw_bits(self, seed, sweep, n, purpose):
        out = np.empty(n, np.uint32)
        self.L.orc_fill_bits(seed, sweep, 0, n, purpose, out)
        return out

    def init_colors(self, seed, n, nCol):
        out = np.empty(n, np.uint32)
        self.L.orc_init_colors(seed, 0, n, nCol, out)
        return out

    # -- graph ------------------------------------------------------------------------------------
    def setup_rnd2(self, n, prob, srand=1):
        """Graph::setupRnd2 with libc rand(); srand(1) == glibc's initial state."""
        if srand is not None:
            self.libc.srand(srand)
        cumul = np.zeros(n + 1, np.uint32)
        nnz = C.c_uint64()
        rc = self.L.orc_setup_rnd2(n, prob, cumul, None, 0, C.byref(nnz))
        assert rc == 0, rc
        neighs = np.zeros(max(nnz.value, 1), np.uint32)
        rc = self.L.orc_setup_rnd2(n, prob, cumul, neighs.ctypes.data, nnz.value, C.byref(nnz))
        assert rc == 0, rc
        return cumul, neighs[:nnz.value].copy()

    # -- counting -----------------------------------------------------------------------------------
    def violation_count(self, cumul, neighs, colors, vb=0, ve=None, want_flags=False):
        n = len(cumul) - 1
        ve = n if ve is None else ve
        flags = np.zeros(n, np.uint8) if want_flags else None
        r = self.L.orc_violation_count(n, cumul, neighs, colors, vb, ve,
                                       flags.ctypes.data if want_flags else None)
        return (r, flags) if want_flags else r

    def conflict_edges(self, cumul, neighs, colors, vb=0, ve=None):
        n = len(cumul) - 1
        return self.L.orc_conflict_edges(n, cumul, neighs, colors, vb, n if ve is None else ve)

    def occupancy(self, v, cumul, neighs, colors, nCol):
        occ = np.zeros(nCol, np.uint8)
        free = self.L.orc_occupancy(v, cumul, neighs, colors, nCol, occ)
        return occ, free

    def fill_p(self, nCol, eps, occ, own):
        p = np.zeros(nCol, np.float32)
        self.L.orc_fill_p_uniform(nCol, eps, occ, own, p)
        return p

    # -- sweep ----------------------------------------------------------------------------------------
    def sweep(self, cumul, neighs, nCol, eps, colors, u, proposal=UNIFORM, taboo=None, taboo_iter=0,
              hist=None, vb=0, ve=None):
        """Returns (Cstar, overflowCount).  taboo (uint32[n]) is updated in place."""
        n = len(cumul) - 1
        ve = n if ve is None else ve
        cstar = colors.copy()
        if proposal == DYNAMIC and hist is None:
            hist = self.class_sizes(colors, nCol)
        ov = self.L.orc_sweep(n, cumul, neighs, nCol, eps, taboo_iter, proposal, colors, cstar,
                              taboo.ctypes.data if taboo is not None else None, u,
                              hist.ctypes.data if hist is not None else None, vb, ve)
        return cstar, ov

    def class_sizes(self, colors, nCol):
        hist = np.zeros(nCol, np.uint32)
        self.L.orc_class_sizes(len(colors), colors, nCol, hist)
        return hist

    def color_stats(self, n, nCol, hist, prob):
        st = ColorStats()
        self.L.orc_color_stats(n, nCol, hist, prob, C.byref(st))
        return st

    def run(self, cumul, neighs, nCol, eps, colors, seed, proposal=UNIFORM, taboo_iter=0, max_rip=250, z=0):
        """Free-running chain; returns (colors, sweeps, finalCount, maxIterReached)."""
        n = len(cumul) - 1
        c = colors.copy()
        cnt = C.c_uint64()
        hit = C.c_int()
        sweeps = self.L.orc_run(n, cumul, neighs, nCol, eps, taboo_iter, proposal, seed, max_rip, z, c,
                                C.byref(cnt), C.byref(hit))
        return c, sweeps, cnt.value, bool(hit.value)

    def tailcut(self, cumul, neighs, nCol, colors, max_rounds=64):
        c = colors.copy()
        left = C.c_uint64()
        rounds = self.L.orc_tailcut(len(cumul) - 1, cumul, neighs, nCol, c, max_rounds, C.byref(left))
        return c, rounds, left.value


class Ref:
    """oracle/ref_harness.cpp over the unmodified reference (only where oracle/_ref was built)."""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self):
        L = self.L = C.CDLL(REF_SO)
        vp = C.c_void_p
        L.ref_graph_simulate.restype = vp
        L.ref_graph_simulate.argtypes = [C.c_uint32, C.c_float, C.c_uint32]
        L.ref_graph_from_csr.restype = vp
        L.ref_graph_from_csr.argtypes = [C.c_uint32, C.c_uint32, _u32p, _u32p, C.c_float]
        L.ref_graph_info.argtypes = [vp] + [C.c_void_p] * 5
        L.ref_graph_from_file.restype = vp
        L.ref_graph_from_file.argtypes = [C.c_char_p]
        L.ref_graph_copy_csr.argtypes = [vp, _u32p, _u32p]
        L.ref_graph_free.argtypes = [vp]
        L.ref_mcmc_create.restype = vp
        L.ref_mcmc_create.argtypes = [vp, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint32,
                                      C.c_uint32, C.c_int, C.c_uint32]
        L.ref_mcmc_free.argtypes = [vp]
        L.ref_mcmc_get_colors.argtypes = [vp, _u32p]
        L.ref_mcmc_set_colors.argtypes = [vp, _u32p]
        L.ref_mcmc_get_taboo.argtypes = [vp, _u32p]
        L.ref_mcmc_set_taboo.argtypes = [vp, _u32p]
        L.ref_mcmc_violations.restype = C.c_uint64
        L.ref_mcmc_violations.argtypes = [vp, _u32p, C.c_void_p]
        L.ref_mcmc_occupancy.restype = C.c_uint64
        L.ref_mcmc_occupancy.argtypes = [vp, _u32p, C.c_uint32, _u8p]
        L.ref_mcmc_fill_p.argtypes = [vp, C.c_uint32, _f32p]
        L.ref_mcmc_sweep_tape.restype = C.c_uint64
        L.ref_mcmc_sweep_tape.argtypes = [vp, _f32p, C.POINTER(C.c_uint64)]
        L.ref_mcmc_sweep_range_timed.restype = C.c_double
        L.ref_mcmc_sweep_range_timed.argtypes = [vp, _f32p, C.c_uint64, C.c_uint64]
        L.ref_mcmc_run_native.restype = C.c_uint64
        L.ref_mcmc_run_native.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        L.ref_mcmc_run.argtypes = [vp]
        L.ref_mcmc_scan_overflows.restype = C.c_uint64
        L.ref_mcmc_scan_overflows.argtypes = [vp, C.c_uint64, np.ctypeslib.ndpointer(np.float64), C.c_uint64]
        L.ref_mcmc_iterations.restype = C.c_uint64
        L.ref_mcmc_iterations.argtypes = [vp]
        L.ref_mcmc_save_stats.argtypes = [vp, C.c_uint64, C.c_float, C.c_char_p]
        L.ref_mcmc_save_colors.argtypes = [vp, C.c_char_p]

    def graph_simulate(self, n, prob, srand=1):
        return self.L.ref_graph_simulate(n, prob, srand or 0)

    def graph_from_csr(self, cumul, neighs, prob=0.0):
        return self.L.ref_graph_from_csr(len(cumul) - 1, len(neighs), cumul, neighs, prob)

    def graph_from_file(self, path):
        return self.L.ref_graph_from_file(path.encode())

    def graph_info(self, g):
        n, nnz, mx, mn = (C.c_uint32() for _ in range(4))
        mean = C.c_float()
        self.L.ref_graph_info(g, C.byref(n), C.byref(nnz), C.byref(mx), C.byref(mn), C.byref(mean))
        return dict(n=n.value, nnz=nnz.value, maxDeg=mx.value, minDeg=mn.value, meanDeg=mean.value)

    def graph_csr(self, g):
        info = self.graph_info(g)
        cumul = np.zeros(info["n"] + 1, np.uint32)
        neighs = np.zeros(max(info["nnz"], 1), np.uint32)
        self.L.ref_graph_copy_csr(g, cumul, neighs)
        return cumul, neighs[:info["nnz"]].copy()

    def mcmc(self, g, nCol, seed, eps=1e-8, taboo_iter=0, tailcut=False, max_rip=250, ratio=1.0):
        # hard-coded values of main.cu:160-168: lambda=1, ratioFreezed=1e-2
        return self.L.ref_mcmc_create(g, nCol, ratio, 1.0, eps, 1e-2, max_rip, taboo_iter, int(tailcut), seed)

    def get_colors(self, h, n):
        out = np.zeros(n, np.uint32)
        self.L.ref_mcmc_get_colors(h, out)
        return out

    def set_colors(self, h, colors):
        self.L.ref_mcmc_set_colors(h, np.ascontiguousarray(colors, np.uint32))

    def get_taboo(self, h, n):
        out = np.zeros(n, np.uint32)
        self.L.ref_mcmc_get_taboo(h, out)
        return out

    def violations(self, h, colors):
        flags = np.zeros(len(colors), np.uint8)
        r = self.L.ref_mcmc_violations(h, colors, flags.ctypes.data)
        return r, flags

    def occupancy(self, h, colors, v, nCol):
        occ = np.zeros(nCol, np.uint8)
        free = self.L.ref_mcmc_occupancy(h, colors, v, occ)
        return occ, free

    def fill_p(self, h, v, nCol):
        p = np.zeros(nCol, np.float32)
        self.L.ref_mcmc_fill_p(h, v, p)
        return p

    def sweep_tape(self, h, u):
        ov = C.c_uint64(0)
        before = self.L.ref_mcmc_sweep_tape(h, np.ascontiguousarray(u, np.float32), C.byref(ov))
        return before, ov.value

    def sweep_range_timed(self, h, u, vb, ve):
        return self.L.ref_mcmc_sweep_range_timed(h, np.ascontiguousarray(u, np.float32), vb, ve)

    def scan_overflows(self, h, max_sweeps=260, cap=64):
        """run_native's loop on handle h, returning rows (sweep, vertex, draw, final cdf) of every CDF overflow
        (the events on which the reference calls libc rand(), coloringMCMC_CPU.cpp:517-520)."""
        rec = np.zeros(4 * cap)
        k = self.L.ref_mcmc_scan_overflows(h, max_sweeps, rec, cap)
        return rec[:4 * k].reshape(-1, 4)

    def run_native(self, h):
        sweeps = C.c_uint64()
        hit = C.c_int()
        viol = self.L.ref_mcmc_run_native(h, C.byref(sweeps), C.byref(hit))
        return viol, sweeps.value, bool(hit.value)


def importer_csr(n, src, dst):
    """Plain restatement of Graph::setupImporterNew (graph/graphCPU.cpp:112-170) on vertex-id pairs: first pass counts
    (self-loops skipped, one slot for the edge and one for its back-edge, :121-132), prefix sum (:135-137), second pass fills
    the rows in file order (:150-165).  Duplicate edges are kept.  Python loops: small inputs only.  Pin: the loops are the
    reference's, statement for statement; tests/test_host_layer.py checks the C++ host importer built the same way against
    the reference's own CSR (oracle/_ref) on a file."""
    cumul = np.zeros(n + 1, np.int64)
    for s_, d_ in zip(src, dst):
        if s_ != d_:
            cumul[s_ + 1] += 1
            cumul[d_ + 1] += 1
    for i in range(1, n + 1):
        cumul[i] += cumul[i - 1]
    neighs = np.zeros(int(cumul[n]), np.uint32)
    temp = np.zeros(n, np.int64)
    for s_, d_ in zip(src, dst):
        if s_ != d_:
            neighs[cumul[s_] + temp[s_]] = d_
            temp[s_] += 1
            neighs[cumul[d_] + temp[d_]] = s_
            temp[d_] += 1
    return cumul.astype(np.uint32), neighs


REFGPU_SO = os.path.join(HERE, "_ref_gpu", "libmcmc_refgpu.so")


class RefGpu:
    """oracle/refgpu_harness.cu over the UNMODIFIED reference GPU colourer compiled for sm_100a (oracle/_ref_gpu, built by
    `make -C oracle refgpu` where /root/reference exists; the binary travels to the GPU box).  GPU tests only."""

    @staticmethod
    def available():
        return os.path.exists(REFGPU_SO)

    def __init__(self):
        L = self.L = C.CDLL(REFGPU_SO)
        vp = C.c_void_p
        L.refgpu_create.restype = vp
        L.refgpu_create.argtypes = [C.c_uint32, C.c_uint32, _u32p, _u32p, C.c_float, C.c_uint32, C.c_float, C.c_float,
                                    C.c_uint32, C.c_int, C.c_uint32, C.c_float, C.c_long]
        L.refgpu_destroy.argtypes = [vp]
        for name in ("refgpu_set_colors", "refgpu_get_colors", "refgpu_set_taboo", "refgpu_get_taboo"):
            getattr(L, name).argtypes = [vp, _u32p]
        L.refgpu_peek_draws.argtypes = [vp, _f32p]
        L.refgpu_step_dynamic.argtypes = [vp, C.c_int]
        L.refgpu_conflicts.argtypes = [vp]
        L.refgpu_tailcut.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
        L.refgpu_run.argtypes = [vp, C.c_int, C.c_char_p, C.POINTER(C.c_int)]

    def create(self, cumul, neighs, nCol, prob=0.0, eps=1e-8, taboo_iter=0, tailcut=False, max_rip=250, ratio=1.0, curand_seed=1234):
        n = len(cumul) - 1
        h = self.L.refgpu_create(n, len(neighs), np.ascontiguousarray(cumul, np.uint32), np.ascontiguousarray(neighs, np.uint32),
                                 prob, nCol, eps, 1.0, taboo_iter, int(tailcut), max_rip, 1.0 / ratio, curand_seed)
        assert h, "refgpu_create failed"
        return h

    def destroy(self, h):
        self.L.refgpu_destroy(h)

    def set_colors(self, h, c):
        assert self.L.refgpu_set_colors(h, np.ascontiguousarray(c, np.uint32)) == 0

    def get_colors(self, h, n):
        out = np.zeros(n, np.uint32)
        assert self.L.refgpu_get_colors(h, out) == 0
        return out

    def set_taboo(self, h, t):
        assert self.L.refgpu_set_taboo(h, np.ascontiguousarray(t, np.uint32)) == 0

    def get_taboo(self, h, n):
        out = np.zeros(n, np.uint32)
        assert self.L.refgpu_get_taboo(h, out) == 0
        return out

    def peek_draws(self, h, n):
        out = np.zeros(n, np.float32)
        assert self.L.refgpu_peek_draws(h, out) == 0
        return out

    def step_dynamic(self, h, prefill=True):
        assert self.L.refgpu_step_dynamic(h, int(prefill)) == 0

    def conflicts(self, h):
        return self.L.refgpu_conflicts(h)

    def tailcut(self, h, max_rounds=64):
        r = C.c_int()
        left = self.L.refgpu_tailcut(h, max_rounds, C.byref(r))
        return left, r.value

    def run(self, h, iteration, directory):
        """The reference's own ColoringMCMC::run(iteration): writes <directory>.log and <directory>-colors.txt."""
        mx = C.c_int()
        rip = self.L.refgpu_run(h, iteration, directory.encode(), C.byref(mx))
        return rip, bool(mx.value)
