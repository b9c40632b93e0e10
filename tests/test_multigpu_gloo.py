"""CPU, world_size 2, gloo: the multi-GPU driver (mcmc_colorer_b200.multigpu.DistributedSweeper) over a test double
of the per-rank engine that implements the engine interface with the CPU oracle.  Covers the vertex partition, the
in-place all-gather layout of the colour slices, the all-reduce of the counters / class-size deltas, the stale-counter
protocol and the GPU-count independence of the trajectory -- without a GPU."""
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleEngine:
    """TEST DOUBLE (tests only): GpuEngine's interface on top of the oracle port, numpy/torch-CPU buffers."""

    def __init__(self, port, cumul, neighs, vb, ve, chunk, world, nCol, seed, proposal):
        import torch
        self.P, self.cumul, self.neighs = port, cumul, neighs
        self.n, self.vb, self.ve, self.nCol, self.seed, self.proposal = len(cumul) - 1, vb, ve, nCol, seed, proposal
        self.elem_bytes = 1
        self.bufs = [torch.zeros(chunk * world + 64, dtype=torch.uint8) for _ in range(2)]
        self._counters = torch.zeros(2 + nCol, dtype=torch.int64)
        self.t, self.counts_sweep, self.last = 0, None, (0, 0)
        self.src = np.repeat(np.arange(self.n, dtype=np.uint32), np.diff(cumul.astype(np.int64)))

    def _cur(self):
        return self.bufs[self.t & 1][: self.n].numpy().astype(np.uint32)

    def init_colors(self, colors=None):
        import torch
        c = self.P.init_colors(self.seed, self.n, self.nCol) if colors is None else colors
        self.bufs[0][: self.n] = torch.from_numpy(c.astype(np.uint8))
        self.hist = self.P.class_sizes(c, self.nCol).astype(np.int64)
        self.t, self.counts_sweep = 0, None
        self._counters.zero_()

    def _local_counts(self, c):
        lo, hi = self.cumul[self.vb], self.cumul[self.ve]
        same = c[self.src[lo:hi]] == c[self.neighs[lo:hi]]
        self._counters[0] += int(same.sum())                                   # directed conflicts of the owned rows
        self._counters[1] += int(np.unique(self.src[lo:hi][same]).size)        # violating owned vertices

    def local_sweep(self):
        import torch
        c = self._cur()
        self._local_counts(c)
        u = self.P.tape(self.seed, self.t + 1, self.n, self.proposal)
        cs, _ = self.P.sweep(self.cumul, self.neighs, self.nCol, 1e-8, c, u, self.proposal,
                             hist=self.hist.astype(np.uint32), vb=self.vb, ve=self.ve)
        self.bufs[(self.t + 1) & 1][self.vb:self.ve] = torch.from_numpy(cs[self.vb:self.ve].astype(np.uint8))
        old, new = c[self.vb:self.ve], cs[self.vb:self.ve]
        d = np.bincount(new, minlength=self.nCol).astype(np.int64) - np.bincount(old, minlength=self.nCol).astype(np.int64)
        self._counters[2:] += torch.from_numpy(d)
        self.pending_count = False

    def next_colors(self):
        return self.bufs[(self.t + 1) & 1]

    def cur_colors(self):
        return self.bufs[self.t & 1]

    def init_colors_slice(self, own, sweeper):
        """sliced host interface (GpuEngine.init_colors_slice): this rank provides only the colours it owns"""
        import torch
        self.bufs[0].zero_()
        self.bufs[0][self.vb:self.ve] = torch.from_numpy(own.astype(np.uint8))
        self.t, self.counts_sweep = 0, None
        self._counters.zero_()
        sweeper.gather_current()
        self.hist = self.P.class_sizes(self._cur(), self.nCol).astype(np.int64)

    def counters(self):
        return self._counters

    def finalize(self, advanced):
        self.last = (int(self._counters[0]) // 2, int(self._counters[1]))
        self.counts_sweep = self.t
        if advanced:
            self.hist = self.hist + self._counters[2:].numpy()
            self.t += 1
        self._counters.zero_()

    # ---- distributed tail cutting: the engine's part, in numpy (same contract as mcmcb200_tailcut_dist_*) ----
    def class_sizes(self):
        return np.bincount(self._cur(), minlength=self.nCol).astype(np.int64)

    def _row(self, v):
        return self.neighs[self.cumul[v]:self.cumul[v + 1]]

    def tc_begin(self, order):
        c = self._cur()
        self.order = np.asarray(order, dtype=np.uint32)
        self.flist = [v for v in range(self.vb, self.ve) if any(u > v and c[u] == c[v] for u in self._row(v))]
        self.pend = set()
        return np.array(self.flist, dtype=np.uint32)

    def tc_mark(self, ids):
        self.pend = set(int(i) for i in ids)

    def tc_round(self):
        import torch
        c = self._cur()
        ready = [v for v in self.flist if v in self.pend and not any(u < v and int(u) in self.pend for u in self._row(v))]
        ids, cols, inexact = [], [], False
        for v in ready:
            occ = set(int(c[u]) for u in self._row(v))
            col, j = int(c[v]), 0
            while col in occ and j < self.nCol:                      # coloringMCMC_utils.cu:91-95
                col = int(self.order[j]); j += 1
            inexact = inexact or (j == self.nCol and col in occ)
            ids.append(v); cols.append(col)
        for v, col in zip(ids, cols):
            self.bufs[self.t & 1][v] = col
            self.pend.discard(v)
        left = sum(1 for v in self.flist if v in self.pend)
        return np.array(ids, dtype=np.uint32), np.array(cols, dtype=np.uint32), left, inexact

    def tc_apply(self, ids, cols):
        for v, col in zip(ids, cols):
            self.bufs[self.t & 1][int(v)] = int(col)
            self.pend.discard(int(v))

    def tc_recount(self):
        c = self._cur()
        d = v_ = nf = 0
        for v in range(self.vb, self.ve):
            same = [u for u in self._row(v) if c[u] == c[v]]
            d += len(same); v_ += 1 if same else 0; nf += 1 if any(u > v for u in same) else 0
        return d, v_, nf

    def tc_end(self, directed, viol, exact):
        self.last = (int(directed) // 2, int(viol))
        self.counts_sweep = self.t if exact else None
        self.hist = self.class_sizes()

    def status(self):
        if self.counts_sweep != self.t:
            self._local_counts(self._cur())                                    # local counting pass, like the C ABI in split mode
            return types.SimpleNamespace(sweep=self.t, countsSweep=0xffffffff, conflictEdges=0, violatingVertices=0)
        return types.SimpleNamespace(sweep=self.t, countsSweep=self.t, conflictEdges=self.last[0], violatingVertices=self.last[1])


def _worker(rank, world, port_file, result_dir, proposal):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from mcmc_colorer_b200.graphgen import er_graph_numpy
    from mcmc_colorer_b200.multigpu import DistributedSweeper, partition
    from oracle.pyoracle import Port
    dist.init_process_group("gloo", init_method=f"file://{port_file}", rank=rank, world_size=world)
    P = Port()
    n = 3001
    cumul, neighs = er_graph_numpy(n, 12, seed=5)
    nCol = int(np.diff(cumul.astype(np.int64)).max())
    parts, chunk = partition(n, world, align=256)
    vb, ve = parts[rank]
    eng = OracleEngine(P, cumul, neighs, vb, ve, chunk, world, nCol, seed=11, proposal=proposal)
    sw = DistributedSweeper(eng, rank, world, chunk)
    eng.init_colors(None)
    hist = []
    for s in range(6):
        st = sw.status()
        hist.append((st.sweep, st.conflictEdges, st.violatingVertices))
        sw.sweep(1)
    final = eng.bufs[eng.t & 1][:n].numpy().astype(np.uint32)
    class_sizes = eng.hist.copy()                     # maintained from the all-reduced deltas over the 6 sweeps
    # sliced host interface: restart from a colouring of which every rank holds only its own part
    c0 = P.init_colors(77, n, nCol)
    eng.init_colors_slice(c0[vb:ve], sw)
    st = sw.status()
    sliced = eng.bufs[0][:n].numpy().astype(np.uint32)
    sw.sweep(1)
    after = eng.bufs[eng.t & 1][:n].numpy().astype(np.uint32)
    np.savez(os.path.join(result_dir, f"rank{rank}.npz"), final=final, hist=np.array(hist), class_sizes=class_sizes,
             sliced=sliced, sliced_counts=np.array([st.conflictEdges, st.violatingVertices]), after=after)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("proposal", [0, 1])
def test_two_rank_driver_matches_single_process_oracle(port, tmp_path, proposal):
    import torch.multiprocessing as mp
    from mcmc_colorer_b200.graphgen import er_graph_numpy
    from mcmc_colorer_b200.multigpu import partition
    world = 2
    rdv = str(tmp_path / "rdv")
    mp.spawn(_worker, args=(world, rdv, str(tmp_path), proposal), nprocs=world, join=True)
    n = 3001
    cumul, neighs = er_graph_numpy(n, 12, seed=5)
    nCol = int(np.diff(cumul.astype(np.int64)).max())
    c = port.init_colors(11, n, nCol)
    want_hist = []
    for s in range(6):
        want_hist.append((s, port.conflict_edges(cumul, neighs, c), port.violation_count(cumul, neighs, c)))
        c, _ = port.sweep(cumul, neighs, nCol, 1e-8, c, port.tape(11, s + 1, n, proposal), proposal)
    for r in range(world):
        z = np.load(str(tmp_path / f"rank{r}.npz"))
        assert np.array_equal(z["final"], c), r                       # same trajectory as one process, on every rank
        assert z["hist"].tolist() == [list(x) for x in want_hist]
        assert np.array_equal(z["class_sizes"], port.class_sizes(c, nCol).astype(np.int64))
    c0 = port.init_colors(77, n, nCol)
    c1, _ = port.sweep(cumul, neighs, nCol, 1e-8, c0, port.tape(11, 1, n, proposal), proposal)
    for r in range(world):
        z = np.load(str(tmp_path / f"rank{r}.npz"))
        assert np.array_equal(z["sliced"], c0), r                     # every rank ends up with the whole colouring
        assert z["sliced_counts"].tolist() == [port.conflict_edges(cumul, neighs, c0), port.violation_count(cumul, neighs, c0)]
        assert np.array_equal(z["after"], c1), r
    parts, chunk = partition(n, world)
    assert parts[0][0] == 0 and parts[-1][1] == n and chunk % 256 == 0 and parts[0][1] == parts[1][0]


def test_partition_properties():
    from mcmc_colorer_b200.multigpu import partition
    for n in (1, 255, 256, 257, 1000, 100_000_000):
        for world in (1, 2, 4, 8):
            parts, chunk = partition(n, world)
            assert chunk % 256 == 0 and chunk * world >= n and chunk * world <= n + 256 * world
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert all(0 <= ve - vb <= chunk for vb, ve in parts)


def _tc_worker(rank, world, port_file, result_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from mcmc_colorer_b200.graphgen import er_graph_numpy
    from mcmc_colorer_b200.multigpu import DistributedSweeper, partition
    from oracle.pyoracle import Port
    dist.init_process_group("gloo", init_method=f"file://{port_file}", rank=rank, world_size=world)
    P = Port()
    n = 2500
    cumul, neighs = er_graph_numpy(n, 10, seed=9)
    nCol = int(np.diff(cumul.astype(np.int64)).max()) - 6            # tight palette: conflicts are left after a few sweeps
    parts, chunk = partition(n, world, align=256)
    vb, ve = parts[rank]
    eng = OracleEngine(P, cumul, neighs, vb, ve, chunk, world, nCol, seed=3, proposal=0)
    sw = DistributedSweeper(eng, rank, world, chunk)
    eng.init_colors(None)
    sw.sweep(2)
    before = eng.bufs[eng.t & 1][:n].numpy().astype(np.uint32).copy()
    passes = sw.tailcut(64)
    st = sw.status()
    after = eng.bufs[eng.t & 1][:n].numpy().astype(np.uint32)
    np.savez(os.path.join(result_dir, f"tc{rank}.npz"), before=before, after=after, counts=np.array([st.conflictEdges, st.violatingVertices, passes]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_tailcut_matches_sequential_reference_semantics(port, tmp_path):
    """DistributedSweeper.tailcut over two ranks == the sequential greedy repair (oracle port of coloringMCMC_utils.cu:73-101 +
    coloringMCMC_main.cu:271-290) on the same colouring: colours, what is left, and identical replicas on both ranks."""
    import torch.multiprocessing as mp
    from mcmc_colorer_b200.graphgen import er_graph_numpy
    world = 2
    port_file = str(tmp_path / "rdzv_tc")
    mp.spawn(_tc_worker, args=(world, port_file, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (np.load(tmp_path / f"tc{r}.npz") for r in range(world))
    assert np.array_equal(r0["before"], r1["before"]) and np.array_equal(r0["after"], r1["after"])
    n = 2500
    cumul, neighs = er_graph_numpy(n, 10, seed=9)
    nCol = int(np.diff(cumul.astype(np.int64)).max()) - 6
    assert port.conflict_edges(cumul, neighs, r0["before"]) > 0
    want, rounds, left = port.tailcut(cumul, neighs, nCol, r0["before"])
    assert np.array_equal(r0["after"], want)
    assert int(r0["counts"][0]) == left == port.conflict_edges(cumul, neighs, want)
    assert int(r0["counts"][1]) == port.violation_count(cumul, neighs, want)
