"""GPU parity against the UNMODIFIED reference GPU colourer (oracle/_ref_gpu: ColoringMCMC<float,float> compiled for sm_100a
from /root/reference/src by `make -C oracle refgpu`; the binary travels to the GPU box, the sources do not).

The shipped GPU variant of the reference -- COLOR_BALANCE_DYNAMIC_DISTR: genDynamicDistribution + selectStarColoringBalanceDynamic
(coloringMCMC_utils.cu:64-70, coloringMCMC_balance.cu:79-143) -- and its tail cutting (coloringMCMC_main.cu:271-290,
coloringMCMC_utils.cu:73-101) have no CPU twin, so these tests run the reference's own kernels on the B200:

  * tape replay: before every sweep the harness evaluates curand_uniform on a COPY of the reference's per-vertex XORWOW states --
    the draw each vertex would take -- and that array is the tape of mcmcb200_set_tape.  C*, taboo and class sizes must be
    bit-equal after every sweep (>= 10 sweeps, several palettes, with and without taboo, eps where the FMA question shows);
  * tail cutting: same start colouring, the reference's conflictCounter + tailCutting<<<1,1>>> rounds vs mcmcb200_tailcut;
  * free-running chains: the reference's own run() (its logs parsed) vs ColoringMCMC over libmcmcb200 -- colour count and
    class-size StD within a stated tolerance over seeds (north-star check 3).
"""
import os
import re

import numpy as np
import pytest

from oracle.pyoracle import DYNAMIC, RefGpu

pytestmark = pytest.mark.gpu

EPS = 1e-8


@pytest.fixture(scope="module")
def mc():
    import mcmc_colorer_b200 as m
    return m


@pytest.fixture(scope="module")
def rg():
    if not RefGpu.available():
        pytest.skip("oracle/_ref_gpu/libmcmc_refgpu.so not built (make -C oracle refgpu needs /root/reference)")
    return RefGpu()


@pytest.fixture(scope="module")
def c1_graph(port):
    return port.setup_rnd2(1000, 0.1, srand=1)


def er(n, deg, seed):
    from mcmc_colorer_b200.graphgen import er_graph_numpy
    return er_graph_numpy(n, deg, seed=seed)


KERNEL_FLAGS = {"direct": "FLAG_FORCE_DIRECT", "blocked": "FLAG_FORCE_BLOCKED", "binned": "FLAG_FORCE_BINNED"}


def our_chain(mc, cumul, neighs, nCol, kernel, taboo=0, eps=EPS, tailcut=False, seed=0):
    prm = mc.ColoringMCMCParams(nCol=nCol, proposal=mc.PROPOSAL_DYNAMIC, convergence=mc.CONVERGE_EDGES, tabooIteration=taboo,
                                seed=seed, tailcut=tailcut, epsilon=eps)
    flags = mc.FLAG_NO_EARLY_STOP | getattr(mc, KERNEL_FLAGS[kernel])
    try:
        return mc.Chain(cumul, neighs, prm, device=0, flags=flags, stage_cap_bytes=32768 if kernel == "blocked" else 0)
    except mc.McmcError as e:
        from mcmc_colorer_b200 import capi
        if kernel == "blocked" and e.code == capi.EUNSUPPORTED:
            pytest.skip("tile does not fit the blocked kernel's stage")
        raise


# ------------------------------------------------------------------------------------------------------------
# DYNAMIC proposal: tape replay against the reference's selectStarColoringBalanceDynamic
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nCol,taboo,eps", [(600, 0, 1e-8), (700, 2, 1e-4)])
def test_dynamic_tape_replay_vs_reference_kernel_wide_palettes(mc, rg, port, c1_graph, nCol, taboo, eps):
    """the same replay on the wide-palette kernels (more than 512 colours): the reference kernel takes any palette up to the number of
    vertices (its class-size arrays have n entries, coloringMCMC_main.cu:29-33,211-214: nCol > n overruns them); the chain starts from
    8 colours so that nearly every vertex conflicts and walks"""
    cumul, neighs = c1_graph
    n = 1000
    h = rg.create(cumul, neighs, nCol, prob=0.1, eps=eps, taboo_iter=taboo, curand_seed=99 + nCol)
    prm = mc.ColoringMCMCParams(nCol=nCol, proposal=mc.PROPOSAL_DYNAMIC, convergence=mc.CONVERGE_EDGES, tabooIteration=taboo, seed=0, epsilon=eps)
    ch = mc.Chain(cumul, neighs, prm, device=0, flags=mc.FLAG_NO_EARLY_STOP)
    assert ch.kernel_mode() == "wide-binned"
    rng = np.random.default_rng(nCol)
    c0 = (rng.integers(0, 8, n) * (nCol // 8)).astype(np.uint32)
    rg.set_colors(h, c0)
    ch.init_colors(c0)
    for s in range(8):
        u = rg.peek_draws(h, n)
        rg.step_dynamic(h, prefill=True)
        ch.set_tape(u[None, :])
        ch.sweep(1)
        want = rg.get_colors(h, n)
        got = ch.get_colors()
        assert np.array_equal(got, want), (nCol, taboo, s, np.flatnonzero(got != want)[:8])
        assert np.array_equal(ch.class_sizes().astype(np.uint32), np.bincount(want, minlength=nCol).astype(np.uint32))
    ch.close()
    rg.destroy(h)


@pytest.mark.parametrize("kernel", ["direct", "blocked", "binned"])
@pytest.mark.parametrize("nCol,taboo,eps", [(137, 0, 1e-8), (60, 0, 1e-8), (45, 3, 1e-8), (200, 0, 1e-8), (89, 0, 1e-4), (300, 2, 1e-4)])
def test_dynamic_tape_replay_vs_reference_kernel(mc, rg, port, c1_graph, kernel, nCol, taboo, eps):
    cumul, neighs = c1_graph
    n = 1000
    h = rg.create(cumul, neighs, nCol, prob=0.1, eps=eps, taboo_iter=taboo, curand_seed=4321 + nCol)
    ch = our_chain(mc, cumul, neighs, nCol, kernel, taboo=taboo, eps=eps)
    c0 = port.init_colors(77 + nCol, n, nCol)
    rg.set_colors(h, c0)
    ch.init_colors(c0)
    sweeps = 12
    tapes = []
    for s in range(sweeps):
        u = rg.peek_draws(h, n)                            # the draws the reference kernel is about to consume
        assert u.min() > 0.0 and u.max() <= 1.0            # curand_uniform: (0, 1]
        tapes.append(u)
        rg.step_dynamic(h, prefill=True)                   # the unmodified kernels, reference launch shape
        # our sweep s with the same draws
        ch.set_tape(u[None, :])
        ch.sweep(1)
        want = rg.get_colors(h, n)
        got = ch.get_colors()
        assert np.array_equal(got, want), (kernel, nCol, taboo, s, np.flatnonzero(got != want)[:8])
        assert np.array_equal(ch.class_sizes().astype(np.uint32), np.bincount(want, minlength=nCol).astype(np.uint32))
    # taboo counters at the end: replay the whole tape on the CPU port with its taboo array and compare with the reference's
    if taboo:
        c = c0.copy()
        tb = np.zeros(n, np.uint32)
        for s in range(sweeps):
            c, _ = port.sweep(cumul, neighs, nCol, eps, c, tapes[s], DYNAMIC, taboo=tb, taboo_iter=taboo)
        assert np.array_equal(c, rg.get_colors(h, n))
        assert np.array_equal(tb, rg.get_taboo(h, n))
    ch.close()
    rg.destroy(h)


def test_dynamic_tape_replay_larger_graph(mc, rg, port):
    """n = 65 536 (a multiple of 128: the reference's reduction over-reads otherwise), mean degree 24, default path of this size"""
    n = 65_536
    cumul, neighs = er(n, 24, seed=5)
    nCol = int(np.diff(cumul.astype(np.int64)).max())
    h = rg.create(cumul, neighs, nCol, prob=24.0 / n, curand_seed=99)
    prm = mc.ColoringMCMCParams(nCol=nCol, proposal=mc.PROPOSAL_DYNAMIC, convergence=mc.CONVERGE_EDGES, seed=1)
    ch = mc.Chain(cumul, neighs, prm, device=0, flags=mc.FLAG_NO_EARLY_STOP)
    c0 = port.init_colors(5, n, nCol)
    rg.set_colors(h, c0)
    ch.init_colors(c0)
    for s in range(10):
        u = rg.peek_draws(h, n)
        rg.step_dynamic(h, prefill=True)
        ch.set_tape(u[None, :])
        ch.sweep(1)
        want = rg.get_colors(h, n)
        assert np.array_equal(ch.get_colors(), want), s
        # the reference's conflicting-edge count (conflictCounter kernel) == ours
        assert ch.status().conflictEdges == rg.conflicts(h)
    ch.close()
    rg.destroy(h)


# ------------------------------------------------------------------------------------------------------------
# tail cutting against the reference's tailCutting<<<1,1>>>
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nCol,seed", [(137, 3), (60, 4), (45, 5), (40, 6)])
def test_tailcut_vs_reference_kernel(mc, rg, port, c1_graph, nCol, seed):
    cumul, neighs = c1_graph
    n = 1000
    c0 = port.init_colors(seed, n, nCol)
    h = rg.create(cumul, neighs, nCol, prob=0.1, tailcut=True)
    rg.set_colors(h, c0)
    left, rounds = rg.tailcut(h, max_rounds=64)
    want = rg.get_colors(h, n)
    ch = our_chain(mc, cumul, neighs, nCol, "direct", tailcut=True)
    ch.init_colors(c0)
    ch.tailcut(64)
    got = ch.get_colors()
    assert np.array_equal(got, want), np.flatnonzero(got != want)[:8]
    assert ch.status().conflictEdges == left
    # and the CPU restatement used by the non-reference tests agrees with the reference kernel too
    pw, prounds, pleft = port.tailcut(cumul, neighs, nCol, c0)
    assert np.array_equal(pw, want) and pleft == left
    ch.close()
    rg.destroy(h)


# ------------------------------------------------------------------------------------------------------------
# free-running chains: the reference's own run() on the GPU vs the drop-in
# ------------------------------------------------------------------------------------------------------------
def parse_ref_gpu_log(path):
    txt = open(path).read()
    out = {}
    m = re.search(r"Number of used colors is (\d+) on (\d+) available", txt)
    out["used"], out["nCol"] = int(m.group(1)), int(m.group(2))
    out["std"] = float(re.search(r"StandardDeviation (\S+)", txt).group(1))
    out["bal"] = float(re.search(r"BalancingIndex (\S+)", txt).group(1))
    out["maxiter"] = re.search(r"Max iteration reached (\w+)", txt).group(1) == "yes"
    return out


@pytest.mark.parametrize("ratio", [1.0, 1.5, 2.0])
def test_free_running_statistics_vs_reference_gpu_run(mc, rg, port, tmp_path, ratio):
    """north-star check 3 on an ER graph (n = 20 480, mean degree 40): proper colouring, colour count equal and class-size StD /
    BalancingIndex of the drop-in within 20 % of the reference GPU colourer's mean over 5 seeds each (DYNAMIC proposal, the
    shipped default of both).  The two use different RNGs (XORWOW vs Philox), so only the distributions are comparable."""
    n = 20_480
    cumul, neighs = er(n, 40, seed=11)
    g = mc.Graph(cumul, neighs, prob=40.0 / n)
    nCol = mc.Graph.default_ncol(g.getMaxNodeDeg(), ratio)
    src = np.repeat(np.arange(n, dtype=np.uint32), np.diff(cumul.astype(np.int64)))
    ref_stats, our_stats = [], []
    for seed in range(1, 6):
        h = rg.create(cumul, neighs, nCol, prob=40.0 / n, curand_seed=seed, ratio=ratio)
        d = str(tmp_path / f"ref-{ratio}-{seed}")
        rip, hit = rg.run(h, seed, d)
        assert not hit
        r = parse_ref_gpu_log(d + ".log")
        colors = np.loadtxt(d + "-colors.txt", dtype=np.int64)[:, 1]
        assert not np.any(colors[src] == colors[neighs])                    # the reference's result is proper
        ref_stats.append(r)
        rg.destroy(h)
        prm = mc.ColoringMCMCParams(nCol=nCol, seed=seed, numColorRatio=1.0 / ratio)    # defaults: DYNAMIC, edges
        col = mc.ColoringMCMC(g, None, prm, device=0)
        col.setDirectoryPath(str(tmp_path / f"ours-{ratio}-{seed}"))
        c = col.run(seed)
        assert not np.any(c[src] == c[neighs]) and not col.maxIterReached
        our_stats.append(dict(used=col.stats["used"], std=col.stats["std"], bal=col.stats["balancingIndex"], rip=col.rip))
        col.chain.close()
    ref_used = np.mean([r["used"] for r in ref_stats]); our_used = np.mean([r["used"] for r in our_stats])
    ref_std = np.mean([r["std"] for r in ref_stats]); our_std = np.mean([r["std"] for r in our_stats])
    ref_bal = np.mean([r["bal"] for r in ref_stats]); our_bal = np.mean([r["bal"] for r in our_stats])
    print(f"ratio {ratio} nCol {nCol}: used ref {ref_used} ours {our_used}; StD ref {ref_std:.3f} ours {our_std:.3f}; BalancingIndex ref {ref_bal:.4f} ours {our_bal:.4f}")
    assert abs(our_used - ref_used) <= 1.0                                   # colour count (all colours are in use on both sides)
    assert abs(our_std - ref_std) <= 0.20 * ref_std, (our_std, ref_std)      # tolerance: 20 % of the reference's mean StD
    assert abs(our_bal - ref_bal) <= 0.20 * ref_bal, (our_bal, ref_bal)
