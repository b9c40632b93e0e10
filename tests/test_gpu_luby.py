"""GPU: the Luby cross-check colourer against a numpy restatement of the same rounds, and the
MCMC-vs-Luby comparison BASELINE config 5 asks for (colour count and balance) at a test-sized n."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def luby_numpy(port, cumul, neighs, seed):
    n = len(cumul) - 1
    deg = np.diff(cumul.astype(np.int64))
    src = np.repeat(np.arange(n), deg)
    nb = neighs.astype(np.int64)
    colors = np.zeros(n, np.uint32)
    color = rnd = 0
    while (colors == 0).any():
        color += 1
        cands = colors == 0
        ind = np.zeros(n, bool)
        while cands.any():
            rnd += 1
            u = ((port.draw_bits(seed, rnd, n, 2) >> 8).astype(np.float32) + np.float32(1)) * np.float32(2.0 ** -24)
            chosen = cands & (u < np.float32(0.5))
            bad = chosen[src] & chosen[nb] & (deg[src] <= deg[nb])
            kill = np.zeros(n, bool)
            kill[src[bad]] = True
            keep = chosen & ~kill
            ind |= keep
            cands[keep] = False
            cands[nb[keep[src]]] = False
        colors[ind] = color
    return colors, color, rnd


def test_luby_matches_numpy_restatement_and_is_proper(port):
    import mcmc_colorer_b200 as mc
    from mcmc_colorer_b200.graphgen import er_graph_numpy
    for n, d, seed in [(3000, 10, 1), (20_001, 16, 7)]:
        cumul, neighs = er_graph_numpy(n, d, seed=3)
        got, ncol, rounds = mc.luby_color(cumul, neighs, seed=seed, device=0)
        want, wcol, wr = luby_numpy(port, cumul, neighs, seed)
        assert (ncol, rounds) == (wcol, wr) and np.array_equal(got, want)
        assert got.min() >= 1 and port.conflict_edges(cumul, neighs, got) == 0
        assert ncol <= int(np.diff(cumul.astype(np.int64)).max()) + 1


def test_mcmc_is_better_balanced_than_luby_with_comparable_colours(port):
    """BASELINE config 5 in miniature: Luby's greedy classes shrink geometrically, the MCMC sampler's are balanced."""
    import mcmc_colorer_b200 as mc
    from mcmc_colorer_b200.graphgen import er_graph_numpy
    n = 100_000
    cumul, neighs = er_graph_numpy(n, 16, seed=11)
    lub, lcol, _ = mc.luby_color(cumul, neighs, seed=5, device=0)
    lhist = np.bincount(lub - 1, minlength=lcol)
    maxdeg = int(np.diff(cumul.astype(np.int64)).max())
    prm = mc.ColoringMCMCParams(nCol=mc.Graph.default_ncol(maxdeg, 1.0), seed=5, proposal=mc.PROPOSAL_DYNAMIC,
                                convergence=mc.CONVERGE_EDGES, tailcut=True)
    ch = mc.Chain(cumul, neighs, prm, device=0)
    ch.init_colors(None)
    ch.sweep(60)
    if ch.status().conflictEdges:
        ch.tailcut()
    st = ch.status()
    assert st.conflictEdges == 0 and port.conflict_edges(cumul, neighs, ch.get_colors()) == 0
    mhist = ch.class_sizes().astype(np.int64)
    cv_mcmc = mhist.std() / mhist.mean()
    cv_luby = lhist.std() / lhist.mean()
    assert cv_mcmc < 0.5 * cv_luby, (cv_mcmc, cv_luby)          # balance: coefficient of variation of the class sizes
    ch.close()
