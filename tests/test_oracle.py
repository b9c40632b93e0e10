"""CPU tests of the checker itself: the C port (oracle/mcmc_oracle.c) against the committed golden fixtures
(generated from the unmodified reference by tests/golden/make_golden.py) and, where oracle/_ref is present,
against the reference live."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle.pyoracle import DYNAMIC, UNIFORM


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# Random123 v1.14 known-answer vectors for philox4x32-10 (kat_vectors)
PHILOX_KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def test_philox_known_answers(port):
    for ctr, key, want in PHILOX_KAT:
        assert port.philox(ctr, key).tolist() == want


def test_draw_conventions(port):
    n = 4096
    u0 = port.tape(11, 3, n, UNIFORM)
    u1 = port.tape(11, 3, n, DYNAMIC)
    assert u0.min() >= 0.0 and u0.max() < 1.0
    assert u1.min() > 0.0 and u1.max() <= 1.0
    assert np.array_equal(u1, u0 + np.float32(2.0 ** -24))
    c = port.init_colors(5, n, 43)
    assert c.max() < 43 and len(np.unique(c)) == 43


@pytest.fixture(scope="module")
def pins(golden_dir):
    return json.load(open(os.path.join(golden_dir, "c1_pins.json")))


@pytest.fixture(scope="module")
def c1_graph(port):
    return port.setup_rnd2(1000, 0.1, srand=1)


def test_c1_graph_matches_reference_pin(pins, c1_graph):
    cumul, neighs = c1_graph
    g = pins["graph"]
    assert len(neighs) == g["nnz"] == 99634
    deg = np.diff(cumul)
    assert deg.max() == g["maxDeg"] == 137 and deg.min() == g["minDeg"] == 73
    assert sha(cumul) == g["cumul_sha256"] and sha(neighs) == g["neighs_sha256"]
    # neighbour lists ascending (graphCPU.cpp:374-385)
    for v in range(0, 1000, 37):
        row = neighs[cumul[v]:cumul[v + 1]]
        assert np.all(np.diff(row.astype(np.int64)) > 0)


def test_c1_tape_trajectories_match_reference(port, pins, c1_graph):
    cumul, neighs = c1_graph
    n = 1000
    for tr in pins["tape_trajectories"]:
        nCol, ti = tr["nCol"], tr["tabooIteration"]
        c = port.init_colors(tr["color_seed"], n, nCol)
        assert sha(c) == tr["start_sha256"]
        taboo = np.zeros(n, np.uint32) if ti else None
        for s, step in enumerate(tr["steps"], start=1):
            u = port.tape(tr["tape_seed"], s, n)
            assert port.violation_count(cumul, neighs, c) == step["viol_before"]
            c, ov = port.sweep(cumul, neighs, nCol, 1e-8, c, u, UNIFORM, taboo=taboo, taboo_iter=ti)
            assert ov == step["overflow"]
            assert sha(c) == step["colors_sha256"], (nCol, ti, s)


def test_small_fixture_arrays(port, golden_dir):
    z = np.load(os.path.join(golden_dir, "small_traj.npz"))
    cumul, neighs = z["cumul"], z["neighs"]
    n = len(cumul) - 1
    for tag in ("a", "b", "c", "ovf"):
        nCol, ti = int(z[f"{tag}_nCol"]), int(z[f"{tag}_taboo_iter"])
        c = z[f"{tag}_c0"].copy()
        viol, flags = port.violation_count(cumul, neighs, c, want_flags=True)
        assert viol == int(z[f"{tag}_viol0"]) and np.array_equal(flags, z[f"{tag}_flags0"])
        for v in range(n):
            occ, free = port.occupancy(v, cumul, neighs, c, nCol)
            assert np.array_equal(occ, z[f"{tag}_occ0"][v])
            p = port.fill_p(nCol, 1e-8, occ, int(c[v]))
            assert np.array_equal(p.view(np.uint32), z[f"{tag}_p0"][v].view(np.uint32))  # bit-exact floats
        taboo = np.zeros(n, np.uint32) if ti else None
        for s in range(10):
            assert port.violation_count(cumul, neighs, c) == int(z[f"{tag}_viol_before"][s])
            c, ov = port.sweep(cumul, neighs, nCol, 1e-8, c, z[f"{tag}_tapes"][s], UNIFORM, taboo=taboo, taboo_iter=ti)
            assert ov == int(z[f"{tag}_overflow"][s])
            assert np.array_equal(c, z[f"{tag}_colors"][s]), (tag, s)
            if ti:
                assert np.array_equal(taboo, z[f"{tag}_taboo"][s])
    assert z["ovf_overflow"].sum() > 100  # the overflow rule (idx = nCol-1) is really exercised


def test_conflict_metrics_and_stats(port, pins, c1_graph):
    cumul, neighs = c1_graph
    n, nCol = 1000, 137
    c = port.init_colors(3, n, nCol)
    viol, flags = port.violation_count(cumul, neighs, c, want_flags=True)
    edges = port.conflict_edges(cumul, neighs, c)
    # independent numpy restatement: every conflicting edge once, every violating vertex once
    src = np.repeat(np.arange(n, dtype=np.uint32), np.diff(cumul))
    same = c[src] == c[neighs]
    assert edges == int((same & (src < neighs)).sum()) == int(same.sum()) // 2
    assert viol == len(np.unique(src[same])) == int(flags.sum())
    # split ranges add up (what each rank of a vertex partition computes)
    assert sum(port.conflict_edges(cumul, neighs, c, a, b) for a, b in [(0, 300), (300, 777), (777, n)]) == edges
    # saveStats numbers of the reference's own seed-1234 run (coloringMCMC_CPUutils.cpp:87-101)
    log = pins["saveStats_seed1234"]
    hist = np.array([int(l.split(": ")[1]) for l in log.splitlines() if l.split(": ")[0].isdigit()], np.uint32)
    assert len(hist) == 137 and hist.sum() == 1000
    st = port.color_stats(1000, 137, hist, 0.1)
    assert "Average number of nodes for each color: %g" % st.meanCPU in log
    assert "Variance: %g" % st.varianceCPU in log
    assert "StD: %g" % st.stdCPU in log
    assert st.usedColors == 137


def test_free_running_port_is_proper_and_balanced_like_reference(port, pins, c1_graph):
    """Free-running Philox chains of the port vs the reference's std::default_random_engine chains: same
    colour count, proper colouring, sweeps and class-size StD inside the reference's seed-to-seed spread."""
    cumul, neighs = c1_graph
    n = 1000
    by_ratio = {}
    for r in pins["runs"]:
        if not r["maxIterReached"]:
            by_ratio.setdefault(r["nCol"], []).append(r)
    for nCol, runs in by_ratio.items():
        ref_std = np.array([r["std"] for r in runs])
        ref_sw = np.array([r["sweeps"] for r in runs])
        stds, sws = [], []
        for seed in range(1, 9):
            c0 = port.init_colors(seed, n, nCol)
            c, sweeps, cnt, hit = port.run(cumul, neighs, nCol, 1e-8, c0, seed, UNIFORM)
            assert cnt == 0 and not hit
            assert port.conflict_edges(cumul, neighs, c) == 0
            hist = port.class_sizes(c, nCol)
            assert (hist > 0).sum() == runs[0]["usedColors"]
            stds.append(port.color_stats(n, nCol, hist, 0.1).stdCPU)
            sws.append(sweeps)
        # tolerance: mean of 8 chains within 25% of the reference mean; sweeps within +-3 of the reference range
        assert abs(np.mean(stds) - ref_std.mean()) <= 0.25 * ref_std.mean(), (nCol, stds, ref_std)
        assert ref_sw.min() - 3 <= np.mean(sws) <= ref_sw.max() + 3, (nCol, sws, ref_sw)


def test_dynamic_proposal_restatement_properties(port, c1_graph):
    """DYNAMIC has no CPU twin in the reference (coloringMCMC_balance.cu:79-143 is GPU-only); the restatement is
    checked through its invariants: proper colouring reached, non-violating vertices stay, better balance than UNIFORM."""
    cumul, neighs = c1_graph
    n, nCol = 1000, 68
    std_u, std_d = [], []
    for seed in range(1, 6):
        c0 = port.init_colors(seed, n, nCol)
        cu, _, cntu, _ = port.run(cumul, neighs, nCol, 1e-8, c0, seed, UNIFORM)
        cd, _, cntd, _ = port.run(cumul, neighs, nCol, 1e-8, c0, seed, DYNAMIC)
        assert cntu == 0 and cntd == 0
        std_u.append(port.color_stats(n, nCol, port.class_sizes(cu, nCol), 0.1).stdCPU)
        std_d.append(port.color_stats(n, nCol, port.class_sizes(cd, nCol), 0.1).stdCPU)
    assert np.mean(std_d) <= np.mean(std_u) * 1.05
    # one sweep: a vertex without conflicts keeps its colour unless its draw is in the epsilon tails
    c0 = port.init_colors(1, n, nCol)
    _, flags = port.violation_count(cumul, neighs, c0, want_flags=True)
    u = np.full(n, 0.5, np.float32)
    c1, ov = port.sweep(cumul, neighs, nCol, 1e-8, c0, u, DYNAMIC)
    assert ov == 0 and np.array_equal(c1[flags == 0], c0[flags == 0])
    moved = c1[flags == 1] != c0[flags == 1]
    assert moved.mean() > 0.9


def test_tailcut_repairs_all_conflicts(port, c1_graph):
    cumul, neighs = c1_graph
    n, nCol = 1000, 137
    c0 = port.init_colors(9, n, nCol)
    assert port.conflict_edges(cumul, neighs, c0) > 0
    c, rounds, left = port.tailcut(cumul, neighs, nCol, c0)
    assert left == 0 and rounds >= 1 and port.violation_count(cumul, neighs, c) == 0


def test_edge_cases(port):
    # empty graph, isolated vertices, single colour
    cumul = np.zeros(6, np.uint32)
    neighs = np.zeros(0, np.uint32)
    c = np.array([0, 1, 2, 1, 0], np.uint32)
    assert port.violation_count(cumul, neighs, c) == 0 and port.conflict_edges(cumul, neighs, c) == 0
    u = np.full(5, 0.5, np.float32)
    c1, ov = port.sweep(cumul, neighs, 3, 1e-8, c, u, UNIFORM)
    assert np.array_equal(c1, c) and ov == 0
    # a triangle with 2 colours can never be proper: violating vertices with no free colour stay put (:402-411)
    cumul = np.array([0, 2, 4, 6], np.uint32)
    neighs = np.array([1, 2, 0, 2, 0, 1], np.uint32)
    c = np.array([0, 1, 1], np.uint32)
    c1, _ = port.sweep(cumul, neighs, 2, 1e-8, c, np.full(3, 0.5, np.float32), UNIFORM)
    assert c1.tolist() == [0, 0, 0] or c1[0] == 0  # vertex 0 not violating keeps 0; 1,2 see both colours occupied -> stay
    assert c1[1] == 1 and c1[2] == 1


def test_port_matches_reference_live(port, ref):
    """Where the unmodified reference is built: fresh random tapes/graphs beyond the committed fixtures."""
    g = ref.graph_simulate(300, 0.07, srand=777)
    cumul, neighs = ref.graph_csr(g)
    pc, pn = port.setup_rnd2(300, 0.07, srand=777)
    assert np.array_equal(cumul, pc) and np.array_equal(neighs, pn)
    n = 300
    rng = np.random.default_rng(5)
    for nCol, ti in [(int(np.diff(cumul).max()), 0), (15, 0), (11, 4), (8, 1)]:
        h = ref.mcmc(g, nCol, 1, taboo_iter=ti)
        c = rng.integers(0, nCol, n).astype(np.uint32)
        ref.set_colors(h, c)
        taboo = np.zeros(n, np.uint32) if ti else None
        for s in range(15):
            u = rng.random(n, dtype=np.float32)
            if s % 4 == 3:
                u[rng.integers(0, n, 20)] = np.nextafter(np.float32(1), np.float32(0))
            before, ov = ref.sweep_tape(h, u)
            assert before == port.violation_count(cumul, neighs, c)
            c, ov2 = port.sweep(cumul, neighs, nCol, 1e-8, c, u, UNIFORM, taboo=taboo, taboo_iter=ti)
            assert ov == ov2 and np.array_equal(c, ref.get_colors(h, n))
            if ti:
                assert np.array_equal(taboo, ref.get_taboo(h, n))


def test_port_matches_reference_live_wide_palettes(port, ref):
    """The same live pin for palettes above 512 colours (the wide-palette GPU kernels are compared with the port): the unmodified
    reference CPU colourer takes any nCol (its scratch is nCol-sized per vertex), so port == reference there too -- start from few
    colours so that most vertices conflict and the long CDF walks really run."""
    g = ref.graph_simulate(250, 0.08, srand=4242)
    cumul, neighs = ref.graph_csr(g)
    n = 250
    rng = np.random.default_rng(9)
    for nCol, ti in [(513, 0), (700, 2), (4097, 0)]:
        h = ref.mcmc(g, nCol, 1, taboo_iter=ti)
        c = (rng.integers(0, 6, n) * (nCol // 6)).astype(np.uint32)
        ref.set_colors(h, c)
        taboo = np.zeros(n, np.uint32) if ti else None
        for s in range(6):
            u = rng.random(n, dtype=np.float32)
            if s == 3:
                u[rng.integers(0, n, 20)] = np.nextafter(np.float32(1), np.float32(0))
            before, ov = ref.sweep_tape(h, u)
            assert before == port.violation_count(cumul, neighs, c)
            c, ov2 = port.sweep(cumul, neighs, nCol, 1e-8, c, u, UNIFORM, taboo=taboo, taboo_iter=ti)
            assert ov == ov2 and np.array_equal(c, ref.get_colors(h, n))
