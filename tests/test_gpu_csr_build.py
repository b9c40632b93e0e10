"""GPU-side CSR construction from an edge list (mcmcb200_csr_from_edges) against the restatement of the reference's
Graph::setupImporterNew (graph/graphCPU.cpp:112-170): same arrays, entry for entry."""
import numpy as np
import pytest

from oracle.pyoracle import importer_csr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mc():
    import mcmc_colorer_b200 as m
    return m


def device_to_host(ptr, nelem, dtype="<u4"):
    import torch
    from mcmc_colorer_b200.multigpu import _view
    return _view(ptr, nelem, dtype, "cuda:0").cpu().numpy().astype(np.uint32)


def vectorised_importer(n, src, dst):
    """the same CSR with numpy: stable sort of the doubled edge list (forward entry at 2i, back-edge at 2i+1) by row"""
    keep = src != dst
    s, d = src[keep].astype(np.int64), dst[keep].astype(np.int64)
    rows = np.empty(2 * len(s), np.int64); cols = np.empty(2 * len(s), np.int64)
    rows[0::2], cols[0::2] = s, d
    rows[1::2], cols[1::2] = d, s
    order = np.argsort(rows, kind="stable")
    cumul = np.zeros(n + 1, np.int64)
    np.add.at(cumul, rows + 1, 1)
    return np.cumsum(cumul).astype(np.uint32), cols[order].astype(np.uint32)


@pytest.mark.parametrize("n,m,seed", [(1000, 20000, 1), (37, 400, 2), (5, 0, 3), (3000, 3, 4)])
def test_small_edge_lists_match_the_reference_loops(mc, n, m, seed):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n, m, dtype=np.int64).astype(np.uint32)
    dst = rng.integers(0, n, m, dtype=np.int64).astype(np.uint32)
    if m > 10:
        dst[:5] = src[:5]                               # self-loops
        src[5:8], dst[5:8] = src[8], dst[8]             # duplicates
    want_c, want_n = importer_csr(n, src.tolist(), dst.tolist())
    vc, vn = vectorised_importer(n, src, dst)
    assert np.array_equal(vc, want_c) and np.array_equal(vn, want_n)
    csr = mc.DeviceCsr(n, src, dst)
    assert csr.nnz == len(want_n)
    assert np.array_equal(device_to_host(csr.rowptr.value, n + 1), want_c)
    if csr.nnz:
        assert np.array_equal(device_to_host(csr.neighs.value, csr.nnz), want_n)
    csr.close()


def test_large_edge_list_and_sweep_on_the_device_csr(mc):
    from oracle.pyoracle import Port
    P = Port()
    n, m = 200_000, 2_000_000
    rng = np.random.default_rng(7)
    src = rng.integers(0, n, m, dtype=np.int64).astype(np.uint32)
    dst = rng.integers(0, n, m, dtype=np.int64).astype(np.uint32)
    vc, vn = vectorised_importer(n, src, dst)
    csr = mc.DeviceCsr(n, src, dst)
    assert np.array_equal(device_to_host(csr.rowptr.value, n + 1), vc)
    assert np.array_equal(device_to_host(csr.neighs.value, csr.nnz), vn)
    # a chain on the adopted device CSR behaves like one created from the host arrays (duplicate edges included)
    nCol = int(np.diff(vc.astype(np.int64)).max())
    prm = mc.ColoringMCMCParams(nCol=nCol, seed=5)
    a = mc.Chain(params=prm, device=0, n_global=n, v_begin=0, v_end=n, device_csr=csr.as_tuple())
    b = mc.Chain(vc, vn, prm, device=0)
    for ch in (a, b):
        ch.init_colors(None)
        ch.sweep(3)
    sa, sb = a.status(), b.status()
    assert (sa.sweep, sa.conflictEdges, sa.violatingVertices) == (sb.sweep, sb.conflictEdges, sb.violatingVertices)
    assert np.array_equal(a.get_colors(), b.get_colors())
    c0 = P.init_colors(5, n, nCol)
    assert (P.conflict_edges(vc, vn, c0), P.violation_count(vc, vn, c0)) == a.conflicts_of(c0)
    a.close(); b.close(); csr.close()


def test_out_of_range_endpoint_is_rejected(mc):
    from mcmc_colorer_b200 import capi
    with pytest.raises(mc.McmcError) as e:
        mc.DeviceCsr(10, np.array([1, 2, 30], np.uint32), np.array([2, 3, 4], np.uint32))
    assert e.value.code == capi.EINVAL
