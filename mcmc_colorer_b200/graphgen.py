"""Synthetic graph generators (BASELINE.json configs 2-5).

The reference's --simulate generator (Graph::setupRnd2, graph/graphCPU.cpp:290-404) needs n(n+1)/2 libc rand()
draws and as many bits of memory -- 62 GB at n = 1 M -- so it only exists for small n (host/graph.cpp keeps a
rand()-exact twin for drop-in reproducibility).  For the large configs the same G(n, p) family is sampled in
O(E): m undirected pairs are drawn uniformly, canonicalised, de-duplicated, symmetrised and sorted, which gives a
simple undirected graph with sorted neighbour lists, no self loops (the reference drops them, graphCPU.cpp:126,334)
and mean degree ~ avg_deg.  Two implementations of the same recipe: numpy (CPU tests) and torch on the GPU
(bench-sized graphs are generated where they are used; they never cross PCIe).
"""
import numpy as np


def _csr_from_undirected_numpy(lo, hi, n):
    key = np.unique(lo * n + hi)
    lo, hi = key // n, key % n
    both = np.sort(np.concatenate([lo * n + hi, hi * n + lo]))
    src, dst = both // n, both % n
    cumul = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=cumul[1:])
    return cumul.astype(np.uint32), dst.astype(np.uint32)


def er_graph_numpy(n, avg_deg, seed):
    """Returns (cumulDegs uint32[n+1], neighs uint32[nnz]) of a G(n, p=avg_deg/n)-like simple graph."""
    rng = np.random.default_rng(seed)
    m = int(round(n * avg_deg / 2.0))
    a = rng.integers(0, n, size=m, dtype=np.int64)
    b = rng.integers(0, n, size=m, dtype=np.int64)
    keep = a != b
    return _csr_from_undirected_numpy(np.minimum(a, b)[keep], np.maximum(a, b)[keep], n)


def rmat_graph_numpy(scale, edge_factor, seed, a=0.57, b=0.19, c=0.19, n_keep=None):
    """R-MAT (Chakrabarti et al.) with the Graph500 parameters, symmetrised, de-duplicated, self loops dropped,
    trimmed to the first n_keep vertices (BASELINE config 4)."""
    rng = np.random.default_rng(seed)
    n = 1 << scale
    m = edge_factor * n
    src = np.zeros(m, np.int64)
    dst = np.zeros(m, np.int64)
    ab, abc = a + b, a + b + c
    for _ in range(scale):
        r = rng.random(m)
        src = (src << 1) | (r >= ab)
        dst = (dst << 1) | (((r >= a) & (r < ab)) | (r >= abc))
    if n_keep is not None:
        keep = (src < n_keep) & (dst < n_keep)
        src, dst, n = src[keep], dst[keep], n_keep
    keep = src != dst
    return _csr_from_undirected_numpy(np.minimum(src, dst)[keep], np.maximum(src, dst)[keep], n)


def _csr_from_undirected_torch(lo, hi, n):
    import torch
    key = torch.unique(lo * n + hi)            # sorted, de-duplicated undirected pairs
    del lo, hi
    lo, hi = torch.div(key, n, rounding_mode="floor"), key % n
    del key
    both = torch.cat([lo * n + hi, hi * n + lo])
    del lo, hi
    both, _ = torch.sort(both)
    src = torch.div(both, n, rounding_mode="floor")
    neighs = (both % n).to(torch.int32)
    del both
    deg = torch.bincount(src, minlength=n)
    del src
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=neighs.device)
    torch.cumsum(deg, 0, out=rowptr[1:])
    return rowptr, neighs


def _finish_torch(rowptr, neighs):
    import torch
    nnz = int(rowptr[-1].item())
    max_deg = int((rowptr[1:] - rowptr[:-1]).max().item())
    pad = (-nnz) % 8 + 8                        # the sweep's 256-bit loads may touch up to 7 ids past the end
    neighs = torch.cat([neighs, torch.zeros(pad, dtype=torch.int32, device=neighs.device)])
    return rowptr, neighs, nnz, max_deg


def er_graph_torch(n, avg_deg, seed, device="cuda", chunk=1 << 27):
    """Same recipe on the GPU.  Returns (rowptr int64[n+1], neighs int32[nnz padded], nnz, maxDeg).
    Deterministic for a given (n, avg_deg, seed) and torch build."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    m = int(round(n * avg_deg / 2.0))
    los, his = [], []
    left = m
    while left > 0:
        k = min(left, chunk)
        a = torch.randint(0, n, (k,), generator=g, device=device, dtype=torch.int64)
        b = torch.randint(0, n, (k,), generator=g, device=device, dtype=torch.int64)
        keep = a != b
        los.append(torch.minimum(a, b)[keep])
        his.append(torch.maximum(a, b)[keep])
        left -= k
    lo, hi = torch.cat(los), torch.cat(his)
    del los, his
    return _finish_torch(*_csr_from_undirected_torch(lo, hi, n))


def rmat_graph_torch(scale, edge_factor, seed, n_keep=None, a=0.57, b=0.19, c=0.19, device="cuda"):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n = 1 << scale
    m = edge_factor * n
    src = torch.zeros(m, dtype=torch.int64, device=device)
    dst = torch.zeros(m, dtype=torch.int64, device=device)
    ab, abc = a + b, a + b + c
    for _ in range(scale):
        r = torch.rand(m, generator=g, device=device)
        src = (src << 1) | (r >= ab).to(torch.int64)
        dst = (dst << 1) | (((r >= a) & (r < ab)) | (r >= abc)).to(torch.int64)
    if n_keep is not None:
        keep = (src < n_keep) & (dst < n_keep)
        src, dst, n = src[keep], dst[keep], n_keep
    keep = src != dst
    lo, hi = torch.minimum(src, dst)[keep], torch.maximum(src, dst)[keep]
    del src, dst
    return _finish_torch(*_csr_from_undirected_torch(lo, hi, n))
