"""torchrun --nproc-per-node N scripts/multigpu_check.py : N-GPU trajectories must equal the single-process oracle
(and therefore the 1-GPU run) bit for bit -- counters, class sizes and colours."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mcmc_colorer_b200 import ColoringMCMCParams
from mcmc_colorer_b200.graphgen import er_graph_numpy
from mcmc_colorer_b200.multigpu import DistributedSweeper, GpuEngine, partition, partition_by_nnz
from oracle.pyoracle import Port

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = f"cuda:{lr}"
dist.init_process_group("nccl", device_id=torch.device(dev))
P = Port()
ok = True
# (n, mean degree, proposal, fused exchange (peer colour stores + device-side counter all-reduce), stage bytes: 32768 = the
#  large-partition configuration with pass A || pass B, nnz-balanced cut points instead of equal vertex counts)
for n, deg, proposal, p2p, cap, by_nnz in [(50_001, 12, 0, False, 0, False), (50_001, 12, 1, False, 0, True), (300_000, 20, 0, False, 0, False),
                                           (300_000, 20, 1, True, 0, True), (1_000_003, 16, 0, True, 0, False), (1_000_003, 16, 1, True, 32768, False)]:
    cumul, neighs = er_graph_numpy(n, deg, seed=5)
    nCol = int(np.diff(cumul.astype(np.int64)).max())
    parts, chunk = partition(n, world)
    if by_nnz:
        parts = partition_by_nnz(cumul.astype(np.int64), world)
    vb, ve = parts[rank]
    e0, e1 = int(cumul[vb]), int(cumul[ve])
    rp = torch.from_numpy((cumul[vb:ve + 1].astype(np.int64) - e0).astype(np.int32)).to(dev)
    nb = torch.zeros(e1 - e0 + 16, dtype=torch.int32, device=dev)
    nb[: e1 - e0] = torch.from_numpy(neighs[e0:e1].astype(np.int32)).to(dev)
    prm = ColoringMCMCParams(nCol=nCol, proposal=proposal, convergence=proposal, seed=11)
    eng = GpuEngine(rp, nb, e1 - e0, n, vb, ve, prm, lr, stage_cap_bytes=cap)
    if rank == 0:
        print(f"n={n} proposal={proposal}: sweep mode {eng.chain.kernel_mode()}")
    if p2p:
        got_p2p = eng.enable_p2p(rank, world)
        if rank == 0:
            print(f"n={n}: fused exchange {'ON' if got_p2p else 'not available -> NCCL all-gather + all-reduce'}")
    sw = DistributedSweeper(eng, rank, world, chunk, parts=parts)
    eng.init_colors(None)
    c = P.init_colors(11, n, nCol)
    for s in range(5):
        st = sw.status()
        want = (P.conflict_edges(cumul, neighs, c), P.violation_count(cumul, neighs, c))
        if (st.conflictEdges, st.violatingVertices) != want or st.sweep != s:
            ok = False; print(f"rank {rank}: counters differ at sweep {s}: {(st.conflictEdges, st.violatingVertices)} vs {want}")
        sw.sweep(1)
        c, _ = P.sweep(cumul, neighs, nCol, 1e-8, c, P.tape(11, s + 1, n, proposal), proposal)
        got = eng.colors_host()
        if not np.array_equal(got, c):
            ok = False; print(f"rank {rank}: colours differ at sweep {s + 1}: {np.flatnonzero(got != c)[:8]}")
    sw.sweep(3)                                         # a batch: with the fused exchange these are 6 launches per rank and nothing else
    for s in range(5, 8):
        c, _ = P.sweep(cumul, neighs, nCol, 1e-8, c, P.tape(11, s + 1, n, proposal), proposal)
    if not np.array_equal(eng.colors_host(), c):
        ok = False; print(f"rank {rank}: colours differ after the 3-sweep batch")
    st = sw.status()
    if (st.conflictEdges, st.violatingVertices, st.sweep) != (P.conflict_edges(cumul, neighs, c), P.violation_count(cumul, neighs, c), 8):
        ok = False; print(f"rank {rank}: counters differ after the batch: {(st.conflictEdges, st.violatingVertices, st.sweep)}")
    if not np.array_equal(eng.chain.class_sizes().astype(np.uint32), P.class_sizes(c, nCol)):
        ok = False; print(f"rank {rank}: class sizes differ")
    # sliced host interface: every rank uploads only its own colours, the slices meet on the device
    own = np.ascontiguousarray(c[vb:ve].astype(np.uint32))
    eng.init_colors_slice(own.ctypes.data, sw)
    st = sw.status()
    if (st.conflictEdges, st.violatingVertices, st.sweep) != (P.conflict_edges(cumul, neighs, c), P.violation_count(cumul, neighs, c), 0):
        ok = False; print(f"rank {rank}: sliced init counters differ")
    if not np.array_equal(eng.colors_host(), c):
        ok = False; print(f"rank {rank}: sliced init colours differ")
    sw.sweep(1)
    c2, _ = P.sweep(cumul, neighs, nCol, 1e-8, c, P.tape(11, 1, n, proposal), proposal)
    back = np.zeros(max(ve - vb, 1), dtype=np.uint32)
    eng.chain.get_colors_slice_ptr(back.ctypes.data)
    if not np.array_equal(back[: ve - vb], c2[vb:ve]):
        ok = False; print(f"rank {rank}: sliced download differs")
    # the same two transfers in the device's narrow colour format
    eb = eng.chain.color_bytes()
    ndt = np.uint8 if eb == 1 else np.uint16
    own_n = np.ascontiguousarray(c[vb:ve].astype(ndt)) if ve > vb else np.zeros(1, ndt)
    eng.init_colors_slice(own_n.ctypes.data, sw, eb)
    st = sw.status()
    if (st.conflictEdges, st.violatingVertices, st.sweep) != (P.conflict_edges(cumul, neighs, c), P.violation_count(cumul, neighs, c), 0):
        ok = False; print(f"rank {rank}: narrow sliced init counters differ")
    sw.sweep(1)
    back_n = np.zeros(max(ve - vb, 1), dtype=ndt)
    eng.chain.get_colors_slice_narrow_ptr(back_n.ctypes.data, eb)
    if not np.array_equal(back_n[: ve - vb].astype(np.uint32), c2[vb:ve]):
        ok = False; print(f"rank {rank}: narrow sliced download differs")
    eng.chain.close()
# ---- distributed tail cutting: chain with params.tailcut stops on every rank at <= z violating vertices (global count), the
#      ranks repair their own violators in synchronised rounds; == the oracle's chain + its sequential tail cut ----
for n, deg, p2p, slack in [(300_000, 16, False, 14), (1_000_003, 16, True, 13)]:
    cumul, neighs = er_graph_numpy(n, deg, seed=21)
    nCol = int(np.diff(cumul.astype(np.int64)).max()) - slack       # a tight palette: the chain needs a few sweeps
    z = max(50, n // 2000)
    parts, chunk = partition(n, world)
    vb, ve = parts[rank]
    e0, e1 = int(cumul[vb]), int(cumul[ve])
    rp = torch.from_numpy((cumul[vb:ve + 1].astype(np.int64) - e0).astype(np.int32)).to(dev)
    nb = torch.zeros(e1 - e0 + 16, dtype=torch.int32, device=dev)
    nb[: e1 - e0] = torch.from_numpy(neighs[e0:e1].astype(np.int32)).to(dev)
    prm = ColoringMCMCParams(nCol=nCol, proposal=0, convergence=0, seed=5, tailcut=True)
    eng = GpuEngine(rp, nb, e1 - e0, n, vb, ve, prm, lr, early_stop=True)       # the chain stops on the device at the threshold
    if p2p:
        eng.enable_p2p(rank, world)
    sw = DistributedSweeper(eng, rank, world, chunk, parts=parts)
    eng.init_colors(None)
    c0 = P.init_colors(5, n, nCol)
    want, sweeps, cnt, hit = P.run(cumul, neighs, nCol, 1e-8, c0, 5, 0, z=z)
    for _ in range(sweeps + 3):                                     # sweeps past the threshold are no-ops on every rank
        sw.sweep(1)
    st = sw.status()
    if st.sweep != sweeps or st.violatingVertices != cnt or not np.array_equal(eng.colors_host(), want):
        ok = False; print(f"rank {rank}: tail-cut chain differs before the repair: sweep {st.sweep} vs {sweeps}, viol {st.violatingVertices} vs {cnt}")
    fixed, _, left = P.tailcut(cumul, neighs, nCol, want)
    passes = sw.tailcut(64)
    st = sw.status()
    got = eng.colors_host()
    if not np.array_equal(got, fixed) or st.conflictEdges != left or st.violatingVertices != P.violation_count(cumul, neighs, fixed):
        ok = False; print(f"rank {rank}: distributed tail cut differs: {np.flatnonzero(got != fixed)[:8]}, left {st.conflictEdges} vs {left}")
    if not np.array_equal(eng.chain.class_sizes().astype(np.uint32), P.class_sizes(fixed, nCol)):
        ok = False; print(f"rank {rank}: class sizes differ after the distributed tail cut")
    if rank == 0:
        print(f"n={n}: distributed tail cut after {sweeps} sweeps ({cnt} violating <= z={z}): {passes} pass(es), {left} conflicts left, fused={eng.p2p}")
    eng.chain.close()
flag = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(flag)
if rank == 0:
    print("MULTIGPU_CHECK", "OK" if flag.item() == 0 else "FAILED", "world", world)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 0 else 1)
