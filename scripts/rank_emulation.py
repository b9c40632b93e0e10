"""One rank of an N-GPU run on ONE GPU (no peers, no exchange): the owned rows of rank r of BASELINE config 3 with the full
replicated colour array, i.e. the per-rank kernel time an N-GPU sweep is built on.  Comparing it with the N-GPU step time
(bench.py --gpus N) apportions the rest to the colour exchange, the cross-rank reduction and rank skew.
    python scripts/rank_emulation.py [world=8] [rank=0] [workload=c3] [itemBits,itemBits,...] [stageCapBytes]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import mcmc_colorer_b200 as mc
from mcmc_colorer_b200.multigpu import partition, partition_by_nnz, slice_csr_torch

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
workload = sys.argv[3] if len(sys.argv) > 3 else "c3"
n, deg, desc = bench.WORKLOADS[workload]
rowptr, neighs, nnz, max_deg = bench.gen_graph_device(n, deg, "cuda:0", workload)
nCol = bench.palette_for(workload, max_deg)
parts, chunk = partition(n, world)
if workload.startswith("c4"):
    parts = partition_by_nnz(rowptr.to(torch.int64), world)
vb, ve = parts[rank]
rp, nb, nnz_local = slice_csr_torch(rowptr.to(torch.int64), neighs, vb, ve)
del rowptr, neighs
torch.cuda.empty_cache()
item_bits = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
stage_cap = int(sys.argv[5]) if len(sys.argv) > 5 else 0
prm = mc.ColoringMCMCParams(nCol=nCol, proposal=mc.PROPOSAL_UNIFORM, convergence=mc.CONVERGE_VERTICES, seed=bench.CHAIN_SEED)
# (a lone partition: its counters are local sums, the colours of the other ranks' vertices simply never change)
for ib in item_bits:
    ch = mc.Chain(params=prm, device=0, flags=mc.FLAG_NO_EARLY_STOP, n_global=n, v_begin=vb, v_end=ve, device_csr=(rp.data_ptr(), nb.data_ptr(), nnz_local),
                  item_bits=ib, stage_cap_bytes=stage_cap)
    ms = []
    for i in range(8):
        ch.init_colors(None); torch.cuda.synchronize(); ch.sweep(1); t = ch.last_sweep_ms()
        if i >= 3:
            ms.append(t)
    ch.init_colors(None); ch.sweep(10); chain = ch.last_sweep_ms() / 10
    print(json.dumps({"workload": workload, "world": world, "rank": rank, "owned": [vb, ve], "nnz_local": int(nnz_local), "kernel_mode": ch.kernel_mode(),
                      "item_bits": ib, "stage_cap_bytes": stage_cap, "first_sweep_ms": float(np.mean(ms)), "chain_ms_per_sweep": chain}), flush=True)
    ch.close()
