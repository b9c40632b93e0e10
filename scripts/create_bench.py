"""Times mcmcb200_create (layout build included) on a device-resident graph, several times in one process (the first create also pays
CUDA module loading).   python scripts/create_bench.py [workload=c3] [repeats=4]   (MCMCB200_LIB selects a library variant)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import mcmc_colorer_b200 as mc

workload = sys.argv[1] if len(sys.argv) > 1 else "c3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n, deg, desc = bench.WORKLOADS[workload]
rowptr, neighs, nnz, max_deg = bench.gen_graph_device(n, deg, "cuda:0", workload)
nCol = bench.palette_for(workload, max_deg)
prm = mc.ColoringMCMCParams(nCol=nCol, proposal=mc.PROPOSAL_UNIFORM, convergence=mc.CONVERGE_VERTICES, seed=1)
out = []
for es in [0] * reps + [8, 8]:
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ch = mc.Chain(params=prm, device=0, n_global=n, v_begin=0, v_end=n, device_csr=(rowptr.data_ptr(), neighs.data_ptr(), nnz), expected_sweeps=es)
    ch.synchronize()
    out.append({"expected_sweeps": es, "kernel_mode": ch.kernel_mode(), "create_ms": round(1e3 * (time.perf_counter() - t0), 1), "launches": ch.launch_count()})
    ch.close()
print(json.dumps({"workload": workload, "lib": os.environ.get("MCMCB200_LIB", "").split("/")[-1], "creates": out}))
