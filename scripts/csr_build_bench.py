"""gpurun -- python scripts/csr_build_bench.py [m_edges] : time mcmcb200_csr_from_edges on a random device-resident edge list."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mcmc_colorer_b200 as mc

m = int(float(sys.argv[1])) if len(sys.argv) > 1 else 800_000_000
n = max(2, m // 8)
g = torch.Generator(device="cuda"); g.manual_seed(3)
src = torch.randint(0, n, (m,), generator=g, device="cuda", dtype=torch.int32)
dst = torch.randint(0, n, (m,), generator=g, device="cuda", dtype=torch.int32)
torch.cuda.synchronize()
out = []
for _ in range(3):
    t0 = time.perf_counter()
    csr = mc.DeviceCsr(n, src.data_ptr(), dst.data_ptr(), m=m)
    torch.cuda.synchronize()
    out.append(time.perf_counter() - t0)
    nnz = csr.nnz
    csr.close()
print(json.dumps({"what": "mcmcb200_csr_from_edges (device edge list -> CSR, self-loops dropped, back-edges added, file order)",
                  "n": n, "m_edges": m, "nnz_directed": nnz, "seconds": out, "edges_per_sec": m / min(out)}))
