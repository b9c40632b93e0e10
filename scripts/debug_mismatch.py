import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mcmc_colorer_b200 as mc
from mcmc_colorer_b200.graphgen import er_graph_numpy
from oracle.pyoracle import Port
P = Port()
n = 200_001
cumul, neighs = er_graph_numpy(n, 16, seed=3)
for kernel, flag in (("direct", mc.FLAG_FORCE_DIRECT), ("blocked", mc.FLAG_FORCE_BLOCKED)):
  for nCol in (300, 37):
    prm = mc.ColoringMCMCParams(nCol=nCol, proposal=0, convergence=0, seed=9)
    ch = mc.Chain(cumul, neighs, prm, device=0, flags=flag)
    ch.init_colors(None)
    c = P.init_colors(9, n, nCol)
    for s in range(1, 8):
        u = P.tape(9, s, n, 0)
        ch.sweep(1)
        c2, ov = P.sweep(cumul, neighs, nCol, 1e-8, c, u, 0)
        got = ch.get_colors()
        bad = np.flatnonzero(got != c2)
        print(kernel, nCol, "sweep", s, "mismatches", len(bad), "port overflows", ov, "status sweep", ch.status().sweep)
        for v in bad[:5]:
            occ, free = P.occupancy(int(v), cumul, neighs, c, nCol)
            print("   v", v, "old", c[v], "gpu", got[v], "port", c2[v], "u", repr(float(u[v])), "viol", bool(occ[c[v]]), "free", free,
                  "S[own]~", c[v] * 1e-8)
        if len(bad):
            break
        c = c2
    ch.close()
