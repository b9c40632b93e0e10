#!/usr/bin/env python
"""Statistical pins for BASELINE config 5 (numColRatio sweep) from the UNMODIFIED reference CPU colourer.

    make -C oracle ref && python tests/golden/make_stat_pins.py          (build container; ~10 minutes of CPU)

Graph: Erdos-Renyi n = 1 000 000, mean degree 16 (mcmc_colorer_b200.graphgen.er_graph_numpy, seed 42) -- config 5's law at a
size the single-threaded reference finishes in minutes.  For every ratio in {0.5, 0.75, 1.0, 1.5, 2.0} and seeds 1..5 the
reference's ColoringMCMC_CPU (std::default_random_engine(seed), its own run loop without the non-terminating tail-cut,
oracle/ref_harness.cpp run_native) runs to convergence; recorded per chain: sweeps, used colours, class-size StD
(saveStats formula, coloringMCMC_CPUutils.cpp:87-101) and BalancingIndex (coloringMCMC_prints.cu:146-167).
Writes tests/golden/c5_stat_pins.json.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Port, Ref  # noqa: E402
from mcmc_colorer_b200.graphgen import er_graph_numpy  # noqa: E402

N, DEG, GSEED = 1_000_000, 16, 42
RATIOS = [0.5, 0.75, 1.0, 1.5, 2.0]
SEEDS = [1, 2, 3, 4, 5]


def main():
    P, R = Port(), Ref()
    cumul, neighs = er_graph_numpy(N, DEG, seed=GSEED)
    max_deg = int(np.diff(cumul.astype(np.int64)).max())
    prob = float(len(neighs)) / N / N
    g = R.graph_from_csr(cumul, neighs, prob)
    pins = dict(n=N, deg=DEG, graph_seed=GSEED, nnz=int(len(neighs)), maxDeg=max_deg, prob=prob, chains=[])
    for ratio in RATIOS:
        nCol = int(np.float32(max_deg) * (np.float32(1.0) / np.float32(ratio)))    # main.cu:53,162
        for seed in SEEDS:
            t0 = time.time()
            h = R.mcmc(g, nCol, seed, ratio=float(np.float32(1.0) / np.float32(ratio)))
            viol, sweeps, hit = R.run_native(h)
            col = R.get_colors(h, N)
            R.L.ref_mcmc_free(h)
            hist = P.class_sizes(col, nCol)
            st = P.color_stats(N, nCol, hist, prob)
            rec = dict(ratio=ratio, nCol=nCol, seed=seed, sweeps=int(sweeps), violations=int(viol), maxIterReached=bool(hit),
                       usedColors=int(st.usedColors), std=float(st.stdCPU), balancingIndex=float(st.balancingIndex))
            pins["chains"].append(rec)
            print(rec, "%.0f s" % (time.time() - t0), flush=True)
    R.L.ref_graph_free(g)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "c5_stat_pins.json"), "w") as f:
        json.dump(pins, f, indent=1)


if __name__ == "__main__":
    main()
