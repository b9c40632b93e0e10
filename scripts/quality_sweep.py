#!/usr/bin/env python
"""BASELINE config 5: numColRatio sweep (0.5 - 2.0) -- colour count, balance (class-size StD, BalancingIndex) and sweeps
of the B200 MCMC sampler (both proposals) against the reference CPU MCMC (oracle/_ref, unmodified reference code, its own
std::default_random_engine chains) and the Luby cross-check colourer, on the same Erdos-Renyi graph.

  python scripts/quality_sweep.py [--n 1000000] [--seeds 3] [--big 10000000]

The CPU reference is run at --n (a sweep costs ~0.3 s per million vertices on one core); the GPU rows are repeated at
--big (config 5's n = 10 M) where the CPU path would take hours.  Writes gpurun_out/quality_c5.{json,md}."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcmc_colorer_b200 as mc                                    # noqa: E402
from mcmc_colorer_b200.graphgen import er_graph_numpy            # noqa: E402
from oracle.pyoracle import Port, Ref                            # noqa: E402


def gpu_chain(cumul, neighs, nCol, seed, proposal, prob, max_rip=250):
    prm = mc.ColoringMCMCParams(nCol=nCol, seed=seed, proposal=proposal, maxRip=max_rip,
                                convergence=mc.CONVERGE_EDGES if proposal == mc.PROPOSAL_DYNAMIC else mc.CONVERGE_VERTICES)
    ch = mc.Chain(cumul, neighs, prm, device=0)
    ch.init_colors(None)
    t0 = time.perf_counter()
    st = ch.status()
    while not st.converged and st.sweep < max_rip:
        ch.sweep(5)
        st = ch.status()
    dt = time.perf_counter() - t0
    left = int(st.conflictEdges)
    if left:                                                      # --tailcut style repair of what the cap left over
        ch.tailcut()
        st = ch.status()
    hist = ch.class_sizes()
    s = mc.color_stats(hist, len(cumul) - 1, prob)
    ch.close()
    return dict(sweeps=int(st.sweep), seconds=dt, conflicts_before_repair=left, conflicts=int(st.conflictEdges), used=s["used"],
                std=s["std"], balancingIndex=s["balancingIndex"])


def cpu_chain(R, P, g, n, nCol, seed, prob):
    h = R.mcmc(g, nCol, seed)
    t0 = time.perf_counter()
    viol, sweeps, hit = R.run_native(h)
    dt = time.perf_counter() - t0
    col = R.get_colors(h, n)
    R.L.ref_mcmc_free(h)
    hist = P.class_sizes(col, nCol)
    s = mc.color_stats(hist, n, prob)
    return dict(sweeps=int(sweeps), seconds=dt, violating_vertices=int(viol), max_iter=bool(hit), used=s["used"], std=s["std"],
                balancingIndex=s["balancingIndex"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--big", type=int, default=10_000_000)
    ap.add_argument("--seeds", type=int, default=3)
    ap.add_argument("--deg", type=float, default=16.0)
    args = ap.parse_args()
    P = Port()
    R = Ref() if Ref.available() else None
    out = {"config": vars(args), "rows": []}
    for n, with_cpu in ((args.n, True), (args.big, False)):
        cumul, neighs = er_graph_numpy(n, args.deg, seed=42)
        maxdeg = int(np.diff(cumul.astype(np.int64)).max())
        prob = len(neighs) / float(n) / float(n)
        g = R.graph_from_csr(cumul, neighs, prob) if (R and with_cpu) else None
        lub, lcol, lrounds = mc.luby_color(cumul, neighs, seed=1, device=0)
        lhist = np.bincount(lub - 1, minlength=lcol)
        ls = mc.color_stats(lhist, n, prob)
        out["rows"].append(dict(n=n, algo="luby_gpu", ratio=None, nCol=lcol, used=lcol, std=ls["std"], balancingIndex=ls["balancingIndex"],
                                sweeps=lrounds))
        for ratio in (0.5, 0.75, 1.0, 1.5, 2.0):
            nCol = mc.Graph.default_ncol(maxdeg, ratio)
            for algo, fn in (("mcmc_gpu_dynamic", lambda s: gpu_chain(cumul, neighs, nCol, s, mc.PROPOSAL_DYNAMIC, prob)),
                             ("mcmc_gpu_uniform", lambda s: gpu_chain(cumul, neighs, nCol, s, mc.PROPOSAL_UNIFORM, prob)),
                             ("mcmc_cpu_reference", (lambda s: cpu_chain(R, P, g, n, nCol, s, prob)) if g else None)):
                if fn is None:
                    continue
                runs = [fn(seed) for seed in range(1, args.seeds + 1)]
                row = dict(n=n, algo=algo, ratio=ratio, nCol=nCol)
                for k in runs[0]:
                    vals = [r[k] for r in runs]
                    row[k] = float(np.mean(vals)) if not isinstance(vals[0], bool) else any(vals)
                row["std_spread"] = [min(r["std"] for r in runs), max(r["std"] for r in runs)]
                out["rows"].append(row)
                print(json.dumps(row), flush=True)
        if g:
            R.L.ref_graph_free(g)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "quality_c5.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "gpurun_out", "quality_c5.md"), "w") as f:
        f.write("| n | algorithm | numColRatio | nCol | used colours | class-size StD (mean, [min,max] over seeds) | BalancingIndex | sweeps | conflicts left |\n|---|---|---|---|---|---|---|---|---|\n")
        for r in out["rows"]:
            f.write("| %d | %s | %s | %d | %.1f | %.2f %s | %.3f | %.1f | %s |\n" % (
                r["n"], r["algo"], r["ratio"], r["nCol"], r["used"], r["std"], r.get("std_spread", ""), r["balancingIndex"], r["sweeps"],
                r.get("conflicts", r.get("violating_vertices", 0))))


if __name__ == "__main__":
    main()
