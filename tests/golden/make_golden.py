#!/usr/bin/env python
"""Generate the committed golden fixtures from the UNMODIFIED reference CPU colourer.

Run in the build container (needs /root/reference -> oracle/_ref/libmcmc_ref.so):

    make -C oracle ref && python tests/golden/make_golden.py

Writes
  tests/golden/c1_pins.json      hashes / scalars for BASELINE config 1 (n=1000, p=0.1, setupRnd2 + libc rand())
  tests/golden/small_traj.npz    full arrays for a 200-vertex graph: CSR, start colouring, draw tapes and the
                                 colouring after every sweep (UNIFORM proposal, taboo 0 and 3, and an
                                 all-overflow tape), occupancy rows, violation flags, p vectors.

Every array in the fixtures is produced by the reference's own code (ColoringMCMC_CPU public methods driven by
oracle/ref_harness.cpp); the only non-reference inputs are the start colourings and the draw tapes (Philox).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Port, Ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    P, R = Port(), Ref()
    pins = {}

    # ---- C1 graph: Graph(1000, 0.1f, .) via setupRnd2, libc rand() in its initial state (srand(1)) ----
    g = R.graph_simulate(1000, 0.1, srand=1)
    info = R.graph_info(g)
    cumul, neighs = R.graph_csr(g)
    pins["graph"] = dict(info, cumul_sha256=sha(cumul), neighs_sha256=sha(neighs))
    n = info["n"]

    # ---- free-running reference chains (std::default_random_engine(seed)), SURVEY 8c pins ----
    runs = []
    for seed, ratio in [(1234, 1.0), (1, 1.0), (2, 1.0), (3, 1.0), (1, 1.5), (2, 1.5), (3, 1.5), (1, 2.0), (2, 2.0),
                        (3, 2.0), (1, 3.0), (2, 3.0), (3, 3.0), (1, 4.0)]:
        nCol = int(np.float32(info["maxDeg"]) * np.float32(np.float32(1.0) / np.float32(ratio)))  # main.cu:53,162
        # chains that overflow the CDF walk call libc rand() in the reference (:517-520) and are only
        # reproducible together with the libc state; record how many such events each chain saw
        hs = R.mcmc(g, nCol, seed)
        overflows = len(R.scan_overflows(hs))
        R.L.ref_mcmc_free(hs)
        h = R.mcmc(g, nCol, seed, ratio=float(np.float32(1.0) / np.float32(ratio)))
        viol, sweeps, hit = R.run_native(h)
        col = R.get_colors(h, n)
        hist = P.class_sizes(col, nCol)
        st = P.color_stats(n, nCol, hist, 0.1)
        path = "/tmp/_golden_colors.txt"
        R.L.ref_mcmc_save_colors(h, path.encode())
        md5 = hashlib.md5(open(path, "rb").read()).hexdigest()
        runs.append(dict(seed=seed, ratio=ratio, nCol=nCol, sweeps=sweeps, violations=viol, maxIterReached=hit, overflows=overflows,
                         usedColors=int((hist > 0).sum()), std=float(np.float32(st.stdCPU)), colors_md5=md5,
                         colors_sha256=sha(col)))
        R.L.ref_mcmc_free(h)
    pins["runs"] = runs

    # the reference's own run() + saveStats on the headline pin (seed 1234)
    h = R.mcmc(g, info["maxDeg"], 1234)
    R.L.ref_mcmc_run(h)
    R.L.ref_mcmc_save_stats(h, 0, 0.5, b"/tmp/_golden_stats.log")
    pins["saveStats_seed1234"] = open("/tmp/_golden_stats.log").read()
    R.L.ref_mcmc_free(h)

    # ---- tape-replay trajectories on C1 (hash per sweep) ----
    trajs = []
    for nCol, taboo_iter, cseed, tseed in [(137, 0, 7, 99), (68, 0, 8, 100), (45, 2, 9, 101), (34, 0, 10, 102)]:
        h = R.mcmc(g, nCol, 1, taboo_iter=taboo_iter)
        c0 = P.init_colors(cseed, n, nCol)
        R.set_colors(h, c0)
        steps = []
        for s in range(1, 13):
            u = P.tape(tseed, s, n)
            before, ov = R.sweep_tape(h, u)
            steps.append(dict(viol_before=int(before), overflow=int(ov), colors_sha256=sha(R.get_colors(h, n))))
        trajs.append(dict(nCol=nCol, tabooIteration=taboo_iter, color_seed=cseed, tape_seed=tseed,
                          start_sha256=sha(c0), steps=steps))
        R.L.ref_mcmc_free(h)
    pins["tape_trajectories"] = trajs
    R.L.ref_graph_free(g)

    with open(os.path.join(OUT, "c1_pins.json"), "w") as f:
        json.dump(pins, f, indent=1)

    # ---- small graph with full arrays ----
    sg = R.graph_simulate(200, 0.08, srand=12345)
    sinfo = R.graph_info(sg)
    scumul, sneighs = R.graph_csr(sg)
    sn = sinfo["n"]
    arrays = dict(cumul=scumul, neighs=sneighs)
    for tag, nCol, taboo_iter, tape_kind in [("a", sinfo["maxDeg"], 0, "philox"), ("b", 12, 3, "philox"),
                                             ("c", 9, 0, "philox"), ("ovf", 14, 0, "max")]:
        h = R.mcmc(sg, nCol, 1, taboo_iter=taboo_iter)
        c0 = P.init_colors(1000 + nCol, sn, nCol)
        R.set_colors(h, c0)
        viol0, flags0 = R.violations(h, c0)
        occ = np.stack([R.occupancy(h, c0, v, nCol)[0] for v in range(sn)])
        pvec = np.stack([R.fill_p(h, v, nCol) for v in range(sn)])
        tapes, cols, viols, ovs, taboos = [], [], [], [], []
        for s in range(1, 11):
            if tape_kind == "philox":
                u = P.tape(555 + nCol, s, sn)
            else:  # largest float below 1: forces the CDF walk off the end wherever the sum stays below it
                u = np.full(sn, np.nextafter(np.float32(1.0), np.float32(0.0)), np.float32)
                u[::3] = 0.0
            before, ov = R.sweep_tape(h, u)
            tapes.append(u); cols.append(R.get_colors(h, sn)); viols.append(before); ovs.append(ov)
            taboos.append(R.get_taboo(h, sn))
        arrays.update({f"{tag}_nCol": np.uint32(nCol), f"{tag}_taboo_iter": np.uint32(taboo_iter), f"{tag}_c0": c0,
                       f"{tag}_viol0": np.uint64(viol0), f"{tag}_flags0": flags0, f"{tag}_occ0": occ,
                       f"{tag}_p0": pvec, f"{tag}_tapes": np.stack(tapes), f"{tag}_colors": np.stack(cols),
                       f"{tag}_viol_before": np.array(viols, np.uint64), f"{tag}_overflow": np.array(ovs, np.uint64),
                       f"{tag}_taboo": np.stack(taboos)})
        R.L.ref_mcmc_free(h)
    R.L.ref_graph_free(sg)
    np.savez_compressed(os.path.join(OUT, "small_traj.npz"), **arrays)
    print("wrote c1_pins.json, small_traj.npz; small graph:", sinfo,
          "overflows in ovf case:", arrays["ovf_overflow"].tolist())


if __name__ == "__main__":
    main()
