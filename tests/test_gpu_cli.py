"""GPU coverage of the drop-in itself: the C++ host layer (mcmc_colorer_b200/host/, the classes and flags of the reference's
main.cu:28-215) driving libmcmcb200 through bin/MCMC_Colorer on a B200 -- not the Python mirror.

  bin/MCMC_Colorer --simulate 0.1 -n 1000 --seed 1234 --mcmcgpu --mcmccpu --lubygpu --tailcut

* colours files equal to what the ctypes Chain / the CPU oracle give for the same seed (bit-exact);
* the log files carry exactly the keys the reference's own parser reads (pyScripts/logParser.py:17-54, restated below field
  by field because /root/reference does not exist on the GPU box);
* --lubygpu is wired to mcmcb200_luby_color (coloringLuby.cu:364-501 semantics, 1-based colours, log format :179-219).
"""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "mcmc_colorer_b200", "bin", "MCMC_Colorer")
EPS = 1e-8


def reference_line_parser(path):
    """pyScripts/logParser.py:17-54 (lineParser + clusterParser :8-15): same substring tests, same token positions."""
    d = {}
    with open(path) as f:
        it = iter(f)
        for line in it:
            if 'Color histogram:' in line:                       # :20-21 -> clusterParser :8-15
                cluster = []
                for line in it:
                    if 'Number of colors:' in line or 'Average number of nodes' in line or 'end_used_colors' in line:
                        break
                    cluster.append(int(line.split(sep=' ')[1]))
                d["colorClusters"] = cluster
            if 'Max deg:' in line:                               # :22-25
                t = line.split(sep=' ')
                d["maxDeg"], d["minDeg"], d["avgDeg"] = int(t[2]), int(t[6]), float(t[10])
            if 'Nodes:' in line:                                 # :26-28
                d["nnodes"], d["nedges"] = int(line.split(sep=' ')[1]), int(line.split(sep=' ')[4])
            if 'Edge probability' in line:                       # :29-30
                d["edgeProb"] = float(line.split(sep=' ')[6])
            if 'Repetition:' in line:                            # :31-32
                d["repet"] = int(line.split(sep=' ')[1])
            if 'Iteration performed:' in line:                   # :33-34
                d["performedIter"] = int(line.split(sep=' ')[2])
            if 'Max iteration' in line:                          # :35-39
                d["convergence"] = 'no' in line
            if 'Execution time:' in line:                        # :40-41
                d["execTime"] = float(line.split(sep=' ')[2])
            if 'Number of colors:' in line:                      # :42-43
                d["numColors"] = int(line.split(sep=' ')[3])
            if 'Used colors:' in line:                           # :44-45
                d["usedColors"] = int(line.split(sep=' ')[7])
            if 'Color ratio:' in line:                           # :46-47
                d["colorRatio"] = float(line.split(sep=' ')[2])
            if 'Average number' in line:                         # :48-49
                d["avgNodesPerColor"] = float(line.split(sep=' ')[7])
            if 'Variance:' in line:                              # :50-51
                d["varNodesPerColor"] = float(line.split(sep=' ')[1])
            if 'StD:' in line:                                   # :52-53
                d["stdNodesPerColor"] = float(line.split(sep=' ')[1])
    return d


def read_colors(path, n):
    a = np.loadtxt(path, dtype=np.int64)
    assert a.shape == (n, 2) and np.array_equal(a[:, 0], np.arange(n))
    return a[:, 1].astype(np.uint32)


@pytest.fixture(scope="module")
def cli_run(tmp_path_factory):
    out = tmp_path_factory.mktemp("cli") / "out"
    r = subprocess.run([EXE, "--quiet", "--simulate", "0.1", "-n", "1000", "--seed", "1234", "--mcmcgpu", "--mcmccpu", "--lubygpu",
                        "--tailcut", "-o", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    return str(out), r.stdout


def test_cli_graph_and_stdout(cli_run):
    out, stdout = cli_run
    # Graph(1000, 0.1f, .) via setupRnd2 with libc rand() in its initial state: the SURVEY pins (8c)
    assert "Nodes: 1000 - Edges: 99634" in stdout
    assert "Min Degree: 73 - Max Degree: 137" in stdout
    assert "LubyGPU - number of colors:" in stdout and "MCMC GPU elapsed time:" in stdout and "MCMC_CPU elapsed time:" in stdout
    assert "not part of this build" not in stdout


def test_cli_mcmcgpu_colours_equal_ctypes_chain(cli_run, port):
    import mcmc_colorer_b200 as mc
    out, _ = cli_run
    name = "1000_0.100000_1.000000"
    cumul, neighs = port.setup_rnd2(1000, 0.1, srand=1)
    got = read_colors(os.path.join(out, name + "-MCMC_GPU-0-colors.txt"), 1000)
    # the same run through the ctypes binding: shipped defaults (DYNAMIC proposal, conflicting-edge test), --tailcut, seed 1234
    g = mc.Graph(cumul, neighs, prob=0.1)
    prm = mc.ColoringMCMCParams(nCol=137, seed=1234, tailcut=True)
    col = mc.ColoringMCMC(g, None, prm, device=0)
    want = col.run(0)
    col.chain.close()
    assert np.array_equal(got, want)
    assert port.conflict_edges(cumul, neighs, got) == 0                    # proper
    log = open(os.path.join(out, name + "-MCMC_GPU-0.log")).read()
    for key in ("numCol: 137", "epsilon: 1e-08", "maxRip: 250", "***** Tentativo numero: 1", "conflitti rilevati: ", "nuovi conflitti rilevati: ",
                "COLORAZIONE FINALE", "Max iteration reached no", "Number of used colors is 137 on 137 available", "Average 7.29927",
                "StandardDeviation ", "BalancingIndex "):
        assert key in log, key
    st = mc.color_stats(np.bincount(got, minlength=137), 1000, 0.1)
    assert ("StandardDeviation %g" % st["std"]) in log and ("BalancingIndex %g" % st["balancingIndex"]) in log


def test_cli_mcmccpu_log_parses_with_the_reference_parser_rules(cli_run, port):
    from oracle.pyoracle import UNIFORM
    out, _ = cli_run
    name = "1000_0.100000_1.000000"
    cumul, neighs = port.setup_rnd2(1000, 0.1, srand=1)
    d = reference_line_parser(os.path.join(out, name + "-MCMC_CPU-0.log"))
    # every key of lineParser is present
    for k in ("colorClusters", "maxDeg", "minDeg", "avgDeg", "nnodes", "nedges", "edgeProb", "repet", "performedIter", "convergence", "execTime",
              "numColors", "usedColors", "colorRatio", "avgNodesPerColor", "varNodesPerColor", "stdNodesPerColor"):
        assert k in d, k
    assert (d["nnodes"], d["nedges"], d["maxDeg"], d["minDeg"], d["repet"], d["numColors"]) == (1000, 99634, 137, 73, 0, 137)
    got = read_colors(os.path.join(out, name + "-MCMC_CPU-0-colors.txt"), 1000)
    # --mcmccpu = the CPU class's semantics (uniform proposal, violating-vertex test, z = 50 with --tailcut) on the device: the oracle's chain + repair
    c0 = port.init_colors(1234, 1000, 137)
    want, sweeps, cnt, hit = port.run(cumul, neighs, 137, EPS, c0, 1234, UNIFORM, z=50)
    if cnt:
        want, _, _ = port.tailcut(cumul, neighs, 137, want)
    assert np.array_equal(got, want)
    assert d["performedIter"] == sweeps and d["convergence"] is True
    hist = np.bincount(got, minlength=137)
    assert d["colorClusters"] == hist.tolist() and d["usedColors"] == int((hist > 0).sum())
    st = port.color_stats(1000, 137, hist.astype(np.uint32), 0.1)
    assert abs(d["stdNodesPerColor"] - st.stdCPU) < 1e-4 * max(1.0, st.stdCPU)


def test_cli_lubygpu_is_wired(cli_run, port):
    import mcmc_colorer_b200 as mc
    out, _ = cli_run
    name = "1000_0.100000_1.000000"
    cumul, neighs = port.setup_rnd2(1000, 0.1, srand=1)
    got = read_colors(os.path.join(out, name + "-LUBY-0-colors.txt"), 1000)
    want, ncol, rounds = mc.luby_color(cumul, neighs, seed=1234, device=0)
    assert np.array_equal(got, want) and got.min() == 1 and got.max() == ncol
    src = np.repeat(np.arange(1000), np.diff(cumul.astype(np.int64)))
    assert not np.any(got[src] == got[neighs])                            # a proper colouring
    d = reference_line_parser(os.path.join(out, name + "-LUBY-0.log"))
    assert d["numColors"] == ncol and d["nnodes"] == 1000 and d["nedges"] == 99634
    assert d["colorClusters"] == np.bincount(got - 1, minlength=ncol).tolist()
