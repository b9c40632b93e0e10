"""CPU tests of the C++ host layer (mcmc_colorer_b200/host): --simulate and --graph CSR construction against the
reference's own Graph code (oracle/_ref), the CLI surface, and the loud failure without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "mcmc_colorer_b200", "bin")
HOST_SO = os.path.join(ROOT, "mcmc_colorer_b200", "libmcmcb200_host.so")


@pytest.fixture(scope="module")
def host():
    if not (os.path.exists(HOST_SO) and os.path.exists(os.path.join(BIN, "MCMC_Colorer"))):
        import __graft_entry__
        __graft_entry__.build()
    L = C.CDLL(HOST_SO)
    for f in ("mcmchost_graph_simulate", "mcmchost_graph_simulate_fast"):
        getattr(L, f).restype = C.c_void_p
        getattr(L, f).argtypes = [C.c_uint32, C.c_float, C.c_uint32]
    L.mcmchost_graph_from_file.restype = C.c_void_p
    L.mcmchost_graph_from_file.argtypes = [C.c_char_p]
    L.mcmchost_graph_info.argtypes = [C.c_void_p] + [C.c_void_p] * 6
    L.mcmchost_graph_copy_csr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.mcmchost_graph_free.argtypes = [C.c_void_p]
    return L


def host_csr(L, g):
    n, nnz, mx, mn = (C.c_uint32() for _ in range(4))
    mean, prob = C.c_float(), C.c_float()
    L.mcmchost_graph_info(g, *(C.byref(x) for x in (n, nnz, mx, mn, mean, prob)))
    cumul = np.zeros(n.value + 1, np.uint32)
    neighs = np.zeros(max(nnz.value, 1), np.uint32)
    L.mcmchost_graph_copy_csr(g, cumul.ctypes.data, neighs.ctypes.data)
    return cumul, neighs[:nnz.value], dict(n=n.value, nnz=nnz.value, maxDeg=mx.value, minDeg=mn.value,
                                           meanDeg=mean.value, prob=prob.value)


def test_simulate_is_rand_exact_like_reference(host, port, golden_dir):
    import hashlib, json
    pins = json.load(open(os.path.join(golden_dir, "c1_pins.json")))["graph"]
    port.libc.srand(1)
    g = host.mcmchost_graph_simulate(1000, 0.1, 1234)
    cumul, neighs, info = host_csr(host, g)
    host.mcmchost_graph_free(g)
    assert (info["nnz"], info["maxDeg"], info["minDeg"]) == (pins["nnz"], pins["maxDeg"], pins["minDeg"])
    assert abs(info["meanDeg"] - pins["meanDeg"]) < 1e-6
    assert hashlib.sha256(cumul.tobytes()).hexdigest() == pins["cumul_sha256"]
    assert hashlib.sha256(neighs.tobytes()).hexdigest() == pins["neighs_sha256"]


def test_fast_generator_is_a_simple_symmetric_graph(host):
    n, p = 60000, 16.0 / 60000
    g = host.mcmchost_graph_simulate_fast(n, p, 7)
    cumul, neighs, info = host_csr(host, g)
    host.mcmchost_graph_free(g)
    assert info["n"] == n and abs(info["meanDeg"] - 16.0) < 0.3
    src = np.repeat(np.arange(n, dtype=np.int64), np.diff(cumul.astype(np.int64)))
    dst = neighs.astype(np.int64)
    assert not np.any(src == dst)                                        # no self loops
    key = src * n + dst
    assert np.all(np.diff(key) > 0)                                      # rows ascending, no duplicates
    assert np.array_equal(np.sort(dst * n + src), key)                   # symmetric


def test_dataset_file_roundtrip_matches_reference_importer(host, ref, tmp_path):
    path = str(tmp_path / "net.txt")
    subprocess.run([os.path.join(BIN, "datasetGen"), "300", "0.05", path], check=True, capture_output=True)
    lines = open(path).read().splitlines()
    nn, ne = map(int, lines[0].split("\t"))
    assert nn == 300 and ne == len(lines) - 1 and len(lines[1].split("\t")) == 3
    # add a self loop and a blank line: both must be ignored like the reference does
    name = lines[1].split("\t")[0]
    with open(path, "a") as f:
        f.write(f"\n{name}\t{name}\t0.5\n")
    g = host.mcmchost_graph_from_file(path.encode())
    cumul, neighs, info = host_csr(host, g)
    host.mcmchost_graph_free(g)
    rg = ref.graph_from_file(path)
    rc, rn = ref.graph_csr(rg)
    rinfo = ref.graph_info(rg)
    assert np.array_equal(cumul, rc) and np.array_equal(neighs, rn)
    assert info["nnz"] == rinfo["nnz"] == 2 * ne and info["maxDeg"] == rinfo["maxDeg"]


def test_cli_surface_and_loud_failure_without_gpu(host, tmp_path):
    exe = os.path.join(BIN, "MCMC_Colorer")
    out = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert out.returncode == 0
    for flag in ("--graph", "--outDir", "--simulate", "-n N", "--mcmccpu", "--mcmcgpu", "--lubygpu", "--nCol",
                 "--numColRatio", "--tabooIterations", "--tailcut", "--repet", "--seed"):
        assert flag in out.stdout, flag
    assert subprocess.run([exe, "--cite-me"], capture_output=True, text=True).stdout.count("colorerGbR2019") == 1
    bad = subprocess.run([exe, "--quiet", "--simulate", "0.1"], capture_output=True, text=True)
    assert bad.returncode != 0 and "number of nodes" in bad.stdout
    bad = subprocess.run([exe, "--quiet", "--simulate", "0.1", "-n", "50", "--numColRatio", "17"], capture_output=True, text=True)
    assert bad.returncode != 0
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "--quiet", "--simulate", "0.1", "-n", "200", "--mcmcgpu", "--seed", "3", "-o",
                            str(tmp_path / "o")], capture_output=True, text=True)
        assert r.returncode != 0 and "no usable sm_100a CUDA device" in r.stderr
        assert "Nodes: 200" in r.stdout


def test_integration_option_a_compiles_against_reference_headers(tmp_path):
    """INTEGRATION.md Option A: host/coloringMCMC.{h,cpp} dropped into a translation unit that uses the REFERENCE's graph.h /
    graphCPU.cpp / coloring.h (tests/integration/option_a_main.cpp = main.cu:160-198).  Compiles, links against libmcmcb200.so and
    -- without a B200 -- fails loudly at mcmcb200_create.  Needs the reference tree (build container only)."""
    ref = "/root/reference/src"
    if not os.path.isdir(ref):
        pytest.skip("no /root/reference here")
    import shutil
    host_dir = os.path.join(ROOT, "mcmc_colorer_b200", "host")
    for f in ("coloringMCMC.h", "coloringMCMC.cpp"):
        shutil.copy(os.path.join(host_dir, f), tmp_path / f)          # away from host/graph.h and host/coloring.h
    shutil.copy(os.path.join(ROOT, "tests", "integration", "option_a_main.cpp"), tmp_path / "option_a_main.cpp")
    src = open(tmp_path / "coloringMCMC.cpp").read().replace('#include "../../include/mcmcb200.h"', '#include "mcmcb200.h"')
    open(tmp_path / "coloringMCMC.cpp", "w").write(src)
    exe = str(tmp_path / "option_a")
    lib_dir = os.path.join(ROOT, "mcmc_colorer_b200")
    cmd = ["g++", "-std=c++14", "-O1", "-w", "-o", exe, str(tmp_path / "option_a_main.cpp"), str(tmp_path / "coloringMCMC.cpp"),
           ref + "/utils/fileImporter.cpp", ref + "/utils/timer.cpp", ref + "/utils/miscUtils.cpp",
           "-I" + os.path.join(ROOT, "oracle", "shims"), "-I" + ref, "-I" + ref + "/graph", "-I" + ref + "/graph_coloring", "-I" + ref + "/utils",
           "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "include"), "-L" + lib_dir, "-lmcmcb200", "-Wl,-rpath," + lib_dir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    run = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    import torch
    if torch.cuda.is_available():
        assert run.returncode == 0 and "option_a ok" in run.stdout, run.stderr
        assert os.path.exists(tmp_path / "optA-MCMC_GPU-0.log") and os.path.exists(tmp_path / "optA-MCMC_CPU-0-colors.txt")
    else:
        assert run.returncode == 2 and "no usable sm_100a CUDA device" in run.stderr, run.stderr
