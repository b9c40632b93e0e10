"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every sweep kernel family on small graphs, a
tail-cut repair, the narrow and uint32 host interfaces.  Usage (GPU box):
    compute-sanitizer --tool memcheck  python scripts/sanitize_driver.py
    compute-sanitizer --tool racecheck python scripts/sanitize_driver.py
Results are checked against the oracle as in smoke(), so a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcmc_colorer_b200 import (Chain, ColoringMCMCParams, PROPOSAL_UNIFORM, PROPOSAL_DYNAMIC, CONVERGE_VERTICES, CONVERGE_EDGES,   # noqa: E402
                               FLAG_FORCE_BLOCKED, FLAG_FORCE_BINNED, FLAG_FORCE_DIRECT, FLAG_NO_OVERLAP, FLAG_NO_EARLY_STOP)
from mcmc_colorer_b200.graphgen import er_graph_numpy   # noqa: E402
from oracle.pyoracle import Port, UNIFORM, DYNAMIC   # noqa: E402

P = Port()
n = int(os.environ.get("SANITIZE_N", "20000"))
cumul, neighs = er_graph_numpy(n, 16, seed=3)
maxdeg = int(np.diff(cumul.astype(np.int64)).max())
cases = [("direct", FLAG_FORCE_DIRECT, {}, maxdeg, PROPOSAL_UNIFORM, CONVERGE_VERTICES),
         ("blocked-serial", FLAG_FORCE_BLOCKED | FLAG_NO_OVERLAP, {}, maxdeg, PROPOSAL_DYNAMIC, CONVERGE_EDGES),
         # the overlapped pair needs both kernels resident at once; a tool that serialises launches runs pass A to completion first
         # (pass B then never waits), so this also exercises the hand-over flags
         ("blocked-overlapped", FLAG_FORCE_BLOCKED, dict(stage_cap_bytes=32768), maxdeg, PROPOSAL_UNIFORM, CONVERGE_VERTICES),
         ("blocked-2buf", FLAG_FORCE_BLOCKED, dict(stage_cap_bytes=22528, stage_buffers=2), maxdeg, PROPOSAL_UNIFORM, CONVERGE_VERTICES),
         ("binned", FLAG_FORCE_BINNED, {}, 200, PROPOSAL_DYNAMIC, CONVERGE_EDGES),
         ("binned-wide-masks", FLAG_FORCE_BINNED, {}, 300, PROPOSAL_UNIFORM, CONVERGE_VERTICES),
         ("wide", 0, {}, 700, PROPOSAL_UNIFORM, CONVERGE_VERTICES),
         ("wide-dynamic", 0, {}, 5000, PROPOSAL_DYNAMIC, CONVERGE_EDGES)]
for name, flags, tune, nCol, prop, conv in cases:
    prm = ColoringMCMCParams(nCol=nCol, proposal=prop, convergence=conv, seed=9, tailcut=True)
    ch = Chain(cumul, neighs, prm, device=0, flags=flags | FLAG_NO_EARLY_STOP, **tune)
    rng = np.random.default_rng(1)
    c = (rng.integers(0, 10, n) * (nCol // 10)).astype(np.uint32)     # 10 colours in use: plenty of conflicts and walks
    ch.init_colors(c)
    oprop = UNIFORM if prop == PROPOSAL_UNIFORM else DYNAMIC
    for s in range(1, 3):
        ch.sweep(1)
        c, _ = P.sweep(cumul, neighs, nCol, 1e-8, c, P.tape(9, s, n, oprop), oprop)
        assert np.array_equal(ch.get_colors(), c), (name, s)
    st = ch.status()
    assert st.conflictEdges == P.conflict_edges(cumul, neighs, c), name
    fixed, _, left = P.tailcut(cumul, neighs, nCol, c)
    ch.tailcut(64)
    assert np.array_equal(ch.get_colors(), fixed) and ch.status().conflictEdges == left, name
    narrow = ch.get_colors_narrow()
    ch.init_colors_narrow(narrow)
    assert np.array_equal(ch.get_colors(), fixed), name
    print("sanitize_driver:", name, ch.kernel_mode(), "ok", flush=True)
    ch.close()
print("sanitize_driver: all ok")
