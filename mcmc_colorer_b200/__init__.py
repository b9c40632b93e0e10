"""mcmc_colorer_b200 -- B200-native MCMC balanced graph colouring (hot path of Topopiccione/MCMC_Colorer).

Layout:  csrc/  CUDA kernels + the C ABI (include/mcmcb200.h) -> libmcmcb200.so
         host/  C++ host layer mirroring the reference classes and CLI over that ABI
         capi.py / colorer.py   ctypes binding + Python mirror of the reference's colourer interface
         graphgen.py            synthetic graph generators (Erdos-Renyi, R-MAT)
         multigpu.py            one-process-per-GPU driver (torch.distributed / NCCL plumbing)
"""
from .capi import (CONVERGE_EDGES, CONVERGE_VERTICES, FLAG_FORCE_BINNED, FLAG_FORCE_BLOCKED, FLAG_FORCE_DIRECT, FLAG_NO_EARLY_STOP,
                   FLAG_NO_FUSED_FINALIZE, FLAG_NO_OVERLAP, PROPOSAL_DYNAMIC, PROPOSAL_UNIFORM,
                   McmcError)
from .colorer import Chain, ColoringMCMC, ColoringMCMCParams, DeviceCsr, Graph, color_stats, luby_color, occupancy_bits

__all__ = ["Chain", "DeviceCsr", "ColoringMCMC", "ColoringMCMCParams", "Graph", "McmcError", "color_stats", "occupancy_bits", "luby_color",
           "PROPOSAL_UNIFORM", "PROPOSAL_DYNAMIC", "CONVERGE_VERTICES", "CONVERGE_EDGES", "FLAG_NO_FUSED_FINALIZE", "FLAG_NO_EARLY_STOP", "FLAG_FORCE_DIRECT", "FLAG_FORCE_BLOCKED", "FLAG_FORCE_BINNED", "FLAG_NO_OVERLAP"]
