"""cpic_b200 — B200-native replacement for the per-timestep hot path of the cpic
particle-in-cell simulator (reference: rodarima/cpic, src/sim.c:481-581).

The product is `libcpic_b200.so` (hand-written sm_100a CUDA kernels + cuFFT + NCCL behind
a C ABI, include/cpic_b200.h). This package is only the ctypes view of that ABI used by
the tests and the benchmark; it contains no numerical code and no CPU fallback: every
stage call fails loudly when the library or a GPU is missing.
"""
from ._lib import lib, lib_path, build, Cpic_b200Error, EXPORTS
from .sim import Sim, Params, load_conf, init_particles

__all__ = ["lib", "lib_path", "build", "Cpic_b200Error", "EXPORTS", "Sim", "Params", "load_conf",
           "init_particles"]
