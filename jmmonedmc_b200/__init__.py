"""jmmonedmc_b200 — B200-native hot path of mmansell7/jmmOneDMC.

The product is the C-ABI shared library ``libjmmgpu.so`` (hand-written sm_100a CUDA kernels behind
``include/jmm_gpu.h``) plus the C++ host programs in ``csrc/host``.  This Python package is only the
thin ctypes view of that ABI used by the tests and by ``bench.py``; it contains no arithmetic and no
CPU fallback — every compute call fails loudly if the CUDA library or a GPU is missing.
"""
from .capi import (  # noqa: F401
    JmmError, Config, Deck, Handle, Comm, comm_unique_id, SUMMARY_DOUBLES, SUMMARY_FIELDS, lib, lib_path, read_input, rng_selftest, accept_selftest, declared_symbols,
    POT_LJ, POT_LJCUT, POT_HARMONIC, ENS_NPT, ENS_NLT, RNG_TAUS2, RNG_PHILOX, RNG_RECORDED,
    MODE_TABLE, MODE_RECOMPUTE, MODE_CHECKERBOARD, ADAPT_HOST, ADAPT_DEVICE, ADAPT_CALLER, ARITH_REFERENCE, ARITH_FAST, FLAG_CONSISTENT_VIRIAL,
    LOG_ACCEPTED, LOG_VOLUME, LOG_WALL,
)
from .build import build  # noqa: F401

__version__ = "0.1.0"
