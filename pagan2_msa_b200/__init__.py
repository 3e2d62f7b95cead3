"""pagan2_msa_b200 -- B200-native drop-in for PAGAN2's pairwise sequence-graph Viterbi engine.

Only the hot path lives here: csrc/ holds the CUDA kernels and the C-ABI
(include/pagan2_b200.h); the Python modules are the host-side harness used by tests and bench.py
(ctypes bindings, job-stream I/O, synthetic workloads, the launch-batch scheduler).
"""
from . import abi  # noqa: F401

__all__ = ["abi"]
