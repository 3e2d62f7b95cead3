"""pagan2_msa_b200 -- B200-native drop-in for PAGAN2's pairwise sequence-graph Viterbi engine.

Only the hot path lives here: csrc/ holds the CUDA kernels and the C-ABI
(include/pagan2_b200.h); the Python modules are the host-side harness used by tests and bench.py
(ctypes bindings, job-stream I/O, synthetic workloads, the launch-batch scheduler).
"""
import os

# The engine keeps up to 8 chunks of a pipelined pg2_align_batch call in flight on two streams each.  With the driver's default
# of 8 hardware work queues, streams share queues and a chunk's upload waits behind another chunk's traceback (read when the
# process creates its CUDA context, so it is set before torch or the library can do that).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import abi  # noqa: E402,F401

__all__ = ["abi"]
