"""icpslam_b200 — Blackwell-native ICP scan-matching engine (hot path of YoshuaNava/icpslam).

The product is ``libb2icp.so`` (hand-written CUDA for sm_100a behind the C ABI in
``include/b2icp.h``).  This package holds its sources (``csrc/``), a thin ctypes mirror of the
reference's registration call surface (``registration.py``) and the seeded synthetic workloads
(``synth.py``).  There is no CPU fallback: importing ``registration`` without the built library,
or creating a handle without a CUDA device, fails loudly.
"""
__version__ = "0.1.0"
