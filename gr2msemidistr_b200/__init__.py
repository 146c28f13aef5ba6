"""gr2msemidistr_b200 -- B200-native (sm_100a) GR2MSemiDistr hot path: GR2M recursion over
cells x parameter sets, NSE/KGE/RMSE objective, Weighted-Flow-Accumulation routing.

The compute lives in csrc/ (hand-written CUDA behind the C ABI of include/gr2m_b200.h); this package
holds the ctypes binding and the host-side mirror of the reference's exported functions.
There is no CPU fallback: every entry point raises if libgr2m_b200.so is missing.
"""
__version__ = "0.1.0"
