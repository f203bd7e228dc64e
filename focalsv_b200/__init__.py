"""focalsv_b200 — B200-native (sm_100a) implementation of FocalSV's alignment-DP hot path.

Only what the path needs lives here:
  csrc/      hand-written CUDA kernels + the C ABI (libfocalsv_cuda.so, include/focalsv_cuda.h)
  api.py     ctypes binding (no CPU fallback: raises if the library or a GPU is missing)
  presets.py scoring / band parameters of the reference's call sites
  synth.py   synthetic chr21-/hg38-shaped workloads of BASELINE.json's five configs
  shard.py   length-balanced sharding of independent tasks across GPUs (no collective)
"""
from . import _abi  # noqa: F401

__version__ = "0.1.0"
