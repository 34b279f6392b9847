"""modimizer_b200 - modimizer's data-parallel hot path on NVIDIA B200 (sm_100a).

Host-side mirror of the reference's seqhash / modset / modmap interface for the
path (richarddurbin/modimizer seqhash.h, modset.h, modutils.c:19-51,
modmap.c:93-134,188-231) over the C ABI of libmodgpu.so (include/modgpu.h).
All compute happens in hand-written CUDA kernels; importing this package on a
box without the built library or without a GPU works, but every operation then
raises ModgpuError - nothing falls back to the CPU.
"""
from ._lib import ModgpuError, LIB_PATH, SYMBOLS, load, require_device  # noqa: F401
from .modset import Seqhash, Modset, Reference, kmer_string, seqio_pack  # noqa: F401
from . import synth  # noqa: F401

__all__ = ["ModgpuError", "Seqhash", "Modset", "Reference", "kmer_string", "seqio_pack", "synth", "load", "require_device"]
