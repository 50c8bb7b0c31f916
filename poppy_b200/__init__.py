"""poppy_b200 — B200 (sm_100a) morph renderer behind the reference's morph_images()/Settings API.

Only what the hot path needs lives here:
  csrc/            hand-written CUDA kernels + the extern "C" layer (include/poppy_cuda.h) and the C++ host stages
  _lib.py          ctypes loader of the in-tree libpoppy_cuda.so (no fallback)
  renderer.py      Python mirror of the device boundary
  api.py           Python mirror of the reference interface: Settings, morph_images(), morph_sequence()
  synth.py         deterministic synthetic inputs for the benchmark workloads
"""
__all__ = ["renderer", "synth"]
