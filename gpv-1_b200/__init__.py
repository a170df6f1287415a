"""gpv-1_b200: B200-native (sm_100a) implementation of the GPV-1 data-parallel forward/backward hot path.

Layout
  csrc/      hand-written CUDA kernels + the C-ABI (`include/gpvb200.h`), built into lib/libgpvb200.so
  _C.py      ctypes binding of that C-ABI (fails loudly when the library or a Blackwell GPU is missing)
  ops.py     torch.autograd.Function wrappers around the kernels
  model/     host-side mirror of the reference module surface (GPV, HungarianMatcher, SetCriterion, ...)
"""
__version__ = "0.1.0"
