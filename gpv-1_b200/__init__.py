"""gpv-1_b200: B200-native (sm_100a) implementation of the GPV-1 data-parallel forward/backward hot path.

Layout
  csrc/        hand-written CUDA kernels + the C-ABI (`include/gpvb200.h`), built into lib/libgpvb200.so
  _C.py        ctypes binding of that C-ABI (fails loudly when the library or a Blackwell GPU is missing)
  kernels.py   tensor-level wrappers (raw pointers + current stream -> C-ABI)
  model/       host-side mirror of the reference module surface (GPV, HungarianMatcher, SetCriterion, GPVCriterion),
               the explicit forward/backward engine and the CUDA-graph capture
  parallel.py  stage-bucketed NCCL gradient all-reduce;  optim.py  fused clip + AdamW;  data.py  H2D prefetcher
  train.py / inference.py   entry points mirroring exp/gpv/train_distr.py and inference.py
"""
__version__ = "0.1.0"
