"""graphecho_b200 — B200-native (sm_100a) implementation of GraphEcho's data-parallel hot path.

Layout
  csrc/        hand-written CUDA kernels + the C-ABI (include/graphecho_b200.h)
  _cabi.py     ctypes binding of libgraphecho_b200.so (raw pointers, sizes, stream)
  functional.py torch.autograd.Function shells over the kernels
  models/, utils/  host-side mirror of the reference's module API
               (models.fpnseg / vig / graph_matching / affinity_layer / transformer / TGCN /
               gradient_reversal, utils.sinkhorn_distance / losses): same class names,
               constructor / forward signatures and state_dict keys
  engine.py    the restated training step (train_cardiac_uda.py:223-325) + DDP plumbing
  synth.py     synthetic inputs of the BASELINE shapes

Importing the package does not load the shared library; the first kernel call does, and it
raises if the library is missing (there is no CPU / eager fallback).
"""

__version__ = "0.1.0"
