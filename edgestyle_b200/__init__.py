"""edgestyle_b200: B200-native (sm_100a) implementation of EdgeStyle's per-step denoise hot path.

Host side mirrors the reference's model/controllora.py, model/edgestyle_multicontrolnet.py and
model/edgestyle_pipeline.py call surfaces; arithmetic runs in hand-written CUDA kernels behind the
C-ABI declared in include/edgestyle_b200.h.  No CPU fallback, no Triton, no torch.compile.
"""
__version__ = "0.1.0"
