"""ich_b200 -- B200-native engine behind the reference's U-Net / loss-module API.

Host code is Python + PyTorch (device memory, streams, autograd plumbing, torch.distributed); every
arithmetic step of the hot path runs in libich_b200.so (hand-written sm_100a CUDA, C ABI in
include/ich_b200.h).  No CPU fallback, no cuDNN convolution, no Triton.
"""
from . import config            # noqa: F401
from ._lib import build, lib    # noqa: F401
