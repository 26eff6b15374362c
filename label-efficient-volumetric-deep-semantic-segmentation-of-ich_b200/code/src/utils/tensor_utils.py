"""Drop-in for the reference's `src.utils.tensor_utils`: the confusion-matrix reduction that
follows the forward in `UNet2D.evaluate` (models/optim/UNet2D.py:220-222) as one fused CUDA pass."""
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from ich_b200 import ops  # noqa: E402


def batch_binary_confusion_matrix(pred, target):
    """tn, fp, fn, tp per batch element (reference utils/tensor_utils.py:12-36)."""
    return ops.confusion_matrix(pred, target)
