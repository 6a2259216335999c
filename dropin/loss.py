"""Drop-in `loss`: MatchLoss, PoseLoss and ClassificationLoss over the CUDA path (same names and signatures);
anything else the reference module defines is re-exported untouched."""
from _bootstrap import load_reference_module

_ref = load_reference_module("loss")
globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})

from differentiable_ransac_b200.loss import ClassificationLoss, MatchLoss, PoseLoss  # noqa: E402,F401
# loss.py:4 star-imports cv_utils, so train.py's `from loss import *` would put the host versions back
from differentiable_ransac_b200.cv_utils import AUC, eval_essential_matrix, recoverPose  # noqa: E402,F401
