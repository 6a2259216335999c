"""Drop-in `model_cl` for the reference's scripts (train.py:6, test.py:3 do `from model_cl import *`).

Everything the reference module defines (CLNet backbone, DS_Block, ...) is re-exported untouched;
`RANSACLayer` / `RANSACLayer3D` and the pose evaluation (`eval_essential_matrix`, `recoverPose`, `AUC`) are
replaced by the B200 host mirror and `DeepRansac_CLNet.forward` hands the WHOLE batch to one launch per stage
instead of the python loop of model_cl.py:488-510.
Checkpoints are unaffected: RANSACLayer has no parameters."""
import time

import torch
import torch.nn.functional as F

from _bootstrap import load_reference_module

_ref = load_reference_module("model_cl")
globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})

from differentiable_ransac_b200.model_cl import RANSACLayer, RANSACLayer3D, batch_episym  # noqa: E402,F401
# test.py:71-76 evaluates every pair with eval_essential_matrix, which reaches the scripts through this module's
# `from cv_utils import *` (model_cl.py:10): pose recovery and the angular errors on the device instead of the
# torch + cv2.triangulatePoints host loop of cv_utils.py:48-189
from differentiable_ransac_b200.cv_utils import AUC, eval_essential_matrix, recoverPose  # noqa: E402,F401


class DeepRansac_CLNet(_ref.DeepRansac_CLNet):
    def __init__(self, opt):
        # build the reference module, then swap the (parameter-free) RANSAC layer
        _ref.RANSACLayer, saved = RANSACLayer, _ref.RANSACLayer
        try:
            super().__init__(opt)
        finally:
            _ref.RANSACLayer = saved

    def forward(self, points, K1, K2, im_size1, im_size2, prob_type=0, gt=None, predict=True):
        B = points.shape[0]
        logit = self.ds_0(points)
        if torch.isnan(logit).any():
            raise Exception("the predicted weights are nan")
        log_probs = F.logsigmoid(logit).view(B, -1)
        weights = log_probs.exp()
        if prob_type == 0:
            out_w = weights / weights.sum(-1, keepdim=True)      # model_cl.py:470-474
        elif prob_type == 1:
            out_w = weights
        else:
            out_w = log_probs
        if not predict:
            return out_w, 0.0
        pts = points.squeeze(-1)[:, 0:4].transpose(1, 2).contiguous()         # [B,N,4]
        torch.cuda.synchronize()
        t0 = time.time()
        ret = self.ransac_layer.forward_batched(pts, out_w, K1, K2, im_size1, im_size2, gt)
        torch.cuda.synchronize()
        return ret, out_w, (time.time() - t0) / B
