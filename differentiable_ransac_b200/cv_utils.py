"""Pose recovery and pose error -- host mirror of the evaluation helpers the reference scripts star-import from
`cv_utils.py` (`recoverPose` :48-80, `eval_essential_matrix` :503-525, `AUC` :528-546), over `drb_recover_pose`.

The reference decomposes E with torch, then loops over the four candidate poses calling
`cv2.triangulatePoints` on the host; here the decomposition, the DLT triangulation of every correspondence, the
cheirality vote and the angular errors run in one kernel launch for B pairs x M models.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _matches(p1, p2, device):
    p1 = torch.as_tensor(np.asarray(p1) if not torch.is_tensor(p1) else p1)
    p2 = torch.as_tensor(np.asarray(p2) if not torch.is_tensor(p2) else p2)
    return torch.cat((p1.reshape(-1, 2), p2.reshape(-1, 2)), -1).to(device=device, dtype=torch.float32)


def recoverPose(model, p1, p2, svd=True, distanceThreshold=50):
    """cv_utils.py:48-80: (R [3,3], t [3,1]) of the pose that puts most correspondences in front of both cameras.
    `svd` selects between two decompositions of the same E in the reference; both give the same four poses."""
    dev = model.device if torch.is_tensor(model) and model.is_cuda else torch.device("cuda")
    m = _matches(p1, p2, dev)
    out = ops.recover_pose(torch.as_tensor(model).to(dev).float().reshape(1, 1, 3, 3), m[None], dist=distanceThreshold,
                           want_mask=False)
    dt = model.dtype if torch.is_tensor(model) else torch.float32
    return out["R"][0, 0].to(dt), out["t"][0, 0].to(dt).unsqueeze(1)


def pose_errors(E, matches, R_gt, t_gt, npts=None, distanceThreshold=50):
    """Batched `eval_essential_matrix`: E [B,M,3,3] | [B,3,3], matches [B,N,4], R_gt [B,3,3], t_gt [B,3] ->
    err [B,M,2] degrees (rotation, translation), all on the device, one launch."""
    return ops.recover_pose(E, matches, npts, R_gt, t_gt, dist=distanceThreshold, want_mask=False)["err"]


def eval_essential_matrix(p1n, p2n, E, dR, dt, svd=True):
    """cv_utils.py:503-525 -> (err_R, err_t) in degrees; (180, 90) when there is nothing to evaluate."""
    if len(p1n) != len(p2n):
        raise RuntimeError("Size mismatch in the keypoint lists")
    if len(p1n) < 5 or E is None:
        return 180.0, 90.0
    dev = E.device if torch.is_tensor(E) and E.is_cuda else torch.device("cuda")
    m = _matches(p1n, p2n, dev)
    err = pose_errors(torch.as_tensor(E).to(dev).float().reshape(1, 1, 3, 3), m[None],
                      torch.as_tensor(dR).to(dev).float().reshape(1, 3, 3),
                      torch.as_tensor(dt).to(dev).float().reshape(1, 3))[0, 0]
    return float(err[0]), float(err[1])


def gt_inlier_mask(gt_E, matches, npts=None, distanceThreshold=50):
    """What MatchLoss asks cv2.recoverPose for (loss.py:126-135): the correspondences in front of both cameras
    under the ground-truth pose.  gt_E [B,3,3], matches [B,N,4] -> [B,N] bool, on the device."""
    return ops.recover_pose(gt_E, matches, npts, dist=distanceThreshold)["mask"][:, 0]


def AUC(losses, thresholds=(5, 10, 20), binsize=5):
    """cv_utils.py:528-546 (NG-RANSAC's cumulative-histogram AUC); a host-side metric over a list of numbers."""
    bins = np.arange(int(max(thresholds) / binsize) + 1) * binsize
    hist, _ = np.histogram(np.asarray(losses, dtype=np.float64), bins)
    hist = np.cumsum(hist.astype(np.float32) / len(losses))
    return [np.mean(hist[: int(t / binsize)]) for t in thresholds]
