"""Oracle: pose error of an essential matrix and the AUC metric (evaluation only).

Restates `cv_utils.py:503-525` (`eval_essential_matrix`: decompose E, pick the pose with the
points in front of both cameras, rotation / translation angular errors as `evaluate_R_t_tensor`
`cv_utils.py:361-378`) and `cv_utils.py:528-546` (`AUC`, NG-RANSAC's cumulative-histogram AUC).
The decomposition + cheirality test is OpenCV's `recoverPose`, which the reference itself calls
for the same purpose in its loss (`loss.py:126-131`).  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import math

import numpy as np


def pose_error_deg(E, matches, R_gt, t_gt, mask=None):
    """max-able (err_R, err_t) in degrees of the pose recovered from E.  matches [N,4] normalised
    coordinates (numpy); mask selects the correspondences used for the cheirality vote."""
    import cv2

    E = np.asarray(E, dtype=np.float64)
    m = np.asarray(matches, dtype=np.float64)
    if mask is not None and np.asarray(mask).sum() >= 5:
        m = m[np.asarray(mask, dtype=bool)]
    if m.shape[0] < 5 or not np.isfinite(E).all():
        return 180.0, 90.0
    _, R, t, _ = cv2.recoverPose(E, m[:, None, 0:2].copy(), m[:, None, 2:4].copy(), np.eye(3))
    R_gt = np.asarray(R_gt, dtype=np.float64)
    t_gt = np.asarray(t_gt, dtype=np.float64).ravel()
    cos_r = min(1.0, max(-1.0, (np.trace(R @ R_gt.T) - 1.0) * 0.5))
    err_r = math.degrees(math.acos(cos_r))
    t = t.ravel() / (np.linalg.norm(t) + 1e-15)
    tg = t_gt / (np.linalg.norm(t_gt) + 1e-15)
    loss_t = max(1e-15, 1.0 - float(t @ tg) ** 2)            # sign-free, as cv_utils.py:371-372
    err_t = math.degrees(math.acos(math.sqrt(max(0.0, 1.0 - loss_t))))
    return err_r, err_t


def auc(losses, thresholds=(5, 10, 20), binsize=5):
    """cv_utils.py:528-546."""
    bins = np.arange(int(max(thresholds) / binsize) + 1) * binsize
    hist, _ = np.histogram(losses, bins)
    hist = np.cumsum(hist.astype(np.float32) / max(len(losses), 1))
    return [float(np.mean(hist[: int(t / binsize)])) for t in thresholds]
