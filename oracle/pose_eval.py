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


def recover_pose_ref(E, p1, p2, dist=50.0):
    """`cv_utils.recoverPose` with `svd=True` (cv_utils.py:48-116) and `cheirality_check` (:179-189) restated
    with numpy: SVD of E, R1 = U W V^T, R2 = U W^T V^T, t = U[:, 2]; the four poses in the reference's order;
    homogeneous DLT triangulation (what cv2.triangulatePoints does) of every correspondence; the pose with the
    most points in front of both cameras and closer than `dist`.  -> (R, t, mask of that pose, counts[4])."""
    E = np.asarray(E, dtype=np.float64)
    p1 = np.asarray(p1, dtype=np.float64)
    p2 = np.asarray(p2, dtype=np.float64)
    u, _, vt = np.linalg.svd(E)
    w = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], dtype=np.float64)
    u_ = -u if np.linalg.det(u) < 0 else u
    vt_ = -vt if np.linalg.det(vt) < 0 else vt
    R1, R2, t = u_ @ w @ vt_, u_ @ w.T @ vt_, u[:, 2]
    poses = [(R1, t), (R2, t), (R1, -t), (R2, -t)]
    masks = []
    for R, tt in poses:
        P = np.concatenate((R, tt[:, None]), 1)
        ok = np.zeros(p1.shape[0], dtype=bool)
        for n in range(p1.shape[0]):
            A = np.stack((np.array([-1.0, 0, p1[n, 0], 0]), np.array([0, -1.0, p1[n, 1], 0]),
                          p2[n, 0] * P[2] - P[0], p2[n, 1] * P[2] - P[1]))
            Q = np.linalg.svd(A)[2][-1]
            Qh = Q / Q[3]
            z2 = P[2] @ Qh
            ok[n] = (Q[2] * Q[3] > 0) and (Qh[2] < dist) and (z2 > 0) and (z2 < dist)
        masks.append(ok)
    counts = np.array([m.sum() for m in masks])
    best = int(np.argmax(counts))
    return poses[best][0], poses[best][1], masks[best], counts


def pose_error_ref(R, t, R_gt, t_gt):
    """`evaluate_R_t_tensor` (cv_utils.py:361-378) + the degree conversion of `eval_essential_matrix` (:525)."""
    eps = 1e-8
    R_gt = np.asarray(R_gt, dtype=np.float64)
    t_gt = np.asarray(t_gt, dtype=np.float64).ravel()
    err_q = math.acos(max(-1.0, min(1.0, (np.trace(R @ R_gt.T) - 1) * 0.5)))
    tg = t_gt / (np.linalg.norm(t_gt) + eps)
    loss_t = max(eps, 1.0 - float(np.ravel(t) @ tg) ** 2)
    err_t = math.acos(math.sqrt(1 - loss_t + eps))
    return math.degrees(err_q), math.degrees(err_t)
