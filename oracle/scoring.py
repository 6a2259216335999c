"""Oracle: residual-and-score functions.

Restates `scorings/msac_score.py:12-55` (Sampson / soft-MSAC), `model_cl.py:13-26`
(`batch_episym`) with the clamped mean of `loss.py:138-151`, and the rigid
point residual `estimators/rigid_transformation_SVD_based_solver.py:76-89`.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import torch


def msac_score(matches: torch.Tensor, models: torch.Tensor, threshold: float = 0.75):
    """matches [N,4], models [M,3,3] -> (scores [M], masks [M,N] bool).

    d2 = (x2^T M x1)^2 / ((M x1)_0^2 + (M x1)_1^2 + (M^T x2)_0^2 + (M^T x2)_1^2);
    thr2 = (1.5 t)^2; score = sum_n max(0, 1 - d2/thr2); mask = d2 < thr2.
    Materialises the same [M,3,N] / [M,N] temporaries as msac_score.py:33-48.
    """
    thr2 = (3 / 2 * threshold) ** 2
    n = matches.shape[0]
    one = torch.ones((n, 1), dtype=matches.dtype)
    h1 = torch.cat((matches[:, 0:2], one), dim=-1)
    h2 = torch.cat((matches[:, 2:4], one), dim=-1)
    Mx1 = models.matmul(h1.transpose(-1, -2))                        # [M,3,N]
    Mtx2 = models.transpose(-1, -2).matmul(h2.transpose(-1, -2))     # [M,3,N]
    jj = Mx1[:, 0] ** 2 + Mx1[:, 1] ** 2 + Mtx2[:, 0] ** 2 + Mtx2[:, 1] ** 2
    x1Mtx2 = h1.T.unsqueeze(0).mul(Mtx2).sum(-2)                     # == x2^T M x1
    d2 = x1Mtx2.square().div(jj)
    masks = d2 < thr2
    scores = torch.sum(torch.clamp(1 - d2 / thr2, min=0.0), dim=-1)
    return scores, masks


def episym(x1: torch.Tensor, x2: torch.Tensor, F: torch.Tensor) -> torch.Tensor:
    """Symmetric epipolar distance, model_cl.py:13-26.  x1, x2 [K,P,2], F [K,3,3] -> [K,P]."""
    K, P = x1.shape[0], x1.shape[1]
    h1 = torch.cat([x1, x1.new_ones(K, P, 1)], dim=-1).reshape(K, P, 3, 1)
    h2 = torch.cat([x2, x2.new_ones(K, P, 1)], dim=-1).reshape(K, P, 3, 1)
    Fr = F.reshape(-1, 1, 3, 3).repeat(1, P, 1, 1)
    x2Fx1 = torch.matmul(h2.transpose(2, 3), torch.matmul(Fr, h1)).reshape(K, P)
    Fx1 = torch.matmul(Fr, h1).reshape(K, P, 3)
    Ftx2 = torch.matmul(Fr.transpose(2, 3), h2).reshape(K, P, 3)
    return x2Fx1 ** 2 * (1.0 / (Fx1[:, :, 0] ** 2 + Fx1[:, :, 1] ** 2 + 1e-15)
                         + 1.0 / (Ftx2[:, :, 0] ** 2 + Ftx2[:, :, 1] ** 2 + 1e-15))


def match_loss(models: torch.Tensor, pts1: torch.Tensor, pts2: torch.Tensor, gt_mask: torch.Tensor,
               topk: int | None = None) -> torch.Tensor:
    """loss.py:138-151 for one pair, with the GT-inlier mask given (the
    reference gets it from cv2.recoverPose on the host, loss.py:126-135):
    mean over K x P_inl of min(episym, 1) [or mean of the k best rows]."""
    K = models.shape[0]
    p1 = pts1[gt_mask].repeat(K, 1, 1)
    p2 = pts2[gt_mask].repeat(K, 1, 1)
    geod = episym(p1, p2, models)
    e_l = torch.min(geod, geod.new_ones(geod.shape))
    if topk is not None:
        idx = torch.topk(e_l.mean(1), k=topk, largest=False).indices
        return e_l[idx].mean()
    return e_l.mean()


def rigid_squared_residual(pts1: torch.Tensor, pts2: torch.Tensor, descriptor: torch.Tensor,
                           threshold: float = 0.03):
    """rigid_transformation_SVD_based_solver.py:76-89.  pts1, pts2 [N,3];
    descriptor [K,4,3] = model[:, :3, :]^T -> (sum_n d2 [K], mean d2 (scalar), mask [K,N])."""
    pts_t = torch.cat((pts1, torch.ones((pts1.shape[0], 1), dtype=pts1.dtype)), dim=1)
    t = pts_t @ descriptor
    d2 = torch.sum((pts2[None, :, :] - t) ** 2, dim=-1)
    return d2.sum(-1), d2.mean(), d2 < threshold
