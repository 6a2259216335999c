"""MatchLoss (`-w2`, loss.py:107-153), PoseLoss (`-w0`, :11-68) and ClassificationLoss (`-w1`, :71-104) -- host
mirrors over the CUDA path; none of them leaves the device."""
from __future__ import annotations

import numpy as np
import torch

from . import cv_utils, engine, ops


def denormalize_pts(pts, im_size):
    """cv_utils.py:35-45."""
    return pts * max(im_size) + torch.stack((im_size[1] / 2, im_size[0] / 2))


def normalize_keypoints_tensor(pts, K):
    """Pixel -> normalised camera coordinates: (x - c) / f."""
    return (pts - K[:2, 2]) / torch.stack((K[0, 0], K[1, 1]))


def _as_E(gt_E, device):
    E = torch.as_tensor(np.asarray(gt_E, dtype=np.float32) if not torch.is_tensor(gt_E) else gt_E)
    return E.to(device).float().reshape(-1, 3, 3)


def gt_inlier_mask(gt_E, pts1, pts2):
    """loss.py:126-135 asks cv2.recoverPose, on the host, for the correspondences that lie in front of both
    cameras under the ground-truth pose; `drb_recover_pose` returns the same mask without leaving the device."""
    m = torch.cat((pts1, pts2), -1).detach().float()[None]
    return cv_utils.gt_inlier_mask(_as_E(gt_E, m.device), m)[0]


def gt_inlier_masks(gt_E, p1, p2):
    """The masks of all pairs of a batch: one launch when the pairs hold the same number of correspondences
    (the reference's loaders pad to a fixed count), else one per pair."""
    if len({tuple(a.shape) for a in p1}) == 1:
        m = torch.stack([torch.cat((a, b), -1) for a, b in zip(p1, p2)]).detach().float()
        E = torch.cat([_as_E(gt_E[b], m.device) for b in range(len(p1))])
        return list(cv_utils.gt_inlier_mask(E, m))
    return [gt_inlier_mask(gt_E[b], p1[b], p2[b]) for b in range(len(p1))]


class MatchLoss(object):
    def __init__(self, fmat):
        self.fmat = fmat

    def forward(self, models, gt_E, pts1, pts2, K1, K2, im_size1, im_size2, topk_flag=False, k=1, gt_masks=None):
        """models: list over pairs of [K_b,3,3]; returns the scalar of loss.py:152-153."""
        Es, p1, p2 = [], [], []
        for b in range(len(models)):
            if self.fmat:
                Es.append(K2[b].transpose(-1, -2) @ models[b] @ K1[b])
                p1.append(normalize_keypoints_tensor(denormalize_pts(pts1[b].clone(), im_size1[b]), K1[b]))
                p2.append(normalize_keypoints_tensor(denormalize_pts(pts2[b].clone(), im_size2[b]), K2[b]))
            else:
                Es.append(models[b])
                p1.append(pts1[b])
                p2.append(pts2[b])
        if gt_masks is None:
            gt_masks = gt_inlier_masks(gt_E, p1, p2)
        losses = []
        for b in range(len(models)):
            mask = gt_masks[b]
            inl = torch.cat((p1[b][mask], p2[b][mask]), -1).float()[None]
            row = engine.EpisymLoss.apply(inl, Es[b][None].float())[0] / max(inl.shape[1], 1)   # e_l.mean(1)
            if topk_flag:
                losses.append(torch.topk(row, k=k, largest=False).values.mean())
            else:
                losses.append(row.mean())
        return sum(losses) / len(models)


class _PoseErrorTerm(torch.autograd.Function):
    """(err_R + err_t) / 2 in degrees per model, gradient to the models from `drb_pose_loss` (forward-mode duals
    in the same launch, so the backward is one multiply)."""

    @staticmethod
    def forward(ctx, models, matches, R_gt, t_gt):
        err, grad = ops.pose_loss(models.detach(), matches, R_gt, t_gt)
        ctx.save_for_backward(grad)
        return err.mean(-1)

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return g[..., None, None] * grad, None, None, None


class PoseLoss(torch.nn.Module):
    """`-w0`: average rotation / translation error of every model, host mirror of loss.py:11-68."""

    def __init__(self, fmat=False):
        super().__init__()
        self.fmat = fmat

    def forward_average(self, estimated_models, pts1, pts2, gt_R, gt_t, K1=None, K2=None, im_size1=None,
                        im_size2=None, svd=False):
        """estimated_models: list over pairs of [K_b,3,3].  The reference evaluates `models[i]` -- the models as
        handed in, also when `fmat` (loss.py:59 ignores the `Es` it computes at :36) -- against points that are
        K-normalised when `fmat`; mirrored as is.  `svd=True` differentiates through torch.linalg.svd in the
        reference, which is singular for an essential matrix (two equal singular values); only the closed-form
        branch the loss defaults to is provided."""
        if svd:
            raise NotImplementedError("PoseLoss differentiates Horn's closed form (svd=False, the reference default)")
        total = 0.0
        for b, models in enumerate(estimated_models):
            if self.fmat:
                p1 = normalize_keypoints_tensor(denormalize_pts(pts1[b].clone(), im_size1[b]), K1[b])
                p2 = normalize_keypoints_tensor(denormalize_pts(pts2[b].clone(), im_size2[b]), K2[b])
            else:
                p1, p2 = pts1[b], pts2[b]
            m = torch.cat((p1, p2), -1).detach().float()[None]
            dev = m.device
            term = _PoseErrorTerm.apply(models[None].float(), m, torch.as_tensor(gt_R[b]).to(dev).float()[None],
                                        torch.as_tensor(gt_t[b]).to(dev).float().reshape(1, 3))
            total = total + term.sum() / models.shape[0]
        return total / len(estimated_models)


class ClassificationLoss(torch.nn.Module):
    """`-w1`: binary cross-entropy of the predicted weights against the ground-truth inlier mask, host mirror of
    loss.py:71-104; the mask comes from `drb_recover_pose` instead of cv2.recoverPose on the host."""

    def __init__(self, fmat):
        super().__init__()
        self.fmat = fmat

    def forward(self, gt_E, pts1, pts2, logits, K1, K2, im_size1, im_size2):
        p1, p2 = [], []
        for b in range(len(logits)):
            if self.fmat:   # cv2.undistortPoints without distortion = (x - c) / f  (loss.py:81-92); K arrives as numpy
                Ka = torch.as_tensor(K1[b]).to(device=pts1[b].device, dtype=pts1[b].dtype)
                Kb = torch.as_tensor(K2[b]).to(device=pts2[b].device, dtype=pts2[b].dtype)
                p1.append(normalize_keypoints_tensor(denormalize_pts(pts1[b].clone(), im_size1[b]), Ka))
                p2.append(normalize_keypoints_tensor(denormalize_pts(pts2[b].clone(), im_size2[b]), Kb))
            else:
                p1.append(pts1[b])
                p2.append(pts2[b])
        masks = gt_inlier_masks(gt_E, p1, p2)
        return torch.nn.BCELoss()(logits, torch.stack(masks).to(logits.dtype))
