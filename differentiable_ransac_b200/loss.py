"""MatchLoss -- host mirror of `loss.py:107-153` (the `-w2` symmetric-epipolar training loss)."""
from __future__ import annotations

import numpy as np
import torch

from . import cv_utils, engine


def denormalize_pts(pts, im_size):
    """cv_utils.py:35-45."""
    return pts * max(im_size) + torch.stack((im_size[1] / 2, im_size[0] / 2))


def normalize_keypoints_tensor(pts, K):
    """Pixel -> normalised camera coordinates: (x - c) / f."""
    return (pts - K[:2, 2]) / torch.stack((K[0, 0], K[1, 1]))


def gt_inlier_mask(gt_E, pts1, pts2):
    """loss.py:126-135 asks cv2.recoverPose, on the host, for the correspondences that lie in front of both
    cameras under the ground-truth pose; `drb_recover_pose` returns the same mask without leaving the device."""
    m = torch.cat((pts1, pts2), -1).detach().float()[None]
    E = torch.as_tensor(np.asarray(gt_E, dtype=np.float32) if not torch.is_tensor(gt_E) else gt_E)
    return cv_utils.gt_inlier_mask(E.to(m.device).float().reshape(1, 3, 3), m)[0]


class MatchLoss(object):
    def __init__(self, fmat):
        self.fmat = fmat

    def forward(self, models, gt_E, pts1, pts2, K1, K2, im_size1, im_size2, topk_flag=False, k=1, gt_masks=None):
        """models: list over pairs of [K_b,3,3]; returns the scalar of loss.py:152-153."""
        losses = []
        for b in range(len(models)):
            if self.fmat:
                Es = K2[b].transpose(-1, -2) @ models[b] @ K1[b]
                p1 = normalize_keypoints_tensor(denormalize_pts(pts1[b].clone(), im_size1[b]), K1[b])
                p2 = normalize_keypoints_tensor(denormalize_pts(pts2[b].clone(), im_size2[b]), K2[b])
            else:
                Es, p1, p2 = models[b], pts1[b], pts2[b]
            mask = gt_masks[b] if gt_masks is not None else gt_inlier_mask(gt_E[b], p1, p2)
            inl = torch.cat((p1[mask], p2[mask]), -1).float()[None]
            row = engine.EpisymLoss.apply(inl, Es[None].float())[0] / max(inl.shape[1], 1)   # e_l.mean(1)
            if topk_flag:
                losses.append(torch.topk(row, k=k, largest=False).values.mean())
            else:
                losses.append(row.mean())
        return sum(losses) / len(models)
