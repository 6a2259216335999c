"""MSACScore -- host mirror of `scorings/msac_score.py:4-55` over the fused CUDA scorer."""
from __future__ import annotations

import torch

from .. import ops


def sampson_sq(matches, model):
    """d2 [N] of one model (torch ops; used only for the winner's mask)."""
    n = matches.shape[0]
    one = torch.ones(n, 1, device=matches.device, dtype=matches.dtype)
    h1 = torch.cat((matches[:, 0:2], one), -1)
    h2 = torch.cat((matches[:, 2:4], one), -1)
    e = h1 @ model.T
    f = h2 @ model
    r = (h2 * e).sum(-1)
    return r * r / (e[:, 0] ** 2 + e[:, 1] ** 2 + f[:, 0] ** 2 + f[:, 1] ** 2)


class LazyMasks:
    """Stands in for the reference's dense `masks [M,N]` (msac_score.py:44).  The driver only ever
    indexes one row (`ransac.py:117`), so rows are produced on demand; `.dense()` builds them all."""

    def __init__(self, matches, models, thr2):
        self.matches, self.models, self.thr2 = matches, models, thr2
        self.shape = (models.shape[0], matches.shape[0])

    def __getitem__(self, i):
        if isinstance(i, torch.Tensor) and i.dim() == 0:
            i = int(i)
        if isinstance(i, int):
            return sampson_sq(self.matches, self.models[i]) < self.thr2
        return self.dense()[i]

    def dense(self):
        return torch.stack([sampson_sq(self.matches, m) < self.thr2 for m in self.models])


class MSACScore:
    def __init__(self, device="cuda"):
        self.device = device
        self.provides_inliers = True

    def score(self, matches, models, threshold=0.75):
        """matches [N,4], models [M,3,3] -> (scores [M], masks)."""
        thr = torch.tensor([float(threshold)], device=matches.device)
        scores, _ = ops.score_msac(matches[None], models.reshape(1, -1, 9), thr)
        return scores[0], LazyMasks(matches.float(), models.float(), (1.5 * float(threshold)) ** 2)
