"""EssentialMatrixEstimatorNister -- host mirror of
`estimators/essential_matrix_estimator_nister.py:30-408` over the CUDA 5-point kernel."""
from __future__ import annotations

import torch

from .. import ops
from ._autograd import E5AllSlots


class EssentialMatrixEstimatorNister:
    def __init__(self, device="cuda"):
        self.sample_size = 5
        self.device = device
        self.last_nsol = None

    def estimate_model(self, matches, weights=None, K1=None, K2=None, inlier_indices=None, best_model=None,
                       unnormalzied_threshold=None, best_score=0):
        """matches [K,5,4] -> [K*10,3,3] (ten slots per sample, nister.py:400-407; slots beyond the
        number of real roots hold the identity, `self.last_nsol[k]` says how many are genuine).
        `weights` scale the rows of a 5 x 9 system whose null space they cannot change, so they are
        accepted and ignored there.  Non-minimal input [K,n>5,4] takes the reference's no-pymagsac branch
        (nister.py:64-65: the same polynomial system on the four smallest right singular vectors of
        A^T A) through `drb_refit_e5` (accumulation and solve in double, no gradient); the pymagsac
        bundle adjustment itself (nister.py:10-24) is an external C++ module and is not reproduced."""
        if matches.shape[1] < self.sample_size:
            return None
        if matches.shape[1] > self.sample_size:
            models, nsol = ops.refit_e5(matches.detach().float(), None, weights)
            self.last_nsol = nsol
            return models.reshape(-1, 3, 3).to(matches.dtype)
        models, nsol = E5AllSlots.apply(matches.float())
        self.last_nsol = nsol
        return models.reshape(-1, 3, 3).to(matches.dtype)

    estimate_minimal_model = estimate_model
