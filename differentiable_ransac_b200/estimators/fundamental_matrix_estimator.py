"""FundamentalMatrixEstimatorNew -- host mirror of `estimators/fundamental_matrix_estimator.py:161-308`."""
from __future__ import annotations

import torch

from .. import ops
from ._autograd import F8Solve


class FundamentalMatrixEstimatorNew:
    def __init__(self, device="cuda", weighted=0):
        self.sample_size = 7
        self.device = device
        self.weighted = weighted
        self.eps = 1e-8
        self.last_nsol = None

    def estimate_model(self, matches, weights=None):
        s = matches.shape[1]
        if s == 8:                                    # normalised 8-point (:230-260)
            models, _ = F8Solve.apply(matches.float())
            return models.to(matches.dtype)
        if s == 7:                                    # correct 7-point; the reference's is broken
            models, nsol = ops.solve_f7(matches.float())
            self.last_nsol = nsol[0]
            return models[0].reshape(-1, 3, 3).to(matches.dtype)
        if s > 8:                                     # normalise + eight-point on all n rows (:169-175), no gradient
            models, _ = ops.refit_f8(matches.detach().float(), None, weights)
            return models.reshape(-1, 3, 3).to(matches.dtype)
        return None
