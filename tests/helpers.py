"""Shared comparison helpers for the parity tests."""
import torch


def trace_constraint_residual(E: torch.Tensor) -> torch.Tensor:
    """|| 2 E E^T E - tr(E E^T) E ||_F per model: ~0 only for genuine essential
    matrices, i.e. for the real-root slots of the 5-point solvers."""
    E = E.double()
    EEt = E @ E.transpose(-1, -2)
    tr = EEt.diagonal(dim1=-2, dim2=-1).sum(-1)
    return (2 * EEt @ E - tr[:, None, None] * E).flatten(1).norm(dim=1)


def match_up_to_sign(cand: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """cand [K,S,3,3], ref [K,R,3,3] (unit norm).  For every ref model the
    distance to the closest candidate of the same sample, up to sign -> [K,R]."""
    c = cand.double().flatten(2)
    r = ref.double().flatten(2)
    d_pos = (c[:, :, None, :] - r[:, None, :, :]).norm(dim=-1)
    d_neg = (c[:, :, None, :] + r[:, None, :, :]).norm(dim=-1)
    return torch.minimum(d_pos, d_neg).min(dim=1).values


def unit(M: torch.Tensor) -> torch.Tensor:
    return M / M.flatten(-2).norm(dim=-1)[..., None, None]
