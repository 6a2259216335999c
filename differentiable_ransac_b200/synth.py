"""Deterministic synthetic workloads for the hypothesize-and-score path.

The reference ships no data and no seeds, so every test and
benchmark in this repository draws its inputs from the generators below.  They
follow SURVEY.md section 8(d): a random rigid motion, 3-D points in front of
both cameras, a fixed fraction of the second-image points replaced by uniform
outliers, and the two "logit" regimes the reference produces
(`model_cl.py:461-480`: normalised probabilities ~1/N, or raw U(0,1)).

Everything is generated on the CPU with an explicit torch.Generator so the
same tensors can be rebuilt on the GPU box without shipping fixtures.
"""
from __future__ import annotations

import math

import torch


def _rotation(gen: torch.Generator, dtype=torch.float64) -> torch.Tensor:
    q, r = torch.linalg.qr(torch.randn(3, 3, generator=gen, dtype=dtype))
    q = q * torch.sign(torch.diagonal(r)).unsqueeze(0)
    if torch.linalg.det(q) < 0:
        q[:, 2] = -q[:, 2]
    return q


def skew(t: torch.Tensor) -> torch.Tensor:
    z = torch.zeros((), dtype=t.dtype)
    return torch.stack(
        (
            torch.stack((z, -t[2], t[1])),
            torch.stack((t[2], z, -t[0])),
            torch.stack((-t[1], t[0], z)),
        )
    )


def relative_pose_pair(
    n_points: int = 2000,
    inlier_ratio: float = 0.4,
    seed: int = 1234,
    noise: float = 0.0,
    small_motion: bool = True,
    dtype=torch.float32,
    return_pose: bool = False,
):
    """One image pair in normalised camera coordinates.

    Returns (matches[N,4] = [x1,y1,x2,y2], E_gt[3,3] with x2^T E x1 = 0 and
    ||E||_F = 1, inlier_mask[N] bool) and, with `return_pose`, also (R[3,3], t[3])
    with X2 = R X1 + t.
    """
    gen = torch.Generator().manual_seed(seed)
    if small_motion:
        # moderate rotation keeps all points in front of both cameras
        axis = torch.randn(3, generator=gen, dtype=torch.float64)
        axis = axis / axis.norm()
        ang = 0.35 * (torch.rand((), generator=gen, dtype=torch.float64) - 0.5) * 2
        kx = skew(axis)
        rot = torch.eye(3, dtype=torch.float64) + math.sin(ang) * kx + (1 - math.cos(ang)) * (kx @ kx)
    else:
        rot = _rotation(gen)
    t = torch.randn(3, generator=gen, dtype=torch.float64)
    t = t / t.norm()
    X = torch.randn(n_points, 3, generator=gen, dtype=torch.float64)
    X[:, 2] = X[:, 2] + 5.0
    x1 = X[:, :2] / X[:, 2:3]
    X2 = X @ rot.T + t
    x2 = X2[:, :2] / X2[:, 2:3]
    if noise > 0:
        x1 = x1 + noise * torch.randn(n_points, 2, generator=gen, dtype=torch.float64)
        x2 = x2 + noise * torch.randn(n_points, 2, generator=gen, dtype=torch.float64)
    n_out = int(math.floor((1.0 - inlier_ratio) * n_points))
    x2[:n_out] = torch.rand(n_out, 2, generator=gen, dtype=torch.float64) - 0.5
    inl = torch.ones(n_points, dtype=torch.bool)
    inl[:n_out] = False
    E = skew(t) @ rot
    E = E / E.norm()
    matches = torch.cat((x1, x2), dim=1).to(dtype)
    if return_pose:
        return matches, E.to(dtype), inl, rot, t
    return matches, E.to(dtype), inl


def relative_pose_batch(batch: int, n_points: int = 2000, seed: int = 1234, noise: float = 0.0,
                        dtype=torch.float32):
    """`batch` pairs, inlier ratio round-robin over {0.2, 0.4, 0.6}."""
    ratios = (0.2, 0.4, 0.6)
    ms, es, ins = [], [], []
    for b in range(batch):
        m, e, i = relative_pose_pair(n_points, ratios[b % 3], seed + b, noise, dtype=dtype)
        ms.append(m)
        es.append(e)
        ins.append(i)
    return torch.stack(ms), torch.stack(es), torch.stack(ins)


def logits_regime(batch: int, n_points: int, regime: str = "L0", seed: int = 99,
                  dtype=torch.float32) -> torch.Tensor:
    """L0: sigmoid(N(0,1)) normalised to sum 1 (what `model_cl.py:470-474`
    feeds the sampler when prob_type == 0); L1: U(0,1)."""
    gen = torch.Generator().manual_seed(seed)
    if regime == "L0":
        w = torch.sigmoid(torch.randn(batch, n_points, generator=gen, dtype=torch.float64))
        w = w / w.sum(-1, keepdim=True)
    elif regime == "L1":
        w = torch.rand(batch, n_points, generator=gen, dtype=torch.float64)
    else:
        raise ValueError(regime)
    return w.to(dtype)


def gumbel_noise(shape, seed: int = 7, dtype=torch.float32) -> torch.Tensor:
    """G = -log(-log(U)), U ~ Uniform(tiny, 1 - eps): the distribution that
    `torch.distributions.Gumbel(0, 1).sample` draws (`gumbel_sampler.py:20-22`)."""
    gen = torch.Generator().manual_seed(seed)
    fi = torch.finfo(dtype)
    u = torch.rand(shape, generator=gen, dtype=dtype)
    u = u * ((1 - fi.eps) - fi.tiny) + fi.tiny
    return -torch.log(-torch.log(u))


def pixel_pair(n_points: int = 2000, inlier_ratio: float = 0.5, seed: int = 4321,
               focal: float = 800.0, im_size=(480.0, 640.0), dtype=torch.float32):
    """Fundamental-matrix flavour: returns pixel coordinates (what
    `RANSACLayer.forward` hands the driver after `denormalize_pts`,
    `model_cl.py:239-242`), plus K and F_gt with x2^T F x1 = 0."""
    m, E, inl = relative_pose_pair(n_points, inlier_ratio, seed, dtype=torch.float64)
    K = torch.tensor([[focal, 0.0, im_size[1] / 2], [0.0, focal, im_size[0] / 2], [0.0, 0.0, 1.0]],
                     dtype=torch.float64)
    p1 = m[:, :2] * focal + K[:2, 2]
    p2 = m[:, 2:] * focal + K[:2, 2]
    Kinv = torch.linalg.inv(K)
    F = Kinv.T @ E @ Kinv
    F = F / F.norm()
    return torch.cat((p1, p2), 1).to(dtype), F.to(dtype), K.to(dtype), inl


def rigid_pair(n_points: int = 50000, outlier_ratio: float = 0.7, seed: int = 777,
               noise: float = 0.01, dtype=torch.float32):
    """3D-3D registration pair: Q = R P + t + noise, a fraction
    of Q replaced by uniform points in the bounding box.  Returns
    (points[N,6] = [P | Q], pose[4,4], inlier_mask)."""
    gen = torch.Generator().manual_seed(seed)
    rot = _rotation(gen)
    t = torch.randn(3, generator=gen, dtype=torch.float64)
    P = torch.randn(n_points, 3, generator=gen, dtype=torch.float64)
    Q = P @ rot.T + t + noise * torch.randn(n_points, 3, generator=gen, dtype=torch.float64)
    n_out = int(math.floor(outlier_ratio * n_points))
    lo, hi = Q.min(0).values, Q.max(0).values
    Q[:n_out] = lo + (hi - lo) * torch.rand(n_out, 3, generator=gen, dtype=torch.float64)
    inl = torch.ones(n_points, dtype=torch.bool)
    inl[:n_out] = False
    pose = torch.eye(4, dtype=torch.float64)
    pose[:3, :3] = rot
    pose[:3, 3] = t
    return torch.cat((P, Q), 1).to(dtype), pose.to(dtype), inl
