"""Oracle: Gumbel-softmax / Gumbel-top-s sampler and the minimal-sample gather.

Restates `samplers/gumbel_sampler.py:25-42` (sample) and the driver gather
`ransac.py:63-65`.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import torch


def gumbel_keys(logits: torch.Tensor, noise: torch.Tensor, tau: float = 1.0) -> torch.Tensor:
    """keys[k, n] = (logits[n] + G[k, n]) / tau   (gumbel_sampler.py:30,34)."""
    return (logits.unsqueeze(0).expand_as(noise) + noise) / tau


def sample(logits: torch.Tensor, noise: torch.Tensor, num_samples: int, tau: float = 1.0):
    """Straight-through Gumbel top-s draw.

    logits [N], noise [K, N] (the Gumbel(0,1) draw the reference takes from the
    global generator at gumbel_sampler.py:33 -- injected here).
    Returns (ret [K,N], y_soft [K,N], idx_sorted [K,s]): `ret` is
    y_hard - y_soft.detach() + y_soft (gumbel_sampler.py:35-38); idx_sorted are
    the selected point indices in ascending order, which is the order in which
    the boolean-mask gather of ransac.py:65 emits the minimal sample.
    """
    keys = gumbel_keys(logits, noise, tau)
    y_soft = keys.softmax(-1)
    top = torch.topk(keys, num_samples, dim=-1)
    y_hard = torch.zeros_like(keys).scatter_(-1, top.indices, 1.0)
    ret = y_hard - y_soft.detach() + y_soft
    idx_sorted = top.indices.sort(-1).values
    return ret, y_soft, idx_sorted


def gather_minimal(matches: torch.Tensor, ret: torch.Tensor) -> torch.Tensor:
    """ransac.py:64-65: points = matches.repeat(K,1,1) * ret[...,None];
    minimal = points[ret != 0].view(K, -1, D).  Materialises [K,N,D] exactly as
    the reference does (that temporary is part of what the CPU baseline pays)."""
    K = ret.shape[0]
    points = matches.repeat([K, 1, 1]) * ret.unsqueeze(-1)
    return points[ret != 0].view(K, -1, matches.shape[-1])


def logits_grad_from_sample_grad(logits, noise, tau, idx_sorted, g_sel):
    """Closed form of the straight-through backward:
    g_sel[k, j] = dL/d ret[k, idx_sorted[k, j]] (zero for unselected n);
    dL/dlogits[n] = (1/tau) sum_k y[k,n] (g[k,n] - sum_m y[k,m] g[k,m])."""
    keys = gumbel_keys(logits, noise, tau)
    y = keys.softmax(-1)
    g = torch.zeros_like(y).scatter_(-1, idx_sorted, g_sel)
    inner = (y * g).sum(-1, keepdim=True)
    return ((y * (g - inner)).sum(0)) / tau
