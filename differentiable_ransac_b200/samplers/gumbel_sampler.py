"""GumbelSoftmaxSampler -- host mirror of `samplers/gumbel_sampler.py:9-42` over the CUDA sampler.

Two entry points:
  * `sample_indices(logits[B,N], K)`  -> the s selected point indices per hypothesis (what the
    B200 driver uses; nothing of size K x N is created);
  * `sample(logits[N])` -> `(ret[K,N], y_soft[K,N])`, the reference's dense return, for callers
    that still run the reference's own `ransac.py` gather (`ransac.py:63-65`).  The noise and the
    top-s come from the kernel; the dense soft one-hot is formed with torch ops so autograd
    reaches `logits` exactly as in the reference (gumbel_sampler.py:34-38).
"""
from __future__ import annotations

import torch

from .. import ops


class GumbelSoftmaxSampler:
    def __init__(self, batch_size, num_samples, tau=1.0, device="cuda", data_type=torch.float32, seed=0):
        self.batch_size = batch_size
        self.num_samples = num_samples
        self.tau = tau
        self.device = device
        self.dtype = data_type
        self.seed = int(seed)
        self.offset = 0              # advanced on every draw: successive calls are independent
        self.injected_noise = None   # test seam: a [K,N] / [B,K,N] Gumbel tensor replaces Philox

    def _next_offset(self):
        o = self.offset
        self.offset += 1
        return o

    def sample_indices(self, logits, K=None, want_lse=False):
        """logits [B,N] (or [N]) -> idx [B,K,s] int32 ascending (+ lse, sel_key when want_lse)."""
        K = K or self.batch_size
        lg = logits if logits.dim() == 2 else logits[None]
        noise = self.injected_noise
        if noise is not None and noise.dim() == 2:
            noise = noise[None]
        return ops.sample(lg.to(self.device), K, self.num_samples, self.tau, noise=noise, seed=self.seed,
                          offset=self._next_offset(), want_lse=want_lse)

    def sample(self, logits=None, num_points=2000, selected=None):
        if logits is None:
            logits = torch.ones(num_points, device=self.device, dtype=torch.float32, requires_grad=True)
        lg = logits.to(self.device).float().reshape(1, -1)
        noise = self.injected_noise
        if noise is not None and noise.dim() == 2:
            noise = noise[None]
        idx, _, _, g = ops.sample(lg, self.batch_size, self.num_samples, self.tau, noise=noise, seed=self.seed,
                                  offset=self._next_offset(), want_noise=True)
        keys = (lg + g[0]) / self.tau                       # [K,N], differentiable w.r.t. logits
        y_soft = keys.softmax(-1)
        y_hard = torch.zeros_like(keys).scatter_(-1, idx[0].long(), 1.0)
        ret = y_hard - y_soft.detach() + y_soft
        return ret.to(self.dtype), y_soft.to(self.dtype)
