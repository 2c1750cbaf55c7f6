"""Correlation volume, pyramid and lookup (reference: ``CorrBlock`` modules/raft.py:12-48 and
the volume handling in ``RaftFlow.forward`` raft.py:183-186, :208, :217-240).
"""
from __future__ import annotations

import torch

from . import ops


class CorrPyramid:
    """bf16 all-pairs volume produced by the fused tcgen05 kernel.

    ``volume0`` (B, rows_total, h*w): row ``offset(l) + q`` is the h x w source map of driving
    query q at driving resolution (h/2^l) x (w/2^l) -- what the reference obtains with
    ``avg_pool2d`` over the driving dims plus two ``rearrange`` copies (raft.py:208,219,235-236).
    ``volume1`` (B, rows_total, h*w/4): the same maps 2x2-average-pooled over the source dims,
    i.e. level 1 of every ``CorrBlock`` the reference rebuilds per iteration (raft.py:20,238).
    """

    def __init__(self, q_d: torch.Tensor, k_s: torch.Tensor, scale: float):
        self.B, self.C, self.h, self.w = q_d.shape
        self.volume0, self.volume1 = torch.ops.mrfa.corr_pyramid(q_d, k_s, float(scale))
        self.rows_total = self.volume0.shape[1]

    def block(self, pool_log2: int = 0, radius: int = 3) -> "CorrBlock":
        """CorrBlock over the driving plane pooled by 2^pool_log2 (0 = basic resolution)."""
        return CorrBlock.from_pyramid(self, pool_log2, radius)

    def dense(self, pool_log2: int = 0) -> torch.Tensor:
        """fp32 copy of one driving level as the reference's (B*Q,1,h,w) ``corr`` (tests)."""
        off = ops.corr_row_offset(self.h, self.w, pool_log2)
        q = (self.h >> pool_log2) * (self.w >> pool_log2)
        return self.volume0[:, off:off + q].float().reshape(self.B * q, 1, self.h, self.w)


class CorrBlock:
    """Drop-in ``CorrBlock(corr, num_levels=2, radius=3)`` (raft.py:12-48).

    ``corr`` is the reference's (B*h1*w1, 1, H, W) fp32 tensor; level 1 is built by the
    ``mrfa::avg_pool2x2`` kernel.  ``__call__(coords)`` takes (B,2,h1,w1) pixel coordinates and
    returns (B, 2*(2r+1)^2, h1, w1) float32 with channel ``lvl*49 + a*7 + b`` sampled at
    ``(x/2^lvl + a - r, y/2^lvl + b - r)``.  The ``batch_bilinear_sampler`` chunking of
    raft.py:39-40 is a memory workaround with identical results and is not needed here.
    """

    def __init__(self, corr, num_levels=2, radius=3):
        if num_levels != 2:
            raise RuntimeError("mrfa_b200: CorrBlock is specialised for num_levels=2 (raft.py:13 default, never overridden)")
        if corr.dim() != 4 or corr.shape[1] != 1:
            raise RuntimeError("mrfa_b200: CorrBlock expects corr of shape (B*h1*w1, 1, H, W)")
        self.num_levels = num_levels
        self.radius = radius
        corr = corr.float().contiguous()
        self.corr_pyramid = [corr, torch.ops.mrfa.avg_pool2x2(corr)]
        self._H, self._W = corr.shape[-2:]
        self._stride = None          # maps per sample = queries per sample (set at call time)
        self._offset = 0

    @classmethod
    def from_pyramid(cls, pyr: CorrPyramid, pool_log2: int, radius: int = 3) -> "CorrBlock":
        self = cls.__new__(cls)
        self.num_levels, self.radius = 2, radius
        self.corr_pyramid = [pyr.volume0, pyr.volume1]
        self._H, self._W = pyr.h, pyr.w
        self._stride = pyr.rows_total
        self._offset = ops.corr_row_offset(pyr.h, pyr.w, pool_log2)
        return self

    def __call__(self, coords, channels_last: bool = False):
        B, _, h1, w1 = coords.shape
        stride = self._stride if self._stride is not None else h1 * w1
        if self._stride is None and self.corr_pyramid[0].shape[0] != B * h1 * w1:
            raise RuntimeError("mrfa_b200: corr has %d maps but coords address %d queries"
                               % (self.corr_pyramid[0].shape[0], B * h1 * w1))
        return torch.ops.mrfa.corr_lookup(self.corr_pyramid[0], self.corr_pyramid[1], coords, self._H, self._W,
                                          stride, self._offset, self.radius, channels_last)
