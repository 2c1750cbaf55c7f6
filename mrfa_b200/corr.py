"""Correlation volume, pyramid and lookup (reference: ``CorrBlock`` modules/raft.py:12-48 and
the volume handling in ``RaftFlow.forward`` raft.py:183-186, :208, :217-240).
"""
from __future__ import annotations

import torch

from . import _lib, ops


class _GradHolder:
    """fp32 gradient accumulators of one pyramid.  Every lookup's backward scatter-adds straight
    into them (no per-lookup dense gradient tensors, no bf16 round trip); the pyramid's own
    backward consumes them once all lookups are done."""

    def __init__(self, vol0, vol1):
        self.g0 = torch.zeros(vol0.shape, device=vol0.device, dtype=torch.float32)
        self.g1 = torch.zeros(vol1.shape, device=vol1.device, dtype=torch.float32)


class _CorrPyramidFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q_d, k_s, scale, holder_box):
        vol0, vol1 = torch.ops.mrfa.corr_pyramid(q_d, k_s, scale)
        ctx.save_for_backward(q_d, k_s)
        ctx.scale, ctx.holder_box = scale, holder_box
        ctx.set_materialize_grads(False)
        return vol0, vol1

    @staticmethod
    def backward(ctx, _g0, _g1):
        # the volume gradients arrive through the holder (see _PyramidLookupFn), not through autograd
        q_d, k_s = ctx.saved_tensors
        holder = ctx.holder_box[0]
        if holder is None:
            return torch.zeros_like(q_d), torch.zeros_like(k_s), None, None
        # G = scale * (g0 + unpool(g1) / 4) -> dA = G Bm, dB = G^T A on tcgen05, then the avg-pool backward of the pooled
        # driving rows and the un-permutation of the source rows (csrc/corr_bwd.cu)
        d_q, d_k = torch.ops.mrfa.corr_pyramid_bwd(holder.g0, holder.g1, q_d.detach(), k_s.detach(), ctx.scale)
        ctx.holder_box[0] = None
        return d_q, d_k, None, None


class _PyramidLookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coords, vol0, vol1, holder_box, H, W, stride, offset, radius, layout, channels_last):
        out = torch.ops.mrfa.corr_lookup(vol0, vol1, coords, H, W, stride, offset, radius, layout, channels_last)
        ctx.save_for_backward(coords, vol0, vol1)
        ctx.cfg, ctx.holder_box = (H, W, stride, offset, radius, layout), holder_box
        return out

    @staticmethod
    def backward(ctx, g):
        coords, vol0, vol1 = ctx.saved_tensors
        H, W, stride, offset, radius, layout = ctx.cfg
        holder = ctx.holder_box[0]
        if holder is None:
            holder = ctx.holder_box[0] = _GradHolder(vol0, vol1)
        g = g.contiguous()
        coords = coords.contiguous()
        B, _, h1, w1 = coords.shape
        gc = torch.empty_like(coords)
        with torch.cuda.device(coords.device):
            ops.check(ops.lib.mrfa_corr_lookup_bwd(ops._p(g), ops._p(vol0), ops._p(vol1), 1, ops._p(coords), ops._p(holder.g0),
                                                   ops._p(holder.g1), ops._p(gc), B, h1 * w1, H, W, stride, offset, radius,
                                                   layout, ops._stream()), "mrfa_corr_lookup_bwd")
        return gc, None, None, None, None, None, None, None, None, None, None


class CorrPyramid:
    """bf16 all-pairs volume produced by the fused tcgen05 kernel.

    ``volume0`` (B, rows_total, h*w): row ``offset(l) + q`` is the h x w source map of driving
    query q at driving resolution (h/2^l) x (w/2^l) -- what the reference obtains with
    ``avg_pool2d`` over the driving dims plus two ``rearrange`` copies (raft.py:208,219,235-236).
    ``volume1`` (B, rows_total, h*w/4): the same maps 2x2-average-pooled over the source dims,
    i.e. level 1 of every ``CorrBlock`` the reference rebuilds per iteration (raft.py:20,238).
    """

    def __init__(self, q_d: torch.Tensor, k_s: torch.Tensor, scale: float, q_bias=None, k_bias=None):
        self.B, self.C, self.h, self.w = q_d.shape
        self._holder_box = [None]
        if torch.is_grad_enabled() and (q_d.requires_grad or k_s.requires_grad or
                                        any(b is not None and b.requires_grad for b in (q_bias, k_bias))):
            if q_bias is not None:
                q_d = q_d + q_bias.view(1, -1, 1, 1)
            if k_bias is not None:
                k_s = k_s + k_bias.view(1, -1, 1, 1)
            self.volume0, self.volume1 = _CorrPyramidFn.apply(q_d, k_s, float(scale), self._holder_box)
        else:
            # q_bias / k_bias: the biases of the 1x1 heads that produced q_d / k_s, added inside the pack kernel
            self.volume0, self.volume1 = torch.ops.mrfa.corr_pyramid(q_d, k_s, float(scale), q_bias, k_bias)
        self.rows_total = self.volume0.shape[1]
        self.layout = ops.corr_map_layout(self.h, self.w)      # _lib.MAP_TILED for w = 64 / 128, else row-major

    def block(self, pool_log2: int = 0, radius: int = 3) -> "CorrBlock":
        """CorrBlock over the driving plane pooled by 2^pool_log2 (0 = basic resolution)."""
        return CorrBlock.from_pyramid(self, pool_log2, radius)

    def dense(self, pool_log2: int = 0, level: int = 0) -> torch.Tensor:
        """fp32 row-major copy of one driving level as the reference's (B*Q,1,H,W) ``corr`` (level 0) or its 2x2 source
        pool (level 1, raft.py:20), whatever the stored map layout (tests / debugging)."""
        off = ops.corr_row_offset(self.h, self.w, pool_log2)
        q = (self.h >> pool_log2) * (self.w >> pool_log2)
        H, W = self.h >> level, self.w >> level
        rows = (self.volume1 if level else self.volume0)[:, off:off + q].float()
        if self.layout != _lib.MAP_ROWMAJOR:
            rows = rows.index_select(2, ops.corr_map_permutation(self.layout, level, H, W, rows.device))
        return rows.reshape(self.B * q, 1, H, W)


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
        self._layout = _lib.MAP_ROWMAJOR
        self._holder_box = None

    @classmethod
    def from_pyramid(cls, pyr: CorrPyramid, pool_log2: int, radius: int = 3) -> "CorrBlock":
        self = cls.__new__(cls)
        self.num_levels, self.radius = 2, radius
        self.corr_pyramid = [pyr.volume0, pyr.volume1]
        self._H, self._W = pyr.h, pyr.w
        self._stride = pyr.rows_total
        self._offset = ops.corr_row_offset(pyr.h, pyr.w, pool_log2)
        self._layout = pyr.layout
        self._holder_box = pyr._holder_box
        return self

    def __call__(self, coords, channels_last: bool = False):
        B, _, h1, w1 = coords.shape
        stride = self._stride if self._stride is not None else h1 * w1
        if self._stride is None and self.corr_pyramid[0].shape[0] != B * h1 * w1:
            raise RuntimeError("mrfa_b200: corr has %d maps but coords address %d queries"
                               % (self.corr_pyramid[0].shape[0], B * h1 * w1))
        if self._holder_box is not None and torch.is_grad_enabled() and \
                (coords.requires_grad or self.corr_pyramid[0].requires_grad):
            return _PyramidLookupFn.apply(coords, self.corr_pyramid[0], self.corr_pyramid[1], self._holder_box, self._H,
                                          self._W, stride, self._offset, self.radius, self._layout, channels_last)
        return torch.ops.mrfa.corr_lookup(self.corr_pyramid[0], self.corr_pyramid[1], coords, self._H, self._W,
                                          stride, self._offset, self.radius, self._layout, channels_last)
