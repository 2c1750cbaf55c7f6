"""RaftFlow: the coarse-to-fine non-prior motion refinement decoder (reference:
modules/raft.py:50-311), with the correlation volume / pyramid / lookup and every feature warp
running on the hand-written kernels.  Same constructor kwargs (config/vox1.yaml:45-64), forward
signature, return values and state_dict keys as the reference class.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn.functional as F
from torch import nn

from . import ops, sampling
from .blocks import Hourglass, OcclusionAwareGenerator, _Cache, _small7_ok, conv_relu, fast_path, install_cache_hooks, invalidate_caches
from .corr import CorrPyramid


SPLIT_K = os.environ.get("MRFA_SPLIT_K", "1") != "0"            # conv(cat([a, b])) as two accumulating convolutions
CAT_SLICES = os.environ.get("MRFA_CAT_SLICES", "1") != "0"       # A/B switch: coarse warps written into the decoder's cat buffers
FUSED_CARRY = os.environ.get("MRFA_FUSED_CARRY", "1") != "0"     # A/B switch for the fused level hand-over


def _resize(x, size):
    """F.interpolate(bilinear, align_corners=True); on the inference path one fused kernel."""
    if x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled():
        return torch.ops.mrfa.resize_bilinear(x, int(size[0]), int(size[1]), 0)
    return F.interpolate(x, size=size, mode="bilinear", align_corners=True)


def _like(t, ref):
    """Contiguous copy of a weight slice in the memory format of `ref`."""
    fmt = torch.channels_last if (ref.dim() == 4 and ref.is_contiguous(memory_format=torch.channels_last)) else torch.contiguous_format
    return t.contiguous(memory_format=fmt)


def _cat(module, a, b):
    """torch.cat([a, b], 1); one vectorised pass for channels_last maps on the inference path."""
    if fast_path(module, a) and ops.cat2_ok(a, b):
        return torch.ops.mrfa.cat2(a, b)
    return torch.cat([a, b], dim=1)


class BasicMotionEncoder(nn.Module):
    """raft.py:50-68: (flow, correlation features) -> 128-channel motion feature."""

    def __init__(self, num_levels=2, radius=3):
        super().__init__()
        planes = num_levels * (2 * radius + 1) ** 2
        self.convc1 = nn.Conv2d(planes, 128, 1, padding=0)
        self.convc2 = nn.Conv2d(128, 96, 3, padding=1)
        self.convf1 = nn.Conv2d(2, 128, 7, padding=3)
        self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
        self.conv = nn.Conv2d(64 + 96, 128 - 2, 3, padding=1)

    def _padded_conv(self):
        """self.conv with two zero filters appended (126 -> 128 output channels), cached."""
        if not hasattr(self, "_pad"):
            self._pad = _Cache()
        cv = self.conv

        def build():
            w = torch.cat([cv.weight, cv.weight.new_zeros((2,) + tuple(cv.weight.shape[1:]))], dim=0)
            b = torch.cat([cv.bias, cv.bias.new_zeros(2)])
            if cv.weight.is_contiguous(memory_format=torch.channels_last):
                w = w.contiguous(memory_format=torch.channels_last)
            return w, b

        return self._pad.get((cv.weight, cv.bias), build)

    def forward(self, delta_flow, corr):
        if corr.shape[-2:] != delta_flow.shape[-2:]:
            # correlation features still at the basic resolution (levels above it, raft.py:241-243):
            # the 1x1 convc1 commutes with the align_corners bilinear resize, so apply it first and
            # let one kernel produce relu(resize(.)) -- the 98-channel upsampled tensor is never built
            # (the bias commutes as well -- the bilinear weights sum to one -- and rides in the same kernel instead of a
            # separate elementwise pass behind the GEMM)
            c1 = F.conv2d(corr, self.convc1.weight)
            c = torch.ops.mrfa.resize_bilinear(c1, delta_flow.shape[-2], delta_flow.shape[-1], 1, self.convc1.bias)
            c = conv_relu(self.convc2, c)
        else:
            c = conv_relu(self.convc2, conv_relu(self.convc1, corr))
        f = conv_relu(self.convf2, conv_relu(self.convf1, delta_flow))
        if SPLIT_K and fast_path(self, c) and self.conv.out_channels + delta_flow.shape[1] == 128:
            # conv(cat([c, f])) = conv(c; W[:, :96]) + conv(f; W[:, 96:]): the second convolution accumulates onto the
            # first (cuDNN's fused add + bias + ReLU), so the cat pass disappears into two compute-bound kernels
            w, b = self._padded_conv()
            if not hasattr(self, "_split"):
                self._split = _Cache()
            nc = c.shape[1]
            wa, wb = self._split.get((w,), lambda: (_like(w[:, :nc], w), _like(w[:, nc:], w)))
            cv = self.conv
            z = torch.cudnn_convolution(c, wa, cv.padding, cv.stride, cv.dilation, cv.groups, torch.backends.cudnn.benchmark, False, torch.backends.cudnn.allow_tf32)
            y = torch.cudnn_convolution_add_relu(f, wb, z, 1.0, b, cv.stride, cv.padding, cv.dilation, cv.groups)
            y[:, 126:128] = delta_flow
            return y
        cf = _cat(self, c, f)
        if fast_path(self, cf) and self.conv.out_channels + delta_flow.shape[1] == 128:
            # 126 output channels force cuDNN through pad / un-pad copies of the whole map: run the
            # convolution with two zero filters appended (128 channels) and drop the flow into
            # those two channels -- same values as cat([relu(conv(.)), flow]) without the cat.
            cv = self.conv
            w, b = self._padded_conv()
            y = torch.cudnn_convolution_relu(cf, w, b, cv.stride, cv.padding, cv.dilation, cv.groups)
            y[:, 126:128] = delta_flow
            return y
        y = conv_relu(self.conv, cf)
        return torch.cat([y, delta_flow], dim=1)


class RefineFlow(nn.Module):
    """raft.py:70-87: motion feature + warped-source context -> (d_flow(2) ++ d_occ(1)).

    Returns ``(out, inp)`` like the reference.  ``inp`` (the concatenated 256-channel input, which the reference returns
    but never uses: raft.py:256 discards it) is ``None`` on the inference fast path with MRFA_SPLIT_K=1, where the two
    halves feed two accumulating convolutions and the concatenation is never materialised; set MRFA_SPLIT_K=0 (or run
    with grad enabled / in train mode) to get the tensor."""

    def __init__(self):
        super().__init__()
        self.convc1 = nn.Conv2d(192, 128, 3, padding=1)
        self.conv1 = nn.Conv2d(256, 128, 3, padding=1)
        self.conv2 = nn.Conv2d(128, 2, 3, padding=1)
        self.convo1 = nn.Conv2d(256, 128, 3, padding=1)
        self.convo2 = nn.Conv2d(128, 1, 3, padding=1)

    def forward(self, m_f, warp_f):
        ctx = conv_relu(self.convc1, warp_f)
        split = SPLIT_K and fast_path(self, ctx)
        inp = None if split else _cat(self, m_f, ctx)
        if fast_path(self, ctx):
            # conv1 | convo1 read the same 256-channel input: run them as one 256 -> 256 convolution
            # (one pass over `inp`), then conv2 / convo2 as one block-diagonal 256 -> 4 convolution
            # whose channels are [flow_x, flow_y, occlusion, 0] -- identical arithmetic per output.
            if not hasattr(self, "_merged"):
                self._merged = _Cache()

            def build():
                cl = self.conv1.weight.is_contiguous(memory_format=torch.channels_last)
                w1 = torch.cat([self.conv1.weight, self.convo1.weight], dim=0)
                b1 = torch.cat([self.conv1.bias, self.convo1.bias])
                z2 = self.conv2.weight.new_zeros(self.conv2.weight.shape)
                zo = self.convo2.weight.new_zeros(self.convo2.weight.shape)
                w2 = torch.cat([torch.cat([self.conv2.weight, z2], dim=1), torch.cat([zo, self.convo2.weight], dim=1),
                                self.conv2.weight.new_zeros((1, 256) + tuple(self.conv2.weight.shape[2:]))], dim=0)
                b2 = torch.cat([self.conv2.bias, self.convo2.bias, self.conv2.bias.new_zeros(1)])
                if cl:
                    w1, w2 = w1.contiguous(memory_format=torch.channels_last), w2.contiguous(memory_format=torch.channels_last)
                return w1, b1, w2, b2

            w1, b1, w2, b2 = self._merged.get((self.conv1.weight, self.conv1.bias, self.convo1.weight, self.convo1.bias,
                                               self.conv2.weight, self.conv2.bias, self.convo2.weight, self.convo2.bias), build)
            c1 = self.conv1
            if split:
                # (conv1 | convo1)(cat([m_f, ctx])) as two accumulating convolutions: no cat pass (see BasicMotionEncoder)
                if not hasattr(self, "_split"):
                    self._split = _Cache()
                nm = m_f.shape[1]
                wa, wb = self._split.get((w1,), lambda: (_like(w1[:, :nm], w1), _like(w1[:, nm:], w1)))
                z = torch.cudnn_convolution(m_f, wa, c1.padding, c1.stride, c1.dilation, 1, torch.backends.cudnn.benchmark, False, torch.backends.cudnn.allow_tf32)
                hdn = torch.cudnn_convolution_add_relu(ctx, wb, z, 1.0, b1, c1.stride, c1.padding, c1.dilation, 1)
            else:
                hdn = torch.cudnn_convolution_relu(inp, w1, b1, c1.stride, c1.padding, c1.dilation, 1)
            out = F.conv2d(hdn, w2, b2, self.conv2.stride, self.conv2.padding)
            return out[:, :3], inp
        flow = self.conv2(conv_relu(self.conv1, inp))
        occ = self.convo2(conv_relu(self.convo1, inp))
        return torch.cat([flow, occ], dim=1), inp


class RaftFlow(nn.Module):
    def __init__(self, prior_only=False, num_kp=10, dim=256, size=256, generator=None, driving_encoder=None,
                 source_encoder=None):
        super().__init__()
        self.scale = dim ** -0.5
        self.size = size
        self.h = self.w = size // 4                       # basic flow resolution of the prior
        self.prior_only = prior_only
        self.generator = OcclusionAwareGenerator(**generator)
        widths = (512, 512, 512, 256, 128, 64)             # raft.py:105-113, coarsest first
        self.total_iter = self.num_iter = int(math.log(2 ** 5, 2)) + 1
        self.basic_res_index = int(math.log(self.h // (size // 32), 2))
        install_cache_hooks(self)
        if self.prior_only:
            return
        self.kp = Hourglass(**driving_encoder)
        self.kp_img = Hourglass(**source_encoder)
        self.kp_head = nn.Conv2d(self.kp.out_filters, dim, kernel_size=1, padding=0)
        self.kp_img_head = nn.Conv2d(self.kp_img.out_filters, dim, kernel_size=1, padding=0)
        self.pos_embedding = nn.Parameter(torch.zeros(1, num_kp, self.h, self.w))
        nn.init.trunc_normal_(self.pos_embedding, std=.02)
        self.corr_enc = BasicMotionEncoder()
        self.refine = RefineFlow()
        self.to_context = nn.ModuleList(nn.Conv2d(widths[i], 192, 1, padding=0) for i in range(self.num_iter))

    channels_last = False
    auto_channels_last = True      # inference on CUDA switches to NHWC memory on first use (values unchanged; sticky: the
                                   # module stays channels_last afterwards -- set False to keep the reference's NCHW memory)

    def train(self, mode: bool = True):
        invalidate_caches(self)                   # folded / merged inference weights are rebuilt from the live parameters
        return super().train(mode)

    def channels_last_(self, enable: bool = True):
        """Run the decoder in NHWC memory (torch.channels_last): the layout the sm_100 tensor-core
        convolutions produce natively and the one in which a bilinear tap is C contiguous floats,
        so the warp / lookup kernels take their vectorised path.  Values are unchanged."""
        self.to(memory_format=torch.channels_last if enable else torch.contiguous_format)
        self.channels_last = enable
        return self

    # ------------------------------------------------------------------ raft.py:155-173
    def _forward_prior_only(self, feature, dense_motion, img_full):
        grid, occ = dense_motion["deformation"], dense_motion["occlusion"]
        warps, occs = [], []
        g = grid
        for f in feature:
            if grid.shape[2] != f.shape[2]:
                g = _resize(grid.permute(0, 3, 1, 2), f.shape[2:]).permute(0, 2, 3, 1)
                o = _resize(occ, f.shape[2:])
            else:
                g, o = grid, occ
            warps.append(sampling.grid_sample(f, g))
            occs.append(torch.sigmoid(o))
        warp_img = sampling.grid_sample(img_full, g)
        out = self.generator.decode(warps, warp_img, occs)
        occlusion = torch.cat([_resize(o, (self.size, self.size)) for o in occs], dim=3)
        return out, warp_img, occlusion

    def structure_features(self, kp_s, kp_d, img, fused_bias: bool = False):
        """raft.py:177-182: Gaussian key-point maps (+ positional embedding) -> q_d, k_s.
        fused_bias=True (forward's own call) returns (q_d, k_s, q_bias, k_bias) where, on the channels-last inference
        path, q_d / k_s are the head outputs WITHOUT their biases and the biases ride into CorrPyramid."""
        h, w = img.shape[2:]
        if kp_s.shape == kp_d.shape and not (torch.is_grad_enabled() and (kp_s.requires_grad or kp_d.requires_grad)):
            g = torch.ops.mrfa.kp2gaussian(torch.cat([kp_s, kp_d], dim=0), self.pos_embedding, h, w, 0.1)   # one launch for both
            g_s, g_d = g[:kp_s.shape[0]], g[kp_s.shape[0]:]
        else:
            g_s = torch.ops.mrfa.kp2gaussian(kp_s, self.pos_embedding, h, w, 0.1)
            g_d = torch.ops.mrfa.kp2gaussian(kp_d, self.pos_embedding, h, w, 0.1)
        f_s, f_d = self.kp_img(torch.cat([g_s, img], dim=1)), self.kp(g_d)
        if fused_bias and fast_path(self, f_d) and self.channels_last:
            # inference: the 1x1 heads run as plain GEMMs; their biases are added while CorrPyramid packs the operands
            # (one fewer read + write of each (B,256,h,w) map)
            return F.conv2d(f_d, self.kp_head.weight), F.conv2d(f_s, self.kp_img_head.weight), self.kp_head.bias, self.kp_img_head.bias
        q_d, k_s = self.kp_head(f_d), self.kp_img_head(f_s)
        return (q_d, k_s, None, None) if fused_bias else (q_d, k_s)

    def forward(self, kp_s, kp_d, dense_motion, img, img_full):
        if self.auto_channels_last and not self.channels_last and not self.training and img_full.is_cuda:
            self.channels_last_()
        cl = self.channels_last
        if cl and not (fast_path(self, img_full) and _small7_ok(self.generator.first.conv, img_full)):
            # (the tcgen05 small-channel convolution reads any strides and writes NHWC: the 3-channel image then stays NCHW,
            # which is also what the few-channel image warp at the end wants -- no NHWC copy and no copy back)
            img_full = img_full.contiguous(memory_format=torch.channels_last)
        feature = self.generator.encode(img_full)
        if img is None:
            raise RuntimeError("RaftFlow.forward needs `img` (the reference's self.down is commented out, raft.py:102)")
        if self.prior_only:
            return self._forward_prior_only(feature, dense_motion, img_full)
        B = img.shape[0]
        dev = img.device
        h, w, base = self.h, self.w, self.basic_res_index

        # structure correlation volume + pyramid at the basic resolution (raft.py:177-186, :208)
        q_d, k_s, q_bias, k_bias = self.structure_features(kp_s, kp_d, img, fused_bias=True)
        pyramid = CorrPyramid(q_d, k_s, self.scale, q_bias, k_bias)

        prior = dense_motion["deformation"]
        prior_occ = dense_motion["occlusion"]
        init_flow = torch.ops.mrfa.prior_to_flow(prior, float(self.h - 1))                    # raft.py:189-190
        r0 = (self.h // 8, self.w // 8)                       # scale_factor 1/8 == size h//8 under align_corners=True
        flow = _resize(init_flow, r0) / 8.0
        occlusion = _resize(prior_occ, r0)
        prior_nchw = prior.permute(0, 3, 1, 2)
        ident_basic = sampling.coords_grid(B, h, w, dev)

        out_warp_f, out_occlusion, out_warp_f_c, out_occlusion_c = [], [], [], []
        cat_bufs = [None] * self.total_iter if (CAT_SLICES and fast_path(self, img) and cl) else None
        d_f_pre = d_occ_pre = None
        fused_small = img.is_cuda and not torch.is_grad_enabled() and FUSED_CARRY
        for i in range(self.total_iter):
            R = self.size // 32 * 2 ** i
            # ---- correlation features at this level (raft.py:217-243) ----
            if i < base:
                k = 2 ** (base - i)
                coords = (flow + sampling.coords_grid(B, R, R, dev)) * k
                corr = pyramid.block(base - i)(coords, cl)
            elif i == base:
                corr = pyramid.block(0)(flow + ident_basic, cl)
            else:
                flow_sample = _resize(flow, (self.h, self.h)) * 0.5 ** (i - base)
                corr = pyramid.block(0)(flow_sample + ident_basic, cl)
                if not fast_path(self, corr):                   # reference order (differentiable path)
                    corr = _resize(corr, (R, R))
            m_f = self.corr_enc(flow, corr)

            # ---- warps of feature[i]: refined (at flow) and coarse (prior grid) in one pass ----
            occ_res_sig = None
            if i != base:
                prior_grid = _resize(prior_nchw, (R, R)).permute(0, 2, 3, 1).contiguous()
                if fused_small:
                    occ_res, occ_res_sig = None, torch.ops.mrfa.resize_bilinear(prior_occ, R, R, 2)   # sigmoid(resize(.))
                else:
                    occ_res = _resize(prior_occ, (R, R))
            else:
                prior_grid, occ_res = prior, prior_occ
            if cat_bufs is not None and 0 < i < self.num_iter - 1 and feature[i].shape[1] % 4 == 0 \
                    and feature[i].is_contiguous(memory_format=torch.channels_last) and not feature[i].is_contiguous():
                # the coarse warp lands in the upper half of the decoder's cat([y, warp_c]) buffer of this level
                warp_f, buf = torch.ops.mrfa.dual_warp_cat(feature[i], flow, prior_grid)
                warp_c = buf[:, feature[i].shape[1]:]
                cat_bufs[i] = buf
            else:
                warp_f, warp_c = torch.ops.mrfa.dual_warp(feature[i], flow, prior_grid)   # raft.py:247, :271
            warp_f = conv_relu(self.to_context[i], warp_f)

            d_flow, _ = self.refine(m_f, warp_f)
            if fused_small:
                # raft.py:256-262: both adds and the sigmoid in one pass over the (strided) update
                flow_w, occlusion, occ_sig = torch.ops.mrfa.flow_update(flow, occlusion, d_flow)
            else:
                flow_w = flow + d_flow[:, 0:2]
                d_occ = d_flow[:, 2:]
                occlusion = occlusion + d_occ
                occ_sig = torch.sigmoid(occlusion)

            out_warp_f.append(sampling.warp_by_flow(feature[i], flow_w))                     # raft.py:260
            out_occlusion.append(occ_sig)
            out_warp_f_c.append(warp_c)
            out_occlusion_c.append(occ_res_sig if occ_res_sig is not None else torch.sigmoid(occ_res))

            # ---- carry flow / occlusion to the next resolution (raft.py:276-295) ----
            if i < self.num_iter - 1:
                R2 = 2 * R
                scale = 2 ** (base - i) / 2.0
                if flow.is_cuda and not torch.is_grad_enabled() and FUSED_CARRY:
                    # the whole update below as one kernel
                    flow, occlusion, d_f_pre, d_occ_pre = torch.ops.mrfa.flow_carry(
                        d_flow, init_flow, prior_occ, d_f_pre, d_occ_pre, scale, bool(cl))
                    continue
                d_f = _resize(d_flow[:, 0:2], (R2, R2)) * 2
                flow = d_f + _resize(init_flow, (R2, R2)) / scale
                d_o = _resize(d_occ, (R2, R2))
                occlusion = d_o + _resize(prior_occ, (R2, R2))
                if i == 0:
                    d_f_pre, d_occ_pre = d_f, d_o
                else:
                    up_f = _resize(d_f_pre, (R2, R2)) * 2
                    up_o = _resize(d_occ_pre, (R2, R2))
                    flow = flow + up_f
                    occlusion = occlusion + up_o
                    d_f_pre, d_occ_pre = d_f + up_f, d_o + up_o

        warp_img = sampling.warp_by_flow(img_full, flow)                                   # raft.py:302
        out = self.generator.decode(out_warp_f, warp_img, out_occlusion, out_warp_f_c, out_occlusion_c, coarse_cat=cat_bufs)
        vis = out_occlusion + [torch.sigmoid(prior_occ)]
        if img_full.is_cuda and not torch.is_grad_enabled() and self.size % 4 == 0:
            occlusion = torch.ops.mrfa.resize_strip(vis, self.size, self.size)           # raft.py:304-306 in one strip
        else:
            occlusion = torch.cat([_resize(o, (self.size, self.size)) for o in vis], dim=3)
        return out, warp_img, occlusion
