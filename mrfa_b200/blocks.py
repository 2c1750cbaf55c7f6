"""Dense-convolution building blocks that surround the hot path (they stay cuDNN / stock
PyTorch: SURVEY.md section 2 marks them out of scope for hand-written kernels).

Only the *parameter names and shapes* are part of the contract: reference checkpoints
(``logger.py:50-58``) must load into the drop-in modules unchanged, so the attribute names
below (``conv``, ``norm``, ``conv1``, ``norm1``, ``down_blocks`` ...) follow the state_dict of
the reference blocks (``modules/util.py:111-326``, ``modules/generator.py:8-32``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
import os

from torch import nn

from . import ops

# Inference fast path (CUDA, eval mode, grad disabled): BatchNorm folded into the preceding
# convolution, bias + ReLU fused into the cuDNN convolution (aten::cudnn_convolution_relu) and
# the remaining per-channel passes done by one vectorised kernel (mrfa::channel_affine).  The
# parameters themselves are never modified, so state_dicts stay reference-compatible; folded
# tensors are cached per module and refreshed when any source tensor changes.
FAST_INFERENCE = os.environ.get("MRFA_FAST_CONV", "1") != "0"


def _conv(cin, cout, k, pad, groups=1):
    return nn.Conv2d(cin, cout, kernel_size=k, padding=pad, groups=groups)


def fast_path(module: nn.Module, x: torch.Tensor) -> bool:
    return FAST_INFERENCE and x.is_cuda and x.dtype == torch.float32 and not module.training and not torch.is_grad_enabled()


class _Cache:
    __slots__ = ("key", "val")

    def __init__(self):
        self.key, self.val = None, None

    def get(self, tensors, build):
        key = tuple((t.data_ptr(), t._version) for t in tensors if t is not None)
        if key != self.key:
            with torch.no_grad():
                self.val = build()
            self.key = key
        return self.val


def invalidate_caches(module: nn.Module) -> None:
    """Drop every folded / merged / packed weight cached for the inference fast path under `module`.

    The caches key on (data_ptr, _version) of their source tensors, which in-place writes through ``.data`` (EMA, weight
    averaging) do not bump.  Called automatically after ``load_state_dict`` and on ``train()`` of the drop-in networks;
    call it by hand after editing parameters through ``.data``."""
    for m in module.modules():
        for v in vars(m).values():
            if isinstance(v, _Cache):
                v.key, v.val = None, None


def install_cache_hooks(module: nn.Module) -> None:
    """load_state_dict post-hook + train() invalidation for a top-level drop-in network."""
    module.register_load_state_dict_post_hook(lambda mod, _incompatible: invalidate_caches(mod))


def _bn_affine(bn: nn.BatchNorm2d):
    """Eval-mode BatchNorm as y = x * scale + shift."""
    scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
    return scale, bn.bias - bn.running_mean * scale


def _fold(conv: nn.Conv2d, bn: nn.BatchNorm2d):
    """conv followed by eval-mode BatchNorm == one conv with these weights / bias."""
    scale, shift = _bn_affine(bn)
    w = conv.weight * scale.view(-1, 1, 1, 1)
    if conv.weight.dim() == 4 and conv.weight.is_contiguous(memory_format=torch.channels_last):
        w = w.contiguous(memory_format=torch.channels_last)
    b = shift if conv.bias is None else conv.bias * scale + shift
    return w, b.contiguous()


S2D_FINAL = os.environ.get("MRFA_S2D_FINAL", "1") != "0"      # A/B switch for the space-to-depth final convolution
S2D_BLOCK = 4
HG_SUBPIXEL = os.environ.get("MRFA_HG_SUBPIXEL", "1") != "0"  # A/B switch for the hourglass sub-pixel up-blocks
SMALL_CONV = os.environ.get("MRFA_SMALL_CONV", "1") != "0"    # A/B switch for mrfa::conv7x7_small
FUSED_TAIL = os.environ.get("MRFA_FUSED_TAIL", "1") != "0"      # A/B switch: pixel shuffle + bias + sigmoid + last blend as one kernel
FOLD_CB_BIAS = os.environ.get("MRFA_FOLD_CB_BIAS", "1") != "0"  # A/B switch: ChannelBlock2d bias folded into the ResBlock2d behind it


def _small7_ok(conv: nn.Conv2d, x: torch.Tensor) -> bool:
    """7x7 / pad 3 / stride 1 with 2-3 input channels: served by the tcgen05 TF32 kernel instead of cuDNN's
    legacy indexed path (raft.py:56,63 convf1, generator.py:13 first)."""
    # the kernel multiplies in TF32 (tcgen05 kind::tf32): only when the caller allows TF32 convolutions, like cuDNN
    return (SMALL_CONV and torch.backends.cudnn.allow_tf32 and tuple(conv.kernel_size) == (7, 7) and tuple(conv.padding) == (3, 3)
            and tuple(conv.stride) == (1, 1) and tuple(conv.dilation) == (1, 1) and conv.groups == 1
            and ops.conv7x7_small_ok(x, conv.in_channels, conv.out_channels))


def conv_relu(conv: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """relu(conv(x)); bias and ReLU ride in the convolution epilogue on the fast path."""
    if fast_path(conv, x):
        if _small7_ok(conv, x):
            if not hasattr(conv, "_pk"):
                conv._pk = _Cache()
            wp = conv._pk.get((conv.weight,), lambda: ops.conv7x7_small_pack(conv.weight))
            return torch.ops.mrfa.conv7x7_small(x, wp, conv.bias, True)
        return torch.cudnn_convolution_relu(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)
    return F.relu(conv(x))


class _ConvNormAct(nn.Module):
    """conv -> BatchNorm -> ReLU, optionally preceded by x2 nearest upsampling or followed by
    a 2x2 average pool.  Base of the Same/Up/Down blocks (util.py:160-214)."""

    pre_upsample = False
    post_pool = False

    def __init__(self, in_features, out_features, kernel_size=3, padding=1, groups=1):
        super().__init__()
        self.conv = _conv(in_features, out_features, kernel_size, padding, groups)
        self.norm = nn.BatchNorm2d(out_features, affine=True)
        self._folded = _Cache()

    def forward(self, x):
        if self.pre_upsample:
            x = F.interpolate(x, scale_factor=2)
        if fast_path(self, x):
            c, n = self.conv, self.norm
            w, b = self._folded.get((c.weight, c.bias, n.weight, n.bias, n.running_mean, n.running_var), lambda: _fold(c, n))
            if _small7_ok(c, x):
                if not hasattr(self, "_pk"):
                    self._pk = _Cache()
                x = torch.ops.mrfa.conv7x7_small(x, self._pk.get((w,), lambda: ops.conv7x7_small_pack(w)), b, True)
            else:
                x = torch.cudnn_convolution_relu(x, w, b, c.stride, c.padding, c.dilation, c.groups)
        else:
            x = F.relu(self.norm(self.conv(x)))
        if self.post_pool:
            if (FAST_INFERENCE and x.is_cuda and x.dtype == torch.float32          # training too: the op has a backward
                    and x.shape[1] % 4 == 0 and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0
                    and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()):
                x = torch.ops.mrfa.avg_pool2x2_nhwc(x)
            else:
                x = F.avg_pool2d(x, (2, 2))
        return x


class SameBlock2d(_ConvNormAct):
    def __init__(self, in_features, out_features, groups=1, kernel_size=3, padding=1):
        super().__init__(in_features, out_features, kernel_size, padding, groups)


class UpBlock2d(_ConvNormAct):
    pre_upsample = True

    def subpixel_ok(self, x):
        c = self.conv
        return (fast_path(self, x) and tuple(c.kernel_size) == (3, 3) and tuple(c.padding) == (1, 1)
                and tuple(c.stride) == (1, 1) and c.groups == 1 and c.out_channels % 4 == 0
                and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous())

    def subpixel_weights(self):
        """(4*Cout, Cin, 2, 2) weights and (4*Cout,) bias of the sub-pixel form (BatchNorm folded), cached."""
        c, n = self.conv, self.norm
        if not hasattr(self, "_sub"):
            self._sub = _Cache()

        def build():
            w, b = _fold(c, n)
            R = [torch.tensor([[1., 0., 0.], [0., 1., 1.]], device=w.device), torch.tensor([[1., 1., 0.], [0., 0., 1.]], device=w.device)]
            w2 = torch.cat([torch.einsum("pi,ocij,qj->ocpq", R[a], w, R[q]) for a in (0, 1) for q in (0, 1)], dim=0)
            return w2.contiguous(memory_format=torch.channels_last), b.repeat(4).contiguous()

        return self._sub.get((c.weight, c.bias, n.weight, n.bias, n.running_mean, n.running_var), build)

    def forward_subpixel(self, x):
        """relu(norm(conv(upsample_x2_nearest(x)))) as ONE 2x2 convolution on the padded low-res
        input with 4*Cout phase-major outputs (B, 4*Cout, H+1, W+1): output parity (a,b) of pixel
        (Y,X) lives at [Y//2 + a, X//2 + b, (2a+b)*Cout + c].  3x3 rows {0,1,2} over a nearest x2
        map collapse to low-res rows {y-1, y, y} (a=0) or {y, y, y+1} (a=1): 16/36 of the FLOPs and
        no upsampled tensor.  Consumed by mrfa::occlusion_blend_subpixel."""
        w2, b2 = self.subpixel_weights()
        # 2x2 kernel with padding 1: (H, W) -> (H+1, W+1), the zero border supplied by the convolution itself
        return torch.cudnn_convolution_relu(x, w2, b2, (1, 1), (1, 1), (1, 1), 1)


class DownBlock2d(_ConvNormAct):
    post_pool = True


class ResBlock2d(nn.Module):
    """Pre-activation residual block (util.py:135-157)."""

    def __init__(self, in_features, kernel_size, padding):
        super().__init__()
        self.conv1 = _conv(in_features, in_features, kernel_size, padding)
        self.conv2 = _conv(in_features, in_features, kernel_size, padding)
        self.norm1 = nn.BatchNorm2d(in_features, affine=True)
        self.norm2 = nn.BatchNorm2d(in_features, affine=True)
        self._pre, self._folded = _Cache(), _Cache()

    def forward(self, x):
        if fast_path(self, x):
            n1, n2, c1, c2 = self.norm1, self.norm2, self.conv1, self.conv2
            s1, h1 = self._pre.get((n1.weight, n1.bias, n1.running_mean, n1.running_var), lambda: _bn_affine(n1))
            w, b = self._folded.get((c1.weight, c1.bias, n2.weight, n2.bias, n2.running_mean, n2.running_var), lambda: _fold(c1, n2))
            t = torch.ops.mrfa.channel_affine(x, s1, h1, None, 1)                       # norm1 + relu
            t = torch.cudnn_convolution_relu(t, w, b, c1.stride, c1.padding, c1.dilation, c1.groups)   # conv1 + norm2 + relu
            t = F.conv2d(t, c2.weight, None, c2.stride, c2.padding, c2.dilation, c2.groups)
            return torch.ops.mrfa.channel_affine(t, None, c2.bias, x, 0)                # + bias + residual
        y = self.conv1(F.relu(self.norm1(x)))
        y = self.conv2(F.relu(self.norm2(y)))
        return y + x

    def forward_prebias(self, t, bias):
        """forward(t + bias[c]) without materialising the sum (inference fast path): the per-channel bias of the producing
        convolution folds into the norm1 shift (relu(s (t + b) + h) = relu(s t + (s b + h))) and into the shift of the
        closing residual pass (v + b2 + (t + b) = v + (b2 + b) + t), which saves one read + write of the whole map."""
        n1, n2, c1, c2 = self.norm1, self.norm2, self.conv1, self.conv2
        if not hasattr(self, "_prebias"):
            self._prebias = _Cache()

        def build():
            s1, h1 = _bn_affine(n1)
            return s1, s1 * bias + h1, c2.bias + bias

        s1, h1b, b2b = self._prebias.get((n1.weight, n1.bias, n1.running_mean, n1.running_var, c2.bias, bias), build)
        w, b = self._folded.get((c1.weight, c1.bias, n2.weight, n2.bias, n2.running_mean, n2.running_var), lambda: _fold(c1, n2))
        a = torch.ops.mrfa.channel_affine(t, s1, h1b, None, 1)                          # norm1(t + bias) + relu
        a = torch.cudnn_convolution_relu(a, w, b, c1.stride, c1.padding, c1.dilation, c1.groups)
        a = F.conv2d(a, c2.weight, None, c2.stride, c2.padding, c2.dilation, c2.groups)
        return torch.ops.mrfa.channel_affine(a, None, b2b, t, 0)                        # + conv2 bias + (t + bias)


class ChannelBlock2d(nn.Module):
    """BN -> ReLU -> conv halving the channel count (util.py:111-133)."""

    def __init__(self, in_features, kernel_size, padding):
        super().__init__()
        self.conv1 = _conv(in_features, in_features // 2, kernel_size, padding)
        self.norm1 = nn.BatchNorm2d(in_features, affine=True)
        self._pre = _Cache()

    def forward(self, x):
        if fast_path(self, x):
            n1, c1 = self.norm1, self.conv1
            s1, h1 = self._pre.get((n1.weight, n1.bias, n1.running_mean, n1.running_var), lambda: _bn_affine(n1))
            t = torch.ops.mrfa.channel_affine(x, s1, h1, None, 1)
            t = F.conv2d(t, c1.weight, None, c1.stride, c1.padding, c1.dilation, c1.groups)
            return torch.ops.mrfa.channel_affine(t, None, c1.bias, None, 0)
        return self.conv1(F.relu(self.norm1(x)))

    def forward_nobias(self, x):
        """(conv1(relu(norm1(x))) WITHOUT its bias, the bias): the ResBlock2d behind it absorbs the bias (forward_prebias)."""
        n1, c1 = self.norm1, self.conv1
        s1, h1 = self._pre.get((n1.weight, n1.bias, n1.running_mean, n1.running_var), lambda: _bn_affine(n1))
        t = torch.ops.mrfa.channel_affine(x, s1, h1, None, 1)
        return F.conv2d(t, c1.weight, None, c1.stride, c1.padding, c1.dilation, c1.groups), c1.bias


class _HGEncoder(nn.Module):
    def __init__(self, block_expansion, in_features, num_blocks, max_features):
        super().__init__()
        widths = [in_features] + [min(max_features, block_expansion * 2 ** (i + 1)) for i in range(num_blocks)]
        self.down_blocks = nn.ModuleList(
            DownBlock2d(widths[i], widths[i + 1], kernel_size=3, padding=1) for i in range(num_blocks))

    def forward(self, x):
        feats = [x]
        for blk in self.down_blocks:
            feats.append(blk(feats[-1]))
        return feats


class _HGDecoder(nn.Module):
    def __init__(self, block_expansion, in_features, num_blocks, max_features):
        super().__init__()
        ups = []
        for i in reversed(range(num_blocks)):
            cin = min(max_features, block_expansion * 2 ** (i + 1)) * (1 if i == num_blocks - 1 else 2)
            ups.append(UpBlock2d(cin, min(max_features, block_expansion * 2 ** i), kernel_size=3, padding=1))
        self.up_blocks = nn.ModuleList(ups)
        self.out_filters = block_expansion + in_features

    def forward(self, feats):
        feats = list(feats)
        y = feats.pop()
        for blk in self.up_blocks:
            skip = feats.pop()
            if HG_SUBPIXEL and blk.subpixel_ok(y) and tuple(skip.shape[2:]) == (2 * y.shape[2], 2 * y.shape[3]):
                # sub-pixel up-convolution; its de-interleave and the cat are one kernel
                y = torch.ops.mrfa.subpixel_shuffle_cat(blk.forward_subpixel(y), skip)
            else:
                y = torch.cat([blk(y), skip], dim=1)
        return y


class Hourglass(nn.Module):
    """U-net without the final projection (util.py:217-278)."""

    def __init__(self, block_expansion, in_features, num_blocks=3, max_features=256):
        super().__init__()
        self.encoder = _HGEncoder(block_expansion, in_features, num_blocks, max_features)
        self.decoder = _HGDecoder(block_expansion, in_features, num_blocks, max_features)
        self.out_filters = self.decoder.out_filters

    def forward(self, x):
        return self.decoder(self.encoder(x))


class AntiAliasInterpolation2d(nn.Module):
    """Gaussian low-pass followed by nearest sub-sampling (util.py:282-326)."""

    def __init__(self, channels, scale):
        super().__init__()
        sigma = (1 / scale - 1) / 2
        ksize = 2 * round(sigma * 4) + 1
        self.ka = ksize // 2
        self.kb = self.ka - 1 if ksize % 2 == 0 else self.ka
        ax = torch.arange(ksize, dtype=torch.float32)
        mean = (ksize - 1) / 2
        g1 = torch.exp(-(ax - mean) ** 2 / (2 * sigma ** 2)) if sigma > 0 else torch.ones(1)
        kernel = g1[:, None] * g1[None, :]
        kernel = kernel / torch.sum(kernel)
        self.register_buffer("weight", kernel[None, None].repeat(channels, 1, 1, 1))
        self.groups = channels
        self.scale = scale

    def forward(self, x):
        if self.scale == 1.0:
            return x
        stride = round(1 / self.scale)
        if (fast_path(self, x) and self.ka == self.kb and abs(stride * self.scale - 1) < 1e-9
                and x.shape[2] % stride == 0 and x.shape[3] % stride == 0):
            # evaluate the Gaussian only at the pixels the nearest sub-sampling keeps
            return torch.ops.mrfa.antialias_down(x, self.weight, self.ka, stride)
        y = F.pad(x, (self.ka, self.kb, self.ka, self.kb))
        y = F.conv2d(y, weight=self.weight, groups=self.groups)
        return F.interpolate(y, scale_factor=(self.scale, self.scale))


class OcclusionAwareGenerator(nn.Module):
    """Feature-pyramid encoder and occlusion-blended decoder that *consume* the hot-path
    warps (generator.py:8-64).  Convolutions only; the warps are produced by RaftFlow."""

    def __init__(self, num_channels, block_expansion, max_features, num_up_blocks):
        super().__init__()
        self.num_up_blocks = num_up_blocks
        self.first = SameBlock2d(num_channels, block_expansion, kernel_size=(7, 7), padding=(3, 3))
        widths = [min(max_features, block_expansion * 2 ** i) for i in range(num_up_blocks + 1)]
        self.down_blocks = nn.ModuleList(
            DownBlock2d(widths[i], widths[i + 1], kernel_size=(3, 3), padding=(1, 1)) for i in range(num_up_blocks))
        rev = list(reversed(range(num_up_blocks)))
        self.up_blocks = nn.ModuleList(
            UpBlock2d(widths[i + 1], widths[i], kernel_size=(3, 3), padding=(1, 1)) for i in rev)
        self.resblock = nn.ModuleList(ResBlock2d(widths[i + 1], kernel_size=(3, 3), padding=(1, 1)) for i in rev)
        self.channel_block = nn.ModuleList(
            ChannelBlock2d(widths[i + 1] * 2, kernel_size=(3, 3), padding=(1, 1)) for i in rev)
        self.final = _conv(block_expansion, num_channels, (7, 7), (3, 3))

    def encode(self, x):
        feats = [self.first(x)]
        for blk in self.down_blocks:
            feats.append(blk(feats[-1]))
        return feats[::-1]          # coarsest first: R = S/32 ... S

    def decode(self, warp_f, warp_img, occlusion, warp_f_c=None, occlusion_c=None, coarse_cat=None):
        """`coarse_cat[i]` (optional, inference): the (N,2C,H,W) buffer of mrfa::dual_warp_cat whose upper half is
        warp_f_c[i]; the blend of that level is then written into its lower half instead of a cat."""
        if fast_path(self, warp_img):
            return self._decode_fast(warp_f, warp_img, occlusion, warp_f_c, coarse_cat)
        use_coarse = warp_f_c is not None
        y = warp_f[0] * occlusion[0]
        if use_coarse:
            y = torch.cat([y, warp_f_c[0]], dim=1)
        for i in range(self.num_up_blocks):
            if use_coarse:
                y = self.channel_block[i](y)
            y = self.up_blocks[i](self.resblock[i](y))
            occ = occlusion[i + 1]
            y = warp_f[i + 1] * occ + y * (1 - occ)
            if use_coarse and i != self.num_up_blocks - 1:
                y = torch.cat([y, warp_f_c[i + 1]], dim=1)
        y = torch.sigmoid(self.final(y))
        return y * (1 - occlusion[-1]) + warp_img * occlusion[-1]

    def _final_s2d(self, ys, fused_tail=None):
        """self.final (7x7, C -> 3, pad 3) on a 4x4 space-to-depth input (N, 16C, H/4, W/4): the same sums as a
        3x3 / pad 1 convolution with 48 outputs (weights re-indexed, taps outside the 7x7 window zero) followed by
        a pixel shuffle.  The library's 7x7 kernel tiles N = 3 outputs into a 64-wide tile (3.9 ms per batch of
        64 at 256x256); the 3x3 form runs in 0.55 ms."""
        c = self.final
        if not hasattr(self, "_s2d"):
            self._s2d = _Cache()

        def build():
            w, r, K = c.weight, S2D_BLOCK, 7
            Co, Ci = w.shape[:2]
            w2 = w.new_zeros(Co, r, r, r, r, Ci, 3, 3)       # [co, oy, ox, iy, ix, ci, by, bx]
            for by in range(3):
                for iy in range(r):
                    for oy in range(r):
                        ky = r * (by - 1) + iy - oy + K // 2
                        if 0 <= ky < K:
                            for bx in range(3):
                                for ix in range(r):
                                    for ox in range(r):
                                        kx = r * (bx - 1) + ix - ox + K // 2
                                        if 0 <= kx < K:
                                            w2[:, oy, ox, iy, ix, :, by, bx] = w[:, :, ky, kx]
            w2 = w2.reshape(Co * r * r, r * r * Ci, 3, 3).contiguous(memory_format=torch.channels_last)
            return w2, c.bias.repeat_interleave(r * r).contiguous()

        w2, b2 = self._s2d.get((c.weight, c.bias), build)
        if fused_tail is not None:
            # pixel shuffle + bias + sigmoid + the last occlusion blend in one pass over the 48-channel convolution output
            warp_img, occ = fused_tail
            return torch.ops.mrfa.final_blend_s2d(F.conv2d(ys, w2, None, padding=1), c.bias, warp_img, occ, S2D_BLOCK)
        return F.pixel_shuffle(F.conv2d(ys, w2, b2, padding=1), S2D_BLOCK)

    def _decode_fast(self, warp_f, warp_img, occlusion, warp_f_c, coarse_cat=None):
        """Same dataflow as decode(); the occlusion blends are single fused passes."""
        blend = torch.ops.mrfa.occlusion_blend
        use_coarse = warp_f_c is not None
        c = self.final
        s2d_final = (S2D_FINAL and tuple(c.kernel_size) == (7, 7) and tuple(c.padding) == (3, 3) and c.bias is not None
                     and tuple(c.stride) == (1, 1) and c.groups == 1)
        y = blend(warp_f[0], None, occlusion[0])
        if use_coarse:
            y = torch.cat([y, warp_f_c[0]], dim=1)
        for i in range(self.num_up_blocks):
            cb = self.channel_block[i]
            if use_coarse and FOLD_CB_BIAS and cb.conv1.bias is not None and self.resblock[i].conv2.bias is not None:
                t, cbias = cb.forward_nobias(y)                  # the ChannelBlock's bias pass folds into the ResBlock
                y = self.resblock[i].forward_prebias(t, cbias)
            else:
                if use_coarse:
                    y = cb(y)
                y = self.resblock[i](y)
            up = self.up_blocks[i]
            if up.subpixel_ok(y) and warp_f[i + 1].is_contiguous(memory_format=torch.channels_last) \
                    and not warp_f[i + 1].is_contiguous():
                last = i == self.num_up_blocks - 1
                if last and s2d_final and warp_f[i + 1].shape[2] % S2D_BLOCK == 0 and warp_f[i + 1].shape[3] % S2D_BLOCK == 0:
                    # the last blend writes its result in 4x4 space-to-depth order for the final convolution
                    ys = torch.ops.mrfa.occlusion_blend_subpixel(warp_f[i + 1], up.forward_subpixel(y), occlusion[i + 1], S2D_BLOCK)
                    if FUSED_TAIL and warp_img.is_contiguous():              # NCHW planes (the few-channel image warp's output)
                        return self._final_s2d(ys, (warp_img, occlusion[-1]))
                    return blend(warp_img, torch.sigmoid(self._final_s2d(ys)), occlusion[-1])
                buf = coarse_cat[i + 1] if (coarse_cat is not None and use_coarse and not last) else None
                if buf is not None and buf.shape[1] == 2 * warp_f[i + 1].shape[1]:
                    torch.ops.mrfa.occlusion_blend_subpixel_into(warp_f[i + 1], up.forward_subpixel(y), occlusion[i + 1], buf)
                    y = buf                                   # == cat([blend, warp_f_c[i + 1]], 1), no copy
                    continue
                y = torch.ops.mrfa.occlusion_blend_subpixel(warp_f[i + 1], up.forward_subpixel(y), occlusion[i + 1], 1)
            else:
                y = blend(warp_f[i + 1], up(y), occlusion[i + 1])
            if use_coarse and i != self.num_up_blocks - 1:
                y = torch.cat([y, warp_f_c[i + 1]], dim=1)
        y = torch.sigmoid(self.final(y))
        return blend(warp_img, y, occlusion[-1])

    def forward(self, x):
        return self.decode(self.encode(x))
