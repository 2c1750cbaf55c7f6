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
from torch import nn


def _conv(cin, cout, k, pad, groups=1):
    return nn.Conv2d(cin, cout, kernel_size=k, padding=pad, groups=groups)


class _ConvNormAct(nn.Module):
    """conv -> BatchNorm -> ReLU, optionally preceded by x2 nearest upsampling or followed by
    a 2x2 average pool.  Base of the Same/Up/Down blocks (util.py:160-214)."""

    pre_upsample = False
    post_pool = False

    def __init__(self, in_features, out_features, kernel_size=3, padding=1, groups=1):
        super().__init__()
        self.conv = _conv(in_features, out_features, kernel_size, padding, groups)
        self.norm = nn.BatchNorm2d(out_features, affine=True)

    def forward(self, x):
        if self.pre_upsample:
            x = F.interpolate(x, scale_factor=2)
        x = F.relu(self.norm(self.conv(x)))
        if self.post_pool:
            x = F.avg_pool2d(x, (2, 2))
        return x


class SameBlock2d(_ConvNormAct):
    def __init__(self, in_features, out_features, groups=1, kernel_size=3, padding=1):
        super().__init__(in_features, out_features, kernel_size, padding, groups)


class UpBlock2d(_ConvNormAct):
    pre_upsample = True


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

    def forward(self, x):
        y = self.conv1(F.relu(self.norm1(x)))
        y = self.conv2(F.relu(self.norm2(y)))
        return y + x


class ChannelBlock2d(nn.Module):
    """BN -> ReLU -> conv halving the channel count (util.py:111-133)."""

    def __init__(self, in_features, kernel_size, padding):
        super().__init__()
        self.conv1 = _conv(in_features, in_features // 2, kernel_size, padding)
        self.norm1 = nn.BatchNorm2d(in_features, affine=True)

    def forward(self, x):
        return self.conv1(F.relu(self.norm1(x)))


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
            y = torch.cat([blk(y), feats.pop()], dim=1)
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

    def decode(self, warp_f, warp_img, occlusion, warp_f_c=None, occlusion_c=None):
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

    def forward(self, x):
        return self.decode(self.encode(x))
