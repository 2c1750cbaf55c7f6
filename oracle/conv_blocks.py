"""Plain-PyTorch dense-convolution blocks for the CPU oracle (TEST INFRASTRUCTURE ONLY).

The oracle must not depend on the product package (importing ``mrfa_b200`` loads the CUDA
library), so the convolution stacks that surround the hot path are restated here with stock
``nn.Conv2d`` / ``nn.BatchNorm2d`` only -- no fast paths, no custom ops.  What has to agree with
the reference is (i) the arithmetic of each block and (ii) the state_dict key names, so the
name-keyed synthetic weights (``synthetic_inputs.fill_state_dict_``) and reference checkpoints
load unchanged:

    block kinds            modules/util.py:111-214   (ChannelBlock2d, ResBlock2d, Up/Down/SameBlock2d)
    Hourglass              modules/util.py:217-278   (encoder.down_blocks.*, decoder.up_blocks.*)
    AntiAliasInterpolation modules/util.py:282-326   (buffer ``weight``)
    generator              modules/generator.py:8-64 (first, down_blocks, up_blocks, resblock, channel_block, final)

Pinned by tests/test_oracle_golden.py against outputs of the unmodified reference
(tests/golden/*.npz) and, in the build container, against oracle/_ref directly
(tests/test_reference_vendored.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


def _width(block_expansion: int, max_features: int, level: int) -> int:
    return min(max_features, block_expansion << level)


class ConvUnit(nn.Module):
    """conv -> BatchNorm -> ReLU with an optional resampling step; keys ``conv.*`` / ``norm.*``.

    resample: None (SameBlock2d, util.py:196-214), "up" = nearest x2 *before* the convolution
    (UpBlock2d, util.py:160-176), "down" = 2x2 average pool *after* the ReLU (DownBlock2d,
    util.py:179-194)."""

    def __init__(self, cin, cout, kernel_size=3, padding=1, resample=None):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=kernel_size, padding=padding)
        self.norm = nn.BatchNorm2d(cout, affine=True)
        self.resample = resample

    def forward(self, x):
        if self.resample == "up":
            x = F.interpolate(x, scale_factor=2)
        x = torch.relu(self.norm(self.conv(x)))
        return F.avg_pool2d(x, (2, 2)) if self.resample == "down" else x


class PreActResidual(nn.Module):
    """x + conv2(relu(norm2(conv1(relu(norm1(x)))))) (ResBlock2d, util.py:135-157)."""

    def __init__(self, width, kernel_size, padding):
        super().__init__()
        for i in (1, 2):
            setattr(self, f"conv{i}", nn.Conv2d(width, width, kernel_size=kernel_size, padding=padding))
            setattr(self, f"norm{i}", nn.BatchNorm2d(width, affine=True))

    def forward(self, x):
        y = x
        for i in (1, 2):
            y = getattr(self, f"conv{i}")(torch.relu(getattr(self, f"norm{i}")(y)))
        return x + y


class PreActHalve(nn.Module):
    """conv1(relu(norm1(x))) with half as many output channels (ChannelBlock2d, util.py:111-133)."""

    def __init__(self, width, kernel_size, padding):
        super().__init__()
        self.conv1 = nn.Conv2d(width, width // 2, kernel_size=kernel_size, padding=padding)
        self.norm1 = nn.BatchNorm2d(width, affine=True)

    def forward(self, x):
        return self.conv1(torch.relu(self.norm1(x)))


class _Down(nn.Module):
    def __init__(self, block_expansion, in_features, num_blocks, max_features):
        super().__init__()
        cins = [in_features] + [_width(block_expansion, max_features, i) for i in range(1, num_blocks)]
        self.down_blocks = nn.ModuleList(
            ConvUnit(cins[i], _width(block_expansion, max_features, i + 1), resample="down") for i in range(num_blocks))


class _Up(nn.Module):
    def __init__(self, block_expansion, in_features, num_blocks, max_features):
        super().__init__()
        self.up_blocks = nn.ModuleList()
        for lvl in range(num_blocks - 1, -1, -1):
            cin = _width(block_expansion, max_features, lvl + 1) * (2 if lvl < num_blocks - 1 else 1)
            self.up_blocks.append(ConvUnit(cin, _width(block_expansion, max_features, lvl), resample="up"))


class Hourglass(nn.Module):
    """U-net whose output is cat(up-path, input) (util.py:217-278); ``out_filters`` as the reference."""

    def __init__(self, block_expansion, in_features, num_blocks=3, max_features=256):
        super().__init__()
        self.encoder = _Down(block_expansion, in_features, num_blocks, max_features)
        self.decoder = _Up(block_expansion, in_features, num_blocks, max_features)
        self.out_filters = block_expansion + in_features

    def forward(self, x):
        skips = [x]
        for blk in self.encoder.down_blocks:
            skips.append(blk(skips[-1]))
        y = skips.pop()
        for blk in self.decoder.up_blocks:
            y = torch.cat([blk(y), skips.pop()], dim=1)
        return y


class AntiAliasInterpolation2d(nn.Module):
    """Depthwise Gaussian blur + nearest sub-sampling (util.py:282-326)."""

    def __init__(self, channels, scale):
        super().__init__()
        sigma = (1 / scale - 1) / 2
        k = 2 * round(sigma * 4) + 1
        self.pad = (k // 2, k // 2 - 1 if k % 2 == 0 else k // 2)
        # separable product of two 1-D Gaussians evaluated on the integer taps, normalised to sum 1
        # (same op order as the reference: product accumulated axis by axis, then one division)
        taps = torch.arange(k, dtype=torch.float32)
        yy, xx = torch.meshgrid(taps, taps, indexing="ij")
        kernel = torch.ones(k, k)
        if sigma > 0:                                      # scale == 1 never filters (forward returns x)
            for axis in (yy, xx):
                kernel = kernel * torch.exp(-(axis - (k - 1) / 2) ** 2 / (2 * sigma ** 2))
        kernel = kernel / kernel.sum()
        self.register_buffer("weight", kernel.view(1, 1, k, k).repeat(channels, 1, 1, 1))
        self.channels, self.scale = channels, scale

    def forward(self, x):
        if self.scale == 1.0:
            return x
        a, b = self.pad
        y = F.conv2d(F.pad(x, (a, b, a, b)), self.weight, groups=self.channels)
        return F.interpolate(y, scale_factor=(self.scale, self.scale))


class OcclusionAwareGenerator(nn.Module):
    """Feature-pyramid encoder + occlusion-blended decoder (generator.py:8-64)."""

    def __init__(self, num_channels, block_expansion, max_features, num_up_blocks):
        super().__init__()
        n = self.num_up_blocks = num_up_blocks
        w = [_width(block_expansion, max_features, i) for i in range(n + 1)]
        self.first = ConvUnit(num_channels, block_expansion, kernel_size=(7, 7), padding=(3, 3))
        self.down_blocks = nn.ModuleList(ConvUnit(w[i], w[i + 1], (3, 3), (1, 1), "down") for i in range(n))
        order = range(n - 1, -1, -1)                       # decoder runs coarse -> fine
        self.up_blocks = nn.ModuleList(ConvUnit(w[i + 1], w[i], (3, 3), (1, 1), "up") for i in order)
        self.resblock = nn.ModuleList(PreActResidual(w[i + 1], (3, 3), (1, 1)) for i in order)
        self.channel_block = nn.ModuleList(PreActHalve(2 * w[i + 1], (3, 3), (1, 1)) for i in order)
        self.final = nn.Conv2d(block_expansion, num_channels, kernel_size=(7, 7), padding=(3, 3))

    def encode(self, x):
        """-> feature maps, coarsest first (generator.py:34-42)."""
        feats = [self.first(x)]
        for blk in self.down_blocks:
            feats.append(blk(feats[-1]))
        feats.reverse()
        return feats

    def decode(self, warp_f, warp_img, occlusion, warp_f_c=None, occlusion_c=None):
        """generator.py:44-64: blend the refined warps level by level, coarse warps concatenated
        (then halved by channel_block) when given."""
        coarse = warp_f_c is not None
        y = warp_f[0] * occlusion[0]
        for i in range(self.num_up_blocks):
            if coarse:
                y = self.channel_block[i](torch.cat([y, warp_f_c[i]], dim=1))
            y = self.up_blocks[i](self.resblock[i](y))
            o = occlusion[i + 1]
            y = warp_f[i + 1] * o + y * (1 - o)
        y = torch.sigmoid(self.final(y))
        return y * (1 - occlusion[-1]) + warp_img * occlusion[-1]
