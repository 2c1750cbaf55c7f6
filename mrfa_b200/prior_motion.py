"""Prior dense-motion networks with the sparse-motion construction and deformed-source
stacking running as fused kernels (reference: modules/dense_motion.py).

Same constructor kwargs (config/vox1.yaml:17-24, :37-43), forward signature, output dict keys
and state_dict keys as the reference classes, so reference checkpoints load unchanged.
The hourglass / mask / occlusion convolutions stay stock PyTorch (cuDNN).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .blocks import AntiAliasInterpolation2d, Hourglass, install_cache_hooks, invalidate_caches


def _dropout_softmax(X, P):
    """Softmax with transformation dropout (dense_motion.py:87-102 / :245-260; Eq. 7-8 of the
    TPSM paper).  Training-time only; plain tensor ops."""
    drop = (torch.rand(X.shape[0], X.shape[1], device=X.device) < (1 - P)).to(X.dtype)
    drop[..., 0] = 1
    drop = drop[:, :, None, None].expand_as(X)
    X = X - X.max(1, keepdim=True).values
    X_exp = X.exp().masked_fill(drop == 0, 0)
    return X_exp / (X_exp.sum(dim=1, keepdim=True) + 1e-6)


def _combine(motions, weights):
    """deformation = sum_k weights[:,k] * motions[:,k]  (dense_motion.py:132-136, :292-295)."""
    return (motions * weights.unsqueeze(-1)).sum(dim=1)


class DenseMotionNetwork(nn.Module):
    """dense_motion.py:8-146 (FOMM / MTIA prior)."""

    def __init__(self, block_expansion, num_blocks, max_features, num_kp, num_channels, estimate_occlusion_map=True,
                 scale_factor=1, kp_variance=0.01):
        super().__init__()
        self.infeatures = num_kp + 1
        self.hourglass = Hourglass(block_expansion=block_expansion, in_features=self.infeatures * (num_channels + 1),
                                   max_features=max_features, num_blocks=num_blocks)
        self.mask = nn.Conv2d(self.hourglass.out_filters, self.infeatures, kernel_size=(7, 7), padding=(3, 3))
        self.occlusion = nn.Conv2d(self.hourglass.out_filters, 1, kernel_size=(7, 7), padding=(3, 3)) \
            if estimate_occlusion_map else None
        self.num_kp = num_kp
        self.scale_factor = scale_factor
        self.kp_variance = kp_variance
        if self.scale_factor != 1:
            self.down = AntiAliasInterpolation2d(num_channels, self.scale_factor)
        install_cache_hooks(self)

    channels_last = False
    auto_channels_last = True

    def train(self, mode: bool = True):
        invalidate_caches(self)                   # folded inference weights are rebuilt from the live parameters
        return super().train(mode)

    def channels_last_(self, enable: bool = True):
        """Run the hourglass convolutions in NHWC memory (see RaftFlow.channels_last_)."""
        self.to(memory_format=torch.channels_last if enable else torch.contiguous_format)
        self.channels_last = enable
        return self

    # --- the three reference helper methods, each backed by the fused kernel -------------------
    def _prior(self, source_image, kp_driving, kp_source, bg_param):
        jd, js = kp_driving.get("jacobian"), kp_source.get("jacobian")
        if jd is None or js is None:
            jd = js = None
        # differentiable: mrfa::dense_motion_prior carries an autograd formula whose backward is one fused pass
        # (csrc/motion_bwd.cu) -- key-points, Jacobians, background affine and the source all receive gradients
        return torch.ops.mrfa.dense_motion_prior(kp_driving["kp"], kp_source["kp"], jd, js, bg_param, source_image,
                                                 float(self.kp_variance))

    def create_heatmap_representations(self, source_image, kp_driving, kp_source):
        B, C, h, w = source_image.shape
        _, hg = self._prior(source_image, {"kp": kp_driving["kp"]}, {"kp": kp_source["kp"]}, None)
        return hg.view(B, self.num_kp + 1, C + 1, h, w)[:, :, :1]

    def create_sparse_motions(self, source_image, kp_driving, kp_source, bg_param=None):
        return self._prior(source_image, kp_driving, kp_source, bg_param)[0]

    def create_deformed_source_image(self, source_image, sparse_motions):
        B, _, h, w = source_image.shape
        K1 = self.num_kp + 1
        out = torch.ops.mrfa.grid_sample(source_image, sparse_motions.reshape(B * K1, h, w, 2), _lib.COORD_NORM_ACF,
                                         _lib.PAD_ZEROS, False, K1)
        return out.view(B, K1, -1, h, w)

    def forward(self, source_image, kp_driving, kp_source, bg_param=None, dropout_flag=False, dropout_p=0):
        if self.auto_channels_last and not self.channels_last and not self.training and source_image.is_cuda:
            self.channels_last_()
        if self.scale_factor != 1:
            source_image = self.down(source_image)
        B, C, h, w = source_image.shape
        K1 = self.num_kp + 1
        motions, hg_input = self._prior(source_image, kp_driving, kp_source, bg_param)
        out_dict = {"sparse_deformed": hg_input.view(B, K1, C + 1, h, w)[:, :, 1:]}
        prediction = self.hourglass(hg_input)
        mask = self.mask(prediction)
        out_dict["logit_mask"] = mask
        mask = _dropout_softmax(mask, dropout_p) if dropout_flag else F.softmax(mask, dim=1)
        out_dict["mask"] = mask
        out_dict["deformation"] = _combine(motions, mask)
        if self.occlusion is not None:
            out_dict["occlusion"] = self.occlusion(prediction)
        return out_dict


class TPSDenseMotionNetwork(nn.Module):
    """dense_motion.py:150-312 (TPSM prior; ``multi_mask=True`` is unusable in the reference --
    un-imported names at :174,:179 -- and is rejected here)."""

    def __init__(self, block_expansion, num_blocks, max_features, num_tps, num_channels, scale_factor=0.25, bg=False,
                 multi_mask=False, kp_variance=0.01):
        super().__init__()
        if multi_mask:
            raise NotImplementedError("multi_mask=True does not run in the reference either (dense_motion.py:174)")
        if scale_factor != 1:
            self.down = AntiAliasInterpolation2d(num_channels, scale_factor)
        self.scale_factor = scale_factor
        self.multi_mask = multi_mask
        self.hourglass = Hourglass(block_expansion=block_expansion,
                                   in_features=(num_channels * (num_tps + 1) + num_tps * 5 + 1),
                                   max_features=max_features, num_blocks=num_blocks)
        self.maps = nn.Conv2d(self.hourglass.out_filters, num_tps + 1, kernel_size=(7, 7), padding=(3, 3))
        self.occlusion = nn.ModuleList([nn.Conv2d(self.hourglass.out_filters, 1, kernel_size=(7, 7), padding=(3, 3))])
        self.num_tps = num_tps
        self.bg = bg
        self.kp_variance = kp_variance

    def _prior(self, source_image, kp_driving, kp_source, bg_param):
        # solve + synthesis as one op; its backward reduces the motion / heat-map / warp gradients per transformation
        # and runs the adjoint 8x8 solve in fp64 (csrc/motion_bwd.cu)
        motions, hg_input, _, _ = torch.ops.mrfa.tps_prior(kp_driving["kp"], kp_source["kp"], bg_param, source_image,
                                                           float(self.kp_variance))
        return motions, hg_input

    def create_heatmap_representations(self, source_image, kp_driving, kp_source):
        _, hg = self._prior(source_image, kp_driving, kp_source, None)
        return hg[:, :self.num_tps * 5 + 1]

    def create_transformations(self, source_image, kp_driving, kp_source, bg_param):
        return self._prior(source_image, kp_driving, kp_source, bg_param)[0]

    def create_deformed_source_image(self, source_image, transformations):
        B, _, h, w = source_image.shape
        G1 = self.num_tps + 1
        out = torch.ops.mrfa.grid_sample(source_image, transformations.reshape(B * G1, h, w, 2), _lib.COORD_NORM_ACT,
                                         _lib.PAD_ZEROS, False, G1)
        return out.view(B, G1, -1, h, w)

    def forward(self, source_image, kp_driving, kp_source, bg_param=None, dropout_flag=False, dropout_p=0):
        if self.scale_factor != 1:
            source_image = self.down(source_image)
        B, C, h, w = source_image.shape
        G1 = self.num_tps + 1
        motions, hg_input = self._prior(source_image, kp_driving, kp_source, bg_param)
        out_dict = {"deformed_source": hg_input[:, self.num_tps * 5 + 1:].view(B, G1, C, h, w)}
        prediction = self.hourglass(hg_input)
        maps = self.maps(prediction)
        maps = _dropout_softmax(maps, dropout_p) if dropout_flag else F.softmax(maps, dim=1)
        out_dict["contribution_maps"] = maps
        out_dict["mask"] = maps
        out_dict["deformation"] = _combine(motions, maps)
        out_dict["occlusion"] = self.occlusion[0](prediction)
        return out_dict
