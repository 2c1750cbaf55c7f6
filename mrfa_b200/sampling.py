"""Drop-in replacements for the sampling / grid primitives of the reference's
``modules/util.py`` (same names, argument meaning and error behaviour), running on the
hand-written kernels of libmrfa_b200.  CUDA tensors only: there is no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib, ops


def _device_of_type_string(type_str, default=None):
    """The reference passes legacy type strings (``'torch.cuda.FloatTensor'``, util.py:67,95)."""
    if isinstance(type_str, torch.device):
        return type_str
    if isinstance(type_str, str):
        if "cuda" in type_str:
            return torch.device("cuda", torch.cuda.current_device())
        if default is not None:
            return default
        raise RuntimeError(f"mrfa_b200: tensor type {type_str!r} is not a CUDA type (there is no CPU fallback)")
    if isinstance(type_str, torch.dtype):
        return torch.device("cuda", torch.cuda.current_device())
    raise TypeError(type_str)


def bilinear_sampler(img, coords, mode="bilinear", mask=False):
    """util.py:26-38 -- bilinear warp at *pixel* coordinates (align_corners=True, zeros padding).

    img (N,C,H,W), coords (N,Ho,Wo,2) -> (N,C,Ho,Wo).  ``mode`` is ignored, as in the reference.
    """
    out = torch.ops.mrfa.grid_sample(img, coords, _lib.COORD_PIXEL, _lib.PAD_ZEROS, False, 1)
    if mask:
        H, W = img.shape[-2:]
        xg = 2 * coords[..., 0:1] / (W - 1) - 1
        yg = 2 * coords[..., 1:2] / (H - 1) - 1
        m = (xg > -1) & (yg > -1) & (xg < 1) & (yg < 1)
        return out, m.float()
    return out


def batch_bilinear_sampler(img, coords, mode="bilinear", mask=False, h=256, w=256, mini_batch=4):
    """util.py:40-51.  The reference chunks the call to bound grid_sample's memory and silently
    drops the remainder chunk when ``batch % mini_batch != 0``; one kernel launch over the kept
    prefix reproduces that exactly."""
    batch = img.shape[0] // (h * w)
    keep = (batch // mini_batch) * mini_batch * h * w
    return bilinear_sampler(img[:keep], coords[:keep])


def warp_by_flow(feature, flow):
    """``bilinear_sampler(feature, (flow + coords_grid).permute(0,2,3,1))`` (raft.py:247,260,302)
    with the identity-grid add fused into the kernel; flow (B,2,H,W) in pixels."""
    return torch.ops.mrfa.grid_sample(feature, flow.permute(0, 2, 3, 1), _lib.COORD_PIXEL, _lib.PAD_ZEROS, True, 1)


def grid_sample(inp, grid, align_corners=False, padding_mode="zeros"):
    """``F.grid_sample(inp, grid, mode='bilinear')`` for the conventions the reference uses
    (raft.py:166,168,271; dense_motion.py:83,241; model.py:48)."""
    mode = _lib.COORD_NORM_ACT if align_corners else _lib.COORD_NORM_ACF
    pad = {"zeros": _lib.PAD_ZEROS, "reflection": _lib.PAD_REFLECTION}[padding_mode]
    return torch.ops.mrfa.grid_sample(inp, grid, mode, pad, False, 1)


def deform_input(inp, deformation):
    """FOMM-named feature warp: resize the (B,h,w,2) deformation to the feature size if needed
    (bilinear, align_corners=True), then sample with align_corners=False.  MRFA inlines this at
    raft.py:160-166 and :265-271."""
    _, h_old, w_old, _ = deformation.shape
    _, _, h, w = inp.shape
    if h_old != h or w_old != w:
        deformation = F.interpolate(deformation.permute(0, 3, 1, 2), size=(h, w), mode="bilinear",
                                    align_corners=True).permute(0, 2, 3, 1)
    return grid_sample(inp, deformation)


def coords_grid(batch, ht, wd, device):
    """util.py:53-56 -- (B,2,ht,wd) pixel grid, channel 0 = x (bit-exact)."""
    return ops.coords_grid_cuda(batch, ht, wd, device)


def make_coordinate_grid(spatial_size, type):
    """util.py:90-108 -- (h,w,2) grid in [-1,1], last dim (x,y) (bit-exact).  ``type`` is the
    legacy tensor-type string / a device."""
    h, w = spatial_size
    return ops.make_coordinate_grid_cuda(int(h), int(w), _device_of_type_string(type))


def kp2gaussian(kp, spatial_size, kp_variance):
    """util.py:59-87 -- kp (...,2) -> (...,h,w) Gaussian heat-maps."""
    h, w = spatial_size
    return torch.ops.mrfa.kp2gaussian(kp, None, int(h), int(w), float(kp_variance))


def to_homogeneous(coordinates):
    """util.py:329-334."""
    return torch.cat([coordinates, torch.ones_like(coordinates[..., :1])], dim=-1)


def from_homogeneous(coordinates):
    """util.py:337-338."""
    return coordinates[..., :2] / coordinates[..., 2:3]


class TPS:
    """util.py:341-427 -- thin-plate-spline transformation.

    mode 'kp' (Eq. 2 of the TPSM paper) solves for the spline on the GPU (`mrfa::tps_solve`);
    mode 'random' (equivariance loss, training only) draws its parameters like the reference.
    """

    def __init__(self, mode, bs, **kwargs):
        self.bs = bs
        self.mode = mode
        if mode == "random":
            noise = torch.normal(mean=0, std=kwargs["sigma_affine"] * torch.ones([bs, 2, 3]))
            self.theta = noise + torch.eye(2, 3).view(1, 2, 3)
            n = kwargs["points_tps"]
            ax = 2 * (torch.arange(n, dtype=torch.float32) / (n - 1)) - 1
            self.control_points = torch.stack([ax[None, :].expand(n, n), ax[:, None].expand(n, n)], 2).unsqueeze(0)
            self.control_params = torch.normal(mean=0, std=kwargs["sigma_tps"] * torch.ones([bs, 1, n ** 2]))
        elif mode == "kp":
            kp_1, kp_2 = kwargs["kp_1"], kwargs["kp_2"]
            self.gs = kp_1.shape[1]
            self.theta, self.control_params = torch.ops.mrfa.tps_solve(kp_1, kp_2)
            self.control_points = kp_1
        else:
            raise Exception("Error TPS mode")

    def transform_frame(self, frame):
        h, w = frame.shape[2:]
        if self.mode == "random":            # one kernel instead of the (B, HW, P, 2) temporaries
            dev = frame.device
            return torch.ops.mrfa.random_warp_grid(self.theta.to(dev, torch.float32), self.control_points.to(dev, torch.float32),
                                                   self.control_params.to(dev, torch.float32), h, w, _lib.TPS_L2SQ)
        grid = ops.make_coordinate_grid_cuda(h, w, frame.device).view(1, h * w, 2)
        shape = [self.bs, h, w, 2]
        if self.mode == "kp":
            shape.insert(1, self.gs)
        return self.warp_coordinates(grid).view(*shape)

    def warp_coordinates(self, coordinates):
        theta = self.theta.to(coordinates)
        control_points = self.control_points.to(coordinates)
        control_params = self.control_params.to(coordinates)
        if self.mode == "kp":
            out = torch.matmul(theta[:, :, :, :2], coordinates.permute(0, 2, 1)) + theta[:, :, :, 2:]
            d = coordinates.view(coordinates.shape[0], 1, 1, -1, 2) - control_points.view(self.bs, control_points.shape[1], -1, 1, 2)
            r2 = (d ** 2).sum(-1)
            rbf = torch.matmul((r2 * torch.log(r2 + 1e-9)).permute(0, 1, 3, 2), control_params)
            return out.permute(0, 1, 3, 2) + rbf
        if self.mode == "random":
            theta = theta.unsqueeze(1)
            out = (torch.matmul(theta[:, :, :, :2], coordinates.unsqueeze(-1)) + theta[:, :, :, 2:]).squeeze(-1)
            d = coordinates.view(coordinates.shape[0], -1, 1, 2) - control_points.view(1, 1, -1, 2)
            r2 = (d ** 2).sum(-1)
            rbf = (r2 * torch.log(r2 + 1e-9) * control_params).sum(dim=2).view(self.bs, coordinates.shape[1], 1)
            return out + rbf
        raise Exception("Error TPS mode")
