"""Training-only random warps of the equivariance constraint.

``Transform`` mirrors modules/model.py:26-77: a random affine (+ optional thin-plate) warp of a
frame through a reflection-padded ``grid_sample``, the same warp applied to key-points, and its
Jacobian.  The frame path -- the only part that touches image-sized tensors -- runs on two
kernels: ``mrfa::random_warp_grid`` builds the (B,H,W,2) sampling grid in one pass (the
reference materialises (B, HW, P, 2) temporaries) and ``mrfa::grid_sample`` samples it with
reflection padding (forward and backward).  ``warp_coordinates`` / ``jacobian`` act on (B,K,2)
key-points, need double-backward through autograd, and stay torch expressions.
"""
from __future__ import annotations

import torch
from torch.autograd import grad

from . import _lib, sampling


class Transform:
    """Random tps transformation for equivariance constraints (model.py:26-77); same constructor
    kwargs (`sigma_affine`, optional `sigma_tps` + `points_tps`) and the same RNG draws, so a
    shared seed gives the reference's parameters."""

    def __init__(self, bs, **kwargs):
        noise = torch.normal(mean=0, std=kwargs["sigma_affine"] * torch.ones([bs, 2, 3]))
        self.theta = noise + torch.eye(2, 3).view(1, 2, 3)
        self.bs = bs
        if ("sigma_tps" in kwargs) and ("points_tps" in kwargs):
            self.tps = True
            n = kwargs["points_tps"]
            ax = 2 * (torch.arange(n, dtype=torch.float32) / (n - 1)) - 1        # make_coordinate_grid, host side
            self.control_points = torch.stack([ax[None, :].expand(n, n), ax[:, None].expand(n, n)], 2).unsqueeze(0)
            self.control_params = torch.normal(mean=0, std=kwargs["sigma_tps"] * torch.ones([bs, 1, n ** 2]))
        else:
            self.tps = False

    def transform_frame(self, frame):
        if not frame.is_cuda:
            raise RuntimeError("mrfa_b200: Transform.transform_frame needs a CUDA frame (there is no CPU fallback)")
        dev = frame.device
        theta = self.theta.to(dev, torch.float32)
        cp = self.control_points.to(dev, torch.float32) if self.tps else None
        cw = self.control_params.to(dev, torch.float32) if self.tps else None
        grid = torch.ops.mrfa.random_warp_grid(theta, cp, cw, frame.shape[2], frame.shape[3], _lib.TPS_L1)
        return sampling.grid_sample(frame, grid, padding_mode="reflection")

    def warp_coordinates(self, coordinates):
        theta = self.theta.to(coordinates).unsqueeze(1)
        transformed = (torch.matmul(theta[:, :, :, :2], coordinates.unsqueeze(-1)) + theta[:, :, :, 2:]).squeeze(-1)
        if self.tps:
            control_points = self.control_points.to(coordinates)
            control_params = self.control_params.to(coordinates)
            distances = coordinates.view(coordinates.shape[0], -1, 1, 2) - control_points.view(1, 1, -1, 2)
            distances = torch.abs(distances).sum(-1)
            result = distances ** 2 * torch.log(distances + 1e-6) * control_params
            transformed = transformed + result.sum(dim=2).view(self.bs, coordinates.shape[1], 1)
        return transformed

    def jacobian(self, coordinates):
        new_coordinates = self.warp_coordinates(coordinates)
        grad_x = grad(new_coordinates[..., 0].sum(), coordinates, create_graph=True)
        grad_y = grad(new_coordinates[..., 1].sum(), coordinates, create_graph=True)
        return torch.cat([grad_x[0].unsqueeze(-2), grad_y[0].unsqueeze(-2)], dim=-2)
