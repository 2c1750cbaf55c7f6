"""Deterministic synthetic inputs and weights.

Everything is drawn from numpy's frozen ``RandomState`` streams keyed by a name, so the same
bits are produced in the build container (where the golden vectors are generated from the
reference), on the GPU box, and for any module layout with the same state_dict keys.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch


def _rs(name: str, seed: int = 0) -> np.random.RandomState:
    return np.random.RandomState((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31 - 1))


def tensor(name: str, shape, kind: str = "normal", scale: float = 1.0, shift: float = 0.0, seed: int = 0):
    rs = _rs(name, seed)
    if kind == "normal":
        a = rs.standard_normal(size=shape)
    elif kind == "uniform":
        a = rs.uniform(0.0, 1.0, size=shape)
    else:
        raise ValueError(kind)
    return torch.from_numpy((a * scale + shift).astype(np.float32))


def fill_state_dict_(module: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    """Overwrite every parameter / BN statistic with a name-keyed deterministic value.

    Convolution weights are He-scaled so activations stay O(1) through the stacks; the last
    layers that emit flow / occlusion / mask logits are damped so the iterative refinement
    stays inside the image (a few pixels of motion per level).
    """
    damped = ("refine.conv2.", "refine.convo2.", "mask.", "occlusion.", "maps.")
    sd = module.state_dict()
    new = {}
    for key in sorted(sd):
        ref = sd[key]
        shape = tuple(ref.shape)
        if key.endswith("num_batches_tracked"):
            new[key] = torch.zeros_like(ref)
        elif key.endswith("down.weight") and ref.dim() == 4 and shape[1] == 1:
            new[key] = ref.clone()                         # fixed anti-alias Gaussian buffer
        elif key.endswith("running_var"):
            new[key] = tensor(key, shape, "uniform", 1.0, 0.5, seed)
        elif key.endswith("running_mean"):
            new[key] = tensor(key, shape, "normal", 0.1, 0.0, seed)
        elif key.endswith("pos_embedding"):
            new[key] = tensor(key, shape, "normal", 0.02, 0.0, seed)
        elif ref.dim() == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            s = (2.0 / fan_in) ** 0.5
            if any(d in key for d in damped):
                s *= 0.25
            new[key] = tensor(key, shape, "normal", s, 0.0, seed)
        elif ".norm" in key and key.endswith("weight"):
            new[key] = tensor(key, shape, "uniform", 0.4, 0.8, seed)
        else:                                              # biases
            new[key] = tensor(key, shape, "normal", 0.05, 0.0, seed)
        new[key] = new[key].to(ref.dtype)
    module.load_state_dict(new)
    return module


def frame_pairs(batch: int, size: int, seed: int = 0):
    """U[0,1) source / driving frames (B,3,S,S)."""
    return (tensor("source", (batch, 3, size, size), "uniform", seed=seed),
            tensor("driving", (batch, 3, size, size), "uniform", seed=seed))


def keypoints(batch: int, num_kp: int = 10, seed: int = 0, jacobian: bool = True, spread: float = 0.8,
              motion: float = 0.15):
    """Source key-points ~U(-spread, spread); driving = source + N(0, motion); Jacobians
    I + 0.1 N(0,1) (random-init detectors emit exact identity)."""
    kp_s = tensor("kp_s", (batch, num_kp, 2), "uniform", 2 * spread, -spread, seed)
    kp_d = kp_s + tensor("kp_d", (batch, num_kp, 2), "normal", motion, 0.0, seed)
    src, drv = {"kp": kp_s}, {"kp": kp_d}
    if jacobian:
        eye = torch.eye(2).view(1, 1, 2, 2)
        src["jacobian"] = eye + tensor("jac_s", (batch, num_kp, 2, 2), "normal", 0.1, 0.0, seed)
        drv["jacobian"] = eye + tensor("jac_d", (batch, num_kp, 2, 2), "normal", 0.1, 0.0, seed)
    return src, drv


def bg_affine(batch: int, seed: int = 0):
    """Identity 3x3 with the top two rows perturbed by 0.05 N(0,1) (celebvhq bg path)."""
    p = torch.eye(3).repeat(batch, 1, 1)
    p[:, :2] += tensor("bg_param", (batch, 2, 3), "normal", 0.05, 0.0, seed)
    return p
