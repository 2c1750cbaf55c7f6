"""torch custom ops (`torch.ops.mrfa.*`) over the C ABI of libmrfa_b200.so.

Each op has a CUDA implementation (a ctypes call that launches hand-written sm_100a kernels on
the current stream), a fake implementation (shape inference) and, where the reference
differentiates through it, an autograd formula.  There is no CPU implementation on purpose:
calling an op with CPU tensors raises.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import GridStrides, check, lib

_c = _lib.ctypes


def _p(t: Optional[Tensor]):
    return None if t is None else _c.c_void_p(t.data_ptr())


def _stream():
    return _c.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t: Tensor, name: str, dtype=torch.float32) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"mrfa_b200: `{name}` must be a CUDA tensor (there is no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"mrfa_b200: `{name}` must be {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _is_channels_last(t: Tensor) -> bool:
    """NHWC in memory (and not simultaneously NCHW-contiguous), with a vectorisable channel count."""
    return (t.dim() == 4 and t.shape[1] % 4 == 0 and not t.is_contiguous()
            and t.is_contiguous(memory_format=torch.channels_last))


def _suggest_channels_last(t: Tensor) -> bool:
    """Layout a result should take to stay in the producer's memory format (like ATen's
    suggest_memory_format): NHWC if the tensor is channels_last or a channel slice of one."""
    if t.dim() != 4 or t.is_contiguous():
        return False
    return t.is_contiguous(memory_format=torch.channels_last) or (t.stride(1) == 1 and t.shape[1] > 1)


def _req_image(t: Tensor, name: str):
    """float32 CUDA image in either NCHW or NHWC memory; returns (tensor, channels_last flag)."""
    if not t.is_cuda:
        raise RuntimeError(f"mrfa_b200: `{name}` must be a CUDA tensor (there is no CPU fallback)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"mrfa_b200: `{name}` must be float32, got {t.dtype}")
    if _is_channels_last(t):
        return t, True
    return (t if t.is_contiguous() else t.contiguous()), False


def _like_layout(t: Tensor, channels_last: bool) -> Tensor:
    return t.contiguous(memory_format=torch.channels_last) if channels_last else t.contiguous()


def _empty_image(shape, device, channels_last: bool) -> Tensor:
    return torch.empty(shape, device=device, dtype=torch.float32,
                       memory_format=torch.channels_last if channels_last else torch.contiguous_format)


def _grid_strides(grid: Tensor) -> GridStrides:
    s = grid.stride()
    return GridStrides(s[0], s[1], s[2], s[3])


class KernelTimer:
    """Per-kernel CUDA-event timing on the launching stream (bench.py roofline accounting).

    While installed (``with KernelTimer() as t:``) every C-ABI launch is bracketed by two events
    recorded on the current stream and tagged with its algorithmic bytes / flops."""

    def __init__(self, only=None):
        """`only`: names to bracket with events; every other launch is merely counted (no events, so the
        instrumentation does not perturb a timed region)."""
        self.records = {}
        self.counts = {}
        self.only = None if only is None else frozenset(only)

    def __enter__(self):
        global _TIMER
        self._prev, _TIMER = _TIMER, self
        return self

    def __exit__(self, *exc):
        global _TIMER
        _TIMER = self._prev

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, recs in self.records.items():
            ms = [a.elapsed_time(b) for a, b, _, _, _ in recs]
            out[name] = {"launches": sum(r[4] for r in recs), "calls": len(recs), "total_ms": sum(ms),
                         "bytes": sum(r[2] for r in recs), "flops": sum(r[3] for r in recs)}
        return out


_TIMER = None


class _timed:
    __slots__ = ("name", "nbytes", "flops", "launches", "start")

    def __init__(self, name, nbytes=0, flops=0, launches=1):
        self.name, self.nbytes, self.flops, self.launches = name, nbytes, flops, launches

    def __enter__(self):
        self.start = None
        if _TIMER is not None:
            _TIMER.counts[self.name] = _TIMER.counts.get(self.name, 0) + self.launches
            if _TIMER.only is None or self.name in _TIMER.only:
                self.start = torch.cuda.Event(enable_timing=True)
                self.start.record()

    def __exit__(self, *exc):
        if _TIMER is not None and self.start is not None and exc[0] is None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            _TIMER.records.setdefault(self.name, []).append((self.start, end, self.nbytes, self.flops, self.launches))


_SM_COUNT = {}


def sm_count(device) -> int:
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


# ------------------------------------------------------------------------------------------
# bilinear warps
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("mrfa::grid_sample", mutates_args=(), device_types="cuda")
def grid_sample(inp: Tensor, grid: Tensor, coord_mode: int, padding_mode: int, add_identity: bool,
                in_batch_div: int) -> Tensor:
    inp, cl = _req_image(inp, "input")
    if not grid.is_cuda or grid.dtype != torch.float32 or grid.dim() != 4 or grid.shape[-1] != 2:
        raise RuntimeError("mrfa_b200: grid must be a CUDA float32 tensor of logical shape (N,Ho,Wo,2)")
    N, Ho, Wo, _ = grid.shape
    Nin, C, H, W = inp.shape
    if Nin * in_batch_div != N:
        raise RuntimeError(f"mrfa_b200: grid batch {N} != input batch {Nin} * {in_batch_div}")
    out = _empty_image((N, C, Ho, Wo), inp.device, cl)
    if out.numel() == 0:
        return out
    with torch.cuda.device(inp.device):
        with _timed("grid_sample_fwd", 4 * (N * C * Ho * Wo + Nin * C * H * W + 2 * N * Ho * Wo)):
            check(lib.mrfa_grid_sample_fwd(_p(inp), _p(grid), _grid_strides(grid), _p(out), N, C, H, W, Ho, Wo,
                                           in_batch_div, coord_mode, padding_mode, int(add_identity), int(cl), _stream()),
                  "mrfa_grid_sample_fwd")
    return out


@grid_sample.register_fake
def _(inp, grid, coord_mode, padding_mode, add_identity, in_batch_div):
    out = inp.new_empty((grid.shape[0], inp.shape[1], grid.shape[1], grid.shape[2]))
    return out.contiguous(memory_format=torch.channels_last) if _is_channels_last(inp) else out


@torch.library.custom_op("mrfa::grid_sample_bwd", mutates_args=(), device_types="cuda")
def grid_sample_bwd(grad_out: Tensor, inp: Tensor, grid: Tensor, coord_mode: int, padding_mode: int,
                    add_identity: bool, in_batch_div: int, need_input: bool, need_grid: bool) -> Tuple[Tensor, Tensor]:
    inp, cl = _req_image(inp, "input")
    grad_out = _like_layout(_req_image(grad_out, "grad_out")[0], cl)
    N, Ho, Wo, _ = grid.shape
    _, C, H, W = inp.shape
    g_in = torch.zeros_like(inp) if need_input else inp.new_empty(0)
    g_grid = torch.empty((N, Ho, Wo, 2), device=inp.device, dtype=torch.float32) if need_grid else inp.new_empty(0)
    with torch.cuda.device(inp.device):
        with _timed("grid_sample_bwd", 4 * (N * C * Ho * Wo + 3 * inp.numel() + 4 * N * Ho * Wo)):
            check(lib.mrfa_grid_sample_bwd(_p(grad_out), _p(inp), _p(grid), _grid_strides(grid),
                                           _p(g_in) if need_input else None, _p(g_grid) if need_grid else None,
                                           N, C, H, W, Ho, Wo, in_batch_div, coord_mode, padding_mode, int(add_identity),
                                           int(cl), _stream()), "mrfa_grid_sample_bwd")
    return g_in, g_grid


@grid_sample_bwd.register_fake
def _(grad_out, inp, grid, coord_mode, padding_mode, add_identity, in_batch_div, need_input, need_grid):
    return (torch.empty_like(inp) if need_input else inp.new_empty(0),
            inp.new_empty(tuple(grid.shape)) if need_grid else inp.new_empty(0))


def _gs_setup(ctx, inputs, output):
    inp, grid, coord_mode, padding_mode, add_identity, in_batch_div = inputs
    ctx.save_for_backward(inp, grid)
    ctx.cfg = (coord_mode, padding_mode, add_identity, in_batch_div)


def _gs_backward(ctx, grad_out):
    inp, grid = ctx.saved_tensors
    need_input, need_grid = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    g_in, g_grid = torch.ops.mrfa.grid_sample_bwd(grad_out, inp, grid, *ctx.cfg, need_input, need_grid)
    return (g_in if need_input else None, g_grid if need_grid else None, None, None, None, None)


grid_sample.register_autograd(_gs_backward, setup_context=_gs_setup)


@torch.library.custom_op("mrfa::dual_warp", mutates_args=(), device_types="cuda")
def dual_warp(inp: Tensor, flow: Tensor, prior_grid: Tensor) -> Tuple[Tensor, Tensor]:
    (inp, cl), flow, prior_grid = _req_image(inp, "input"), _req(flow, "flow"), _req(prior_grid, "prior_grid")
    N, C, H, W = inp.shape
    if tuple(flow.shape) != (N, 2, H, W) or tuple(prior_grid.shape) != (N, H, W, 2):
        raise RuntimeError("mrfa_b200: dual_warp expects flow (N,2,H,W) and prior_grid (N,H,W,2) at the feature size")
    out_r, out_c = torch.empty_like(inp), torch.empty_like(inp)
    with torch.cuda.device(inp.device):
        with _timed("dual_warp_fwd", 4 * (3 * inp.numel() + 4 * N * H * W)):
            check(lib.mrfa_dual_warp_fwd(_p(inp), _p(flow), _p(prior_grid), _p(out_r), _p(out_c), N, C, H, W, int(cl), 0, _stream()),
                  "mrfa_dual_warp_fwd")
    return out_r, out_c


@torch.library.custom_op("mrfa::dual_warp_cat", mutates_args=(), device_types="cuda")
def dual_warp_cat(inp: Tensor, flow: Tensor, prior_grid: Tensor) -> Tuple[Tensor, Tensor]:
    """dual_warp whose coarse result is written into channels [C, 2C) of a fresh (N,2C,H,W) channels_last buffer --
    the tensor the decoder would build with cat([y, warp_c], 1) (generator.py:51,60); channels [0, C) are left for
    mrfa::occlusion_blend_subpixel_into.  channels_last input only (inference path)."""
    (inp, cl), flow, prior_grid = _req_image(inp, "input"), _req(flow, "flow"), _req(prior_grid, "prior_grid")
    N, C, H, W = inp.shape
    if not cl or tuple(flow.shape) != (N, 2, H, W) or tuple(prior_grid.shape) != (N, H, W, 2):
        raise RuntimeError("mrfa_b200: dual_warp_cat expects a channels_last input, flow (N,2,H,W) and prior_grid (N,H,W,2)")
    out_r = torch.empty_like(inp)
    buf = _empty_image((N, 2 * C, H, W), inp.device, True)
    if inp.numel() == 0:
        return out_r, buf
    with torch.cuda.device(inp.device):
        with _timed("dual_warp_fwd", 4 * (3 * inp.numel() + 4 * N * H * W)):
            check(lib.mrfa_dual_warp_fwd(_p(inp), _p(flow), _p(prior_grid), _p(out_r), buf.data_ptr() + 4 * C, N, C, H, W, 1,
                                         2 * C, _stream()), "mrfa_dual_warp_fwd")
    return out_r, buf


@dual_warp_cat.register_fake
def _(inp, flow, prior_grid):
    N, C, H, W = inp.shape
    return torch.empty_like(inp), inp.new_empty((N, 2 * C, H, W)).contiguous(memory_format=torch.channels_last)


@dual_warp.register_fake
def _(inp, flow, prior_grid):
    return torch.empty_like(inp), torch.empty_like(inp)


def _dw_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _dw_backward(ctx, g_r, g_c):
    inp, flow, prior = ctx.saved_tensors
    ni, nf, npg = ctx.needs_input_grad
    gi = gf = gp = None
    a, b = torch.ops.mrfa.grid_sample_bwd(g_r, inp, flow.permute(0, 2, 3, 1), _lib.COORD_PIXEL, _lib.PAD_ZEROS, True, 1,
                                          ni, nf)
    c, d = torch.ops.mrfa.grid_sample_bwd(g_c, inp, prior, _lib.COORD_NORM_ACF, _lib.PAD_ZEROS, False, 1, ni, npg)
    if ni:
        gi = a + c
    if nf:
        gf = b.permute(0, 3, 1, 2)
    if npg:
        gp = d
    return gi, gf, gp


dual_warp.register_autograd(_dw_backward, setup_context=_dw_setup)


# ------------------------------------------------------------------------------------------
# grids / heat-maps (no autograd needed for the grids; kp2gaussian differentiates w.r.t. kp)
# ------------------------------------------------------------------------------------------
def coords_grid_cuda(batch: int, ht: int, wd: int, device) -> Tensor:
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("mrfa_b200: coords_grid needs a CUDA device (there is no CPU fallback)")
    out = torch.empty((batch, 2, ht, wd), device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        with _timed("coords_grid", 4 * out.numel()):
            check(lib.mrfa_coords_grid(_p(out), batch, ht, wd, _stream()), "mrfa_coords_grid")
    return out


def make_coordinate_grid_cuda(h: int, w: int, device) -> Tensor:
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("mrfa_b200: make_coordinate_grid needs a CUDA device (there is no CPU fallback)")
    out = torch.empty((h, w, 2), device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        with _timed("make_coordinate_grid", 4 * out.numel()):
            check(lib.mrfa_make_coordinate_grid(_p(out), h, w, _stream()), "mrfa_make_coordinate_grid")
    return out


@torch.library.custom_op("mrfa::kp2gaussian", mutates_args=(), device_types="cuda")
def kp2gaussian(kp: Tensor, add: Optional[Tensor], h: int, w: int, variance: float) -> Tensor:
    kp = _req(kp, "kp")
    P = kp.numel() // 2
    period = 0
    if add is not None:
        add = _req(add, "add")
        period = add.numel() // (h * w)
    out = torch.empty(tuple(kp.shape[:-1]) + (h, w), device=kp.device, dtype=torch.float32)
    with torch.cuda.device(kp.device):
        with _timed("kp2gaussian", 4 * (out.numel() + kp.numel())):
            check(lib.mrfa_kp2gaussian(_p(kp), _p(add), period, _p(out), P, h, w, variance, _stream()), "mrfa_kp2gaussian")
    return out


@kp2gaussian.register_fake
def _(kp, add, h, w, variance):
    return kp.new_empty(tuple(kp.shape[:-1]) + (h, w))


def _kp2g_setup(ctx, inputs, output):
    kp, add, h, w, variance = inputs
    ctx.save_for_backward(kp)
    ctx.cfg = (add is not None and tuple(add.shape), h, w, variance)


@torch.library.custom_op("mrfa::kp2gaussian_bwd", mutates_args=(), device_types="cuda")
def kp2gaussian_bwd(grad: Tensor, kp: Tensor, h: int, w: int, variance: float) -> Tensor:
    """d kp of util.kp2gaussian (util.py:59-87): per heat-map block reduction, one atomic per block and component."""
    grad, kp = _req(grad, "grad"), _req(kp, "kp")
    P = kp.numel() // 2
    gk = torch.zeros_like(kp)
    with torch.cuda.device(kp.device):
        with _timed("kp2gaussian_bwd", 4 * (grad.numel() + 2 * kp.numel())):
            check(lib.mrfa_kp2gaussian_bwd(_p(grad), _p(kp), _p(gk), P, h, w, variance, _stream()), "mrfa_kp2gaussian_bwd")
    return gk


@kp2gaussian_bwd.register_fake
def _(grad, kp, h, w, variance):
    return torch.empty_like(kp)


def _kp2g_backward(ctx, g):
    (kp,) = ctx.saved_tensors
    add_shape, h, w, variance = ctx.cfg
    gk = ga = None
    if ctx.needs_input_grad[0]:
        gk = torch.ops.mrfa.kp2gaussian_bwd(g, kp, h, w, variance)
    if ctx.needs_input_grad[1] and add_shape:
        ga = g.reshape((-1,) + add_shape).sum(0)
    return gk, ga, None, None, None


kp2gaussian.register_autograd(_kp2g_backward, setup_context=_kp2g_setup)


@torch.library.custom_op("mrfa::prior_to_flow", mutates_args=(), device_types="cuda")
def prior_to_flow(deformation: Tensor, hm1: float) -> Tensor:
    deformation = _req(deformation, "deformation")
    B, h, w, _ = deformation.shape
    out = torch.empty((B, 2, h, w), device=deformation.device, dtype=torch.float32)
    with torch.cuda.device(deformation.device):
        with _timed("prior_to_flow", 8 * out.numel()):
            check(lib.mrfa_prior_to_flow(_p(deformation), _p(out), B, h, w, hm1, _stream()), "mrfa_prior_to_flow")
    return out


@prior_to_flow.register_fake
def _(deformation, hm1):
    B, h, w, _ = deformation.shape
    return deformation.new_empty((B, 2, h, w))


def _p2f_setup(ctx, inputs, output):
    ctx.hm1 = inputs[1]


def _p2f_backward(ctx, g):
    return g.permute(0, 2, 3, 1) * (ctx.hm1 / 2.0), None


prior_to_flow.register_autograd(_p2f_backward, setup_context=_p2f_setup)


# ------------------------------------------------------------------------------------------
# prior dense motion
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("mrfa::dense_motion_prior", mutates_args=(), device_types="cuda")
def dense_motion_prior(kp_d: Tensor, kp_s: Tensor, jac_d: Optional[Tensor], jac_s: Optional[Tensor],
                       bg_param: Optional[Tensor], source: Tensor, variance: float) -> Tuple[Tensor, Tensor]:
    kp_d, kp_s, source = _req(kp_d, "kp_driving"), _req(kp_s, "kp_source"), _req(source, "source")
    jac_d = None if jac_d is None else _req(jac_d, "jac_driving")
    jac_s = None if jac_s is None else _req(jac_s, "jac_source")
    bg_param = None if bg_param is None else _req(bg_param, "bg_param")
    B, K, _ = kp_d.shape
    _, C, h, w = source.shape
    motions = torch.empty((B, K + 1, h, w, 2), device=source.device, dtype=torch.float32)
    hg_input = torch.empty((B, (K + 1) * (C + 1), h, w), device=source.device, dtype=torch.float32)
    with torch.cuda.device(source.device):
        with _timed("dense_motion_prior", 4 * (motions.numel() + hg_input.numel() + source.numel())):
            check(lib.mrfa_dense_motion_prior(_p(kp_d), _p(kp_s), _p(jac_d), _p(jac_s), _p(bg_param), _p(source),
                                              _p(motions), _p(hg_input), B, K, C, h, w, variance, _stream()),
                  "mrfa_dense_motion_prior")
    return motions, hg_input


@dense_motion_prior.register_fake
def _(kp_d, kp_s, jac_d, jac_s, bg_param, source, variance):
    B, K, _ = kp_d.shape
    _, C, h, w = source.shape
    return source.new_empty((B, K + 1, h, w, 2)), source.new_empty((B, (K + 1) * (C + 1), h, w))


@torch.library.custom_op("mrfa::dense_motion_prior_bwd", mutates_args=(), device_types="cuda")
def dense_motion_prior_bwd(grad_motions: Optional[Tensor], grad_hg: Tensor, kp_d: Tensor, kp_s: Tensor,
                           jac_d: Optional[Tensor], jac_s: Optional[Tensor], bg_param: Optional[Tensor], source: Tensor,
                           variance: float, need_source: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Gradients of mrfa::dense_motion_prior w.r.t. (kp_d, kp_s, jac_d, jac_s, bg_param, source); absent ones are
    empty tensors.  One pass over the (K+1) x h x w plane + a finishing launch (csrc/motion_bwd.cu)."""
    kp_d, kp_s, source, grad_hg = _req(kp_d, "kp_driving"), _req(kp_s, "kp_source"), _req(source, "source"), _req(grad_hg, "grad_hg")
    jac_d = None if jac_d is None else _req(jac_d, "jac_driving")
    jac_s = None if jac_s is None else _req(jac_s, "jac_source")
    bg_param = None if bg_param is None else _req(bg_param, "bg_param")
    grad_motions = None if grad_motions is None else _req(grad_motions, "grad_motions")
    B, K, _ = kp_d.shape
    _, C, h, w = source.shape
    dev = source.device
    ws = torch.zeros(int(lib.mrfa_dense_motion_prior_bwd_workspace(B, K)), device=dev, dtype=torch.float32)
    g_kd, g_ks = torch.zeros_like(kp_d), torch.zeros_like(kp_s)
    g_jd = torch.empty_like(jac_d) if jac_d is not None else source.new_empty(0)
    g_js = torch.empty_like(jac_s) if jac_s is not None else source.new_empty(0)
    g_bg = torch.empty_like(bg_param) if bg_param is not None else source.new_empty(0)
    g_src = torch.zeros_like(source) if need_source else source.new_empty(0)
    with torch.cuda.device(dev):
        with _timed("dense_motion_prior_bwd", 4 * (grad_hg.numel() + source.numel() * (2 if need_source else 1)
                                                   + (0 if grad_motions is None else grad_motions.numel())), launches=2):
            check(lib.mrfa_dense_motion_prior_bwd(_p(grad_motions), _p(grad_hg), _p(kp_d), _p(kp_s), _p(jac_d), _p(jac_s),
                                                  _p(bg_param), _p(source), _p(ws), _p(g_kd), _p(g_ks),
                                                  _p(g_jd) if jac_d is not None else None,
                                                  _p(g_js) if jac_s is not None else None,
                                                  _p(g_bg) if bg_param is not None else None,
                                                  _p(g_src) if need_source else None,
                                                  B, K, C, h, w, variance, _stream()), "mrfa_dense_motion_prior_bwd")
    return g_kd, g_ks, g_jd, g_js, g_bg, g_src


@dense_motion_prior_bwd.register_fake
def _(grad_motions, grad_hg, kp_d, kp_s, jac_d, jac_s, bg_param, source, variance, need_source):
    e = source.new_empty(0)
    return (torch.empty_like(kp_d), torch.empty_like(kp_s), torch.empty_like(jac_d) if jac_d is not None else e,
            torch.empty_like(jac_s) if jac_s is not None else e, torch.empty_like(bg_param) if bg_param is not None else e,
            torch.empty_like(source) if need_source else e)


def _dmp_setup(ctx, inputs, output):
    kp_d, kp_s, jac_d, jac_s, bg_param, source, variance = inputs
    ctx.save_for_backward(kp_d, kp_s, jac_d, jac_s, bg_param, source)
    ctx.variance = variance


def _dmp_backward(ctx, g_motions, g_hg):
    kp_d, kp_s, jac_d, jac_s, bg_param, source = ctx.saved_tensors
    need = ctx.needs_input_grad
    if g_hg is None:
        g_hg = torch.zeros((source.shape[0], (kp_d.shape[1] + 1) * (source.shape[1] + 1)) + tuple(source.shape[2:]),
                           device=source.device, dtype=source.dtype)
    g_kd, g_ks, g_jd, g_js, g_bg, g_src = torch.ops.mrfa.dense_motion_prior_bwd(
        g_motions, g_hg, kp_d, kp_s, jac_d, jac_s, bg_param, source, ctx.variance, bool(need[5]))
    return (g_kd if need[0] else None, g_ks if need[1] else None,
            g_jd if (need[2] and jac_d is not None) else None, g_js if (need[3] and jac_s is not None) else None,
            g_bg if (need[4] and bg_param is not None) else None, g_src if need[5] else None, None)


dense_motion_prior.register_autograd(_dmp_backward, setup_context=_dmp_setup)


@torch.library.custom_op("mrfa::tps_solve", mutates_args=(), device_types="cuda")
def tps_solve(kp_1: Tensor, kp_2: Tensor) -> Tuple[Tensor, Tensor]:
    kp_1, kp_2 = _req(kp_1, "kp_1"), _req(kp_2, "kp_2")
    B, G, n, _ = kp_1.shape
    if n != 5:
        raise RuntimeError("mrfa_b200: tps_solve is specialised for 5 control points per transformation")
    theta = torch.empty((B, G, 2, 3), device=kp_1.device, dtype=torch.float32)
    params = torch.empty((B, G, n, 2), device=kp_1.device, dtype=torch.float32)
    with torch.cuda.device(kp_1.device):
        with _timed("tps_solve", 4 * (2 * kp_1.numel() + theta.numel() + params.numel())):
            check(lib.mrfa_tps_solve(_p(kp_1), _p(kp_2), _p(theta), _p(params), B * G, _stream()), "mrfa_tps_solve")
    return theta, params


@tps_solve.register_fake
def _(kp_1, kp_2):
    B, G, n, _ = kp_1.shape
    return kp_1.new_empty((B, G, 2, 3)), kp_1.new_empty((B, G, n, 2))


@torch.library.custom_op("mrfa::tps_motion_prior", mutates_args=(), device_types="cuda")
def tps_motion_prior(kp_d: Tensor, kp_s: Tensor, theta: Tensor, control_params: Tensor, bg_param: Optional[Tensor],
                     source: Tensor, variance: float) -> Tuple[Tensor, Tensor]:
    kp_d, kp_s, source = _req(kp_d, "kp_driving"), _req(kp_s, "kp_source"), _req(source, "source")
    theta, control_params = _req(theta, "theta"), _req(control_params, "control_params")
    bg_param = None if bg_param is None else _req(bg_param, "bg_param")
    B, G = theta.shape[:2]
    _, C, h, w = source.shape
    motions = torch.empty((B, G + 1, h, w, 2), device=source.device, dtype=torch.float32)
    hg_input = torch.empty((B, G * 5 + 1 + (G + 1) * C, h, w), device=source.device, dtype=torch.float32)
    with torch.cuda.device(source.device):
        with _timed("tps_motion_prior", 4 * (motions.numel() + hg_input.numel() + source.numel()), launches=2):
            check(lib.mrfa_tps_motion_prior(_p(kp_d), _p(kp_s), _p(theta), _p(control_params), _p(bg_param), _p(source),
                                            _p(motions), _p(hg_input), B, G, C, h, w, variance, _stream()),
                  "mrfa_tps_motion_prior")
    return motions, hg_input


@tps_motion_prior.register_fake
def _(kp_d, kp_s, theta, control_params, bg_param, source, variance):
    B, G = theta.shape[:2]
    _, C, h, w = source.shape
    return source.new_empty((B, G + 1, h, w, 2)), source.new_empty((B, G * 5 + 1 + (G + 1) * C, h, w))


@torch.library.custom_op("mrfa::tps_prior", mutates_args=(), device_types="cuda")
def tps_prior(kp_d: Tensor, kp_s: Tensor, bg_param: Optional[Tensor], source: Tensor,
              variance: float) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """mrfa::tps_solve + mrfa::tps_motion_prior as one differentiable op (dense_motion.py:200-243, util.py:355-410).
    Returns (motions, hg_input, theta, control_params); the last two are saved for the backward."""
    B = source.shape[0]
    theta, params = tps_solve(kp_d.reshape(B, -1, 5, 2), kp_s.reshape(B, -1, 5, 2))
    motions, hg = tps_motion_prior(kp_d, kp_s, theta, params, bg_param, source, variance)
    return motions, hg, theta, params


@tps_prior.register_fake
def _(kp_d, kp_s, bg_param, source, variance):
    B, KP, _ = kp_d.shape
    G = KP // 5
    _, C, h, w = source.shape
    return (source.new_empty((B, G + 1, h, w, 2)), source.new_empty((B, KP + 1 + (G + 1) * C, h, w)),
            source.new_empty((B, G, 2, 3)), source.new_empty((B, G, 5, 2)))


@torch.library.custom_op("mrfa::tps_prior_bwd", mutates_args=(), device_types="cuda")
def tps_prior_bwd(grad_motions: Optional[Tensor], grad_hg: Tensor, kp_d: Tensor, kp_s: Tensor, theta: Tensor,
                  control_params: Tensor, bg_param: Optional[Tensor], source: Tensor, variance: float,
                  need_source: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Gradients of mrfa::tps_prior w.r.t. (kp_d, kp_s, bg_param, source): heat-map, motion and warp gradients
    reduced per transformation, then the adjoint 8x8 solve in fp64 (csrc/motion_bwd.cu)."""
    kp_d, kp_s, source, grad_hg = _req(kp_d, "kp_driving"), _req(kp_s, "kp_source"), _req(source, "source"), _req(grad_hg, "grad_hg")
    theta, control_params = _req(theta, "theta"), _req(control_params, "control_params")
    bg_param = None if bg_param is None else _req(bg_param, "bg_param")
    grad_motions = None if grad_motions is None else _req(grad_motions, "grad_motions")
    B, G = theta.shape[:2]
    _, C, h, w = source.shape
    dev = source.device
    ws = torch.zeros(int(lib.mrfa_tps_motion_prior_bwd_workspace(B, G)), device=dev, dtype=torch.float32)
    g_kd, g_ks = torch.zeros_like(kp_d), torch.zeros_like(kp_s)
    g_bg = torch.empty_like(bg_param) if bg_param is not None else source.new_empty(0)
    g_src = torch.zeros_like(source) if need_source else source.new_empty(0)
    with torch.cuda.device(dev):
        with _timed("tps_prior_bwd", 4 * (grad_hg.numel() + source.numel()), launches=3):
            check(lib.mrfa_tps_motion_prior_bwd(_p(grad_motions), _p(grad_hg), _p(kp_d), _p(kp_s), _p(theta),
                                                _p(control_params), _p(bg_param), _p(source), _p(ws), _p(g_kd), _p(g_ks),
                                                _p(g_bg) if bg_param is not None else None,
                                                _p(g_src) if need_source else None,
                                                B, G, C, h, w, variance, _stream()), "mrfa_tps_motion_prior_bwd")
    return g_kd, g_ks, g_bg, g_src


@tps_prior_bwd.register_fake
def _(grad_motions, grad_hg, kp_d, kp_s, theta, control_params, bg_param, source, variance, need_source):
    e = source.new_empty(0)
    return (torch.empty_like(kp_d), torch.empty_like(kp_s), torch.empty_like(bg_param) if bg_param is not None else e,
            torch.empty_like(source) if need_source else e)


def _tpsp_setup(ctx, inputs, output):
    kp_d, kp_s, bg_param, source, variance = inputs
    ctx.save_for_backward(kp_d, kp_s, bg_param, source, output[2], output[3])
    ctx.variance = variance
    ctx.hg_shape = tuple(output[1].shape)
    ctx.mark_non_differentiable(output[2], output[3])


def _tpsp_backward(ctx, g_motions, g_hg, _g_theta, _g_params):
    kp_d, kp_s, bg_param, source, theta, params = ctx.saved_tensors
    need = ctx.needs_input_grad
    if g_hg is None:
        g_hg = torch.zeros(ctx.hg_shape, device=source.device, dtype=source.dtype)
    g_kd, g_ks, g_bg, g_src = torch.ops.mrfa.tps_prior_bwd(g_motions, g_hg, kp_d, kp_s, theta, params, bg_param, source,
                                                           ctx.variance, bool(need[3]))
    return (g_kd if need[0] else None, g_ks if need[1] else None, g_bg if (need[2] and bg_param is not None) else None,
            g_src if need[3] else None, None)


tps_prior.register_autograd(_tpsp_backward, setup_context=_tpsp_setup)


# ------------------------------------------------------------------------------------------
# correlation volume + pyramid, lookup
# ------------------------------------------------------------------------------------------
def corr_rows_total(h: int, w: int) -> int:
    return int(lib.mrfa_corr_rows_total(h, w))


def corr_row_offset(h: int, w: int, pool_log2: int) -> int:
    return int(lib.mrfa_corr_row_offset(h, w, pool_log2))


def corr_map_layout(h: int, w: int) -> int:
    """Layout of the maps mrfa_corr_volume writes for an h x w plane (_lib.MAP_ROWMAJOR / _lib.MAP_TILED)."""
    return int(lib.mrfa_corr_map_layout(h, w))


_MAP_PERM = {}


def corr_map_permutation(layout: int, level: int, H: int, W: int, device) -> Tensor:
    """index[y * W + x] = position of element (y, x) inside a map stored in `layout` (int64, cached per device);
    ``map.index_select(-1, index)`` turns a stored map back into row-major order."""
    key = (layout, level, H, W, str(device))
    if key not in _MAP_PERM:
        idx = [int(lib.mrfa_corr_map_offset(layout, level, y, x, W)) for y in range(H) for x in range(W)]
        _MAP_PERM[key] = torch.tensor(idx, dtype=torch.int64, device=device)
    return _MAP_PERM[key]


@torch.library.custom_op("mrfa::corr_pyramid", mutates_args=(), device_types="cuda")
def corr_pyramid(q_d: Tensor, k_s: Tensor, scale: float, q_bias: Optional[Tensor] = None,
                 k_bias: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """(B,C,h,w) x2 -> volume0 (B, rows_total, h*w) bf16, volume1 (B, rows_total, h*w/4) bf16.
    q_bias / k_bias (C): per-channel biases added to q_d / k_s while they are packed (channels_last inputs)."""
    (q_d, cl), (k_s, cl2) = _req_image(q_d, "q_d"), _req_image(k_s, "k_s")
    if cl != cl2:
        k_s = _like_layout(k_s, cl)
    if q_bias is not None or k_bias is not None:
        if not cl:
            raise RuntimeError("mrfa_b200: corr_pyramid biases need channels_last q_d / k_s")
        q_bias = None if q_bias is None else _req(q_bias, "q_bias")
        k_bias = None if k_bias is None else _req(k_bias, "k_bias")
    B, C, h, w = q_d.shape
    rows = corr_rows_total(h, w)
    dev = q_d.device
    a_op = torch.empty((B, rows, C), device=dev, dtype=torch.bfloat16)
    b_op = torch.empty((B, h * w, C), device=dev, dtype=torch.bfloat16)
    vol0 = torch.empty((B, rows, h * w), device=dev, dtype=torch.bfloat16)
    vol1 = torch.empty((B, rows, (h * w) // 4), device=dev, dtype=torch.bfloat16)
    with torch.cuda.device(dev):
        st = _stream()
        with _timed("corr_pack", 8 * q_d.numel() + 2 * (a_op.numel() + b_op.numel())):
            if q_bias is not None or k_bias is not None:
                check(lib.mrfa_corr_pack_bias(_p(q_d), _p(q_bias), _p(k_s), _p(k_bias), _p(a_op), _p(b_op), B, C, h, w, st),
                      "mrfa_corr_pack_bias")
            else:
                check(lib.mrfa_corr_pack(_p(q_d), _p(k_s), _p(a_op), _p(b_op), B, C, h, w, int(cl), st), "mrfa_corr_pack")
        # algorithmic FLOPs: the basic-resolution contraction only; bytes: operands + pyramid
        with _timed("corr_volume", 2 * (a_op.numel() + b_op.numel() + vol0.numel() + vol1.numel()),
                    2 * B * (h * w) ** 2 * C):
            check(lib.mrfa_corr_volume(_p(a_op), _p(b_op), _p(vol0), _p(vol1), B, C, h, w, scale, sm_count(dev), st),
                  "mrfa_corr_volume")
    return vol0, vol1


@corr_pyramid.register_fake
def _(q_d, k_s, scale, q_bias=None, k_bias=None):
    B, C, h, w = q_d.shape
    rows = h * w + (h * w) // 4 + (h * w) // 16 + (h * w) // 64
    return (q_d.new_empty((B, rows, h * w), dtype=torch.bfloat16),
            q_d.new_empty((B, rows, (h * w) // 4), dtype=torch.bfloat16))


def corr_pack_debug(q_d: Tensor, k_s: Tensor) -> Tuple[Tensor, Tensor]:
    """Packed bf16 operands only (test / profiling helper)."""
    (q_d, cl), (k_s, cl2) = _req_image(q_d, "q_d"), _req_image(k_s, "k_s")
    if cl != cl2:
        k_s = _like_layout(k_s, cl)
    B, C, h, w = q_d.shape
    a_op = torch.empty((B, corr_rows_total(h, w), C), device=q_d.device, dtype=torch.bfloat16)
    b_op = torch.empty((B, h * w, C), device=q_d.device, dtype=torch.bfloat16)
    with torch.cuda.device(q_d.device):
        check(lib.mrfa_corr_pack(_p(q_d), _p(k_s), _p(a_op), _p(b_op), B, C, h, w, int(cl), _stream()), "mrfa_corr_pack")
    return a_op, b_op


@torch.library.custom_op("mrfa::corr_pyramid_bwd", mutates_args=(), device_types="cuda")
def corr_pyramid_bwd(g0: Tensor, g1: Tensor, q_d: Tensor, k_s: Tensor, scale: float) -> Tuple[Tensor, Tensor]:
    """Gradients of mrfa::corr_pyramid w.r.t. (q_d, k_s) from the fp32 volume gradients g0 (B, rows_total, hw) and
    g1 (B, rows_total, hw/4): two tcgen05 GEMMs (dA = G Bm, dB = G^T A) plus pack / transpose / un-pool passes.
    Returns channels_last (B,C,h,w) tensors."""
    (q_d, cl), (k_s, cl2) = _req_image(q_d, "q_d"), _req_image(k_s, "k_s")
    if cl != cl2:
        k_s = _like_layout(k_s, cl)
    g0, g1 = _req(g0, "g0"), _req(g1, "g1")
    B, C, h, w = q_d.shape
    N, rows = h * w, corr_rows_total(h, w)
    rows_pad = int(lib.mrfa_corr_bwd_rows_pad(h, w))
    if tuple(g0.shape) != (B, rows, N) or tuple(g1.shape) != (B, rows, N // 4):
        raise RuntimeError("mrfa_b200: corr_pyramid_bwd gradient shapes do not match the pyramid of (q_d, k_s)")
    dev = q_d.device
    bf = dict(device=dev, dtype=torch.bfloat16)
    a_op, b_op = torch.empty((B, rows, C), **bf), torch.empty((B, N, C), **bf)
    G, GT = torch.empty((B, rows, N), **bf), torch.empty((B, N, rows_pad), **bf)
    aT, bT = torch.empty((B, C, rows_pad), **bf), torch.empty((B, C, N), **bf)
    dA = torch.empty((B, rows, C), device=dev, dtype=torch.float32)
    dB = torch.empty((B, N, C), device=dev, dtype=torch.float32)
    d_q, d_k = _empty_image((B, C, h, w), dev, True), _empty_image((B, C, h, w), dev, True)
    with torch.cuda.device(dev):
        st = _stream()
        nsm = sm_count(dev)
        with _timed("corr_pack", 8 * q_d.numel() + 2 * (a_op.numel() + b_op.numel())):
            check(lib.mrfa_corr_pack(_p(q_d), _p(k_s), _p(a_op), _p(b_op), B, C, h, w, int(cl), st), "mrfa_corr_pack")
        with _timed("corr_bwd_pack", 4 * (g0.numel() + g1.numel()) + 2 * (G.numel() + B * N * rows)):
            check(lib.mrfa_corr_bwd_pack(_p(g0), _p(g1), _p(G), _p(GT), B, h, w, scale, st), "mrfa_corr_bwd_pack")
        with _timed("corr_bwd_transpose", 4 * (a_op.numel() + b_op.numel()), launches=2):
            check(lib.mrfa_transpose_bf16(_p(a_op), _p(aT), B, rows, C, rows_pad, st), "mrfa_transpose_bf16")
            check(lib.mrfa_transpose_bf16(_p(b_op), _p(bT), B, N, C, N, st), "mrfa_transpose_bf16")
        with _timed("corr_bwd_gemm", 2 * (2 * G.numel() + a_op.numel() + b_op.numel()) + 4 * (dA.numel() + dB.numel()),
                    2 * 2 * B * rows * N * C, launches=2):
            check(lib.mrfa_corr_bwd_gemm(_p(G), _p(bT), _p(dA), B, rows, C, N, N, N, nsm, st), "mrfa_corr_bwd_gemm")
            check(lib.mrfa_corr_bwd_gemm(_p(GT), _p(aT), _p(dB), B, N, C, rows, rows_pad, rows_pad, nsm, st), "mrfa_corr_bwd_gemm")
        with _timed("corr_bwd_unpack", 4 * (dA.numel() + dB.numel() + 2 * d_q.numel())):
            check(lib.mrfa_corr_bwd_unpack(_p(dA), _p(dB), _p(d_q), _p(d_k), B, C, h, w, st), "mrfa_corr_bwd_unpack")
    return d_q, d_k


@corr_pyramid_bwd.register_fake
def _(g0, g1, q_d, k_s, scale):
    f = lambda t: torch.empty_like(t).contiguous(memory_format=torch.channels_last)
    return f(q_d), f(k_s)


@torch.library.custom_op("mrfa::avg_pool2x2", mutates_args=(), device_types="cuda")
def avg_pool2x2(x: Tensor) -> Tensor:
    x = _req(x, "x")
    H, W = x.shape[-2:]
    P = x.numel() // (H * W)
    out = torch.empty(tuple(x.shape[:-2]) + (H // 2, W // 2), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        with _timed("avg_pool2x2", 4 * (x.numel() + out.numel())):
            check(lib.mrfa_avg_pool2x2(_p(x), _p(out), P, H, W, _stream()), "mrfa_avg_pool2x2")
    return out


@avg_pool2x2.register_fake
def _(x):
    return x.new_empty(tuple(x.shape[:-2]) + (x.shape[-2] // 2, x.shape[-1] // 2))


def _ap_setup(ctx, inputs, output):
    ctx.shape = tuple(inputs[0].shape)


def _ap_backward(ctx, g):
    H, W = ctx.shape[-2:]
    up = (g * 0.25).repeat_interleave(2, dim=-2).repeat_interleave(2, dim=-1)
    if up.shape[-2] != H or up.shape[-1] != W:
        up = torch.nn.functional.pad(up, (0, W - up.shape[-1], 0, H - up.shape[-2]))
    return up


avg_pool2x2.register_autograd(_ap_backward, setup_context=_ap_setup)


@torch.library.custom_op("mrfa::corr_lookup", mutates_args=(), device_types="cuda")
def corr_lookup(level0: Tensor, level1: Tensor, coords: Tensor, H: int, W: int, map_batch_stride: int,
                row_offset: int, radius: int, map_layout: int, channels_last: bool) -> Tensor:
    """coords (B,2,h1,w1) -> (B, 2*(2r+1)^2, h1, w1).  level maps fp32 or bf16, row-major or tiled (see the header)."""
    coords = _req(coords, "coords")
    if level0.dtype != level1.dtype or level0.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("mrfa_b200: correlation levels must both be float32 or both bfloat16")
    if not (level0.is_cuda and level0.is_contiguous() and level1.is_contiguous()):
        raise RuntimeError("mrfa_b200: correlation levels must be contiguous CUDA tensors")
    B, _, h1, w1 = coords.shape
    n = 2 * radius + 1
    out = _empty_image((B, 2 * n * n, h1, w1), coords.device, channels_last)
    with torch.cuda.device(coords.device):
        with _timed("corr_lookup_fwd", B * h1 * w1 * (2 * (2 * radius + 2) ** 2 * level0.element_size() + 8 + 4 * 2 * n * n)):
            check(lib.mrfa_corr_lookup_fwd(_p(level0), _p(level1), int(level0.dtype == torch.bfloat16), _p(coords), _p(out),
                                           B, h1 * w1, H, W, map_batch_stride, row_offset, radius, map_layout,
                                           int(channels_last), _stream()), "mrfa_corr_lookup_fwd")
    return out


@corr_lookup.register_fake
def _(level0, level1, coords, H, W, map_batch_stride, row_offset, radius, map_layout, channels_last):
    B, _, h1, w1 = coords.shape
    out = coords.new_empty((B, 2 * (2 * radius + 1) ** 2, h1, w1))
    return out.contiguous(memory_format=torch.channels_last) if channels_last else out


@torch.library.custom_op("mrfa::corr_lookup_bwd", mutates_args=(), device_types="cuda")
def corr_lookup_bwd(grad_out: Tensor, level0: Tensor, level1: Tensor, coords: Tensor, H: int, W: int,
                    map_batch_stride: int, row_offset: int, radius: int, map_layout: int, need_levels: bool,
                    need_coords: bool) -> Tuple[Tensor, Tensor, Tensor]:
    grad_out, coords = _req(grad_out, "grad_out"), _req(coords, "coords")
    B, _, h1, w1 = coords.shape
    dev = coords.device
    g0 = torch.zeros(level0.shape, device=dev, dtype=torch.float32) if need_levels else coords.new_empty(0)
    g1 = torch.zeros(level1.shape, device=dev, dtype=torch.float32) if need_levels else coords.new_empty(0)
    gc = torch.empty_like(coords) if need_coords else coords.new_empty(0)
    with torch.cuda.device(dev):
        with _timed("corr_lookup_bwd", B * h1 * w1 * (4 * (2 * radius + 2) ** 2 * 4 + 16 + 4 * 2 * (2 * radius + 1) ** 2)):
            check(lib.mrfa_corr_lookup_bwd(_p(grad_out), _p(level0), _p(level1), int(level0.dtype == torch.bfloat16),
                                           _p(coords), _p(g0) if need_levels else None, _p(g1) if need_levels else None,
                                           _p(gc) if need_coords else None, B, h1 * w1, H, W, map_batch_stride,
                                           row_offset, radius, map_layout, _stream()), "mrfa_corr_lookup_bwd")
    return g0, g1, gc


@corr_lookup_bwd.register_fake
def _(grad_out, level0, level1, coords, H, W, map_batch_stride, row_offset, radius, map_layout, need_levels, need_coords):
    f = lambda t: coords.new_empty(tuple(t.shape)) if need_levels else coords.new_empty(0)
    return f(level0), f(level1), torch.empty_like(coords) if need_coords else coords.new_empty(0)


def _cl_setup(ctx, inputs, output):
    level0, level1, coords, H, W, mbs, ro, radius, layout, _cl = inputs
    ctx.save_for_backward(level0, level1, coords)
    ctx.cfg = (H, W, mbs, ro, radius, layout)


def _cl_backward(ctx, g):
    level0, level1, coords = ctx.saved_tensors
    need_levels = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
    need_coords = ctx.needs_input_grad[2]
    g0, g1, gc = torch.ops.mrfa.corr_lookup_bwd(g, level0, level1, coords, *ctx.cfg, need_levels, need_coords)
    return (g0.to(level0.dtype) if ctx.needs_input_grad[0] else None,
            g1.to(level1.dtype) if ctx.needs_input_grad[1] else None,
            gc if need_coords else None, None, None, None, None, None, None, None)


corr_lookup.register_autograd(_cl_backward, setup_context=_cl_setup)


# ------------------------------------------------------------------------------------------
# fused elementwise passes (inference fast path of the conv blocks and the decoder blending)
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("mrfa::channel_affine", mutates_args=(), device_types="cuda")
def channel_affine(x: Tensor, scale: Optional[Tensor], shift: Optional[Tensor], residual: Optional[Tensor],
                   act: int) -> Tensor:
    """act(x * scale[c] + shift[c] + residual); act 0 none / 1 relu / 2 sigmoid."""
    x, cl = _req_image(x, "x")
    if residual is not None:
        residual = _like_layout(_req_image(residual, "residual")[0], cl)
    scale = None if scale is None else _req(scale, "scale")
    shift = None if shift is None else _req(shift, "shift")
    N, C, H, W = x.shape
    y = torch.empty_like(x)
    if y.numel() == 0:
        return y
    with torch.cuda.device(x.device):
        with _timed("channel_affine", 4 * x.numel() * (2 + (residual is not None))):
            check(lib.mrfa_channel_affine(_p(x), _p(scale), _p(shift), _p(residual), _p(y), N * H * W, C, H * W, int(cl),
                                          act, _stream()), "mrfa_channel_affine")
    return y


@channel_affine.register_fake
def _(x, scale, shift, residual, act):
    return torch.empty_like(x)


@torch.library.custom_op("mrfa::occlusion_blend", mutates_args=(), device_types="cuda")
def occlusion_blend(a: Tensor, b: Optional[Tensor], occ: Tensor) -> Tensor:
    """a * occ + b * (1 - occ) (b None: a * occ); occ (N,1,H,W) broadcast over channels."""
    a, cl = _req_image(a, "a")
    if b is not None:
        b = _like_layout(_req_image(b, "b")[0], cl)
    occ = _req(occ, "occ")
    N, C, H, W = a.shape
    if tuple(occ.shape) != (N, 1, H, W):
        raise RuntimeError("mrfa_b200: occlusion_blend expects occ of shape (N,1,H,W)")
    y = torch.empty_like(a)
    if y.numel() == 0:
        return y
    with torch.cuda.device(a.device):
        with _timed("occlusion_blend", 4 * (a.numel() * (2 + (b is not None)) + occ.numel())):
            check(lib.mrfa_occlusion_blend(_p(a), _p(b), _p(occ), _p(y), N * H * W, C, H * W, int(cl), _stream()),
                  "mrfa_occlusion_blend")
    return y


@occlusion_blend.register_fake
def _(a, b, occ):
    return torch.empty_like(a)


@torch.library.custom_op("mrfa::final_blend_s2d", mutates_args=(), device_types="cuda")
def final_blend_s2d(conv: Tensor, bias: Tensor, a: Tensor, occ: Tensor, r: int) -> Tensor:
    """a * occ + sigmoid(pixel_shuffle(conv, r) + bias[c]) * (1 - occ) in one pass (generator.py:61-63 behind the
    space-to-depth final convolution).  conv (B, C*r*r, H/r, W/r) channels_last, a (B,C,H,W) NCHW, occ (B,1,H,W)."""
    a, bias, occ = _req(a, "a"), _req(bias, "bias"), _req(occ, "occlusion")
    if not (conv.is_cuda and conv.dtype == torch.float32 and conv.dim() == 4
            and conv.is_contiguous(memory_format=torch.channels_last)):
        raise RuntimeError("mrfa_b200: final_blend_s2d expects a channels_last float32 CUDA convolution output")
    B, C, H, W = a.shape
    if conv.shape != (B, C * r * r, H // r, W // r) or H % r or W % r or bias.numel() != C:
        raise RuntimeError("mrfa_b200: final_blend_s2d shape mismatch")
    y = torch.empty_like(a)
    with torch.cuda.device(a.device):
        with _timed("final_blend_s2d", 4 * (conv.numel() + 2 * a.numel() + occ.numel())):
            check(lib.mrfa_final_blend_s2d(_p(conv), _p(bias), _p(a), _p(occ), _p(y), B, C, H, W, r, _stream()),
                  "mrfa_final_blend_s2d")
    return y


@final_blend_s2d.register_fake
def _(conv, bias, a, occ, r):
    return torch.empty_like(a)


@torch.library.custom_op("mrfa::resize_bilinear", mutates_args=(), device_types="cuda")
def resize_bilinear(x: Tensor, Ho: int, Wo: int, act: int, bias: Optional[Tensor] = None) -> Tensor:
    """act(F.interpolate(x, (Ho,Wo), mode='bilinear', align_corners=True) + bias[None, :, None, None]) with act 0 / 1 relu /
    2 sigmoid.  The memory format of `x` (NCHW or channels_last, any C) is preserved."""
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 4:
        raise RuntimeError("mrfa_b200: resize_bilinear expects a 4-D float32 CUDA tensor (there is no CPU fallback)")
    if bias is not None:
        bias = _req(bias, "bias")
        if bias.numel() != x.shape[1]:
            raise RuntimeError("mrfa_b200: resize_bilinear bias must have one value per channel")
    cl = _suggest_channels_last(x)
    x = x.contiguous(memory_format=torch.channels_last) if cl else x.contiguous()
    N, C, H, W = x.shape
    y = _empty_image((N, C, Ho, Wo), x.device, cl)
    if y.numel() == 0:
        return y
    with torch.cuda.device(x.device):
        with _timed("resize_bilinear", 4 * (x.numel() + y.numel())):
            check(lib.mrfa_resize_bilinear(_p(x), _p(bias), _p(y), N, C, H, W, Ho, Wo, int(cl), act, _stream()), "mrfa_resize_bilinear")
    return y


@resize_bilinear.register_fake
def _(x, Ho, Wo, act, bias=None):
    y = x.new_empty((x.shape[0], x.shape[1], Ho, Wo))
    return y.contiguous(memory_format=torch.channels_last) if _suggest_channels_last(x) else y


@torch.library.custom_op("mrfa::resize_strip", mutates_args=(), device_types="cuda")
def resize_strip(maps: Sequence[Tensor], Ho: int, Wo: int) -> Tensor:
    """torch.cat([F.interpolate(m, (Ho,Wo), mode='bilinear', align_corners=True) for m in maps], dim=3) (raft.py:304-306)
    without the intermediate maps and the cat pass: every map is resized straight into its column block of the strip.
    maps: (N, C, H_i, W_i) float32 CUDA tensors, plain contiguous (or C == 1); Wo % 4 == 0."""
    if len(maps) == 0:
        raise RuntimeError("mrfa_b200: resize_strip needs at least one map")
    N, C = maps[0].shape[:2]
    for m in maps:
        if not m.is_cuda or m.dtype != torch.float32 or m.dim() != 4 or m.shape[0] != N or m.shape[1] != C:
            raise RuntimeError("mrfa_b200: resize_strip expects 4-D float32 CUDA maps with equal N, C (there is no CPU fallback)")
    if Wo % 4 != 0:
        raise RuntimeError("mrfa_b200: resize_strip needs Wo % 4 == 0")
    y = torch.empty((N, C, Ho, Wo * len(maps)), device=maps[0].device, dtype=torch.float32)
    if y.numel() == 0:
        return y
    with torch.cuda.device(y.device):
        for i, m in enumerate(maps):
            m = m.contiguous()                          # planar (N*C, H, W); a no-op for C == 1 maps in either memory format
            with _timed("resize_bilinear", 4 * (m.numel() + N * C * Ho * Wo)):
                check(lib.mrfa_resize_bilinear_strip(_p(m), _p(y), N * C, m.shape[2], m.shape[3], Ho, Wo, Wo * len(maps), i * Wo, 0,
                                                     _stream()), "mrfa_resize_bilinear_strip")
    return y


@resize_strip.register_fake
def _(maps, Ho, Wo):
    return maps[0].new_empty((maps[0].shape[0], maps[0].shape[1], Ho, Wo * len(maps)))


def conv7x7_small_pack(weight: Tensor) -> Tensor:
    """(Cout, Cin, 7, 7) -> the (Cout, KP) K-major operand of mrfa_conv7x7_small (k = (ky*7+kx)*Cin + c)."""
    Cout, Cin = weight.shape[:2]
    kp = lib.mrfa_conv7x7_small_kpad(Cin)
    w = weight.detach().permute(0, 2, 3, 1).reshape(Cout, 49 * Cin)
    return torch.nn.functional.pad(w, (0, kp - 49 * Cin)).contiguous()


def conv7x7_small_ok(x: Tensor, Cin: int, Cout: int) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and (Cin, Cout) in ((2, 128), (3, 64))
            and x.shape[1] == Cin and x.shape[3] % 128 == 0 and x.shape[0] * x.shape[2] * x.shape[3] < 2 ** 31)


@torch.library.custom_op("mrfa::conv7x7_small", mutates_args=(), device_types="cuda")
def conv7x7_small(x: Tensor, w_packed: Tensor, bias: Optional[Tensor], relu: bool) -> Tensor:
    """relu?(conv2d(x, w, bias, padding=3)) for 7x7 kernels with 2-3 input channels, TF32 tensor cores;
    x in any strides, result channels_last.  See include/mrfa_b200.h."""
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 4:
        raise RuntimeError("mrfa_b200: conv7x7_small expects a 4-D float32 CUDA tensor (there is no CPU fallback)")
    w_packed = _req(w_packed, "w_packed")
    B, Cin, H, W = x.shape
    Cout = w_packed.shape[0]
    if w_packed.shape[1] != lib.mrfa_conv7x7_small_kpad(Cin):
        raise RuntimeError("mrfa_b200: conv7x7_small weight operand does not match the input channel count")
    if bias is not None:
        bias = _req(bias, "bias")
    y = _empty_image((B, Cout, H, W), x.device, True)
    if y.numel() == 0:
        return y
    st = x.stride()
    with torch.cuda.device(x.device):
        with _timed("conv7x7_small", 4 * (x.numel() + y.numel())):
            check(lib.mrfa_conv7x7_small(_p(x), GridStrides(st[0], st[2], st[3], st[1]), _p(w_packed),
                                         _p(bias) if bias is not None else None, _p(y), B, Cin, Cout, H, W, int(relu),
                                         sm_count(x.device), _stream()), "mrfa_conv7x7_small")
    return y


@conv7x7_small.register_fake
def _(x, w_packed, bias, relu):
    return x.new_empty((x.shape[0], w_packed.shape[0], x.shape[2], x.shape[3])).contiguous(memory_format=torch.channels_last)


@torch.library.custom_op("mrfa::avg_pool2x2_nhwc_bwd", mutates_args=(), device_types="cuda")
def avg_pool2x2_nhwc_bwd(grad_y: Tensor, H: int, W: int) -> Tensor:
    """Gradient of avg_pool2x2_nhwc w.r.t. its (N,C,H,W) channels_last input (H, W even)."""
    if not grad_y.is_cuda or grad_y.dtype != torch.float32 or grad_y.dim() != 4:
        raise RuntimeError("mrfa_b200: avg_pool2x2_nhwc_bwd expects a 4-D float32 CUDA tensor (there is no CPU fallback)")
    grad_y = grad_y.contiguous(memory_format=torch.channels_last)
    N, C = grad_y.shape[:2]
    gx = _empty_image((N, C, H, W), grad_y.device, True)
    if gx.numel() == 0:
        return gx
    with torch.cuda.device(grad_y.device):
        with _timed("avg_pool2x2_nhwc_bwd", 4 * (gx.numel() + grad_y.numel())):
            check(lib.mrfa_avg_pool2x2_nhwc_bwd(_p(grad_y), _p(gx), N, C, H, W, _stream()), "mrfa_avg_pool2x2_nhwc_bwd")
    return gx


@avg_pool2x2_nhwc_bwd.register_fake
def _(grad_y, H, W):
    return grad_y.new_empty((grad_y.shape[0], grad_y.shape[1], H, W)).contiguous(memory_format=torch.channels_last)


def _ap_setup(ctx, inputs, output):
    ctx.hw = tuple(inputs[0].shape[2:])


def _ap_backward(ctx, g):
    return torch.ops.mrfa.avg_pool2x2_nhwc_bwd(g, ctx.hw[0], ctx.hw[1])



@torch.library.custom_op("mrfa::cat2", mutates_args=(), device_types="cuda")
def cat2(a: Tensor, b: Tensor) -> Tensor:
    """torch.cat([a, b], 1) for two channels_last maps with channel counts divisible by 4."""
    (a, cla), (b, clb) = _req_image(a, "a"), _req_image(b, "b")
    N, Ca, H, W = a.shape
    Cb = b.shape[1]
    if not (cla and clb) or tuple(b.shape) != (N, Cb, H, W):
        raise RuntimeError("mrfa_b200: cat2 expects two channels_last maps of the same size with C % 4 == 0")
    y = _empty_image((N, Ca + Cb, H, W), a.device, True)
    if y.numel() == 0:
        return y
    with torch.cuda.device(a.device):
        with _timed("cat2", 4 * 2 * y.numel()):
            check(lib.mrfa_cat2_nhwc(_p(a), _p(b), _p(y), N * H * W, Ca, Cb, _stream()), "mrfa_cat2_nhwc")
    return y


@cat2.register_fake
def _(a, b):
    return a.new_empty((a.shape[0], a.shape[1] + b.shape[1], a.shape[2], a.shape[3])).contiguous(memory_format=torch.channels_last)


def cat2_ok(a: Tensor, b: Tensor) -> bool:
    return _is_channels_last(a) and _is_channels_last(b) and a.is_cuda and a.dtype == torch.float32 and b.dtype == torch.float32


@torch.library.custom_op("mrfa::subpixel_shuffle_cat", mutates_args=(), device_types="cuda")
def subpixel_shuffle_cat(b2: Tensor, skip: Tensor) -> Tensor:
    """cat([shuffle(b2), skip], 1) with b2 the phase-major sub-pixel up-conv output (N,4C,H+1,W+1) channels_last and
    skip (N,Cs,2H,2W) in any strides; result (N,C+Cs,2H,2W) channels_last.  See include/mrfa_b200.h."""
    b2, cl = _req_image(b2, "b2")
    if not skip.is_cuda or skip.dtype != torch.float32 or skip.dim() != 4:
        raise RuntimeError("mrfa_b200: subpixel_shuffle_cat `skip` must be a 4-D float32 CUDA tensor (there is no CPU fallback)")
    N, C4, H1, W1 = b2.shape
    C, H, W = C4 // 4, H1 - 1, W1 - 1
    Cs = skip.shape[1]
    if not cl or C4 % 4 or tuple(skip.shape) != (N, Cs, 2 * H, 2 * W):
        raise RuntimeError("mrfa_b200: subpixel_shuffle_cat expects channels_last b2 (N,4C,H+1,W+1) and skip (N,Cs,2H,2W)")
    y = _empty_image((N, C + Cs, 2 * H, 2 * W), b2.device, True)
    if y.numel() == 0:
        return y
    st = skip.stride()
    with torch.cuda.device(b2.device):
        with _timed("subpixel_shuffle_cat", 4 * 2 * y.numel()):
            check(lib.mrfa_subpixel_shuffle_cat(_p(b2), _p(skip), GridStrides(st[0], st[2], st[3], st[1]), _p(y), N, C, Cs, H, W,
                                                _stream()), "mrfa_subpixel_shuffle_cat")
    return y


@subpixel_shuffle_cat.register_fake
def _(b2, skip):
    N, C4, H1, W1 = b2.shape
    return b2.new_empty((N, C4 // 4 + skip.shape[1], 2 * (H1 - 1), 2 * (W1 - 1))).contiguous(memory_format=torch.channels_last)


@torch.library.custom_op("mrfa::random_warp_grid", mutates_args=(), device_types="cuda")
def random_warp_grid(theta: Tensor, control_points: Optional[Tensor], control_params: Optional[Tensor], h: int, w: int,
                     metric: int) -> Tensor:
    """Grid (B,h,w,2) of the random affine + TPS equivariance warp (model.py:44-70 metric 0; util.py TPS 'random' 1)."""
    theta = _req(theta, "theta")
    B = theta.shape[0]
    if tuple(theta.shape) != (B, 2, 3):
        raise RuntimeError("mrfa_b200: random_warp_grid expects theta of shape (B,2,3)")
    P = 0
    if control_params is not None:
        control_points = _req(control_points, "control_points").reshape(-1, 2)
        control_params = _req(control_params, "control_params").reshape(B, -1)
        P = control_points.shape[0]
        if control_params.shape[1] != P:
            raise RuntimeError("mrfa_b200: random_warp_grid control_params / control_points mismatch")
    grid = torch.empty((B, h, w, 2), device=theta.device, dtype=torch.float32)
    if grid.numel() == 0:
        return grid
    with torch.cuda.device(theta.device):
        with _timed("random_warp_grid", 4 * grid.numel()):
            check(lib.mrfa_random_warp_grid(_p(theta), _p(control_points) if P else None, _p(control_params) if P else None,
                                            _p(grid), B, P, h, w, metric, _stream()), "mrfa_random_warp_grid")
    return grid


@random_warp_grid.register_fake
def _(theta, control_points, control_params, h, w, metric):
    return theta.new_empty((theta.shape[0], h, w, 2))


@torch.library.custom_op("mrfa::flow_update", mutates_args=(), device_types="cuda")
def flow_update(flow: Tensor, occ: Tensor, d_flow: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """raft.py:256-262 in one kernel: (flow + d_flow[:, 0:2], occ + d_flow[:, 2:3], sigmoid(occ + d_flow[:, 2:3])).
    flow (B,2,H,W) in either memory format (kept), occ (B,1,H,W), d_flow (B,>=3,H,W) any strides."""
    for t, name in ((flow, "flow"), (occ, "occ"), (d_flow, "d_flow")):
        if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 4:
            raise RuntimeError(f"mrfa_b200: flow_update `{name}` must be a 4-D float32 CUDA tensor (there is no CPU fallback)")
    B, _, H, W = flow.shape
    if flow.shape[1] != 2 or tuple(occ.shape) != (B, 1, H, W) or d_flow.shape[1] < 3 or tuple(d_flow.shape[2:]) != (H, W) \
            or d_flow.shape[0] != B:
        raise RuntimeError("mrfa_b200: flow_update shape mismatch")
    cl = _suggest_channels_last(flow)
    flow = flow.contiguous(memory_format=torch.channels_last) if cl else flow.contiguous()
    occ = occ.contiguous()
    flow_w = _empty_image((B, 2, H, W), flow.device, cl)
    occ_new, occ_sig = torch.empty_like(occ), torch.empty_like(occ)
    if flow_w.numel() == 0:
        return flow_w, occ_new, occ_sig
    st = d_flow.stride()
    with torch.cuda.device(flow.device):
        with _timed("flow_update", 4 * 9 * B * H * W):
            check(lib.mrfa_flow_update(_p(flow), _p(occ), _p(d_flow), GridStrides(st[0], st[2], st[3], st[1]), _p(flow_w),
                                       _p(occ_new), _p(occ_sig), B, H, W, int(cl), _stream()), "mrfa_flow_update")
    return flow_w, occ_new, occ_sig


@flow_update.register_fake
def _(flow, occ, d_flow):
    fw = flow.new_empty(flow.shape)
    if _suggest_channels_last(flow):
        fw = fw.contiguous(memory_format=torch.channels_last)
    return fw, occ.new_empty(occ.shape), occ.new_empty(occ.shape)


@torch.library.custom_op("mrfa::flow_carry", mutates_args=(), device_types="cuda")
def flow_carry(d_flow: Tensor, init_flow: Tensor, prior_occ: Tensor, d_f_pre: Optional[Tensor],
               d_occ_pre: Optional[Tensor], scale: float, channels_last: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """raft.py:276-295 as one kernel: (flow, occlusion, d_f_acc, d_occ_acc) at twice the resolution of
    `d_flow` (B,3,R,R: flow update ++ occlusion update, any strides).  See include/mrfa_b200.h."""
    for t, name in ((d_flow, "d_flow"), (init_flow, "init_flow"), (prior_occ, "prior_occ")):
        if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 4:
            raise RuntimeError(f"mrfa_b200: flow_carry `{name}` must be a 4-D float32 CUDA tensor (there is no CPU fallback)")
    B, Cd, R, R2 = d_flow.shape
    if Cd < 3 or R != R2 or init_flow.shape[1] != 2 or prior_occ.shape[1] != 1 or init_flow.shape[2] != init_flow.shape[3] \
            or tuple(prior_occ.shape[2:]) != tuple(init_flow.shape[2:]):
        raise RuntimeError("mrfa_b200: flow_carry shape mismatch")
    if (d_f_pre is None) != (d_occ_pre is None):
        raise RuntimeError("mrfa_b200: flow_carry needs both or neither of d_f_pre / d_occ_pre")
    h = init_flow.shape[2]
    init_flow, prior_occ = init_flow.contiguous(), prior_occ.contiguous()
    fmt = torch.channels_last if channels_last else torch.contiguous_format
    if d_f_pre is not None:
        if tuple(d_f_pre.shape) != (B, 2, R, R) or tuple(d_occ_pre.shape) != (B, 1, R, R):
            raise RuntimeError("mrfa_b200: flow_carry previous-update shape mismatch")
        if not d_f_pre.is_cuda or d_f_pre.dtype != torch.float32:
            raise RuntimeError("mrfa_b200: flow_carry `d_f_pre` must be a float32 CUDA tensor (there is no CPU fallback)")
        d_f_pre, d_occ_pre = d_f_pre.contiguous(memory_format=fmt), _req(d_occ_pre, "d_occ_pre")
    dev = d_flow.device
    flow, d_f_acc = _empty_image((B, 2, 2 * R, 2 * R), dev, channels_last), _empty_image((B, 2, 2 * R, 2 * R), dev, channels_last)
    occ = torch.empty((B, 1, 2 * R, 2 * R), device=dev, dtype=torch.float32)
    d_occ_acc = torch.empty_like(occ)
    if flow.numel() == 0:
        return flow, occ, d_f_acc, d_occ_acc
    st = d_flow.stride()
    with torch.cuda.device(dev):
        with _timed("flow_carry", 4 * 6 * flow.numel() // 2):
            check(lib.mrfa_flow_carry(_p(d_flow), GridStrides(st[0], st[2], st[3], st[1]), _p(init_flow), _p(prior_occ),
                                      _p(d_f_pre) if d_f_pre is not None else None,
                                      _p(d_occ_pre) if d_occ_pre is not None else None,
                                      _p(flow), _p(occ), _p(d_f_acc), _p(d_occ_acc), B, R, h, float(scale),
                                      int(channels_last), _stream()), "mrfa_flow_carry")
    return flow, occ, d_f_acc, d_occ_acc


@flow_carry.register_fake
def _(d_flow, init_flow, prior_occ, d_f_pre, d_occ_pre, scale, channels_last):
    B, _, R, _ = d_flow.shape
    fmt = torch.channels_last if channels_last else torch.contiguous_format
    f = d_flow.new_empty((B, 2, 2 * R, 2 * R)).contiguous(memory_format=fmt)
    o = d_flow.new_empty((B, 1, 2 * R, 2 * R))
    return f, o, torch.empty_like(f), torch.empty_like(o)


@torch.library.custom_op("mrfa::antialias_down", mutates_args=(), device_types="cuda")
def antialias_down(x: Tensor, weight: Tensor, ka: int, stride: int) -> Tensor:
    """AntiAliasInterpolation2d for scale 1/stride, computing only the kept pixels (NCHW)."""
    x, weight = _req(x, "x"), _req(weight, "weight")
    N, C, H, W = x.shape
    K = weight.shape[-1]
    y = torch.empty((N, C, H // stride, W // stride), device=x.device, dtype=torch.float32)
    if y.numel() == 0:
        return y
    with torch.cuda.device(x.device):
        with _timed("antialias_down", 4 * (x.numel() + y.numel())):
            check(lib.mrfa_antialias_down(_p(x), _p(weight), _p(y), N, C, H, W, K, ka, stride, _stream()), "mrfa_antialias_down")
    return y


@antialias_down.register_fake
def _(x, weight, ka, stride):
    return x.new_empty((x.shape[0], x.shape[1], x.shape[2] // stride, x.shape[3] // stride))


@torch.library.custom_op("mrfa::occlusion_blend_subpixel", mutates_args=(), device_types="cuda")
def occlusion_blend_subpixel(a: Tensor, b2: Tensor, occ: Tensor, out_block: int = 1) -> Tensor:
    """a * occ + shuffle(b2) * (1 - occ) with b2 the phase-major sub-pixel up-conv output (see header).
    out_block r > 1 returns the result in r x r space-to-depth form: (N, r*r*C, 2H/r, 2W/r) channels_last."""
    a, cl = _req_image(a, "a")
    b2, cl2 = _req_image(b2, "b2")
    occ = _req(occ, "occ")
    N, C, H2, W2 = a.shape
    H, W = H2 // 2, W2 // 2
    if not (cl and cl2) or tuple(b2.shape) != (N, 4 * C, H + 1, W + 1) or tuple(occ.shape) != (N, 1, H2, W2):
        raise RuntimeError("mrfa_b200: occlusion_blend_subpixel expects channels_last a (N,C,2H,2W), b2 (N,4C,H+1,W+1), occ (N,1,2H,2W)")
    r = int(out_block)
    if r < 1 or H2 % r or W2 % r:
        raise RuntimeError("mrfa_b200: occlusion_blend_subpixel out_block must divide the output size")
    y = torch.empty_like(a) if r == 1 else _empty_image((N, r * r * C, H2 // r, W2 // r), a.device, True)
    if y.numel() == 0:
        return y
    with torch.cuda.device(a.device):
        with _timed("occlusion_blend", 4 * (2 * a.numel() + a.numel() + occ.numel())):
            check(lib.mrfa_occlusion_blend_subpixel(_p(a), _p(b2), _p(occ), _p(y), N, C, H, W, r, 0, _stream()),
                  "mrfa_occlusion_blend_subpixel")
    return y


@occlusion_blend_subpixel.register_fake
def _(a, b2, occ, out_block=1):
    if out_block == 1:
        return torch.empty_like(a)
    N, C, H2, W2 = a.shape
    r = out_block
    return a.new_empty((N, r * r * C, H2 // r, W2 // r)).contiguous(memory_format=torch.channels_last)


@torch.library.custom_op("mrfa::occlusion_blend_subpixel_into", mutates_args=("dst",), device_types="cuda")
def occlusion_blend_subpixel_into(a: Tensor, b2: Tensor, occ: Tensor, dst: Tensor) -> None:
    """occlusion_blend_subpixel written into channels [0, C) of `dst` (N,C',2H,2W) channels_last, C' >= C: the other
    half of the cat buffer started by mrfa::dual_warp_cat."""
    a, cl = _req_image(a, "a")
    b2, cl2 = _req_image(b2, "b2")
    occ = _req(occ, "occ")
    N, C, H2, W2 = a.shape
    H, W = H2 // 2, W2 // 2
    Ct = dst.shape[1]
    if not (cl and cl2) or tuple(b2.shape) != (N, 4 * C, H + 1, W + 1) or tuple(occ.shape) != (N, 1, H2, W2):
        raise RuntimeError("mrfa_b200: occlusion_blend_subpixel_into expects channels_last a (N,C,2H,2W), b2 (N,4C,H+1,W+1), occ (N,1,2H,2W)")
    if (not dst.is_cuda or dst.dtype != torch.float32 or tuple(dst.shape) != (N, Ct, H2, W2) or Ct < C or Ct % 4
            or not dst.is_contiguous(memory_format=torch.channels_last)):
        raise RuntimeError("mrfa_b200: occlusion_blend_subpixel_into expects a channels_last float32 dst (N,C'>=C,2H,2W)")
    if a.numel() == 0:
        return
    with torch.cuda.device(a.device):
        with _timed("occlusion_blend", 4 * (2 * a.numel() + a.numel() + occ.numel())):
            check(lib.mrfa_occlusion_blend_subpixel(_p(a), _p(b2), _p(occ), _p(dst), N, C, H, W, 1, Ct, _stream()),
                  "mrfa_occlusion_blend_subpixel")


@torch.library.custom_op("mrfa::avg_pool2x2_nhwc", mutates_args=(), device_types="cuda")
def avg_pool2x2_nhwc(x: Tensor) -> Tensor:
    """F.avg_pool2d(x, (2, 2)) for channels_last activations (C % 4 == 0)."""
    x, cl = _req_image(x, "x")
    if not cl:
        raise RuntimeError("mrfa_b200: avg_pool2x2_nhwc expects a channels_last tensor with C % 4 == 0")
    N, C, H, W = x.shape
    y = _empty_image((N, C, H // 2, W // 2), x.device, True)
    if y.numel() == 0:
        return y
    with torch.cuda.device(x.device):
        with _timed("avg_pool2x2_nhwc", 4 * (x.numel() + y.numel())):
            check(lib.mrfa_avg_pool2x2_nhwc(_p(x), _p(y), N, C, H, W, _stream()), "mrfa_avg_pool2x2_nhwc")
    return y


@avg_pool2x2_nhwc.register_fake
def _(x):
    y = x.new_empty((x.shape[0], x.shape[1], x.shape[2] // 2, x.shape[3] // 2))
    return y.contiguous(memory_format=torch.channels_last)


avg_pool2x2_nhwc.register_autograd(_ap_backward, setup_context=_ap_setup)
