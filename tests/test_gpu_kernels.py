"""GPU parity tests: every kernel of libmrfa_b200 (through the C ABI / torch custom ops)
against the CPU oracle and the golden vectors of the unmodified reference.

Tolerances (BASELINE.json north_star): bit-exact for index / coordinate-grid construction,
1e-5 abs for fp32 warps and lookups, 2e-2 relative for the bf16 correlation.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import np_ops as O                      # noqa: E402
from oracle import torch_path as TP                 # noqa: E402

DEV = "cuda"


def mb():
    import mrfa_b200
    return mrfa_b200


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def close(a, b, atol=1e-5, rtol=0.0):
    a = a.detach().float().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().float().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    np.testing.assert_allclose(a.astype(np.float64), b.astype(np.float64), atol=atol, rtol=rtol)


def rel_close(a, b, rel=2e-2):
    """|a-b| <= rel * (|b| + rms(b)): relative tolerance with an rms floor for near-zero entries."""
    a = a.detach().float().cpu().numpy().astype(np.float64) if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().float().cpu().numpy().astype(np.float64) if torch.is_tensor(b) else np.asarray(b, np.float64)
    rms = np.sqrt((b ** 2).mean())
    bad = np.abs(a - b) > rel * (np.abs(b) + rms)
    assert not bad.any(), f"{bad.sum()} / {bad.size} outside {rel} rel; max abs err {np.abs(a - b).max():.4g}, rms {rms:.4g}"


# ------------------------------------------------------------------ no CPU fallback
def test_cpu_tensors_are_rejected():
    m = mb()
    with pytest.raises(Exception):
        m.bilinear_sampler(torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 2, 2))
    with pytest.raises(Exception):
        m.coords_grid(1, 2, 2, "cpu")


# ------------------------------------------------------------------ grids: bit-exact
def test_grids_bit_exact(golden):
    m, g = mb(), golden("grids")
    for (h, w) in ((2, 3), (16, 16), (64, 64), (5, 7), (128, 128)):
        out = m.make_coordinate_grid((h, w), "torch.cuda.FloatTensor").cpu().numpy()
        assert np.array_equal(out, g[f"mcg_{h}x{w}"])
    for (b, h, w) in ((1, 2, 3), (2, 8, 8), (1, 64, 64)):
        assert np.array_equal(m.coords_grid(b, h, w, DEV).cpu().numpy(), g[f"cg_{b}_{h}x{w}"])
    assert np.array_equal(m.coords_grid(3, 256, 256, DEV).cpu().numpy(), O.coords_grid(3, 256, 256))
    assert np.array_equal(m.make_coordinate_grid((256, 192), "torch.cuda.FloatTensor").cpu().numpy(),
                          O.make_coordinate_grid(256, 192))


def test_kp2gaussian(golden):
    m, g = mb(), golden("grids")
    kp = cu(g["kp"])
    close(m.kp2gaussian(kp, (16, 16), 0.1), g["kp2g_16_0.1"], 1e-6)
    close(m.kp2gaussian(kp, (16, 16), 0.01), g["kp2g_16_0.01"], 1e-6)
    close(m.kp2gaussian(kp, (12, 20), 0.01), g["kp2g_12x20_0.01"], 1e-6)
    pos = torch.randn(1, 10, 16, 16, device=DEV) * 0.02
    out = torch.ops.mrfa.kp2gaussian(kp, pos, 16, 16, 0.1)
    close(out, g["kp2g_16_0.1"] + pos.cpu().numpy(), 1e-6)


# ------------------------------------------------------------------ warps
def test_samplers_golden(golden):
    m, s = mb(), golden("samplers")
    img = cu(s["img"])
    close(m.bilinear_sampler(img, cu(s["pix_coords"])), s["bilinear_sampler"])
    _, mask = m.bilinear_sampler(img, cu(s["pix_coords"]), mask=True)
    assert np.array_equal(mask.cpu().numpy(), s["bilinear_sampler_mask"])
    close(m.grid_sample(img, cu(s["norm_grid"])), s["grid_sample_acF"])
    close(m.grid_sample(img, cu(s["norm_grid"]), align_corners=True), s["grid_sample_acT"])
    close(m.grid_sample(img, cu(s["norm_grid"]) * 1.7, padding_mode="reflection"), s["grid_sample_reflect_acF"])
    close(m.batch_bilinear_sampler(cu(s["bimg"]), cu(s["bco"]), h=2, w=2, mini_batch=1), s["batch_bilinear_mb1"])
    out = m.batch_bilinear_sampler(cu(s["bimg"]), cu(s["bco"]), h=2, w=2, mini_batch=2)
    assert tuple(out.shape) == (8, 1, 3, 3)
    close(out, s["batch_bilinear_mb2"])
    close(m.deform_input(cu(s["feat"]), cu(s["prior"])), s["coarse_warp"])


@pytest.mark.parametrize("C,R", [(512, 8), (512, 16), (256, 64), (64, 128), (3, 256), (5, 37)])
def test_feature_warp_shapes(C, R):
    """The six feature-warp shapes of RaftFlow (reduced batch) + a ragged one, all conventions."""
    m = mb()
    torch.manual_seed(C * R)
    B = 2
    feat = torch.randn(B, C, R, R)
    flow = torch.randn(B, 2, R, R) * 2.5               # some taps leave the image
    ident = TP.coords_grid(B, R, R)
    ref = TP.bilinear_sampler(feat, (flow + ident).permute(0, 2, 3, 1))
    close(m.warp_by_flow(feat.to(DEV), flow.to(DEV)), ref)
    close(m.bilinear_sampler(feat.to(DEV), (flow + ident).permute(0, 2, 3, 1).to(DEV)), ref)
    grid = TP.make_coordinate_grid((R, R))[None].repeat(B, 1, 1, 1) + torch.randn(B, R, R, 2) * 0.1
    close(m.grid_sample(feat.to(DEV), grid.to(DEV)), F.grid_sample(feat, grid, align_corners=False))
    close(m.grid_sample(feat.to(DEV), grid.to(DEV), align_corners=True), F.grid_sample(feat, grid, align_corners=True))
    a, b = torch.ops.mrfa.dual_warp(feat.to(DEV), flow.to(DEV), grid.to(DEV))
    close(a, ref)
    close(b, F.grid_sample(feat, grid, align_corners=False))


@pytest.mark.parametrize("C", [1, 2, 3, 4])
def test_few_channel_image_warp(C):
    """Few-channel NCHW maps with many pixels (the full-resolution image warp, raft.py:302) take the four-pixels-per-thread
    kernel: ragged pixel count, every convention, some taps outside the image, a shared input for several grids."""
    m = mb()
    torch.manual_seed(100 + C)
    B, R, S = 3, 150, 151                              # 3 * 150 * 151 pixels: above the switch point, not a multiple of 512
    img = torch.rand(B, C, R, S)
    flow = torch.randn(B, 2, R, S) * 3.0
    ident = TP.coords_grid(B, R, S)
    ref = TP.bilinear_sampler(img, (flow + ident).permute(0, 2, 3, 1))
    close(m.warp_by_flow(img.to(DEV), flow.to(DEV)), ref)
    grid = TP.make_coordinate_grid((R, S))[None].repeat(B, 1, 1, 1) + torch.randn(B, R, S, 2) * 0.1
    close(m.grid_sample(img.to(DEV), grid.to(DEV)), F.grid_sample(img, grid, align_corners=False))
    close(m.grid_sample(img.to(DEV), grid.to(DEV), align_corners=True), F.grid_sample(img, grid, align_corners=True))
    close(torch.ops.mrfa.grid_sample(img.to(DEV), grid.to(DEV), 0, 1, False, 1),
          F.grid_sample(img, grid, padding_mode="reflection", align_corners=False))
    # one input for three grids (in_batch_div): the fused `.repeat` of dense_motion.py:80-81
    got = torch.ops.mrfa.grid_sample(img[:1].to(DEV), grid.to(DEV), 0, 0, False, 3)
    close(got, F.grid_sample(img[:1].expand(3, -1, -1, -1), grid, align_corners=False))


def test_warp_edge_cases():
    m = mb()
    feat = torch.randn(1, 2, 4, 4, device=DEV)
    # everything out of the image -> zeros; NaN / inf coordinates -> zeros, no fault
    far = torch.full((1, 3, 3, 2), 1e6, device=DEV)
    assert m.bilinear_sampler(feat, far).abs().max().item() == 0
    bad = torch.tensor([float("nan"), float("inf")], device=DEV).view(1, 1, 1, 2)
    assert torch.isfinite(m.bilinear_sampler(feat, bad)).all()
    # exact pixel centres reproduce the input
    ident = m.coords_grid(1, 4, 4, DEV).permute(0, 2, 3, 1)
    close(m.bilinear_sampler(feat, ident), feat, 1e-6)
    # empty batch
    assert m.bilinear_sampler(feat[:0], ident[:0]).shape == (0, 2, 4, 4)


@pytest.mark.parametrize("mode", ["pixel", "acF", "acT", "reflect"])
def test_warp_backward(mode):
    m = mb()
    torch.manual_seed(3)
    B, C, H, W, Ho, Wo = 2, 6, 9, 11, 7, 5
    feat = torch.randn(B, C, H, W, dtype=torch.float64)
    if mode == "pixel":
        grid = torch.rand(B, Ho, Wo, 2, dtype=torch.float64) * torch.tensor([W + 2.0, H + 2.0], dtype=torch.float64) - 1.5
    else:
        grid = torch.rand(B, Ho, Wo, 2, dtype=torch.float64) * 2.6 - 1.3
    go = torch.randn(B, C, Ho, Wo, dtype=torch.float64)
    f64, g64 = feat.clone().requires_grad_(), grid.clone().requires_grad_()
    if mode == "pixel":
        gn = torch.stack([2 * g64[..., 0] / (W - 1) - 1, 2 * g64[..., 1] / (H - 1) - 1], -1)
        ref = F.grid_sample(f64, gn, align_corners=True)
    elif mode == "reflect":
        ref = F.grid_sample(f64, g64, padding_mode="reflection", align_corners=False)
    else:
        ref = F.grid_sample(f64, g64, align_corners=(mode == "acT"))
    ref.backward(go)
    f32 = feat.float().to(DEV).requires_grad_()
    g32 = grid.float().to(DEV).requires_grad_()
    if mode == "pixel":
        out = m.bilinear_sampler(f32, g32)
    elif mode == "reflect":
        out = m.grid_sample(f32, g32, padding_mode="reflection")
    else:
        out = m.grid_sample(f32, g32, align_corners=(mode == "acT"))
    out.backward(go.float().to(DEV))
    close(out, ref, 1e-5)
    close(f32.grad, f64.grad, 2e-5)
    close(g32.grad, g64.grad, 5e-4, 1e-4)


def test_dual_warp_backward():
    torch.manual_seed(4)
    B, C, R = 2, 5, 12
    feat = torch.randn(B, C, R, R, dtype=torch.float64)
    flow = torch.randn(B, 2, R, R, dtype=torch.float64) * 1.5
    grid = (TP.make_coordinate_grid((R, R))[None].double() + torch.randn(B, R, R, 2, dtype=torch.float64) * 0.1)
    ga, gb = torch.randn(B, C, R, R, dtype=torch.float64), torch.randn(B, C, R, R, dtype=torch.float64)
    f64, fl64, g64 = feat.clone().requires_grad_(), flow.clone().requires_grad_(), grid.clone().requires_grad_()
    co = (fl64 + TP.coords_grid(B, R, R).double()).permute(0, 2, 3, 1)
    gn = torch.stack([2 * co[..., 0] / (R - 1) - 1, 2 * co[..., 1] / (R - 1) - 1], -1)
    (F.grid_sample(f64, gn, align_corners=True) * ga).sum().add((F.grid_sample(f64, g64, align_corners=False) * gb).sum()).backward()
    f32, fl32, g32 = (t.float().to(DEV).requires_grad_() for t in (feat, flow, grid))
    a, b = torch.ops.mrfa.dual_warp(f32, fl32, g32)
    ((a * ga.float().to(DEV)).sum() + (b * gb.float().to(DEV)).sum()).backward()
    close(f32.grad, f64.grad, 5e-5)
    close(fl32.grad, fl64.grad, 5e-4, 1e-4)
    close(g32.grad, g64.grad, 2e-3, 1e-4)


# ------------------------------------------------------------------ correlation
def test_corr_pack_and_volume_small(golden):
    m, c = mb(), golden("corr")
    q, k = cu(c["q_d"]), cu(c["k_s"])
    B, C, h, w = q.shape
    a_op, b_op = m.ops.corr_pack_debug(q, k)
    qa = c["q_d"].reshape(B, C, -1).transpose(0, 2, 1)
    assert np.array_equal(a_op[:, :h * w].float().cpu().numpy(), O.round_bf16(qa))
    assert np.array_equal(b_op.float().cpu().numpy(), O.round_bf16(c["k_s"].reshape(B, C, -1).transpose(0, 2, 1)))
    off = 0
    for lvl in range(4):
        kk = 2 ** lvl
        pooled = O.avg_pool2d(c["q_d"], kk) if kk > 1 else c["q_d"]
        n = pooled.shape[-1] * pooled.shape[-2]
        exp = O.round_bf16(pooled.reshape(B, C, n).transpose(0, 2, 1))
        got = a_op[:, off:off + n].float().cpu().numpy()
        assert np.abs(got - exp).max() <= 2 ** -7 * np.abs(exp).max()        # within one bf16 ulp of the mean
        assert m.ops.corr_row_offset(h, w, lvl) == off
        off += n
    assert off == m.ops.corr_rows_total(h, w)

    pyr = m.CorrPyramid(q, k, C ** -0.5)
    vol = pyr.volume0.float().cpu().numpy()
    rel_close(vol[:, :h * w], c["volume"])                                     # vs the reference's fp32 einsum
    # against the exact product of the bf16 operands: only the output rounding remains
    exact = O.corr_volume(c["q_d"], c["k_s"], C ** -0.5, bf16_inputs=True)
    rel_close(vol[:, :h * w], exact, 2 ** -7)
    # pooled driving levels and the source-pooled level 1
    ref = c["volume"]
    for lvl in (1, 2):
        rows = O.corr_pyramid_rows(ref, h, w, 2 ** lvl)
        o = m.ops.corr_row_offset(h, w, lvl)
        rel_close(vol[:, o:o + rows.shape[1]], rows)
    l1 = O.avg_pool2d(ref.reshape(B, h * w, 1, h, w), 2).reshape(B, h * w, -1)
    rel_close(pyr.volume1[:, :h * w].float().cpu().numpy(), l1)


@pytest.mark.parametrize("h,w,C,B", [(16, 8, 64, 3), (32, 32, 128, 2), (64, 64, 256, 2), (8, 64, 512, 1)])
def test_corr_volume_sizes(h, w, C, B):
    m = mb()
    torch.manual_seed(h + C)
    q = torch.randn(B, C, h, w, device=DEV)
    k = torch.randn(B, C, h, w, device=DEV)
    pyr = m.CorrPyramid(q, k, C ** -0.5)
    kd = k.flatten(2).transpose(1, 2).double()
    off = 0
    for lvl in range(4):
        kk = 2 ** lvl
        qp = F.avg_pool2d(q.double(), kk) if kk > 1 else q.double()
        r = torch.einsum("bic,bjc->bij", qp.flatten(2).transpose(1, 2), kd) * C ** -0.5
        # dense(): row-major view of the stored maps ((64, 64) is stored tiled, the small shapes row-major)
        rel_close(pyr.dense(lvl).view(B, -1, h * w), r)
        l1 = F.avg_pool2d(r.view(B, -1, h, w), 2).flatten(2)
        rel_close(pyr.dense(lvl, 1).view(B, -1, h * w // 4), l1)
        off += r.shape[1]
    assert off == pyr.rows_total
    assert pyr.layout == (m._lib.MAP_TILED if w in (64, 128) else m._lib.MAP_ROWMAJOR)


def test_corr_shape_errors():
    m = mb()
    with pytest.raises(RuntimeError):                     # h*w not a multiple of the 128-wide tile
        m.CorrPyramid(torch.randn(1, 64, 8, 8, device=DEV), torch.randn(1, 64, 8, 8, device=DEV), 1.0)
    with pytest.raises(RuntimeError):                     # C not a multiple of 64
        m.CorrPyramid(torch.randn(1, 32, 16, 16, device=DEV), torch.randn(1, 32, 16, 16, device=DEV), 1.0)


def test_corr_volume_512_tile_geometry():
    """w = 128 (the 512x512 configuration): 256-wide tiles, one source row pair per tile."""
    m = mb()
    torch.manual_seed(9)
    B, C, h = 1, 64, 128
    q = torch.randn(B, C, h, h, device=DEV)
    k = torch.randn(B, C, h, h, device=DEV)
    pyr = m.CorrPyramid(q, k, C ** -0.5)
    N = h * h
    idx = torch.randint(0, N, (64,), device=DEV)
    qd = q.flatten(2).transpose(1, 2)[:, idx].double()
    kd = k.flatten(2).transpose(1, 2).double()
    ref = torch.einsum("bic,bjc->bij", qd, kd) * C ** -0.5
    assert pyr.layout == m._lib.MAP_TILED
    p0 = m.ops.corr_map_permutation(pyr.layout, 0, h, h, DEV)
    p1 = m.ops.corr_map_permutation(pyr.layout, 1, h // 2, h // 2, DEV)
    rel_close(pyr.volume0[:, idx].float().index_select(2, p0), ref)
    rel_close(pyr.volume1[:, idx].float().index_select(2, p1), F.avg_pool2d(ref.view(B, -1, h, h), 2).flatten(2))


def test_corr_lookup_golden(golden):
    m, c = mb(), golden("corr")
    B, h = 2, 16
    vol = c["volume"]
    for k in (1, 2, 4):
        rows = O.corr_pyramid_rows(vol, h, h, k)
        l0 = cu(rows.reshape(-1, 1, h, h))
        blk = m.CorrBlock(l0)
        close(blk.corr_pyramid[1], c[f"level1_k{k}"], 1e-6)
        close(blk(cu(c[f"coords_k{k}"])), c[f"lookup_k{k}"])


def test_corr_lookup_from_bf16_pyramid(golden):
    m, c = mb(), golden("corr")
    q, k = cu(c["q_d"]), cu(c["k_s"])
    B, C, h, w = q.shape
    pyr = m.CorrPyramid(q, k, C ** -0.5)
    for lvl, kk in ((0, 1), (1, 2), (2, 4)):
        coords = cu(c[f"coords_k{kk}"])
        got = pyr.block(lvl)(coords)
        # same lookup evaluated by the oracle on the very bf16 maps the kernel read
        off = m.ops.corr_row_offset(h, w, lvl)
        Q = (h // kk) * (w // kk)
        l0 = pyr.volume0[:, off:off + Q].float().reshape(B * Q, 1, h, w).cpu().numpy()
        l1 = pyr.volume1[:, off:off + Q].float().reshape(B * Q, 1, h // 2, w // 2).cpu().numpy()
        close(got, O.corr_lookup([l0, l1], c[f"coords_k{kk}"]))
        rel_close(got, c[f"lookup_k{kk}"])                                    # and 2e-2 of the fp32 reference


@pytest.mark.parametrize("h,w,C", [(64, 64, 64), (16, 128, 64)])
def test_corr_lookup_from_tiled_pyramid(h, w, C):
    """w = 64 / 128: the volume is stored in 4 x 8 tiles (include/mrfa_b200.h "Map layouts") and the lookup walks tiles with
    16-byte loads.  Every driving level, both output layouts, windows hanging over every border, against the oracle lookup
    evaluated on the very bf16 maps the kernel read (un-tiled by CorrPyramid.dense)."""
    m = mb()
    torch.manual_seed(h + w)
    B = 2
    q = torch.randn(B, C, h, w, device=DEV)
    k = torch.randn(B, C, h, w, device=DEV)
    pyr = m.CorrPyramid(q.contiguous(memory_format=torch.channels_last), k.contiguous(memory_format=torch.channels_last), C ** -0.5)
    ref = m.CorrPyramid(q, k, C ** -0.5)                                    # NCHW pack path
    assert pyr.layout == m._lib.MAP_TILED and ref.layout == m._lib.MAP_TILED
    assert torch.equal(pyr.volume0[:, :h * w], ref.volume0[:, :h * w]) and torch.equal(pyr.volume1[:, :h * w], ref.volume1[:, :h * w])
    for lvl in range(4):
        kk = 2 ** lvl
        hq, wq = h // kk, w // kk
        coords = torch.rand(B, 2, hq, wq) * torch.tensor([w + 10.0, h + 10.0]).view(1, 2, 1, 1) - 5.0
        coords[0, :, 0, 0] = torch.tensor([0.0, 0.0])
        coords[0, :, 0, 1] = torch.tensor([w - 1.0, h - 1.0])
        coords[1, :, 0, 0] = torch.tensor([-40.0, 3.5])                      # window entirely outside
        coords[1, :, 0, 1] = torch.tensor([8.0, 4.0])                        # footprint aligned to a tile boundary
        l0 = pyr.dense(lvl).cpu().numpy()
        l1 = pyr.dense(lvl, 1).cpu().numpy()
        exp = O.corr_lookup([l0, l1], coords.numpy())
        close(pyr.block(lvl)(coords.to(DEV)), exp)
        got_cl = pyr.block(lvl)(coords.to(DEV), True)
        assert got_cl.is_contiguous(memory_format=torch.channels_last)
        close(got_cl, exp)
    # radius 1 and 2 share the tile walk (smaller footprints), radius 4 falls back to the per-element kernel
    coords = torch.rand(B, 2, h, w) * torch.tensor([w + 6.0, h + 6.0]).view(1, 2, 1, 1) - 3.0
    l0, l1 = pyr.dense(0).cpu().numpy(), pyr.dense(0, 1).cpu().numpy()
    for r in (1, 2, 4):
        exp = O.corr_lookup([l0, l1], coords.numpy(), radius=r)
        close(pyr.block(0, radius=r)(coords.to(DEV)), exp)
        close(pyr.block(0, radius=r)(coords.to(DEV), True), exp)            # channels-last output: the warp-autonomous kernel
    # ragged query counts (Q not a multiple of the 4 queries a warp pass takes; a pass never straddles two samples)
    for hq, wq in ((3, 5), (1, 1), (7, 9)):
        coords = torch.rand(B, 2, hq, wq) * torch.tensor([w + 6.0, h + 6.0]).view(1, 2, 1, 1) - 3.0
        blk = pyr.block(0)
        blk._stride, blk._offset = pyr.rows_total, 0
        exp = O.corr_lookup([l0.reshape(B, h * w, 1, h, w)[:, :hq * wq].reshape(B * hq * wq, 1, h, w),
                             l1.reshape(B, h * w, 1, h // 2, w // 2)[:, :hq * wq].reshape(B * hq * wq, 1, h // 2, w // 2)], coords.numpy())
        close(blk(coords.to(DEV), True), exp)
        close(blk(coords.to(DEV)), exp)


def test_corr_lookup_backward_tiled_pyramid():
    """Gradients through lookups on a tiled pyramid (w = 64): d/d(coords) and d/d(q_d, k_s) against fp64 autograd on the
    dense reference formulation (einsum volume -> avg-pool pyramid -> bilinear windows)."""
    m = mb()
    torch.manual_seed(21)
    B, C, h, w = 1, 64, 8, 64
    q = torch.randn(B, C, h, w, dtype=torch.float64)
    k = torch.randn(B, C, h, w, dtype=torch.float64)
    coords = torch.rand(B, 2, h, w, dtype=torch.float64) * torch.tensor([w + 2.0, h + 2.0]).view(1, 2, 1, 1) - 1.0
    go = torch.randn(B, 98, h, w, dtype=torch.float64)
    q64, k64, x64 = q.clone().requires_grad_(), k.clone().requires_grad_(), coords.clone().requires_grad_()
    vol = torch.einsum("bic,bjc->bij", q64.flatten(2).transpose(1, 2), k64.flatten(2).transpose(1, 2)) * C ** -0.5
    TP.corr_lookup(vol.reshape(B * h * w, 1, h, w), x64).backward(go)
    q32 = q.float().to(DEV).requires_grad_()
    k32 = k.float().to(DEV).requires_grad_()
    x32 = coords.float().to(DEV).requires_grad_()
    pyr = m.CorrPyramid(q32, k32, C ** -0.5)
    assert pyr.layout == m._lib.MAP_TILED
    pyr.block(0)(x32).backward(go.float().to(DEV))
    for name, g, r in (("q_d", q32.grad, q64.grad), ("k_s", k32.grad, k64.grad), ("coords", x32.grad, x64.grad)):
        a, b = g.double().cpu().flatten(), r.flatten()
        rel = float((a - b).norm() / b.norm())
        assert rel < 3e-2, (name, rel)                                        # bf16 volume and bf16 gradient GEMM operands


@pytest.mark.parametrize("h,w,C,cl", [(16, 16, 64, False), (8, 64, 128, True), (16, 64, 256, False), (8, 16, 512, True)])
def test_corr_pyramid_backward_gemms(h, w, C, cl):
    """mrfa::corr_pyramid_bwd (tcgen05 GEMMs dA = G Bm, dB = G^T A + pack / transpose / un-pool kernels) against fp64 autograd
    through the dense formulation the reference differentiates: einsum volume at every pooled driving level (raft.py:185,
    :219), avg_pool2d level 1 (raft.py:20).  Row-major (w = 16) and tiled (w = 64) map layouts, NCHW and NHWC operands."""
    m = mb()
    torch.manual_seed(h * w + C)
    B = 2
    N, rows = h * w, m.ops.corr_rows_total(h, w)
    scale = C ** -0.5
    q = torch.randn(B, C, h, w, dtype=torch.float64)
    k = torch.randn(B, C, h, w, dtype=torch.float64)
    g0 = torch.randn(B, rows, N)                       # gradients in the STORED map layout
    g1 = torch.randn(B, rows, N // 4)
    layout = m.ops.corr_map_layout(h, w)
    assert layout == (m._lib.MAP_TILED if w == 64 else m._lib.MAP_ROWMAJOR)
    p0 = m.ops.corr_map_permutation(layout, 0, h, w, "cpu")
    p1 = m.ops.corr_map_permutation(layout, 1, h // 2, w // 2, "cpu")
    g0d, g1d = g0.index_select(2, p0).double(), g1.index_select(2, p1).double()      # row-major view of the same gradients
    q64, k64 = q.clone().requires_grad_(), k.clone().requires_grad_()
    kd = k64.flatten(2).transpose(1, 2)
    levels = [q64] + [F.avg_pool2d(q64, 2 ** l) for l in (1, 2, 3)]
    a_full = torch.cat([t.flatten(2).transpose(1, 2) for t in levels], dim=1)          # (B, rows_total, C)
    vol0 = torch.einsum("bic,bjc->bij", a_full, kd) * scale
    vol1 = F.avg_pool2d(vol0.view(B, rows, h, w), 2).flatten(2)
    ((vol0 * g0d).sum() + (vol1 * g1d).sum()).backward()
    fmt = torch.channels_last if cl else torch.contiguous_format
    dq, dk = torch.ops.mrfa.corr_pyramid_bwd(g0.to(DEV), g1.to(DEV), q.float().to(DEV).contiguous(memory_format=fmt),
                                             k.float().to(DEV).contiguous(memory_format=fmt), scale)
    assert dq.shape == q.shape and dk.shape == k.shape
    for name, got, ref in (("d_q", dq, q64.grad), ("d_k", dk, k64.grad)):
        a, b = got.double().cpu().flatten(), ref.flatten()
        rel = float((a - b).norm() / b.norm())
        # bf16 G and bf16 operands, fp32 accumulation over K = hw / rows_total
        assert rel < 1e-2, (name, rel)
        rel_close(got, ref, 5e-2)


def test_corr_epilogue_variants_bit_identical():
    """The GEMM epilogue rounds half of its outputs to bf16 on the FMA / ALU pipes (pack_bf16_fma, exact RNE by magic-number
    addition) and half with cvt.rn.bf16x2.f32; store modes 0 / 1 / 2 stage the same values differently.  Every combination
    must give bit-identical volumes (MRFA_CORR_CVT / MRFA_CORR_STORE are read per launch)."""
    import os
    m = mb()
    torch.manual_seed(33)
    q = torch.randn(2, 256, 64, 64, device=DEV) * 3.0
    k = torch.randn(2, 256, 64, 64, device=DEV) * 3.0
    q[0, :, 0, 0] = 0.0                                                        # exact zeros and tiny values
    k[1, :, 5, 7] *= 1e-20
    saved = {v: os.environ.get(v) for v in ("MRFA_CORR_CVT", "MRFA_CORR_STORE")}
    try:
        ref = None
        for cvt in ("0", "1"):
            for store in ("0", "1", "2"):
                os.environ["MRFA_CORR_CVT"], os.environ["MRFA_CORR_STORE"] = cvt, store
                pyr = m.CorrPyramid(q, k, 256 ** -0.5)
                if ref is None:
                    ref = (pyr.volume0.clone(), pyr.volume1.clone())
                    # and the XU-only reference path equals torch's own RNE cast of an fp32 product of the bf16 operands closely
                    continue
                assert torch.equal(pyr.volume0, ref[0]), (cvt, store)
                assert torch.equal(pyr.volume1, ref[1]), (cvt, store)
    finally:
        for v, val in saved.items():
            if val is None:
                os.environ.pop(v, None)
            else:
                os.environ[v] = val


def test_corr_lookup_edge_cases():
    m = mb()
    torch.manual_seed(5)
    B, h1, H = 1, 3, 10                      # ragged query count (9 < 32), odd map size
    corr = torch.randn(B * h1 * h1, 1, H, H)
    coords = torch.rand(B, 2, h1, h1) * (H + 6) - 3
    coords[0, :, 0, 0] = torch.tensor([-50.0, 4.0])         # window entirely outside
    coords[0, :, 0, 1] = torch.tensor([0.0, 0.0])           # corner
    coords[0, :, 0, 2] = torch.tensor([H - 1.0, H - 1.0])
    ref = TP.corr_lookup(corr, coords)
    close(m.CorrBlock(corr.to(DEV))(coords.to(DEV)), ref)


def test_corr_lookup_backward():
    m = mb()
    torch.manual_seed(6)
    B, h1, H = 2, 4, 12
    corr = torch.randn(B * h1 * h1, 1, H, H, dtype=torch.float64)
    coords = torch.rand(B, 2, h1, h1, dtype=torch.float64) * (H + 2) - 1
    go = torch.randn(B, 98, h1, h1, dtype=torch.float64)
    c64, x64 = corr.clone().requires_grad_(), coords.clone().requires_grad_()

    def ref_lookup(level0, co):
        n, r = 7, 3
        d = torch.linspace(-r, r, n, dtype=co.dtype)
        delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, n, n, 2)
        centre = co.permute(0, 2, 3, 1).reshape(-1, 1, 1, 2)
        outs, lvl_map = [], level0
        for lvl in range(2):
            if lvl:
                lvl_map = F.avg_pool2d(lvl_map, 2, stride=2)
            outs.append(TP.bilinear_sampler(lvl_map, centre / 2 ** lvl + delta).view(co.shape[0], h1, h1, n * n))
        return torch.cat(outs, -1).permute(0, 3, 1, 2)

    ref_lookup(c64, x64).backward(go)
    c32, x32 = corr.float().to(DEV).requires_grad_(), coords.float().to(DEV).requires_grad_()
    out = m.CorrBlock(c32)(x32)
    out.backward(go.float().to(DEV))
    close(c32.grad, c64.grad, 5e-5)
    close(x32.grad, x64.grad, 5e-4, 1e-4)


def test_corr_lookup_batched_sampler_branch():
    """h1 = w1 = 128 with batch 2: the reference switches to batch_bilinear_sampler (raft.py:39-40, util.py:40-51), one
    sample per chunk.  The chunking is a memory workaround -- the kernel's single launch must give the same values."""
    m = mb()
    torch.manual_seed(12)
    B, h1, H = 2, 128, 16
    corr = torch.randn(B * h1 * h1, 1, H, H)
    coords = torch.rand(B, 2, h1, h1) * (H + 4) - 2
    ref = TP.corr_lookup(corr, coords)                  # takes the `B > 1 and h1 >= 128` branch
    close(m.CorrBlock(corr.to(DEV))(coords.to(DEV)), ref)
    # and the drop-in batch_bilinear_sampler itself with one sample per chunk (the dropped-remainder case is in the goldens)
    img = torch.randn(B * 9, 1, H, H)
    pts = torch.rand(B * 9, 7, 7, 2) * (H + 2) - 1
    got = m.batch_bilinear_sampler(img.to(DEV), pts.to(DEV), "bilinear", False, 3, 3, 1)
    ref = torch.cat([TP.bilinear_sampler(img[i * 9:(i + 1) * 9], pts[i * 9:(i + 1) * 9]) for i in range(B)], 0)
    close(got, ref)


def test_prior_to_flow_direct():
    """raft.py:189-190: init_flow = (self.h - 1) * (deformation + 1) / 2 - coords_grid, with self.h used for BOTH axes
    (checked on a non-square map), forward at 1e-5 and the backward factor (h - 1) / 2."""
    m = mb()
    torch.manual_seed(13)
    for (B, h, w, hm1) in ((2, 64, 64, 63.0), (3, 6, 10, 5.0), (1, 128, 128, 127.0)):
        d = (torch.rand(B, h, w, 2) * 2.4 - 1.2)
        ref = hm1 * (d.permute(0, 3, 1, 2) + 1) / 2.0 - TP.coords_grid(B, h, w)
        x = d.to(DEV).requires_grad_()
        got = torch.ops.mrfa.prior_to_flow(x, hm1)
        assert got.shape == (B, 2, h, w)
        close(got, ref, 1e-5)
        go = torch.randn(B, 2, h, w)
        got.backward(go.to(DEV))
        close(x.grad, go.permute(0, 2, 3, 1) * (hm1 / 2.0), 1e-6)


# ------------------------------------------------------------------ prior dense motion
def _prior_inputs():
    import synthetic_inputs as syn
    src, _ = syn.frame_pairs(2, 64, seed=1)
    kp_s, kp_d = syn.keypoints(2, 10, seed=1)
    return src, kp_s, kp_d, syn.bg_affine(2, seed=1)


def _to(d):
    return {k: v.to(DEV) for k, v in d.items()}


def _cfg():
    import os
    import yaml
    return yaml.safe_load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vox1.yaml")))


def test_dense_motion_prior_kernels(golden):
    import synthetic_inputs as syn
    m, d = mb(), golden("prior_motion")
    _, kp_s, kp_d, bg = _prior_inputs()
    cfg = _cfg()
    net = m.DenseMotionNetwork(**dict(cfg["dense_motion"], block_expansion=16, max_features=64, num_blocks=3)).to(DEV)
    small = cu(d["source_small"])
    close(net.create_heatmap_representations(small, _to(kp_d), _to(kp_s)), d["heatmap"], 1e-6)
    close(net.create_sparse_motions(small, _to(kp_d), _to(kp_s), bg_param=bg.to(DEV)), d["sparse_motions_jac_bg"])
    close(net.create_sparse_motions(small, {"kp": kp_d["kp"].to(DEV)}, {"kp": kp_s["kp"].to(DEV)}), d["sparse_motions_plain"], 1e-6)
    close(net.create_deformed_source_image(small, cu(d["sparse_motions_jac_bg"])), d["deformed"])
    # fused path writes the same numbers into the hourglass input layout
    motions, hg = torch.ops.mrfa.dense_motion_prior(kp_d["kp"].to(DEV), kp_s["kp"].to(DEV), kp_d["jacobian"].to(DEV),
                                                    kp_s["jacobian"].to(DEV), bg.to(DEV), small, 0.01)
    hg = hg.view(2, 11, 4, 16, 16)
    close(hg[:, :, 1:], d["deformed"])
    close(hg[:, :, :1], d["heatmap"], 1e-6)


def test_tps_kernels(golden):
    m, d = mb(), golden("prior_motion")
    import synthetic_inputs as syn
    bg = syn.bg_affine(2, seed=1).to(DEV)
    kp_d, kp_s = cu(d["tps_kp_d"]), cu(d["tps_kp_s"])
    tps = m.TPS(mode="kp", bs=2, kp_1=kp_d.view(2, -1, 5, 2), kp_2=kp_s.view(2, -1, 5, 2))
    # the fp64 in-kernel solve must agree with the fp64 oracle solve to fp32 resolution ...
    th, cp, cw = O.tps_params(d["tps_kp_d"].reshape(2, -1, 5, 2), d["tps_kp_s"].reshape(2, -1, 5, 2))
    close(tps.theta, th, 1e-5, 1e-5)
    close(tps.control_params, cw, 1e-5, 1e-5)
    # ... and with the reference's own parameters: its fp32 inverse sits within 2e-6 of the fp64 solve on these inputs
    # (tests/test_reference_vendored.py::test_tps_reference_solve_agrees_with_fp64), so no loosened bound is needed
    close(tps.theta, d["tps_theta"], 3e-5, 1e-5)
    close(tps.control_params, d["tps_control_params"], 3e-5, 1e-5)
    cfg = _cfg()
    net = m.TPSDenseMotionNetwork(**dict(cfg["tpsm_dense_motion"], block_expansion=16, max_features=64, num_blocks=3)).to(DEV)
    small = cu(d["source_small"])
    got = net.create_transformations(small, {"kp": kp_d}, {"kp": kp_s}, bg)
    close(got, O.tps_transformations(d["tps_kp_d"], d["tps_kp_s"], 16, 16, bg.cpu().numpy()), 2e-5)
    close(got, d["tps_transformations_bg"], 3e-5)
    close(tps.transform_frame(small), O.tps_transformations(d["tps_kp_d"], d["tps_kp_s"], 16, 16)[:, 1:], 2e-5)


def _grad_close(got, ref, rel=2e-4):
    """Gradients that are sums over the h*w plane: |got - ref| <= rel * max|ref| (fp32 summation-order noise)."""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    scale = float(ref.abs().max()) + 1e-12
    err = float((got - ref).abs().max())
    assert err <= rel * scale, f"max |err| {err:.3e} vs scale {scale:.3e}"


def _leaf(t):
    return t.detach().clone().requires_grad_(True)


@pytest.mark.parametrize("use_jac,use_bg,src_grad", [(True, True, True), (False, False, True), (True, False, False),
                                                     (False, True, False)])
def test_dense_motion_prior_backward(use_jac, use_bg, src_grad):
    """Fused backward of the FOMM / MTIA prior synthesis (csrc/motion_bwd.cu) against CPU autograd through the oracle's
    restatement of dense_motion.py:36-85 (heat-maps, sparse motions, deformed sources), non-square map."""
    import synthetic_inputs as syn
    mb()
    B, K, C, h, w = 3, 10, 3, 24, 20
    g = torch.Generator().manual_seed(11)
    kp_s0, kp_d0 = syn.keypoints(B, K, seed=4)
    bg0 = syn.bg_affine(B, seed=4)
    src0 = torch.rand(B, C, h, w, generator=g)
    w_hg = torch.randn(B, (K + 1) * (C + 1), h, w, generator=g)
    w_mo = torch.randn(B, K + 1, h, w, 2, generator=g)

    def run(dev):
        t = {"kd": _leaf(kp_d0["kp"].to(dev)), "ks": _leaf(kp_s0["kp"].to(dev)), "src": src0.detach().clone().to(dev).requires_grad_(src_grad)}
        if use_jac:
            t["jd"], t["js"] = _leaf(kp_d0["jacobian"].to(dev)), _leaf(kp_s0["jacobian"].to(dev))
        if use_bg:
            t["bg"] = _leaf(bg0.to(dev))
        if dev == "cpu":
            kd = {"kp": t["kd"]}
            ks = {"kp": t["ks"]}
            if use_jac:
                kd["jacobian"], ks["jacobian"] = t["jd"], t["js"]
            heat = TP.kp2gaussian(t["kd"], (h, w), 0.01) - TP.kp2gaussian(t["ks"], (h, w), 0.01)
            heat = torch.cat([torch.zeros_like(heat[:, :1]), heat], dim=1).unsqueeze(2)
            motions = TP.DenseMotionOracle.sparse_motions(None, h, w, kd, ks, t.get("bg"))
            rep = t["src"][:, None].expand(B, K + 1, -1, h, w).reshape(B * (K + 1), -1, h, w)
            deformed = F.grid_sample(rep, motions.view(B * (K + 1), h, w, 2), align_corners=False).view(B, K + 1, -1, h, w)
            hg = torch.cat([heat, deformed], dim=2).view(B, -1, h, w)
        else:
            motions, hg = torch.ops.mrfa.dense_motion_prior(t["kd"], t["ks"], t.get("jd"), t.get("js"), t.get("bg"), t["src"], 0.01)
        ((hg * w_hg.to(dev)).sum() + (motions * w_mo.to(dev)).sum()).backward()
        return t, motions, hg

    ref, mo_r, hg_r = run("cpu")
    got, mo_g, hg_g = run(DEV)
    close(mo_g, mo_r, 2e-5)
    close(hg_g, hg_r, 2e-5)
    for k in ref:
        if ref[k].grad is None:
            assert got[k].grad is None
            continue
        _grad_close(got[k].grad, ref[k].grad)


@pytest.mark.parametrize("use_bg,src_grad", [(True, True), (False, False)])
def test_tps_prior_backward(use_bg, src_grad):
    """Fused backward of the thin-plate-spline prior (motion / heat-map / warp gradients + adjoint 8x8 solve in fp64)
    against fp64 CPU autograd through the oracle's restatement of dense_motion.py:200-243 / util.py:355-410."""
    import synthetic_inputs as syn
    mb()
    B, G, C, h, w = 2, 10, 3, 20, 24
    g = torch.Generator().manual_seed(12)
    kd0 = (torch.rand(B, G * 5, 2, generator=g) * 1.6 - 0.8)
    ks0 = kd0 + 0.05 * torch.randn(B, G * 5, 2, generator=g)
    bg0 = syn.bg_affine(B, seed=5)
    src0 = torch.rand(B, C, h, w, generator=g)
    w_hg = torch.randn(B, G * 5 + 1 + (G + 1) * C, h, w, generator=g)
    w_mo = torch.randn(B, G + 1, h, w, 2, generator=g)

    # fp64 reference: the reference formula evaluated in double (its own fp32 inverse carries ~1e-4 of round-off into
    # these gradients; the kernel solves in fp64)
    dt = torch.float64
    kd, ks = _leaf(kd0.to(dt)), _leaf(ks0.to(dt))
    bg = _leaf(bg0.to(dt)) if use_bg else None
    src = src0.to(dt).requires_grad_(src_grad)
    grid = TP.make_coordinate_grid((h, w), dtype=dt)
    diff_d = grid.view(1, 1, h, w, 2) - kd.view(B, -1, 1, 1, 2)
    diff_s = grid.view(1, 1, h, w, 2) - ks.view(B, -1, 1, 1, 2)
    heat = torch.exp(-0.5 * (diff_d ** 2).sum(-1) / 0.01) - torch.exp(-0.5 * (diff_s ** 2).sum(-1) / 0.01)
    heat = torch.cat([torch.zeros_like(heat[:, :1]), heat], dim=1)
    kp_1, kp_2 = kd.view(B, G, 5, 2), ks.view(B, G, 5, 2)
    Kmat = torch.norm(kp_1[:, :, :, None] - kp_1[:, :, None, :], dim=4, p=2) ** 2
    Kmat = Kmat * torch.log(Kmat + 1e-9)
    kp1p = torch.cat([kp_1, torch.ones(B, G, 5, 1, dtype=dt)], 3)
    L = torch.cat([torch.cat([Kmat, kp1p.permute(0, 1, 3, 2)], 2), torch.cat([kp1p, torch.zeros(B, G, 3, 3, dtype=dt)], 2)], 3)
    L = L + torch.eye(8, dtype=dt).expand(L.shape) * float(np.float32(0.01))
    param = torch.matmul(torch.inverse(L), torch.cat([kp_2, torch.zeros(B, G, 3, 2, dtype=dt)], 2))
    theta, weights = param[:, :, 5:, :].permute(0, 1, 3, 2), param[:, :, :5, :]
    pts = grid.view(1, h * w, 2)
    aff = torch.matmul(theta[..., :2], pts.permute(0, 2, 1)) + theta[..., 2:]
    d2 = ((pts.view(1, 1, 1, -1, 2) - kp_1.view(B, G, -1, 1, 2)) ** 2).sum(-1)
    rbf = torch.matmul((d2 * torch.log(d2 + 1e-9)).permute(0, 1, 3, 2), weights)
    moved = (aff.permute(0, 1, 3, 2) + rbf).view(B, G, h, w, 2)
    bgm = grid.view(1, 1, h, w, 2).repeat(B, 1, 1, 1, 1)
    if use_bg:
        hom = torch.cat([bgm, torch.ones_like(bgm[..., :1])], dim=-1)
        hom = torch.matmul(bg.view(B, 1, 1, 1, 3, 3), hom.unsqueeze(-1)).squeeze(-1)
        bgm = hom[..., :2] / hom[..., 2:3]
    motions = torch.cat([bgm, moved], dim=1)
    rep = src[:, None].expand(B, G + 1, -1, h, w).reshape(B * (G + 1), -1, h, w)
    deformed = F.grid_sample(rep, motions.view(B * (G + 1), h, w, 2), align_corners=True).view(B, G + 1, -1, h, w)
    hg = torch.cat([heat, deformed.reshape(B, -1, h, w)], dim=1)
    ((hg * w_hg.to(dt)).sum() + (motions * w_mo.to(dt)).sum()).backward()

    kd_g, ks_g = _leaf(kd0.to(DEV)), _leaf(ks0.to(DEV))
    bg_g = _leaf(bg0.to(DEV)) if use_bg else None
    src_g = src0.to(DEV).requires_grad_(src_grad)
    mo_g, hg_g, _, _ = torch.ops.mrfa.tps_prior(kd_g, ks_g, bg_g, src_g, 0.01)
    ((hg_g * w_hg.to(DEV)).sum() + (mo_g * w_mo.to(DEV)).sum()).backward()
    close(mo_g, motions, 1e-4)
    close(hg_g, hg, 2e-4)
    # the bilinear-sample derivative is piecewise constant in the position: an fp32 position that rounds across a pixel
    # boundary flips one tap pair, so the warp part of the key-point gradients carries ~1e-3 of the fp32 forward's noise
    _grad_close(kd_g.grad, kd.grad, 2e-3)
    _grad_close(ks_g.grad, ks.grad, 2e-3)
    if use_bg:
        _grad_close(bg_g.grad, bg.grad, 2e-3)
    if src_grad:
        _grad_close(src_g.grad, src.grad, 2e-3)
    else:
        assert src_g.grad is None


def test_kp2gaussian_backward():
    """d kp of util.kp2gaussian through the reduction kernel, and the pos_embedding gradient (raft.py:177-178)."""
    mb()
    g = torch.Generator().manual_seed(13)
    kp0 = torch.rand(4, 10, 2, generator=g) * 1.6 - 0.8
    add0 = 0.02 * torch.randn(1, 10, 12, 18, generator=g)
    wgt = torch.randn(4, 10, 12, 18, generator=g)
    kp, add = _leaf(kp0), _leaf(add0)
    ((TP.kp2gaussian(kp, (12, 18), 0.1) + add) * wgt).sum().backward()
    kp_g, add_g = _leaf(kp0.to(DEV)), _leaf(add0.to(DEV))
    (torch.ops.mrfa.kp2gaussian(kp_g, add_g, 12, 18, 0.1) * wgt.to(DEV)).sum().backward()
    _grad_close(kp_g.grad, kp.grad)
    _grad_close(add_g.grad, add.grad)


def test_dense_motion_networks_forward(golden):
    import synthetic_inputs as syn
    m, d = mb(), golden("prior_motion")
    src, kp_s, kp_d, bg = _prior_inputs()
    cfg = _cfg()
    with torch.no_grad():
        net = syn.fill_state_dict_(m.DenseMotionNetwork(**dict(cfg["dense_motion"], block_expansion=16, max_features=64,
                                                               num_blocks=3))).to(DEV).eval()
        out = net(src.to(DEV), _to(kp_d), _to(kp_s), bg_param=bg.to(DEV))
        for k in ("sparse_deformed", "logit_mask", "mask", "deformation", "occlusion"):
            close(out[k], d["fwd_" + k], 1e-4)
        out = net(src.to(DEV), {"kp": kp_d["kp"].to(DEV)}, {"kp": kp_s["kp"].to(DEV)})
        close(out["deformation"], d["fwd_plain_deformation"], 1e-4)
        tnet = syn.fill_state_dict_(m.TPSDenseMotionNetwork(**dict(cfg["tpsm_dense_motion"], block_expansion=16,
                                                                   max_features=64, num_blocks=3))).to(DEV).eval()
        out = tnet(src.to(DEV), {"kp": cu(d["tps_kp_d"])}, {"kp": cu(d["tps_kp_s"])}, bg_param=bg.to(DEV))
        for k in ("deformed_source", "contribution_maps", "deformation", "occlusion"):
            close(out[k], d["tps_fwd_" + k], 2e-4)


# ------------------------------------------------------------------ seeded random shape sweep
def test_random_shapes_warps_and_lookups():
    """Ragged / odd shapes drawn from a seeded generator: every warp convention in both layouts against the stock ops the
    reference calls, and row-major lookups of every radius against the oracle (non-square maps, window sizes 3..9, query
    planes that are not multiples of the 32-query groups)."""
    m = mb()
    rng = np.random.default_rng(2024)
    for _ in range(12):
        B = int(rng.integers(1, 4))
        C = int(rng.choice([1, 2, 3, 4, 5, 8, 12, 20, 36, 64]))
        H, W = int(rng.integers(3, 41)), int(rng.integers(3, 41))
        Ho, Wo = int(rng.integers(2, 37)), int(rng.integers(2, 37))
        g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
        feat = torch.randn(B, C, H, W, generator=g)
        grid = torch.rand(B, Ho, Wo, 2, generator=g) * 2.6 - 1.3                    # ~12 % of the samples leave the map
        pix = torch.rand(B, Ho, Wo, 2, generator=g) * torch.tensor([W + 6.0, H + 6.0]) - 3.0
        layouts = [feat.to(DEV)]
        if C % 4 == 0:
            layouts.append(feat.to(DEV).contiguous(memory_format=torch.channels_last))
        for f in layouts:
            close(m.grid_sample(f, grid.to(DEV)), F.grid_sample(feat, grid, align_corners=False))
            close(m.grid_sample(f, grid.to(DEV), align_corners=True), F.grid_sample(feat, grid, align_corners=True))
            close(m.grid_sample(f, grid.to(DEV), padding_mode="reflection"),
                  F.grid_sample(feat, grid, padding_mode="reflection", align_corners=False), 2e-5)
            close(m.bilinear_sampler(f, pix.to(DEV)), TP.bilinear_sampler(feat, pix))
    for _ in range(8):
        B = int(rng.integers(1, 3))
        H, W = 2 * int(rng.integers(4, 20)), 2 * int(rng.integers(4, 20))
        h1, w1 = int(rng.integers(1, 12)), int(rng.integers(1, 12))
        radius = int(rng.integers(1, 5))
        g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
        corr = torch.randn(B * h1 * w1, 1, H, W, generator=g)
        coords = torch.rand(B, 2, h1, w1, generator=g) * torch.tensor([W + 8.0, H + 8.0]).view(1, 2, 1, 1) - 4.0
        lv1 = F.avg_pool2d(corr, 2, stride=2)
        exp = O.corr_lookup([corr.numpy(), lv1.numpy()], coords.numpy(), radius=radius)
        blk = m.CorrBlock(corr.to(DEV), radius=radius)
        close(blk(coords.to(DEV)), exp)
        close(blk(coords.to(DEV), True), exp)


def test_random_shapes_warp_backward():
    """Backward of the warps at ragged shapes in both layouts (the NHWC run-walk kernel pre-reduces its scatter-adds along a
    row: odd widths, single rows and channel counts that do not fill a lane group) against fp64 autograd of the stock op.
    Sample positions keep 0.02 px away from cell borders, where the bilinear derivative jumps."""
    m = mb()
    rng = np.random.default_rng(77)
    for it in range(10):
        B = int(rng.integers(1, 3))
        C = int(rng.choice([1, 3, 4, 8, 12, 32, 64]))
        H, W = int(rng.integers(2, 30)), int(rng.integers(2, 30))
        Ho, Wo = int(rng.integers(1, 26)), int(rng.integers(1, 26))
        g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
        feat = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
        pix = torch.rand(B, Ho, Wo, 2, generator=g, dtype=torch.float64) * torch.tensor([W + 3.0, H + 3.0]) - 1.5
        frac = pix - pix.floor()
        pix = pix.floor() + frac.clamp(0.02, 0.98)
        act = bool(it % 2)
        size = torch.tensor([float(W), float(H)], dtype=torch.float64)
        grid = (2 * pix / (size - 1) - 1) if act else ((2 * pix + 1) / size - 1)        # normalised positions of `pix`
        go = torch.randn(B, C, Ho, Wo, generator=g, dtype=torch.float64)
        f64, g64 = feat.clone().requires_grad_(), grid.clone().requires_grad_()
        F.grid_sample(f64, g64, align_corners=act).backward(go)
        for cl in ([False, True] if C % 4 == 0 else [False]):
            f32 = feat.float().to(DEV)
            if cl:
                f32 = f32.contiguous(memory_format=torch.channels_last)
            f32 = f32.requires_grad_()
            g32 = grid.float().to(DEV).requires_grad_()
            m.grid_sample(f32, g32, align_corners=act).backward(go.float().to(DEV))
            scale = float(f64.grad.abs().max()) + 1e-9
            assert float((f32.grad.double().cpu() - f64.grad).abs().max()) <= 1e-4 * max(1.0, scale), (it, cl)
            gs = float(g64.grad.abs().max()) + 1e-9
            assert float((g32.grad.double().cpu() - g64.grad).abs().max()) <= 2e-3 * max(1.0, gs), (it, cl)


# ------------------------------------------------------------------ full-size properties (BASELINE.json config 2: B = 64)
def test_full_size_correlation_and_lookup_properties():
    """Size-independent properties at the benchmarked size (64 pairs, 64 x 64 maps, C = 256, 3.6 GB of volume):
    scaling a factor by a power of two scales the bf16 volume bit-exactly; the driving-pooled rows are the average of the
    basic rows they pool (pooling commutes with the contraction) within bf16 rounding; a lookup at integer coordinates
    returns the stored volume entries (weights 1, 0, 0, 0 up to the coordinate round trip's last ulp) with zeros outside
    the maps."""
    m = mb()
    torch.manual_seed(7)
    B, C, h, w = 64, 256, 64, 64
    q = torch.randn(B, C, h, w, device=DEV).contiguous(memory_format=torch.channels_last)
    k = torch.randn(B, C, h, w, device=DEV).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        pyr = m.CorrPyramid(q, k, C ** -0.5)
        pyr2 = m.CorrPyramid(q * 2.0, k, C ** -0.5)
        assert torch.equal(pyr2.volume0.float(), pyr.volume0.float() * 2.0)
        assert torch.equal(pyr2.volume1.float(), pyr.volume1.float() * 2.0)
        del pyr2
        # level-1 pooled driving rows (32 x 32 queries) of a few samples vs the mean of their four basic rows
        off1 = m.ops.corr_row_offset(h, w, 1)
        for b in (0, 17, 63):
            basic = pyr.volume0[b, :h * w].float().view(h // 2, 2, w // 2, 2, h * w).mean(dim=(1, 3)).view(-1, h * w)
            pooled = pyr.volume0[b, off1:off1 + h * w // 4].float()
            rel_close(pooled, basic, 2e-2)                     # both sides rounded to bf16 separately
        # integer coordinates: every window tap of query (y, x) is a stored cell
        ys, xs = torch.meshgrid(torch.arange(h, device=DEV), torch.arange(w, device=DEV), indexing="ij")
        shift = torch.randint(-6, 7, (B, 2, 1, 1), device=DEV)
        coords = torch.stack([xs, ys]).float()[None] + shift.float()                    # (B,2,h,w), integer valued
        out = pyr.block(0)(coords, True)                                              # NHWC (B,98,h,w)
        # check three samples against the un-tiled maps
        for b in (0, 31, 63):
            sub = m.CorrPyramid(q[b:b + 1], k[b:b + 1], C ** -0.5)
            assert torch.equal(sub.volume0, pyr.volume0[b:b + 1])                      # batch entries are independent
            for lvl in (0, 1):
                maps = sub.dense(0, lvl)[:, 0]                                          # (h*w, H_l, W_l) fp32 copies of the bf16 cells
                Hl, Wl = h >> lvl, w >> lvl
                cx = (coords[b, 0].view(-1) / 2 ** lvl)
                cy = (coords[b, 1].view(-1) / 2 ** lvl)
                integral = (cx == cx.floor()) & (cy == cy.floor())                     # level 1: only even coordinates are cell centres
                for a, bb in ((0, 0), (3, 3), (6, 2), (1, 6)):
                    x = cx.long() + a - 3
                    y = cy.long() + bb - 3
                    inside = (x >= 0) & (x < Wl) & (y >= 0) & (y < Hl)
                    exp = torch.where(inside, maps[torch.arange(h * w, device=DEV), y.clamp(0, Hl - 1), x.clamp(0, Wl - 1)],
                                      torch.zeros((), device=DEV))
                    got = out[b, lvl * 49 + a * 7 + bb].reshape(-1)
                    # not bit for bit: the replayed coordinate round trip returns x only to within an ulp (see the warp test)
                    assert float((got[integral] - exp[integral]).abs().max()) <= 1e-4, (b, lvl, a, bb)


def test_full_size_warp_identity_properties():
    """Zero flow reproduces the feature map at every benchmarked level (B = 64), in both memory layouts, through the single
    and the dual warp kernels and the few-channel image warp; a one-pixel integer shift is a shift with a zero border.
    Not bit for bit: the kernels replay the reference's coordinate round trip x -> 2x/(W-1)-1 -> ((g+1)/2)(W-1)
    (util.py:30-31 + grid_sample), which returns x only to within an ulp of the coordinate (1.5e-5 at x = 255): the blend
    of two neighbouring N(0,1) values then moves by up to ~1e-4, exactly as in the reference."""
    m = mb()
    torch.manual_seed(8)
    B = 64

    def same(a, b):
        assert float((a - b).abs().max()) <= 2e-4

    with torch.no_grad():
        for C, R in ((512, 8), (512, 32), (256, 64), (64, 256), (3, 256)):
            feat = torch.randn(B, C, R, R, device=DEV)
            zero = torch.zeros(B, 2, R, R, device=DEV)
            prior = m.make_coordinate_grid((R, R), "torch.cuda.FloatTensor")[None].repeat(B, 1, 1, 1).contiguous()
            layouts = [feat] if C % 4 else [feat, feat.contiguous(memory_format=torch.channels_last)]
            for f in layouts:
                same(m.warp_by_flow(f, zero), f)
                if C % 4 == 0:
                    same(torch.ops.mrfa.dual_warp(f, zero, prior)[0], f)
                shift = zero.clone()
                shift[:, 0] = 1.0                                                       # sample one pixel to the right
                got = m.warp_by_flow(f, shift)
                same(got[..., :-1], f[..., 1:])
                assert float(got[..., -1].abs().max()) <= 2e-4
            del feat, zero, prior


# ------------------------------------------------------------------ channels-last (NHWC) paths
@pytest.mark.parametrize("C,R", [(512, 8), (256, 64), (64, 128), (4, 19)])
def test_feature_warp_channels_last(C, R):
    m = mb()
    torch.manual_seed(C + R)
    B = 2
    feat = torch.randn(B, C, R, R)
    flow = torch.randn(B, 2, R, R) * 2.5
    ident = TP.coords_grid(B, R, R)
    ref = TP.bilinear_sampler(feat, (flow + ident).permute(0, 2, 3, 1))
    fcl = feat.to(DEV).contiguous(memory_format=torch.channels_last)
    out = m.warp_by_flow(fcl, flow.to(DEV))
    assert out.is_contiguous(memory_format=torch.channels_last)
    close(out, ref)
    grid = TP.make_coordinate_grid((R, R))[None].repeat(B, 1, 1, 1) + torch.randn(B, R, R, 2) * 0.1
    close(m.grid_sample(fcl, grid.to(DEV)), F.grid_sample(feat, grid, align_corners=False))
    close(m.grid_sample(fcl, grid.to(DEV), align_corners=True), F.grid_sample(feat, grid, align_corners=True))
    close(m.grid_sample(fcl, grid.to(DEV) * 1.6, padding_mode="reflection"),
          F.grid_sample(feat, grid * 1.6, padding_mode="reflection", align_corners=False))
    a, b = torch.ops.mrfa.dual_warp(fcl, flow.to(DEV), grid.to(DEV))
    assert a.is_contiguous(memory_format=torch.channels_last)
    close(a, ref)
    close(b, F.grid_sample(feat, grid, align_corners=False))


@pytest.mark.parametrize("C,R,B,flow_kind", [(64, 128, 7, "smooth"), (64, 128, 7, "wild"), (128, 64, 26, "smooth"),
                                             (512, 32, 100, "smooth"), (12, 200, 3, "smooth"), (64, 192, 3, "edges")])
def test_run_walk_warps(C, R, B, flow_kind):
    """>= 100k output pixels: the NHWC warps switch to the run-walk kernels, which keep the right-hand taps of a pixel in
    registers for its neighbour when the sample positions chain (smooth motion) and reload otherwise.  Smooth flows (chains
    hold), per-pixel random flows (chains break everywhere), flows leaving the image, and a channel count that is not a
    multiple of 8 (128-bit path) must all give the plain result to 1e-5 -- refined, coarse and cat-slice outputs."""
    m = mb()
    torch.manual_seed(C + R)
    assert B * R * R >= 100000
    feat = torch.randn(B, C, R, R)
    if flow_kind == "smooth":
        flow = F.interpolate(torch.randn(B, 2, R // 8, R // 8) * 3.0, size=(R, R), mode="bilinear", align_corners=True)
        noise = 0.002
    elif flow_kind == "edges":
        flow = F.interpolate(torch.randn(B, 2, 4, 4) * 0.4 * R, size=(R, R), mode="bilinear", align_corners=True)   # leaves the image
        noise = 0.3
    else:
        flow = torch.randn(B, 2, R, R) * 4.0
        noise = 0.3
    grid = TP.make_coordinate_grid((R, R))[None].repeat(B, 1, 1, 1) + \
        F.interpolate(torch.randn(B, 2, 8, 8) * noise * 10, size=(R, R), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    ident = TP.coords_grid(B, R, R)
    ref_r = TP.bilinear_sampler(feat, (flow + ident).permute(0, 2, 3, 1))
    ref_c = F.grid_sample(feat, grid, align_corners=False)
    fcl = feat.to(DEV).contiguous(memory_format=torch.channels_last)
    # For information: the same stock ops on the GPU.  torch's CUDA `x / (W - 1)` of bilinear_sampler multiplies by a
    # reciprocal, so the stock CUDA result itself differs from the stock CPU result by up to ~1e-4 here; the oracle (and the
    # tolerance) is the CPU arithmetic, which the kernels replay operation by operation.
    cg = (flow + ident).permute(0, 2, 3, 1).to(DEV)
    gn = torch.cat([2 * cg[..., 0:1] / (R - 1) - 1, 2 * cg[..., 1:2] / (R - 1) - 1], dim=-1)
    stock_r = F.grid_sample(feat.to(DEV), gn, align_corners=True)
    stock_c = F.grid_sample(feat.to(DEV), grid.to(DEV), align_corners=False)
    a, b = torch.ops.mrfa.dual_warp(fcl, flow.to(DEV), grid.to(DEV))
    outs = {"warp_by_flow": (m.warp_by_flow(fcl, flow.to(DEV)), ref_r, stock_r), "grid_sample": (m.grid_sample(fcl, grid.to(DEV)), ref_c, stock_c),
            "dual.refined": (a, ref_r, stock_r), "dual.coarse": (b, ref_c, stock_c)}
    if C % 4 == 0:
        a2, buf = torch.ops.mrfa.dual_warp_cat(fcl, flow.to(DEV), grid.to(DEV))
        outs["cat.refined"], outs["cat.coarse"] = (a2, ref_r, stock_r), (buf[:, C:], ref_c, stock_c)
    worst = {}
    for name, (got, ref_cpu, ref_gpu) in outs.items():
        worst[name] = (float((got.cpu() - ref_cpu).abs().max()), float((ref_gpu.cpu() - ref_cpu).abs().max()))
    print(f"run-walk C={C} R={R} {flow_kind}: (max|ours - cpu oracle|, max|stock cuda - cpu oracle|):", worst)
    for name, (e_cpu, _e_stock) in worst.items():
        assert e_cpu <= 1e-5, (name, e_cpu)


@pytest.mark.parametrize("mode", ["pixel", "acF"])
def test_warp_backward_channels_last(mode):
    m = mb()
    torch.manual_seed(7)
    B, C, H, W, Ho, Wo = 2, 8, 9, 11, 7, 5
    feat = torch.randn(B, C, H, W, dtype=torch.float64)
    if mode == "pixel":
        grid = torch.rand(B, Ho, Wo, 2, dtype=torch.float64) * torch.tensor([W + 2.0, H + 2.0], dtype=torch.float64) - 1.5
    else:
        grid = torch.rand(B, Ho, Wo, 2, dtype=torch.float64) * 2.6 - 1.3
    go = torch.randn(B, C, Ho, Wo, dtype=torch.float64)
    f64, g64 = feat.clone().requires_grad_(), grid.clone().requires_grad_()
    if mode == "pixel":
        gn = torch.stack([2 * g64[..., 0] / (W - 1) - 1, 2 * g64[..., 1] / (H - 1) - 1], -1)
        F.grid_sample(f64, gn, align_corners=True).backward(go)
    else:
        F.grid_sample(f64, g64, align_corners=False).backward(go)
    f32 = feat.float().to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_()
    g32 = grid.float().to(DEV).requires_grad_()
    out = m.bilinear_sampler(f32, g32) if mode == "pixel" else m.grid_sample(f32, g32)
    out.backward(go.float().to(DEV))
    close(f32.grad, f64.grad, 2e-5)
    close(g32.grad, g64.grad, 5e-4, 1e-4)


def test_corr_channels_last_inputs_and_lookup_output(golden):
    m, c = mb(), golden("corr")
    q, k = cu(c["q_d"]), cu(c["k_s"])
    B, C, h, w = q.shape
    ref = m.CorrPyramid(q, k, C ** -0.5)
    pyr = m.CorrPyramid(q.contiguous(memory_format=torch.channels_last), k.contiguous(memory_format=torch.channels_last), C ** -0.5)
    # level-0 rows are the same cast; pooled rows may differ by summation order (one bf16 ulp)
    assert torch.equal(pyr.volume0[:, :h * w], ref.volume0[:, :h * w])
    rel_close(pyr.volume0, ref.volume0.float(), 2 ** -6)
    for lvl, kk in ((0, 1), (1, 2)):
        coords = cu(c[f"coords_k{kk}"])
        a = ref.block(lvl)(coords)
        b = ref.block(lvl)(coords, True)
        assert b.is_contiguous(memory_format=torch.channels_last)
        assert torch.equal(a, b.contiguous())
    # biases of the 1x1 heads added while packing (mrfa_corr_pack_bias) == packing the biased maps
    torch.manual_seed(31)
    bq, bk = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
    qcl, kcl = q.contiguous(memory_format=torch.channels_last), k.contiguous(memory_format=torch.channels_last)
    fused = m.CorrPyramid(qcl, kcl, C ** -0.5, bq, bk)
    plain = m.CorrPyramid((qcl + bq.view(1, -1, 1, 1)).contiguous(memory_format=torch.channels_last),
                          (kcl + bk.view(1, -1, 1, 1)).contiguous(memory_format=torch.channels_last), C ** -0.5)
    assert torch.equal(fused.volume0, plain.volume0) and torch.equal(fused.volume1, plain.volume1)
    only_k = m.CorrPyramid(qcl, kcl, C ** -0.5, None, bk)
    assert torch.equal(only_k.volume0[:, :h * w],
                       m.CorrPyramid(qcl, (kcl + bk.view(1, -1, 1, 1)).contiguous(memory_format=torch.channels_last), C ** -0.5).volume0[:, :h * w])
    with pytest.raises(Exception):
        m.CorrPyramid(q, k, C ** -0.5, bq, bk)                     # NCHW inputs: no fused bias


# ------------------------------------------------------------------ fused elementwise passes
@pytest.mark.parametrize("cl", [False, True])
def test_channel_affine_and_blend(cl):
    torch.manual_seed(11)
    B, C, H, W = 2, 8, 5, 7
    fmt = torch.channels_last if cl else torch.contiguous_format
    x = torch.randn(B, C, H, W, device=DEV).contiguous(memory_format=fmt)
    r = torch.randn(B, C, H, W, device=DEV).contiguous(memory_format=fmt)
    s, t = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV)
    sv, tv = s.view(1, C, 1, 1), t.view(1, C, 1, 1)
    close(torch.ops.mrfa.channel_affine(x, s, t, None, 1), torch.relu(x * sv + tv), 1e-6)
    close(torch.ops.mrfa.channel_affine(x, None, t, r, 0), x + tv + r, 1e-6)
    close(torch.ops.mrfa.channel_affine(x, s, None, None, 2), torch.sigmoid(x * sv), 1e-6)
    occ = torch.rand(B, 1, H, W, device=DEV)
    close(torch.ops.mrfa.occlusion_blend(x, r, occ), x * occ + r * (1 - occ), 1e-6)
    close(torch.ops.mrfa.occlusion_blend(x, None, occ), x * occ, 1e-6)
    odd = torch.randn(B, 3, H, W, device=DEV)                       # C % 4 != 0 -> NCHW path
    close(torch.ops.mrfa.occlusion_blend(odd, odd * 2, occ), odd * occ + odd * 2 * (1 - occ), 1e-6)


def test_fast_inference_blocks_match_plain():
    """BN folding / fused conv-bias-ReLU / fused blends are a re-association, not new numerics."""
    from mrfa_b200 import blocks
    import synthetic_inputs as syn
    torch.manual_seed(12)
    gen = syn.fill_state_dict_(blocks.OcclusionAwareGenerator(3, 16, 64, 3)).to(DEV).eval()
    hg = syn.fill_state_dict_(blocks.Hourglass(8, 5, 3, 32)).to(DEV).eval()
    x = torch.rand(2, 3, 32, 32, device=DEV)
    z = torch.randn(2, 5, 16, 16, device=DEV)
    with torch.no_grad():
        for cl in (False, True):
            if cl:
                gen.to(memory_format=torch.channels_last)
                hg.to(memory_format=torch.channels_last)
            feats_fast = gen.encode(x)
            occ = [torch.rand(2, 1, f.shape[2], f.shape[3], device=DEV) for f in feats_fast]
            out_fast = gen.decode(feats_fast, x, occ, [f * 0.5 for f in feats_fast], occ)
            hg_fast = hg(z)
            blocks.FAST_INFERENCE = False
            try:
                feats = gen.encode(x)
                out = gen.decode(feats, x, occ, [f * 0.5 for f in feats], occ)
                hg_plain = hg(z)
            finally:
                blocks.FAST_INFERENCE = True
            for a, b in zip(feats_fast, feats):
                close(a, b, 2e-5, 1e-5)
            close(out_fast, out, 2e-5)
            close(hg_fast, hg_plain, 2e-5, 1e-5)


@pytest.mark.parametrize("cl", [False, True])
def test_resize_bilinear(cl):
    torch.manual_seed(13)
    x = torch.randn(2, 8, 16, 16, device=DEV)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    for size in ((64, 64), (32, 32), (16, 16), (8, 8), (37, 21)):
        ref = F.interpolate(x.cpu().contiguous(), size=size, mode="bilinear", align_corners=True)
        out = torch.ops.mrfa.resize_bilinear(x, size[0], size[1], 0)
        assert out.is_contiguous(memory_format=torch.channels_last if cl else torch.contiguous_format)
        close(out, ref, 2e-6)
        close(torch.ops.mrfa.resize_bilinear(x, size[0], size[1], 1), torch.relu(ref), 2e-6)
        bias = torch.randn(8, device=DEV)                              # per-channel bias between interpolation and activation
        close(torch.ops.mrfa.resize_bilinear(x, size[0], size[1], 1, bias), torch.relu(ref + bias.cpu().view(1, -1, 1, 1)), 3e-6)
    flow = torch.randn(2, 2, 9, 9, device=DEV)                        # 2-channel maps keep their layout too
    if cl:
        flow = flow.contiguous(memory_format=torch.channels_last)
    out = torch.ops.mrfa.resize_bilinear(flow, 18, 18, 0)
    assert out.is_contiguous(memory_format=torch.channels_last if cl else torch.contiguous_format)
    close(out, F.interpolate(flow.cpu().contiguous(), size=(18, 18), mode="bilinear", align_corners=True), 2e-6)


@pytest.mark.parametrize("cl", [False, True])
def test_flow_update_matches_reference_ops(cl):
    """raft.py:256-262: flow + d_flow[:, 0:2], occlusion + d_flow[:, 2:3] and its sigmoid, one kernel vs the torch ops."""
    torch.manual_seed(23)
    B, H, W = 3, 9, 12
    flow = torch.randn(B, 2, H, W, device=DEV) * 4
    occ = torch.randn(B, 1, H, W, device=DEV)
    d4 = torch.randn(B, 4, H, W, device=DEV)
    if cl:
        flow, d4 = flow.contiguous(memory_format=torch.channels_last), d4.contiguous(memory_format=torch.channels_last)
    d = d4[:, :3]                                                    # the strided view RefineFlow returns
    fw, on, osig = torch.ops.mrfa.flow_update(flow, occ, d)
    assert fw.is_contiguous(memory_format=torch.channels_last if cl else torch.contiguous_format)
    assert torch.equal(fw, flow + d[:, 0:2]) and torch.equal(on, occ + d[:, 2:3])
    close(osig, torch.sigmoid(occ + d[:, 2:3]), 1e-6)
    with pytest.raises(Exception):
        torch.ops.mrfa.flow_update(flow, occ, d4[:, :2])


def test_resize_strip_matches_cat_of_resizes():
    """raft.py:304-306: the occlusion strip = cat of align_corners resizes along the width, written by one kernel per map."""
    torch.manual_seed(22)
    maps = [torch.rand(3, 1, r, r, device=DEV) for r in (2, 4, 8, 16, 32, 64, 16)]
    maps[2] = maps[2].contiguous(memory_format=torch.channels_last)
    ref = torch.cat([F.interpolate(m.cpu().contiguous(), size=(64, 64), mode="bilinear", align_corners=True) for m in maps], dim=3)
    out = torch.ops.mrfa.resize_strip(maps, 64, 64)
    assert out.shape == (3, 1, 64, 7 * 64)
    close(out, ref, 2e-6)
    two = [torch.rand(2, 2, 5, 7, device=DEV), torch.rand(2, 2, 9, 3, device=DEV)]
    ref = torch.cat([F.interpolate(m.cpu(), size=(10, 12), mode="bilinear", align_corners=True) for m in two], dim=3)
    close(torch.ops.mrfa.resize_strip(two, 10, 12), ref, 2e-6)
    with pytest.raises(Exception):
        torch.ops.mrfa.resize_strip(two, 10, 10)                       # Wo % 4 != 0


@pytest.mark.parametrize("cl", [False, True])
def test_flow_carry_matches_reference_chain(cl):
    """raft.py:276-295 (the per-level flow / occlusion hand-over) as one kernel vs the torch op chain on CPU."""
    torch.manual_seed(21)
    B, h = 2, 16
    init_flow = torch.randn(B, 2, h, h) * 3
    prior_occ = torch.randn(B, 1, h, h)
    up = lambda t, R: F.interpolate(t, size=(R, R), mode="bilinear", align_corners=True)
    d_f_pre = d_occ_pre = g_pre = go_pre = None
    for i, R in enumerate((2, 4, 8, 16, 32)):
        d4 = torch.randn(B, 4, R, R)                                   # merged conv2|convo2 output: [fx, fy, occ, 0]
        scale = 2 ** (3 - i) / 2.0
        d_f = up(d4[:, 0:2], 2 * R) * 2
        flow = d_f + up(init_flow, 2 * R) / scale
        d_o = up(d4[:, 2:3], 2 * R)
        occ = d_o + up(prior_occ, 2 * R)
        if i == 0:
            d_f_pre, d_occ_pre = d_f, d_o
        else:
            up_f, up_o = up(d_f_pre, 2 * R) * 2, up(d_occ_pre, 2 * R)
            flow, occ = flow + up_f, occ + up_o
            d_f_pre, d_occ_pre = d_f + up_f, d_o + up_o
        g4 = d4.to(DEV)
        if cl:
            g4 = g4.contiguous(memory_format=torch.channels_last)
        got = torch.ops.mrfa.flow_carry(g4[:, :3], init_flow.to(DEV), prior_occ.to(DEV), g_pre, go_pre, scale, cl)
        assert got[0].shape == (B, 2, 2 * R, 2 * R) and got[1].shape == (B, 1, 2 * R, 2 * R)
        assert got[0].is_contiguous(memory_format=torch.channels_last if cl else torch.contiguous_format)
        for a, b in zip(got, (flow, occ, d_f_pre, d_occ_pre)):
            close(a, b, 1e-5, 1e-5)
        g_pre, go_pre = got[2], got[3]
    with pytest.raises(Exception):
        torch.ops.mrfa.flow_carry(torch.zeros(1, 3, 4, 4), init_flow.to(DEV), prior_occ.to(DEV), None, None, 1.0, False)


def test_antialias_down_matches_reference_order():
    from mrfa_b200 import blocks
    torch.manual_seed(14)
    for scale, size in ((0.25, 64), (0.25, 256), (0.25, 100), (0.5, 32)):
        aa = blocks.AntiAliasInterpolation2d(3, scale).to(DEV).eval()
        x = torch.rand(2, 3, size, size, device=DEV)
        with torch.no_grad():
            fast = aa(x)
            blocks.FAST_INFERENCE = False
            try:
                plain = aa(x)
            finally:
                blocks.FAST_INFERENCE = True
        assert fast.shape == plain.shape
        close(fast, plain, 2e-6)


def test_avg_pool2x2_nhwc():
    torch.manual_seed(15)
    x = torch.randn(3, 8, 10, 6, device=DEV).contiguous(memory_format=torch.channels_last)
    out = torch.ops.mrfa.avg_pool2x2_nhwc(x)
    assert out.is_contiguous(memory_format=torch.channels_last)
    close(out, F.avg_pool2d(x.cpu().contiguous(), (2, 2)), 1e-6)


def test_equivariance_transform_matches_reference_golden(golden):
    """model.py:26-77 Transform (reflection-padded random affine + TPS warp, its key-point warp and Jacobian) and
    util.py TPS mode 'random', against vectors produced by the reference (tests/golden/make_golden.py equivariance)."""
    m = mb()
    e = golden("equivariance")
    params = {"sigma_affine": 0.05, "sigma_tps": 0.005, "points_tps": 5}
    frame = torch.from_numpy(e["frame"]).to(DEV)
    kp = torch.from_numpy(e["kp"]).to(DEV)
    torch.manual_seed(7)
    tr = m.Transform(2, **params)
    np.testing.assert_array_equal(tr.theta.numpy(), e["theta"])                     # same RNG draws as the reference
    np.testing.assert_array_equal(tr.control_params.numpy(), e["control_params"])
    np.testing.assert_array_equal(tr.control_points.numpy(), e["control_points"])
    close(tr.transform_frame(frame), e["transform_frame"], 1e-5)
    close(tr.warp_coordinates(kp), e["warp_kp"], 1e-6)
    kpg = kp.clone().requires_grad_(True)
    close(tr.jacobian(kpg), e["jacobian_kp"], 1e-5)
    torch.manual_seed(7)
    aff = m.Transform(2, sigma_affine=0.05)
    close(aff.transform_frame(frame), e["affine_transform_frame"], 1e-5)
    torch.manual_seed(7)
    tps = m.TPS("random", 2, **params)
    close(tps.transform_frame(frame), e["tps_random_grid"], 2e-6)
    close(tps.warp_coordinates(kp), e["tps_random_warp_kp"], 1e-6)
    # the oracle agrees with the kernel at a size the fixtures do not cover, and the warp is differentiable
    cp, cw = e["control_points"].reshape(-1, 2), e["control_params"].reshape(2, -1)
    grid = torch.ops.mrfa.random_warp_grid(cu(e["theta"]), cu(cp), cu(cw), 96, 128, 0)
    close(grid, O.random_warp_grid(e["theta"], cp, cw, 96, 128, "l1"), 1e-6)
    # (a smooth frame: on white noise a 1-ulp grid difference times W/2 pixels already reaches 1e-5)
    big = F.interpolate(torch.rand(2, 3, 12, 16, device=DEV), size=(96, 128), mode="bilinear").requires_grad_(True)
    out = tr.transform_frame(big)
    close(out, O.transform_frame(big.detach().cpu().numpy(), e["theta"], cp, cw), 1e-5)
    out.sum().backward()
    assert big.grad is not None and torch.isfinite(big.grad).all() and float(big.grad.abs().sum()) > 0
    with pytest.raises(RuntimeError):
        tr.transform_frame(torch.zeros(2, 3, 8, 8))


@pytest.mark.parametrize("cin,cout,layout", [(2, 128, "nhwc"), (2, 128, "nchw"), (3, 64, "nchw"), (3, 64, "nhwc")])
def test_conv7x7_small_matches_fp32_convolution(cin, cout, layout):
    """tcgen05 TF32 implicit-GEMM 7x7 convolution (raft.py:56,63 convf1, generator.py:13 first) vs F.conv2d in fp32 on
    the CPU.  TF32 operands (10-bit mantissa, round-to-nearest) with fp32 accumulation: 5e-3 relative."""
    m = mb()
    torch.manual_seed(31 + cin)
    for (B, H, W) in ((2, 5, 128), (1, 37, 256), (3, 130, 128)):            # ragged heights, 1 and 2 tiles per row
        x = torch.randn(B, cin, H, W)
        w = torch.randn(cout, cin, 7, 7) * 0.1
        b = torch.randn(cout)
        ref = F.conv2d(x, w, b, padding=3)
        xg = x.to(DEV)
        if layout == "nhwc":
            xg = xg.contiguous(memory_format=torch.channels_last)
        wp = m.ops.conv7x7_small_pack(w.to(DEV))
        out = torch.ops.mrfa.conv7x7_small(xg, wp, b.to(DEV), False)
        assert out.shape == ref.shape and out.is_contiguous(memory_format=torch.channels_last)
        rel_close(out, ref, 5e-3)
        rel_close(torch.ops.mrfa.conv7x7_small(xg, wp, b.to(DEV), True), torch.relu(ref), 5e-3)
        rel_close(torch.ops.mrfa.conv7x7_small(xg, wp, None, False), F.conv2d(x, w, None, padding=3), 5e-3)
    # a strided (channel-sliced) input, as the refinement loop passes d_flow[:, 0:2]
    big = torch.randn(2, 4, 16, 128, device=DEV).contiguous(memory_format=torch.channels_last)
    if cin == 2:
        rel_close(torch.ops.mrfa.conv7x7_small(big[:, 0:2], wp, None, False), F.conv2d(big[:, 0:2].cpu().contiguous(), w, None, padding=3), 5e-3)
    with pytest.raises(RuntimeError):
        torch.ops.mrfa.conv7x7_small(torch.randn(1, cin, 8, 64, device=DEV), wp, None, False)      # W % 128 != 0


def test_small_conv_blocks_match_cudnn_path():
    """BasicMotionEncoder / SameBlock2d routed through mrfa::conv7x7_small agree with the cuDNN fast path."""
    from mrfa_b200 import blocks
    torch.manual_seed(33)
    blk = blocks.SameBlock2d(3, 64, kernel_size=7, padding=3).to(DEV).eval()
    conv = torch.nn.Conv2d(2, 128, 7, padding=3).to(DEV).eval()
    x3 = torch.rand(2, 3, 128, 128, device=DEV)
    x2 = torch.randn(2, 2, 128, 128, device=DEV)
    with torch.no_grad():
        a3, a2 = blk(x3), blocks.conv_relu(conv, x2)
        blocks.SMALL_CONV = False
        try:
            b3, b2 = blk(x3), blocks.conv_relu(conv, x2)
        finally:
            blocks.SMALL_CONV = True
    rel_close(a3, b3, 5e-3)
    rel_close(a2, b2, 5e-3)


def test_final_blend_s2d_matches_op_chain():
    """pixel shuffle + bias + sigmoid + the last occlusion blend (generator.py:61-63 behind the space-to-depth final
    convolution) as one kernel against the op chain it replaces; non-square, several block sizes."""
    mb()
    torch.manual_seed(21)
    for r, C, (Hb, Wb) in ((4, 3, (16, 20)), (2, 3, (9, 7)), (1, 2, (5, 6))):
        B = 3
        conv = torch.randn(B, C * r * r, Hb, Wb, device=DEV).contiguous(memory_format=torch.channels_last)
        bias = torch.randn(C, device=DEV)
        a = torch.rand(B, C, Hb * r, Wb * r, device=DEV)
        occ = torch.rand(B, 1, Hb * r, Wb * r, device=DEV)
        exp = a * occ + torch.sigmoid(F.pixel_shuffle(conv, r) + bias.view(1, -1, 1, 1)) * (1 - occ)
        close(torch.ops.mrfa.final_blend_s2d(conv, bias, a, occ, r), exp, 1e-6)


def test_blend_subpixel_space_to_depth_output_and_final_conv():
    """The r x r space-to-depth output of the last blend is a pure re-layout, and the generator's final 7x7
    convolution evaluated on it as a 3x3 convolution (blocks._final_s2d) equals the direct one."""
    from mrfa_b200 import blocks
    import synthetic_inputs as syn
    torch.manual_seed(41)
    N, C, H, W = 2, 8, 6, 10
    a = torch.randn(N, C, 2 * H, 2 * W, device=DEV).contiguous(memory_format=torch.channels_last)
    b2 = torch.randn(N, 4 * C, H + 1, W + 1, device=DEV).contiguous(memory_format=torch.channels_last)
    occ = torch.rand(N, 1, 2 * H, 2 * W, device=DEV)
    plain = torch.ops.mrfa.occlusion_blend_subpixel(a, b2, occ, 1)
    for r in (2, 4):
        got = torch.ops.mrfa.occlusion_blend_subpixel(a, b2, occ, r)
        assert got.shape == (N, r * r * C, 2 * H // r, 2 * W // r) and got.is_contiguous(memory_format=torch.channels_last)
        # channel index (iy*r + ix)*C + c of block (Y/r, X/r)  <->  pixel (Y, X), channel c
        want = plain.reshape(N, C, 2 * H // r, r, 2 * W // r, r).permute(0, 3, 5, 1, 2, 4).reshape(N, r * r * C, 2 * H // r, 2 * W // r)
        assert torch.equal(got, want)
    with pytest.raises(RuntimeError):
        torch.ops.mrfa.occlusion_blend_subpixel(a, b2, occ, 8)             # 8 does not divide 12
    gen = syn.fill_state_dict_(blocks.OcclusionAwareGenerator(3, 16, 64, 3)).to(DEV).eval()
    x = torch.randn(2, 16, 24, 40, device=DEV)
    with torch.no_grad():
        ref = gen.final(x)
        xs = x.reshape(2, 16, 6, 4, 10, 4).permute(0, 3, 5, 1, 2, 4).reshape(2, 256, 6, 10).contiguous(memory_format=torch.channels_last)
        close(gen._final_s2d(xs), ref, 1e-5, 1e-5)


@pytest.mark.parametrize("C,Cs,skip_cl", [(8, 12, True), (8, 5, False), (6, 13, True), (16, 0, True)])
def test_subpixel_shuffle_cat(C, Cs, skip_cl):
    """Hourglass decoder step (util.py:239-263) as shuffle + cat in one pass; vector and scalar paths."""
    torch.manual_seed(43)
    N, H, W = 2, 5, 7
    b2 = torch.randn(N, 4 * C, H + 1, W + 1, device=DEV).contiguous(memory_format=torch.channels_last)
    skip = torch.randn(N, Cs, 2 * H, 2 * W, device=DEV)
    if skip_cl:
        skip = skip.contiguous(memory_format=torch.channels_last)
    Y, X = torch.meshgrid(torch.arange(2 * H, device=DEV), torch.arange(2 * W, device=DEV), indexing="ij")
    pa, pb = Y & 1, X & 1
    b2v = b2.reshape(N, 4, C, H + 1, W + 1)                      # channel (2a+b)*C + c
    shuf = b2v[:, (2 * pa + pb), :, (Y >> 1) + pa, (X >> 1) + pb].permute(2, 3, 0, 1)     # (2H,2W,N,C) -> (N,C,2H,2W)
    want = torch.cat([shuf, skip], dim=1)
    got = torch.ops.mrfa.subpixel_shuffle_cat(b2, skip)
    assert got.is_contiguous(memory_format=torch.channels_last) and torch.equal(got, want)


def test_cat_slice_outputs_match_plain_ops():
    """dual_warp_cat + occlusion_blend_subpixel_into fill the decoder's cat([y, warp_c]) buffer in place."""
    torch.manual_seed(47)
    N, C, H, W = 2, 8, 6, 10
    feat = torch.randn(N, C, 2 * H, 2 * W, device=DEV).contiguous(memory_format=torch.channels_last)
    flow = torch.randn(N, 2, 2 * H, 2 * W, device=DEV) * 2
    prior = (torch.rand(N, 2 * H, 2 * W, 2, device=DEV) * 2.2 - 1.1)
    wr, wc = torch.ops.mrfa.dual_warp(feat, flow, prior)
    wr2, buf = torch.ops.mrfa.dual_warp_cat(feat, flow, prior)
    assert buf.shape == (N, 2 * C, 2 * H, 2 * W) and buf.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(wr, wr2) and torch.equal(buf[:, C:], wc)
    b2 = torch.randn(N, 4 * C, H + 1, W + 1, device=DEV).contiguous(memory_format=torch.channels_last)
    occ = torch.rand(N, 1, 2 * H, 2 * W, device=DEV)
    want = torch.ops.mrfa.occlusion_blend_subpixel(wr, b2, occ, 1)
    torch.ops.mrfa.occlusion_blend_subpixel_into(wr, b2, occ, buf)
    assert torch.equal(buf, torch.cat([want, wc], dim=1))
    with pytest.raises(RuntimeError):
        torch.ops.mrfa.dual_warp_cat(feat.contiguous(), flow, prior)                        # NCHW input
    with pytest.raises(RuntimeError):
        torch.ops.mrfa.occlusion_blend_subpixel_into(wr, b2, occ, buf[:, :C].contiguous())   # not channels_last / too narrow


def test_cat2_matches_torch_cat():
    torch.manual_seed(49)
    a = torch.randn(2, 96, 9, 7, device=DEV).contiguous(memory_format=torch.channels_last)
    b = torch.randn(2, 64, 9, 7, device=DEV).contiguous(memory_format=torch.channels_last)
    y = torch.ops.mrfa.cat2(a, b)
    assert y.is_contiguous(memory_format=torch.channels_last) and torch.equal(y, torch.cat([a, b], 1))
    with pytest.raises(RuntimeError):
        torch.ops.mrfa.cat2(a.contiguous(), b)


def test_avg_pool2x2_nhwc_forward_backward():
    torch.manual_seed(51)
    x = torch.randn(2, 8, 6, 10, device=DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    y = torch.ops.mrfa.avg_pool2x2_nhwc(x)
    yr = F.avg_pool2d(xr, (2, 2))
    close(y, yr, 1e-6)
    g = torch.randn_like(yr)
    y.backward(g)
    yr.backward(g)
    close(x.grad, xr.grad, 1e-7)
    torch.library.opcheck(torch.ops.mrfa.avg_pool2x2_nhwc, (x.detach().requires_grad_(True),),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
