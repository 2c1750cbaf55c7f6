"""Pin the CPU oracle (oracle/np_ops.py, oracle/torch_path.py) against golden vectors
produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch
import yaml

import synthetic_inputs as syn
from oracle import np_ops as O
from oracle import torch_path as TP


def close(a, b, atol=1e-5, rtol=0.0):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), atol=atol, rtol=rtol)


# ---------------------------------------------------------------- grids: bit-exact
def test_grids_bit_exact(golden):
    g = golden("grids")
    for (h, w) in ((2, 3), (16, 16), (64, 64), (5, 7), (128, 128)):
        assert np.array_equal(O.make_coordinate_grid(h, w), g[f"mcg_{h}x{w}"])
        assert np.array_equal(TP.make_coordinate_grid((h, w)).numpy(), g[f"mcg_{h}x{w}"])
    for (b, h, w) in ((1, 2, 3), (2, 8, 8), (1, 64, 64)):
        assert np.array_equal(O.coords_grid(b, h, w), g[f"cg_{b}_{h}x{w}"])
        assert np.array_equal(TP.coords_grid(b, h, w).numpy(), g[f"cg_{b}_{h}x{w}"])


def test_known_answers():
    # SURVEY.md section 4 known-answer facts
    assert O.coords_grid(1, 2, 3)[0, 0].tolist() == [[0, 1, 2], [0, 1, 2]]
    assert O.coords_grid(1, 2, 3)[0, 1].tolist() == [[0, 0, 0], [1, 1, 1]]
    m = O.make_coordinate_grid(2, 3)
    assert m[..., 0].tolist() == [[-1, 0, 1], [-1, 0, 1]] and m[..., 1].tolist() == [[-1, -1, -1], [1, 1, 1]]
    kp = np.array([[[-1.0, 1.0]]], np.float32)
    assert O.kp2gaussian(kp, 3, 3, 0.01)[0, 0, 2, 0] == 1.0


def test_kp2gaussian(golden):
    g = golden("grids")
    close(O.kp2gaussian(g["kp"], 16, 16, 0.1), g["kp2g_16_0.1"], 2e-7)
    close(O.kp2gaussian(g["kp"], 16, 16, 0.01), g["kp2g_16_0.01"], 2e-7)
    close(O.kp2gaussian(g["kp"], 12, 20, 0.01), g["kp2g_12x20_0.01"], 2e-7)
    close(TP.kp2gaussian(torch.from_numpy(g["kp"]), (16, 16), 0.1).numpy(), g["kp2g_16_0.1"], 2e-7)


# ---------------------------------------------------------------- samplers
def test_samplers(golden):
    s = golden("samplers")
    close(O.bilinear_sampler(s["img"], s["pix_coords"]), s["bilinear_sampler"], 1e-5)
    _, m = O.bilinear_sampler(s["img"], s["pix_coords"], mask=True)
    assert np.array_equal(m, s["bilinear_sampler_mask"])
    close(O.grid_sample(s["img"], s["norm_grid"], False), s["grid_sample_acF"], 1e-5)
    close(O.grid_sample(s["img"], s["norm_grid"], True), s["grid_sample_acT"], 1e-5)
    close(O.grid_sample(s["img"], s["norm_grid"] * np.float32(1.7), False, "reflection"),
          s["grid_sample_reflect_acF"], 1e-5)
    close(O.batch_bilinear_sampler(s["bimg"], s["bco"], h=2, w=2, mini_batch=1), s["batch_bilinear_mb1"], 1e-5)
    mb2 = O.batch_bilinear_sampler(s["bimg"], s["bco"], h=2, w=2, mini_batch=2)
    assert mb2.shape == s["batch_bilinear_mb2"].shape == (8, 1, 3, 3)     # remainder chunk dropped
    close(mb2, s["batch_bilinear_mb2"], 1e-5)
    close(O.interpolate_bilinear_ac(np.moveaxis(s["prior"], -1, 1), 16, 16), s["prior_resized"], 1e-5)
    close(O.deform_input(s["feat"], s["prior"]), s["coarse_warp"], 1e-5)


# ---------------------------------------------------------------- correlation
def test_corr_volume_and_lookup(golden):
    c = golden("corr")
    vol = O.corr_volume(c["q_d"], c["k_s"], 64 ** -0.5)
    close(vol, c["volume"], 1e-5)
    B, h = 2, 16
    for k in (1, 2, 4):
        rows = O.corr_pyramid_rows(vol, h, h, k)                          # (B, R*R, N)
        l0 = rows.reshape(-1, 1, h, h)
        l1 = O.avg_pool2d(l0, 2)
        close(l1, c[f"level1_k{k}"], 1e-5)
        close(O.corr_lookup([l0, l1], c[f"coords_k{k}"]), c[f"lookup_k{k}"], 1e-5)
        t = TP.corr_lookup(torch.from_numpy(l0), torch.from_numpy(c[f"coords_k{k}"]))
        close(t.numpy(), c[f"lookup_k{k}"], 1e-5)


def test_lookup_channel_order(golden):
    # channel k = lvl*49 + a*7 + b samples at x+(a-3), y+(b-3): first window index moves x
    img = np.zeros((1, 1, 16, 16), np.float32)
    img[0, 0, 8, 5] = 1.0                                     # y=8, x=5
    coords = np.array([8.0, 8.0], np.float32).reshape(1, 2, 1, 1)          # x=8,y=8
    out = O.corr_lookup([img, O.avg_pool2d(img, 2)], coords)
    assert out[0, 0 * 7 + 3, 0, 0] == 1.0                     # a=0 -> x-3 = 5, b=3 -> y = 8
    assert out[0, 3 * 7 + 0, 0, 0] == 0.0


# ---------------------------------------------------------------- prior dense motion
@pytest.fixture(scope="module")
def prior_inputs():
    src, _ = syn.frame_pairs(2, 64, seed=1)
    kp_s, kp_d = syn.keypoints(2, 10, seed=1)
    return src, kp_s, kp_d, syn.bg_affine(2, seed=1)


def _cfg():
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    return yaml.safe_load(open(os.path.join(here, "golden", "vox1.yaml")))


def test_sparse_motion_and_deformed(golden, prior_inputs):
    d = golden("prior_motion")
    _, kp_s, kp_d, bg = prior_inputs
    n = lambda t: t.numpy()
    close(O.heatmap_representation(n(kp_d["kp"]), n(kp_s["kp"]), 16, 16)[:, :, None], d["heatmap"], 5e-7)
    sm = O.sparse_motions(n(kp_d["kp"]), n(kp_s["kp"]), 16, 16, n(kp_d["jacobian"]), n(kp_s["jacobian"]), n(bg))
    close(sm, d["sparse_motions_jac_bg"], 1e-5)
    close(O.sparse_motions(n(kp_d["kp"]), n(kp_s["kp"]), 16, 16), d["sparse_motions_plain"], 1e-6)
    close(O.deformed_source(d["source_small"], d["sparse_motions_jac_bg"], False), d["deformed"], 1e-5)


def test_tps(golden):
    d = golden("prior_motion")
    bg = syn.bg_affine(2, seed=1).numpy()
    theta, cp, cw = O.tps_params(d["tps_kp_d"].reshape(2, -1, 5, 2), d["tps_kp_s"].reshape(2, -1, 5, 2))
    # the reference inverts an 8x8 in fp32; allow its round-off
    close(theta, d["tps_theta"], 2e-3, 2e-3)
    close(O.tps_transformations(d["tps_kp_d"], d["tps_kp_s"], 16, 16, bg), d["tps_transformations_bg"], 5e-3)
    close(O.tps_transformations(d["tps_kp_d"], d["tps_kp_s"], 16, 16), d["tps_transformations"], 5e-3)


def test_dense_motion_forward(golden, prior_inputs):
    d = golden("prior_motion")
    src, kp_s, kp_d, bg = prior_inputs
    cfg = _cfg()
    net = syn.fill_state_dict_(TP.DenseMotionOracle(**dict(cfg["dense_motion"], block_expansion=16,
                                                           max_features=64, num_blocks=3))).eval()
    with torch.no_grad():
        out = net(src, kp_d, kp_s, bg_param=bg)
        for k in ("sparse_deformed", "logit_mask", "mask", "deformation", "occlusion"):
            close(out[k].numpy(), d["fwd_" + k], 2e-5)
        out = net(src, {"kp": kp_d["kp"]}, {"kp": kp_s["kp"]})
        close(out["deformation"].numpy(), d["fwd_plain_deformation"], 2e-5)
        close(out["occlusion"].numpy(), d["fwd_plain_occlusion"], 2e-5)


def test_tps_dense_motion_forward(golden):
    d = golden("prior_motion")
    cfg = _cfg()
    src, _ = syn.frame_pairs(2, 64, seed=1)
    net = syn.fill_state_dict_(TP.TPSDenseMotionOracle(**dict(cfg["tpsm_dense_motion"], block_expansion=16,
                                                              max_features=64, num_blocks=3))).eval()
    kp_s, kp_d = {"kp": torch.from_numpy(d["tps_kp_s"])}, {"kp": torch.from_numpy(d["tps_kp_d"])}
    with torch.no_grad():
        out = net(src, kp_d, kp_s, bg_param=syn.bg_affine(2, seed=1))
    for k in ("deformed_source", "contribution_maps", "deformation", "occlusion"):
        close(out[k].numpy(), d["tps_fwd_" + k], 5e-5)
    assert sorted(net.state_dict().keys()) == list(golden("raft_flow")["tps_state_dict_keys"])


# ---------------------------------------------------------------- RaftFlow end to end
def test_raft_flow_forward(golden, prior_inputs):
    r = golden("raft_flow")
    d = golden("prior_motion")
    src, kp_s, kp_d, bg = prior_inputs
    cfg = _cfg()
    rfc = dict(cfg["raft_flow"], size=64)
    rfc["driving_encoder"] = dict(rfc["driving_encoder"], block_expansion=8, max_features=32, num_blocks=3)
    rfc["source_encoder"] = dict(rfc["source_encoder"], block_expansion=8, max_features=32, num_blocks=3)
    dense = {"deformation": torch.from_numpy(d["fwd_deformation"]), "occlusion": torch.from_numpy(d["fwd_occlusion"])}
    small = torch.from_numpy(d["source_small"])
    with torch.no_grad():
        net = syn.fill_state_dict_(TP.RaftFlowOracle(**rfc)).eval()
        assert sorted(net.state_dict().keys()) == list(r["state_dict_keys"])
        out, warp_img, occ = net(kp_s["kp"], kp_d["kp"], dense, img=small, img_full=src)
        close(out.numpy(), r["out"], 5e-5)
        close(warp_img.numpy(), r["warp_img"], 5e-5)
        close(occ.numpy(), r["occlusion"], 5e-5)
        netp = syn.fill_state_dict_(TP.RaftFlowOracle(**dict(rfc, prior_only=True))).eval()
        out, warp_img, occ = netp(kp_s["kp"], kp_d["kp"], dense, img=small, img_full=src)
        close(out.numpy(), r["prior_only_out"], 5e-5)
        close(warp_img.numpy(), r["prior_only_warp_img"], 5e-5)
        close(occ.numpy(), r["prior_only_occlusion"], 5e-5)


def test_equivariance_warps_match_reference(golden):
    """model.py:26-77 Transform and util.py TPS mode 'random' (SURVEY.md 8(f) N4)."""
    e = golden("equivariance")
    cp = e["control_points"].reshape(-1, 2)
    par = e["control_params"].reshape(2, -1)
    out = O.transform_frame(e["frame"], e["theta"], cp, par)
    np.testing.assert_allclose(out, e["transform_frame"], atol=1e-5)
    np.testing.assert_allclose(O.transform_frame(e["frame"], e["affine_theta"], cp, None), e["affine_transform_frame"], atol=1e-5)
    np.testing.assert_allclose(O.random_warp_coordinates(e["theta"], cp, par, e["kp"], "l1"), e["warp_kp"], atol=1e-6)
    h, w = e["frame"].shape[2:]
    g = O.random_warp_grid(e["tps_random_theta"], cp, e["tps_random_params"].reshape(2, -1), h, w, "l2sq")
    np.testing.assert_allclose(g, e["tps_random_grid"], atol=2e-6)
    np.testing.assert_allclose(O.random_warp_coordinates(e["tps_random_theta"], cp, e["tps_random_params"].reshape(2, -1),
                                                         e["kp"], "l2sq"), e["tps_random_warp_kp"], atol=1e-6)
