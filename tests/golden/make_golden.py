"""Generate golden vectors by running the UNMODIFIED MRFA reference on CPU.

Run in the build container only (needs the reference checkout, default /root/reference or
$MRFA_REF):   python tests/golden/make_golden.py
The GPU box has no reference; it only reads the committed ``tests/golden/*.npz``.

One shim: ``timm`` is not installed and modules/raft.py:5 imports
``timm.models.layers.weight_init.trunc_normal_`` -- mapped to torch.nn.init.trunc_normal_
(the same function upstream).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("MRFA_REF", "/root/reference")


def import_reference():
    for name in ("timm", "timm.models", "timm.models.layers", "timm.models.layers.weight_init"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["timm.models.layers.weight_init"].trunc_normal_ = torch.nn.init.trunc_normal_
    sys.path.insert(0, REF)
    import modules.dense_motion as dm
    import modules.raft as raft
    import modules.util as util
    return util, raft, dm


def npy(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def main():
    import synthetic_inputs as syn
    util, raft, dm = import_reference()
    torch.set_grad_enabled(False)
    torch.manual_seed(0)
    T = syn.tensor
    FT = "torch.FloatTensor"

    # ------------------------------------------------------------------ primitives
    g = {}
    for (h, w) in ((2, 3), (16, 16), (64, 64), (5, 7), (128, 128)):
        g[f"mcg_{h}x{w}"] = npy(util.make_coordinate_grid((h, w), FT))
    for (b, h, w) in ((1, 2, 3), (2, 8, 8), (1, 64, 64)):
        g[f"cg_{b}_{h}x{w}"] = npy(util.coords_grid(b, h, w, "cpu"))
    kp = T("g.kp", (2, 10, 2), "uniform", 1.6, -0.8)
    g["kp"] = npy(kp)
    g["kp2g_16_0.1"] = npy(util.kp2gaussian(kp, (16, 16), 0.1))
    g["kp2g_16_0.01"] = npy(util.kp2gaussian(kp, (16, 16), 0.01))
    g["kp2g_12x20_0.01"] = npy(util.kp2gaussian(kp, (12, 20), 0.01))
    np.savez_compressed(os.path.join(HERE, "grids.npz"), **g)

    # ------------------------------------------------------------------ samplers
    s = {}
    img = T("s.img", (2, 5, 9, 11))
    # pixel coords with ~15% outside the image (zero-padding path), some exactly on pixels
    pc = T("s.pc", (2, 6, 7, 2), "uniform", 1.0, 0.0)
    pc[..., 0] = pc[..., 0] * 14 - 2
    pc[..., 1] = pc[..., 1] * 12 - 2
    pc[0, 0, 0] = torch.tensor([3.0, 4.0])
    pc[0, 0, 1] = torch.tensor([10.0, 8.0])
    pc[0, 0, 2] = torch.tensor([-1.0, 0.0])
    s["img"], s["pix_coords"] = npy(img), npy(pc)
    s["bilinear_sampler"] = npy(util.bilinear_sampler(img, pc))
    out_m, m = util.bilinear_sampler(img, pc, mask=True)
    s["bilinear_sampler_mask"] = npy(m)
    ng = T("s.ng", (2, 6, 7, 2), "uniform", 2.6, -1.3)
    s["norm_grid"] = npy(ng)
    s["grid_sample_acF"] = npy(F.grid_sample(img, ng))
    s["grid_sample_acT"] = npy(F.grid_sample(img, ng, align_corners=True))
    s["grid_sample_reflect_acF"] = npy(F.grid_sample(img, ng * 1.7, padding_mode="reflection"))
    # batch_bilinear_sampler: batch 3, h=w=2, mini_batch 1 and 2 (remainder dropped)
    bimg = T("s.bimg", (3 * 4, 1, 6, 6))
    bco = T("s.bco", (3 * 4, 3, 3, 2), "uniform", 7.0, -1.0)
    s["bimg"], s["bco"] = npy(bimg), npy(bco)
    s["batch_bilinear_mb1"] = npy(util.batch_bilinear_sampler(bimg, bco, h=2, w=2, mini_batch=1))
    s["batch_bilinear_mb2"] = npy(util.batch_bilinear_sampler(bimg, bco, h=2, w=2, mini_batch=2))
    # prior-grid warp with resize (the deform_input-equivalent lines raft.py:265-271)
    feat = T("s.feat", (2, 4, 16, 16))
    prior = T("s.prior", (2, 8, 8, 2), "uniform", 2.2, -1.1)
    pr = F.interpolate(prior.permute(0, 3, 1, 2), size=(16, 16), mode="bilinear", align_corners=True)
    s["feat"], s["prior"] = npy(feat), npy(prior)
    s["prior_resized"] = npy(pr)
    s["coarse_warp"] = npy(F.grid_sample(feat, pr.permute(0, 2, 3, 1)))
    np.savez_compressed(os.path.join(HERE, "samplers.npz"), **s)

    # ------------------------------------------------------------------ correlation + lookup
    c = {}
    B, C, h = 2, 64, 16
    q_d, k_s = T("c.q", (B, C, h, h)), T("c.k", (B, C, h, h))
    f_s = raft.rearrange(k_s, "b c h w -> b (h w) c", h=h, w=h)
    f_d = raft.rearrange(q_d, "b c h w -> b (h w) c", h=h, w=h)
    vol = torch.einsum("bic,bjc->bij", f_d, f_s) * (C ** -0.5)
    c["q_d"], c["k_s"], c["volume"] = npy(q_d), npy(k_s), npy(vol)
    vol_r = raft.rearrange(vol, "b (h w) n -> (b n) h w", h=h, w=h).unsqueeze(1)
    for k in (1, 2, 4):
        cv = F.avg_pool2d(vol_r, k, stride=k) if k > 1 else vol_r
        R = h // k
        cv = raft.rearrange(cv, "(b n) c h w -> (b h w) c n", n=h * h)
        cv = raft.rearrange(cv, "b c (p q) -> b c p q", p=h, q=h)
        blk = raft.CorrBlock(cv)
        coords = raft.coords_grid(B, R, R, "cpu") * k + T(f"c.co{k}", (B, 2, R, R), "normal", 2.5)
        c[f"coords_k{k}"] = npy(coords)
        c[f"lookup_k{k}"] = npy(blk(coords))
        c[f"level1_k{k}"] = npy(blk.corr_pyramid[1])
    np.savez_compressed(os.path.join(HERE, "corr.npz"), **c)

    # ------------------------------------------------------------------ prior dense motion
    d = {}
    cfg = yaml.safe_load(open(os.path.join(REF, "config", "vox1.yaml")))
    dmc = dict(cfg["dense_motion"], block_expansion=16, max_features=64, num_blocks=3)
    net = syn.fill_state_dict_(dm.DenseMotionNetwork(**dmc)).eval()
    src_img, _ = syn.frame_pairs(2, 64, seed=1)
    kp_s, kp_d = syn.keypoints(2, 10, seed=1)
    bg = syn.bg_affine(2, seed=1)
    small = net.down(src_img)
    d["source_small"] = npy(small)
    d["heatmap"] = npy(net.create_heatmap_representations(small, kp_d, kp_s))
    d["sparse_motions_jac_bg"] = npy(net.create_sparse_motions(small, kp_d, kp_s, bg_param=bg))
    nj_s, nj_d = {"kp": kp_s["kp"]}, {"kp": kp_d["kp"]}
    d["sparse_motions_plain"] = npy(net.create_sparse_motions(small, nj_d, nj_s, bg_param=None))
    sm = net.create_sparse_motions(small, kp_d, kp_s, bg_param=bg)
    d["deformed"] = npy(net.create_deformed_source_image(small, sm))
    out = net(src_img, kp_d, kp_s, bg_param=bg)
    for k in ("sparse_deformed", "logit_mask", "mask", "deformation", "occlusion"):
        d["fwd_" + k] = npy(out[k])
    out = net(src_img, nj_d, nj_s, bg_param=None)
    d["fwd_plain_deformation"] = npy(out["deformation"])
    d["fwd_plain_occlusion"] = npy(out["occlusion"])
    # TPS prior
    tc = dict(cfg["tpsm_dense_motion"], block_expansion=16, max_features=64, num_blocks=3)
    tnet = syn.fill_state_dict_(dm.TPSDenseMotionNetwork(**tc)).eval()
    tk_s, tk_d = syn.keypoints(2, 50, seed=2, jacobian=False)
    d["tps_kp_s"], d["tps_kp_d"] = npy(tk_s["kp"]), npy(tk_d["kp"])
    d["tps_transformations_bg"] = npy(tnet.create_transformations(small, tk_d, tk_s, bg))
    d["tps_transformations"] = npy(tnet.create_transformations(small, tk_d, tk_s, None))
    tps = util.TPS(mode="kp", bs=2, kp_1=tk_d["kp"].view(2, -1, 5, 2), kp_2=tk_s["kp"].view(2, -1, 5, 2))
    d["tps_theta"], d["tps_control_params"] = npy(tps.theta), npy(tps.control_params)
    tout = tnet(src_img, tk_d, tk_s, bg_param=bg)
    for k in ("deformed_source", "contribution_maps", "deformation", "occlusion"):
        d["tps_fwd_" + k] = npy(tout[k])
    np.savez_compressed(os.path.join(HERE, "prior_motion.npz"), **d)

    # ------------------------------------------------------------------ RaftFlow end to end (size 64)
    r = {}
    rfc = dict(cfg["raft_flow"], size=64)
    rfc["driving_encoder"] = dict(rfc["driving_encoder"], block_expansion=8, max_features=32, num_blocks=3)
    rfc["source_encoder"] = dict(rfc["source_encoder"], block_expansion=8, max_features=32, num_blocks=3)
    rf = syn.fill_state_dict_(raft.RaftFlow(**rfc)).eval()
    dense = net(src_img, kp_d, kp_s, bg_param=bg)
    out, warp_img, occ = rf(kp_s["kp"], kp_d["kp"], dense, img=small, img_full=src_img)
    r["out"], r["warp_img"], r["occlusion"] = npy(out), npy(warp_img), npy(occ)
    rfp = syn.fill_state_dict_(raft.RaftFlow(**dict(rfc, prior_only=True))).eval()
    out, warp_img, occ = rfp(kp_s["kp"], kp_d["kp"], dense, img=small, img_full=src_img)
    r["prior_only_out"], r["prior_only_warp_img"], r["prior_only_occlusion"] = npy(out), npy(warp_img), npy(occ)
    r["state_dict_keys"] = np.array(sorted(rf.state_dict().keys()))
    r["dense_state_dict_keys"] = np.array(sorted(net.state_dict().keys()))
    r["tps_state_dict_keys"] = np.array(sorted(tnet.state_dict().keys()))
    np.savez_compressed(os.path.join(HERE, "raft_flow.npz"), **r)

    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


def equivariance():
    """Training-only random warps: model.py:26-77 `Transform` and util.py TPS mode 'random'.
    Written to its own file so the other fixtures stay byte-identical:  python make_golden.py equivariance"""
    import synthetic_inputs as syn
    util, _, _ = import_reference()
    from modules.model import Transform
    T = syn.tensor
    e = {}
    params = {"sigma_affine": 0.05, "sigma_tps": 0.005, "points_tps": 5}          # config/vox1.yaml transform_params
    frame = T("eq.frame", (2, 3, 32, 40), "uniform")
    kp = T("eq.kp", (2, 10, 2), "uniform", 1.6, -0.8)
    e["frame"], e["kp"] = npy(frame), npy(kp)
    torch.manual_seed(7)
    tr = Transform(2, **params)
    e["theta"], e["control_points"], e["control_params"] = npy(tr.theta), npy(tr.control_points), npy(tr.control_params)
    with torch.no_grad():
        e["transform_frame"] = npy(tr.transform_frame(frame))
        e["warp_kp"] = npy(tr.warp_coordinates(kp))
    kpg = kp.clone().requires_grad_(True)
    e["jacobian_kp"] = npy(tr.jacobian(kpg))
    torch.manual_seed(7)
    aff = Transform(2, sigma_affine=0.05)                                            # affine-only branch (tps False)
    with torch.no_grad():
        e["affine_theta"] = npy(aff.theta)
        e["affine_transform_frame"] = npy(aff.transform_frame(frame))
    torch.manual_seed(7)
    tps = util.TPS("random", 2, **params)
    with torch.no_grad():
        e["tps_random_theta"], e["tps_random_params"] = npy(tps.theta), npy(tps.control_params)
        e["tps_random_grid"] = npy(tps.transform_frame(frame))
        e["tps_random_warp_kp"] = npy(tps.warp_coordinates(kp))
    np.savez_compressed(os.path.join(HERE, "equivariance.npz"), **e)
    print("equivariance.npz", os.path.getsize(os.path.join(HERE, "equivariance.npz")) // 1024, "KiB")


if __name__ == "__main__":
    if sys.argv[1:] == ["equivariance"]:
        equivariance()
    else:
        main()                                   # main() disables grad globally;
        torch.set_grad_enabled(True)             # Transform.jacobian (model.py:72-77) needs autograd
        equivariance()
