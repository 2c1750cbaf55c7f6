"""End-to-end parity of the drop-in RaftFlow / DenseMotionNetwork on the GPU against the golden
output of the unmodified reference and against the CPU oracle (same weights, same inputs)."""
import os

import numpy as np
import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu

import synthetic_inputs as syn               # noqa: E402
from oracle import torch_path as TP                   # noqa: E402

DEV = "cuda"


def _cfg():
    return yaml.safe_load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vox1.yaml")))


def _rf_cfg(size=64):
    rfc = dict(_cfg()["raft_flow"], size=size)
    rfc["driving_encoder"] = dict(rfc["driving_encoder"], block_expansion=8, max_features=32, num_blocks=3)
    rfc["source_encoder"] = dict(rfc["source_encoder"], block_expansion=8, max_features=32, num_blocks=3)
    return rfc


def _rel(a, b, rel):
    a = a.detach().float().cpu().numpy().astype(np.float64)
    b = np.asarray(b.detach().float().cpu().numpy() if torch.is_tensor(b) else b, np.float64)
    rms = np.sqrt((b ** 2).mean())
    err = np.abs(a - b)
    assert (err <= rel * (np.abs(b) + rms)).all(), f"max abs err {err.max():.4g} (rms {rms:.4g})"
    return err.max()


def _inputs(golden):
    d = golden("prior_motion")
    src, _ = syn.frame_pairs(2, 64, seed=1)
    kp_s, kp_d = syn.keypoints(2, 10, seed=1)
    dense = {"deformation": torch.from_numpy(d["fwd_deformation"]), "occlusion": torch.from_numpy(d["fwd_occlusion"])}
    return src, kp_s, kp_d, dense, torch.from_numpy(d["source_small"])


def test_raft_flow_vs_reference_golden(golden):
    import mrfa_b200
    r = golden("raft_flow")
    src, kp_s, kp_d, dense, small = _inputs(golden)
    with torch.no_grad():
        net = syn.fill_state_dict_(mrfa_b200.RaftFlow(**_rf_cfg())).to(DEV).eval()
        assert sorted(net.state_dict().keys()) == list(r["state_dict_keys"])
        dd = {k: v.to(DEV) for k, v in dense.items()}
        out, warp_img, occ = net(kp_s["kp"].to(DEV), kp_d["kp"].to(DEV), dd, img=small.to(DEV), img_full=src.to(DEV))
        # bf16 correlation -> 2e-2 relative on the predicted frames (north_star tolerance)
        _rel(out, r["out"], 2e-2)
        _rel(warp_img, r["warp_img"], 2e-2)
        _rel(occ, r["occlusion"], 2e-2)
        netp = syn.fill_state_dict_(mrfa_b200.RaftFlow(**dict(_rf_cfg(), prior_only=True))).to(DEV).eval()
        out, warp_img, occ = netp(kp_s["kp"].to(DEV), kp_d["kp"].to(DEV), dd, img=small.to(DEV), img_full=src.to(DEV))
        # no correlation on this path: fp32 warps only
        np.testing.assert_allclose(out.cpu().numpy(), r["prior_only_out"], atol=1e-4)
        np.testing.assert_allclose(warp_img.cpu().numpy(), r["prior_only_warp_img"], atol=1e-5)
        np.testing.assert_allclose(occ.cpu().numpy(), r["prior_only_occlusion"], atol=1e-4)


def test_full_chain_vs_oracle_128(golden):
    """DenseMotionNetwork -> RaftFlow at 128x128 (h=w=32), celebvhq-style bg path on, B=2."""
    import mrfa_b200
    cfg = _cfg()
    dmc = dict(cfg["dense_motion"], block_expansion=16, max_features=64, num_blocks=3)
    src, _ = syn.frame_pairs(2, 128, seed=3)
    kp_s, kp_d = syn.keypoints(2, 10, seed=3)
    bg = syn.bg_affine(2, seed=3)
    with torch.no_grad():
        o_dm = syn.fill_state_dict_(TP.DenseMotionOracle(**dmc)).eval()
        o_rf = syn.fill_state_dict_(TP.RaftFlowOracle(**_rf_cfg(128))).eval()
        small = o_dm.down(src)
        dense = o_dm(src, kp_d, kp_s, bg_param=bg)
        ref_out, ref_warp, ref_occ = o_rf(kp_s["kp"], kp_d["kp"], dense, img=small, img_full=src)

        dm = syn.fill_state_dict_(mrfa_b200.DenseMotionNetwork(**dmc)).to(DEV).eval()
        rf = syn.fill_state_dict_(mrfa_b200.RaftFlow(**_rf_cfg(128))).to(DEV).eval()
        c = lambda d: {k: v.to(DEV) for k, v in d.items()}
        got_dense = dm(src.to(DEV), c(kp_d), c(kp_s), bg_param=bg.to(DEV))
        np.testing.assert_allclose(got_dense["deformation"].cpu().numpy(), dense["deformation"].numpy(), atol=1e-4)
        np.testing.assert_allclose(got_dense["occlusion"].cpu().numpy(), dense["occlusion"].numpy(), atol=1e-4)
        out, warp_img, occ = rf(kp_s["kp"].to(DEV), kp_d["kp"].to(DEV), got_dense, img=dm.down(src.to(DEV)), img_full=src.to(DEV))
        _rel(out, ref_out, 2e-2)
        _rel(warp_img, ref_warp, 2e-2)
        _rel(occ, ref_occ, 2e-2)


def test_channels_last_matches_nchw(golden):
    """NHWC execution of the drop-in modules gives the NCHW results (layout is not numerics)."""
    import mrfa_b200
    src, kp_s, kp_d, dense, small = _inputs(golden)
    with torch.no_grad():
        net = syn.fill_state_dict_(mrfa_b200.RaftFlow(**_rf_cfg())).to(DEV).eval()
        net.auto_channels_last = False                    # first pass in the reference's NCHW memory
        dd = {k: v.to(DEV) for k, v in dense.items()}
        args = (kp_s["kp"].to(DEV), kp_d["kp"].to(DEV), dd)
        out0, warp0, occ0 = net(*args, img=small.to(DEV), img_full=src.to(DEV))
        assert not net.channels_last
        net.channels_last_()
        out1, warp1, occ1 = net(*args, img=small.to(DEV), img_full=src.to(DEV))
        # cuDNN may pick different (still fp32) algorithms per layout: allow conv round-off only
        np.testing.assert_allclose(out1.cpu().numpy(), out0.cpu().numpy(), atol=2e-4)
        np.testing.assert_allclose(warp1.cpu().numpy(), warp0.cpu().numpy(), atol=2e-4)
        np.testing.assert_allclose(occ1.cpu().numpy(), occ0.cpu().numpy(), atol=2e-4)
        r = golden("raft_flow")
        _rel(out1, r["out"], 2e-2)


def test_training_backward_matches_oracle(golden):
    """fwd+bwd through DenseMotionNetwork -> RaftFlow (config 5 path, reduced size): every
    parameter and the key-points receive gradients that match CPU autograd on the oracle."""
    import mrfa_b200
    cfg = _cfg()
    dmc = dict(cfg["dense_motion"], block_expansion=16, max_features=64, num_blocks=3)
    src, drv = syn.frame_pairs(2, 64, seed=4)
    kp_s, kp_d = syn.keypoints(2, 10, seed=4)

    def run(dm, rf, dev):
        ks = {k: v.to(dev).clone().requires_grad_() for k, v in kp_s.items()}
        kd = {k: v.to(dev).clone().requires_grad_() for k, v in kp_d.items()}
        s = src.to(dev)
        dense = dm(s, kd, ks)
        out, warp_img, _ = rf(ks["kp"], kd["kp"], dense, img=dm.down(s), img_full=s)
        loss = (out - drv.to(dev)).abs().mean() + 0.1 * (warp_img - drv.to(dev)).abs().mean()
        loss.backward()
        grads = {"kp_s": ks["kp"].grad, "kp_d": kd["kp"].grad, "jac_d": kd["jacobian"].grad}
        for name, p in list(dm.named_parameters()) + [("rf." + n, p) for n, p in rf.named_parameters()]:
            grads[name] = p.grad
        return float(loss), grads

    o_dm = syn.fill_state_dict_(TP.DenseMotionOracle(**dmc)).eval()
    o_rf = syn.fill_state_dict_(TP.RaftFlowOracle(**_rf_cfg())).eval()
    ref_loss, ref = run(o_dm, o_rf, "cpu")
    dm = syn.fill_state_dict_(mrfa_b200.DenseMotionNetwork(**dmc)).to(DEV).eval()
    rf = syn.fill_state_dict_(mrfa_b200.RaftFlow(**_rf_cfg())).to(DEV).eval()
    loss, got = run(dm, rf, DEV)
    assert abs(loss - ref_loss) < 2e-2 * abs(ref_loss)
    checked = 0
    for name, g_ref in ref.items():
        assert g_ref is not None, name
        g = got[name]
        assert g is not None, f"{name} received no gradient"
        a, b = g.detach().float().cpu().flatten().double(), g_ref.detach().flatten().double()
        if b.norm() < 1e-8:
            continue
        rel = float((a - b).norm() / b.norm())
        cos = float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))
        # bf16 correlation volume + bf16 volume-gradient GEMMs: ~1e-2 relative noise
        assert cos > 0.98 and rel < 0.15, f"{name}: rel {rel:.3g} cos {cos:.4f}"
        checked += 1
    assert checked > 200


def test_cuda_graph_replay_is_bit_identical(golden):
    """The whole refinement forward captures into a CUDA graph (no host syncs in any kernel
    wrapper); replay on new inputs equals the eager call bit for bit."""
    import mrfa_b200
    cfg = _cfg()
    dmc = dict(cfg["dense_motion"], block_expansion=16, max_features=64, num_blocks=3)
    c = lambda d: {k: v.to(DEV) for k, v in d.items()}
    with torch.no_grad():
        dm = syn.fill_state_dict_(mrfa_b200.DenseMotionNetwork(**dmc)).to(DEV).eval().channels_last_()
        rf = syn.fill_state_dict_(mrfa_b200.RaftFlow(**_rf_cfg())).to(DEV).eval().channels_last_()
        src0, _ = syn.frame_pairs(1, 64, seed=10)
        ks0, kd0 = syn.keypoints(1, 10, seed=10)
        g = mrfa_b200.GraphedRefiner(dm, rf, src0.to(DEV), c(ks0), c(kd0))
        for seed in (11, 12):
            src, _ = syn.frame_pairs(1, 64, seed=seed)
            ks, kd = syn.keypoints(1, 10, seed=seed)
            out_g, warp_g, occ_g = (t.clone() for t in g(src.to(DEV), c(ks), c(kd)))
            dense = dm(src.to(DEV), c(kd), c(ks))
            out, warp, occ = rf(ks["kp"].to(DEV), kd["kp"].to(DEV), dense, img=dm.down(src.to(DEV)), img_full=src.to(DEV))
            assert torch.equal(out_g, out) and torch.equal(warp_g, warp) and torch.equal(occ_g, occ)


def test_full_size_vox1_vs_oracle():
    """BASELINE.json configs[0]/[1] at their real size: unchanged vox1.yaml (full network widths),
    one 256x256 pair, product on the GPU vs the CPU oracle with identical weights and inputs."""
    import mrfa_b200
    cfg = _cfg()
    src, drv = syn.frame_pairs(1, 256, seed=6)
    kp_s, kp_d = syn.keypoints(1, 10, seed=6)
    with torch.no_grad():
        o_dm = syn.fill_state_dict_(TP.DenseMotionOracle(**cfg["dense_motion"])).eval()
        o_rf = syn.fill_state_dict_(TP.RaftFlowOracle(**cfg["raft_flow"])).eval()
        dense = o_dm(src, kp_d, kp_s)
        ref_out, ref_warp, ref_occ = o_rf(kp_s["kp"], kp_d["kp"], dense, img=o_dm.down(src), img_full=src)
        dm = mrfa_b200.DenseMotionNetwork(**cfg["dense_motion"]).eval()
        rf = mrfa_b200.RaftFlow(**cfg["raft_flow"]).eval()
        dm.load_state_dict(o_dm.state_dict())
        rf.load_state_dict(o_rf.state_dict())
        dm, rf = dm.to(DEV), rf.to(DEV)
        c = lambda d: {k: v.to(DEV) for k, v in d.items()}
        got = dm(src.to(DEV), c(kp_d), c(kp_s))
        np.testing.assert_allclose(got["deformation"].cpu().numpy(), dense["deformation"].numpy(), atol=2e-4)
        out, warp_img, occ = rf(kp_s["kp"].to(DEV), kp_d["kp"].to(DEV), got, img=dm.down(src.to(DEV)), img_full=src.to(DEV))
        assert rf.channels_last and dm.channels_last               # inference switched itself to NHWC
        # The synthetic frames are white noise (0.5 intensity change per pixel), the harshest case for a
        # warp: a 0.03 px flow difference from the bf16 volume already moves a pixel by 1.5e-2.  Bound
        # the error in norm (2e-2 relative, the north_star tolerance) and against the dynamic range.
        for name, a, b in (("out", out, ref_out), ("warp_img", warp_img, ref_warp), ("occlusion", occ, ref_occ)):
            a, b = a.double().cpu(), b.double()
            rel_l2 = float((a - b).norm() / b.norm())
            max_abs = float((a - b).abs().max())
            print(f"full-size {name}: rel-L2 {rel_l2:.3e}, max|err| {max_abs:.3e}, mean|err| {float((a - b).abs().mean()):.3e}, max|ref| {float(b.abs().max()):.3f}")
            assert rel_l2 < 2e-2, (name, rel_l2)
            assert max_abs < 2e-2 * (1.0 + float(b.abs().max())), (name, max_abs)


def test_custom_op_schemas_and_fake_impls():
    """torch.library.opcheck: schema + fake (meta) implementations agree with the CUDA kernels."""
    from torch.library import opcheck
    torch.manual_seed(20)
    feat = torch.randn(2, 8, 6, 7, device=DEV)
    grid = torch.rand(2, 5, 4, 2, device=DEV) * 2 - 1
    tests = ("test_schema", "test_faketensor")
    opcheck(torch.ops.mrfa.grid_sample.default, (feat, grid, 0, 0, False, 1), test_utils=tests)
    opcheck(torch.ops.mrfa.grid_sample.default, (feat.contiguous(memory_format=torch.channels_last), grid, 2, 0, True, 1), test_utils=tests)
    opcheck(torch.ops.mrfa.dual_warp.default, (feat, torch.randn(2, 2, 6, 7, device=DEV), torch.rand(2, 6, 7, 2, device=DEV)), test_utils=tests)
    corr = torch.randn(2 * 9, 1, 8, 8, device=DEV)
    opcheck(torch.ops.mrfa.corr_lookup.default, (corr, torch.ops.mrfa.avg_pool2x2(corr), torch.rand(2, 2, 3, 3, device=DEV) * 8, 8, 8, 9, 0, 3, 0, False), test_utils=tests)
    opcheck(torch.ops.mrfa.corr_pyramid.default, (torch.randn(1, 64, 16, 16, device=DEV), torch.randn(1, 64, 16, 16, device=DEV), 0.125), test_utils=tests)
    opcheck(torch.ops.mrfa.kp2gaussian.default, (torch.rand(2, 10, 2, device=DEV), None, 8, 8, 0.1), test_utils=tests)
    opcheck(torch.ops.mrfa.resize_bilinear.default, (feat, 12, 9, 1), test_utils=tests)
    opcheck(torch.ops.mrfa.channel_affine.default, (feat, torch.rand(8, device=DEV), None, None, 1), test_utils=tests)


def test_bench_settings_full_size_vs_oracle():
    """The numerics bench.py times are the numerics tested: TF32 convolutions allowed (PyTorch default), cuDNN autotuning on,
    channels_last, every MRFA_* fast path at its default, batch 4 of full-width vox1 256x256 pairs -- against the CPU oracle at
    the north_star tolerance for bf16/TF32 end-to-end frames (2e-2 relative)."""
    import mrfa_b200
    cfg = _cfg()
    B = 4
    src, drv = syn.frame_pairs(B, 256, seed=7)
    kp_s, kp_d = syn.keypoints(B, 10, seed=7)
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.benchmark = True
        with torch.no_grad():
            o_dm = syn.fill_state_dict_(TP.DenseMotionOracle(**cfg["dense_motion"])).eval()
            o_rf = syn.fill_state_dict_(TP.RaftFlowOracle(**cfg["raft_flow"])).eval()
            dense = o_dm(src, kp_d, kp_s)
            ref_out, ref_warp, ref_occ = o_rf(kp_s["kp"], kp_d["kp"], dense, img=o_dm.down(src), img_full=src)
            dm = mrfa_b200.DenseMotionNetwork(**cfg["dense_motion"]).eval()
            rf = mrfa_b200.RaftFlow(**cfg["raft_flow"]).eval()
            dm.load_state_dict(o_dm.state_dict())
            rf.load_state_dict(o_rf.state_dict())
            dm, rf = dm.to(DEV).channels_last_(), rf.to(DEV).channels_last_()      # what bench.py does
            c = lambda d: {k: v.to(DEV) for k, v in d.items()}
            for _ in range(2):                                                      # second pass: autotuned algorithms, warm caches
                got = dm(src.to(DEV), c(kp_d), c(kp_s))
                out, warp_img, occ = rf(kp_s["kp"].to(DEV), kp_d["kp"].to(DEV), got, img=dm.down(src.to(DEV)), img_full=src.to(DEV))
        for name, a, b in (("out", out, ref_out), ("warp_img", warp_img, ref_warp), ("occlusion", occ, ref_occ)):
            a, b = a.double().cpu(), b.double()
            rel_l2 = float((a - b).norm() / b.norm())
            max_abs = float((a - b).abs().max())
            print(f"bench settings (TF32, NHWC, B={B}) {name}: rel-L2 {rel_l2:.3e}, max|err| {max_abs:.3e}")
            assert rel_l2 < 2e-2, (name, rel_l2)
            assert max_abs < 2e-2 * (1.0 + float(b.abs().max())), (name, max_abs)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved


def test_chain_512_vs_oracle():
    """BASELINE.json configs[3]: 512x512 (h = w = 128, N = 16384), B = 2 so the oracle takes the reference's
    batch_bilinear_sampler branch (raft.py:39-40: batch > 1 and h1 >= 128) at the three finest levels.  DenseMotionNetwork ->
    RaftFlow on the GPU (bf16 pyramid 4x the 256x256 one, lookups at H = W = 128) vs the CPU oracle; the structure encoders and
    the dense-motion hourglass are narrowed (the generator keeps its width: raft.py:105-113 hard-codes the pyramid channels)."""
    import mrfa_b200
    cfg = _cfg()
    dmc = dict(cfg["dense_motion"], block_expansion=16, max_features=64, num_blocks=3)
    rfc = _rf_cfg(512)
    src, _ = syn.frame_pairs(2, 512, seed=9)
    kp_s, kp_d = syn.keypoints(2, 10, seed=9)
    with torch.no_grad():
        o_dm = syn.fill_state_dict_(TP.DenseMotionOracle(**dmc)).eval()
        o_rf = syn.fill_state_dict_(TP.RaftFlowOracle(**rfc)).eval()
        dense = o_dm(src, kp_d, kp_s)
        ref_out, ref_warp, ref_occ, trace = o_rf(kp_s["kp"], kp_d["kp"], dense, img=o_dm.down(src), img_full=src, return_trace=True)
        dm = syn.fill_state_dict_(mrfa_b200.DenseMotionNetwork(**dmc)).to(DEV).eval()
        rf = syn.fill_state_dict_(mrfa_b200.RaftFlow(**rfc)).to(DEV).eval()
        c = lambda d: {k: v.to(DEV) for k, v in d.items()}
        got = dm(src.to(DEV), c(kp_d), c(kp_s))
        np.testing.assert_allclose(got["deformation"].cpu().numpy(), dense["deformation"].numpy(), atol=2e-4)
        # the lookup at H = W = 128 in isolation: same fp32 coordinates as the oracle's basic-resolution iteration (i = 3)
        q_d, k_s = rf.structure_features(kp_s["kp"].to(DEV), kp_d["kp"].to(DEV), dm.down(src.to(DEV)))
        pyr = mrfa_b200.CorrPyramid(q_d, k_s, rf.scale)
        coords = trace["flow3"] + TP.coords_grid(2, 128, 128)
        lk = pyr.block(0)(coords.to(DEV))
        _rel(lk, trace["corr3"], 2e-2)
        out, warp_img, occ = rf(kp_s["kp"].to(DEV), kp_d["kp"].to(DEV), got, img=dm.down(src.to(DEV)), img_full=src.to(DEV))
    for name, a, b in (("out", out, ref_out), ("warp_img", warp_img, ref_warp), ("occlusion", occ, ref_occ)):
        a, b = a.double().cpu(), b.double()
        rel_l2 = float((a - b).norm() / b.norm())
        max_abs = float((a - b).abs().max())
        print(f"512x512 chain {name}: rel-L2 {rel_l2:.3e}, max|err| {max_abs:.3e}")
        assert rel_l2 < 2e-2, (name, rel_l2)
        assert max_abs < 2e-2 * (1.0 + float(b.abs().max())), (name, max_abs)
