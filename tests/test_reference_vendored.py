"""The CPU arm of the benchmark is the UNMODIFIED reference (oracle/_ref, vendored by oracle/build_ref.py) and the
oracle restatement equals it bit for bit; neither loads the product package or its CUDA library."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MRFA_REF", "/root/reference")

from oracle import reference_arm as RA          # noqa: E402
from oracle import torch_path as TP             # noqa: E402
import synthetic_inputs as syn                  # noqa: E402

needs_ref = pytest.mark.skipif(not RA.available(), reason="oracle/_ref not vendored (run python oracle/build_ref.py)")


def _cfg(name="vox1"):
    return yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", name + ".yaml")))


def _small(cfg, size):
    dmc = dict(cfg["dense_motion"], block_expansion=16, max_features=64, num_blocks=3)
    rfc = dict(cfg["raft_flow"], size=size)
    rfc["driving_encoder"] = dict(rfc["driving_encoder"], block_expansion=8, max_features=32, num_blocks=3)
    rfc["source_encoder"] = dict(rfc["source_encoder"], block_expansion=8, max_features=32, num_blocks=3)
    return dict(cfg, dense_motion=dmc, raft_flow=rfc)


@needs_ref
def test_manifest_matches_vendored_files_and_source():
    man = json.load(open(os.path.join(RA.REF_DIR, "MANIFEST.json")))
    assert "modules/raft.py" in man["files"] and "modules/util.py" in man["files"] and "config/vox1.yaml" in man["files"]
    for rel, sha in man["files"].items():
        got = hashlib.sha256(open(os.path.join(RA.REF_DIR, rel), "rb").read()).hexdigest()
        assert got == sha, f"{rel} was modified after vendoring"
        src = os.path.join(REF, rel)
        if os.path.exists(src):                                  # build container: the copy equals the checkout
            assert hashlib.sha256(open(src, "rb").read()).hexdigest() == sha, rel
    # the YAMLs the tests / bench read are the reference's own, unchanged
    for name in ("vox1", "celebvhq"):
        a = open(os.path.join(ROOT, "tests", "golden", name + ".yaml"), "rb").read()
        assert hashlib.sha256(a).hexdigest() == man["files"][f"config/{name}.yaml"]


def test_oracle_and_reference_arm_do_not_load_the_product():
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from oracle import torch_path, np_ops, conv_blocks, reference_arm\n"
            "import synthetic_inputs\n"
            "if reference_arm.available(): reference_arm.load()\n"
            "bad = [m for m in sys.modules if m.startswith('mrfa_b200')]\n"
            "assert not bad, bad\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libmrfa_b200' not in maps\n"
            "print('CLEAN')\n" % ROOT)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=240)
    assert res.returncode == 0 and "CLEAN" in res.stdout, res.stderr[-1500:]
    for f in os.listdir(os.path.join(ROOT, "oracle")):
        if f.endswith(".py"):
            src = open(os.path.join(ROOT, "oracle", f)).read()
            assert "import mrfa_b200" not in src and "from mrfa_b200" not in src, f


@needs_ref
@pytest.mark.parametrize("size,batch,config", [(64, 2, "vox1"), (128, 1, "celebvhq")])
def test_oracle_equals_vendored_reference(size, batch, config):
    """Same weights, same inputs: the restated control flow (oracle/torch_path.py + oracle/conv_blocks.py) reproduces the
    reference's DenseMotionNetwork -> RaftFlow outputs exactly (same stock torch CPU ops in the same order)."""
    cfg = _small(_cfg(config), size)
    torch.manual_seed(0)
    nets = RA.build_networks(cfg, size)
    for n in nets[1:]:
        syn.fill_state_dict_(n)
    o_dm = syn.fill_state_dict_(TP.DenseMotionOracle(**cfg["dense_motion"])).eval()
    o_rf = syn.fill_state_dict_(TP.RaftFlowOracle(**cfg["raft_flow"])).eval()
    assert sorted(o_dm.state_dict()) == sorted(nets[1].state_dict())
    assert sorted(o_rf.state_dict()) == sorted(nets[2].state_dict())
    src, _ = syn.frame_pairs(batch, size, seed=2)
    kp_s, kp_d = syn.keypoints(batch, 10, seed=2)
    bg = syn.bg_affine(batch, seed=2) if cfg["train_params"]["bg_start"] == 0 else None
    with torch.no_grad():
        out, warp, occ, dense = RA.forward(nets, src, kp_s, kp_d, bg)
        d = o_dm(src, kp_d, kp_s, bg_param=bg)
        o_out, o_warp, o_occ = o_rf(kp_s["kp"], kp_d["kp"], d, img=o_dm.down(src), img_full=src)
        assert torch.equal(nets[0](src), o_dm.down(src))
    for k in ("deformation", "occlusion", "mask", "sparse_deformed"):
        assert torch.equal(dense[k], d[k]), k
    assert torch.equal(out, o_out) and torch.equal(warp, o_warp) and torch.equal(occ, o_occ)


@needs_ref
def test_oracle_equals_vendored_reference_512_branch():
    """512x512 (h = w = 128), B = 2: the reference takes its batch_bilinear_sampler branch (raft.py:39-40: batch > 1 and
    h1 >= 128, util.py:40-51) at the three finest levels; the oracle's per-sample chunking reproduces it exactly.  Conv widths
    reduced so the CPU run takes seconds; the correlation / lookup path runs at its real 512x512 geometry."""
    cfg = _small(_cfg(), 512)
    cfg["raft_flow"]["generator"] = dict(cfg["raft_flow"]["generator"], block_expansion=8, max_features=32)
    m = RA.load()
    torch.manual_seed(0)
    rf = syn.fill_state_dict_(_NarrowRaft.build(m, cfg["raft_flow"])).eval()
    o_rf = syn.fill_state_dict_(_NarrowRaft.build(TP, cfg["raft_flow"])).eval()
    src, _ = syn.frame_pairs(2, 512, seed=8)
    kp_s, kp_d = syn.keypoints(2, 10, seed=8)
    dense = _NarrowRaft.prior(2, 128, seed=8)
    img = torch.nn.functional.avg_pool2d(src, 4)
    with torch.no_grad():
        a = rf(kp_s["kp"], kp_d["kp"], dense, img=img, img_full=src)
        b = o_rf(kp_s["kp"], kp_d["kp"], dense, img=img, img_full=src)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


class _NarrowRaft:
    """RaftFlow with the feature-pyramid widths of a narrow generator: the reference hard-codes 512/512/512/256/128/64
    (raft.py:105-113), which at 512x512 on CPU is minutes per pair.  The to_context convolutions are rebuilt for the
    narrow widths in BOTH implementations the same way, everything else is the class under test."""

    @staticmethod
    def build(mod, rfc):
        cls = mod.RaftFlowOracle if hasattr(mod, "RaftFlowOracle") else mod["raft"].RaftFlow
        net = cls(**rfc)
        g = rfc["generator"]
        widths = [min(g["max_features"], g["block_expansion"] * 2 ** i) for i in range(g["num_up_blocks"] + 1)][::-1]
        net.to_context = torch.nn.ModuleList(torch.nn.Conv2d(c, 192, 1) for c in widths)
        return net

    @staticmethod
    def prior(B, h, seed):
        ident = TP.make_coordinate_grid((h, h)).view(1, h, h, 2)
        return {"deformation": ident + syn.tensor("prior.def", (B, h, h, 2), "normal", 0.03, 0.0, seed),
                "occlusion": syn.tensor("prior.occ", (B, 1, h, h), "normal", 1.0, 0.0, seed)}


def test_tps_reference_solve_agrees_with_fp64(golden):
    """Round 1 bounded the product's TPS parameters against the reference only to 2e-3 / 5e-3, blaming the reference's fp32
    torch.inverse (util.py:377-379).  Quantified here on the golden inputs, that claim does not hold for the CPU reference:
    its theta / control_params / grids sit within a few 1e-6 of an fp64 solve, so the GPU tests (tests/test_gpu_kernels.py::
    test_tps_kernels) now hold the product to 3e-5 against the reference's own outputs."""
    from oracle import np_ops as O
    d = golden("prior_motion")
    kp_d, kp_s = d["tps_kp_d"].reshape(2, -1, 5, 2), d["tps_kp_s"].reshape(2, -1, 5, 2)
    th64, _cp, cw64 = O.tps_params(kp_d, kp_s)                       # fp64 solve, rounded to fp32
    e_th = float(np.abs(d["tps_theta"] - th64).max())
    e_cw = float(np.abs(d["tps_control_params"] - cw64).max())
    grid = O.tps_transformations(d["tps_kp_d"], d["tps_kp_s"], 16, 16)
    e_grid = float(np.abs(grid - d["tps_transformations"]).max())
    print(f"reference vs fp64 solve: theta {e_th:.2e}, control_params {e_cw:.2e} (max |w| {np.abs(cw64).max():.2f}), grid {e_grid:.2e}")
    assert e_th < 5e-6 and e_cw < 1e-5 and e_grid < 1e-5
    # and on the warped image (dense_motion.py:241, align_corners=True): far inside the 1e-5 warp tolerance
    src = torch.from_numpy(d["source_small"])[:, None].expand(2, grid.shape[1], 3, 16, 16).reshape(-1, 3, 16, 16)
    wa = torch.nn.functional.grid_sample(src, torch.from_numpy(grid).reshape(-1, 16, 16, 2).float(), align_corners=True)
    wb = torch.nn.functional.grid_sample(src, torch.from_numpy(d["tps_transformations"]).reshape(-1, 16, 16, 2), align_corners=True)
    assert float((wa - wb).abs().max()) < 5e-5
