"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the
product refuses CPU tensors (no fallback), state_dict layouts match the reference, and the
multi-GPU sharding / statistics logic works over gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mrfa_b200 import _lib
    header = open(os.path.join(ROOT, "include", "mrfa_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(mrfa_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/mrfa_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.mrfa_abi_version() == _lib.ABI_VERSION
    # pure host helpers of the ABI (no GPU needed)
    assert _lib.lib.mrfa_corr_rows_total(64, 64) == 4096 + 1024 + 256 + 64
    assert _lib.lib.mrfa_corr_row_offset(64, 64, 0) == 0
    assert _lib.lib.mrfa_corr_row_offset(64, 64, 3) == 4096 + 1024 + 256
    assert b"bad argument" in _lib.lib.mrfa_error_string(-1)


def test_no_cpu_fallback():
    import mrfa_b200
    with pytest.raises(Exception):
        mrfa_b200.bilinear_sampler(torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 2, 2))
    with pytest.raises(Exception):
        mrfa_b200.kp2gaussian(torch.zeros(1, 10, 2), (8, 8), 0.1)
    with pytest.raises(Exception):
        mrfa_b200.make_coordinate_grid((4, 4), "torch.FloatTensor")
    with pytest.raises(Exception):
        mrfa_b200.CorrBlock(torch.zeros(4, 1, 8, 8))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "mrfa_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, f)).read(), f"{f} references the oracle"


def test_state_dict_layouts_match_reference(golden):
    import yaml
    import mrfa_b200
    r = golden("raft_flow")
    cfg = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", "vox1.yaml")))
    rfc = dict(cfg["raft_flow"], size=64)
    rfc["driving_encoder"] = dict(rfc["driving_encoder"], block_expansion=8, max_features=32, num_blocks=3)
    rfc["source_encoder"] = dict(rfc["source_encoder"], block_expansion=8, max_features=32, num_blocks=3)
    assert sorted(mrfa_b200.RaftFlow(**rfc).state_dict()) == list(r["state_dict_keys"])
    small = dict(block_expansion=16, max_features=64, num_blocks=3)
    assert sorted(mrfa_b200.DenseMotionNetwork(**dict(cfg["dense_motion"], **small)).state_dict()) == list(r["dense_state_dict_keys"])
    assert sorted(mrfa_b200.TPSDenseMotionNetwork(**dict(cfg["tpsm_dense_motion"], **small)).state_dict()) == list(r["tps_state_dict_keys"])
    # the unchanged reference YAMLs parse and build the full-size modules
    for name in ("vox1", "celebvhq"):
        c = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", name + ".yaml")))
        net = mrfa_b200.RaftFlow(**c["raft_flow"])
        assert len(net.state_dict()) == 361
        assert len(mrfa_b200.DenseMotionNetwork(**c["dense_motion"]).state_dict()) == 75


def test_shard_range():
    from mrfa_b200.dist import shard_range
    for total in (0, 1, 7, 64, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from mrfa_b200.dist import init_from_env, shard_range, reduce_stats
rank, world, local = init_from_env("gloo")
s, e = shard_range(5, rank, world)
# each rank "processes" its shard: L1 sum = sum of pair ids, 10 elements per pair
out = reduce_stats(float(sum(range(s, e))), 10.0 * (e - s), 0.5 * (rank + 1), float(e - s))
if rank == 0:
    print("RESULT", out["l1_mean"], out["pairs"], out["elapsed_s"], out["pairs_per_s"])
dist.destroy_process_group()
"""


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT")][0].split()
    l1_mean, pairs, elapsed, pps = (float(x) for x in line[1:])
    assert pairs == 5.0 and elapsed == 1.0                       # SUM of pairs, MAX of elapsed
    assert abs(l1_mean - (0 + 1 + 2 + 3 + 4) / 50.0) < 1e-12
    assert abs(pps - 5.0) < 1e-12


def test_bench_reference_arm_skips_nonzero_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_patch_reference_rebinds_names():
    """patch_reference() swaps the hot-path names inside an imported reference checkout (only
    runs where the reference is mounted; the GPU box has no reference)."""
    ref = os.environ.get("MRFA_REF", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "modules")):
        pytest.skip("reference checkout not available")
    code = f"""
import sys, types, torch
for n in ("timm", "timm.models", "timm.models.layers", "timm.models.layers.weight_init"):
    sys.modules.setdefault(n, types.ModuleType(n))
sys.modules["timm.models.layers.weight_init"].trunc_normal_ = torch.nn.init.trunc_normal_
sys.path.insert(0, {ref!r}); sys.path.insert(0, {ROOT!r})
import mrfa_b200
mods = mrfa_b200.patch_reference("modules")
import modules.raft as raft, modules.dense_motion as dm, modules.util as util
assert raft.CorrBlock is mrfa_b200.CorrBlock and raft.RaftFlow is mrfa_b200.RaftFlow
assert raft.bilinear_sampler is mrfa_b200.bilinear_sampler and raft.coords_grid is mrfa_b200.coords_grid
assert dm.DenseMotionNetwork is mrfa_b200.DenseMotionNetwork and dm.TPS is mrfa_b200.TPS
assert util.kp2gaussian is mrfa_b200.kp2gaussian and util.deform_input is mrfa_b200.deform_input
print("PATCHED")
"""
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=240)
    assert res.returncode == 0 and "PATCHED" in res.stdout, res.stderr[-1500:]


def test_subpixel_and_space_to_depth_weight_transforms_cpu():
    """Host-side algebra behind two inference rewrites, checked with stock CPU convolutions:
    (1) UpBlock2d (util.py:160-177): conv3x3(nearest_x2(x)) == de-interleave(conv2x2_pad1(x)) with 4*Cout phase-major
        outputs;  (2) the generator's final 7x7 convolution (generator.py:32,61) == pixel_shuffle(conv3x3(space_to_depth_4(x)))."""
    import torch
    import torch.nn.functional as F
    from mrfa_b200 import blocks
    import synthetic_inputs as syn
    torch.manual_seed(5)
    up = syn.fill_state_dict_(blocks.UpBlock2d(6, 8, kernel_size=3, padding=1)).eval()
    x = torch.randn(2, 6, 5, 7)
    with torch.no_grad():
        ref = up(x)                                                          # plain path on CPU
        w2, b2 = up.subpixel_weights()
        y2 = F.relu(F.conv2d(x, w2, b2, padding=1))                          # (2, 32, 6, 8)
        C = 8
        out = torch.empty_like(ref)
        for a in (0, 1):
            for b in (0, 1):
                ph = y2[:, (2 * a + b) * C:(2 * a + b + 1) * C]
                out[:, :, a::2, b::2] = ph[:, :, a:a + 5, b:b + 7]           # pixel (Y,X) <- [Y//2 + a, X//2 + b]
        assert torch.allclose(out, ref, atol=1e-5)
        gen = syn.fill_state_dict_(blocks.OcclusionAwareGenerator(3, 16, 64, 3)).eval()
        z = torch.randn(2, 16, 12, 20)
        zs = z.reshape(2, 16, 3, 4, 5, 4).permute(0, 3, 5, 1, 2, 4).reshape(2, 256, 3, 5)   # channel (iy*4+ix)*16 + c
        assert torch.allclose(gen._final_s2d(zs), gen.final(z), atol=1e-5)


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the vendored reference on the CPU, the one arm that runs without a GPU) prints one JSON line
    carrying the keys of the bench contract; run at 128x128 with one pair so it finishes in seconds."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "128", "--steps", "1",
                          "--warmup", "1", "--ref-batch", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["vs_baseline"] is None and d["higher_is_better"] is True and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["config"]["device"] == "cpu" and d["config"]["pairs_per_step_this_arm"] == 1 and "cuDNN" not in json.dumps(d["config"])
    cb = d["cpu_baseline"]
    from oracle import reference_arm
    assert cb["kind"] == ("reference" if reference_arm.available() else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"] and cb["cpu_model"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
