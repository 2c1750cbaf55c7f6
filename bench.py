#!/usr/bin/env python
"""bench.py -- frame-pairs/s of the MRFA refinement forward on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--size S]

A step = one refinement forward (AntiAlias down -> DenseMotionNetwork -> RaftFlow, vox1.yaml
architecture, random-init weights, eval / no_grad) over one batch of synthetic frame pairs with
synthetic key-points (the key-point detector is upstream of the hot path).
N=1 workload: BASELINE.json configs[1] (batch 64, 256x256, one B200).  N>1: one process per GPU
(torchrun), the batch of pairs is sharded by rank with no data-path collective ("weak" scaling:
64 pairs per GPU); NCCL only all-reduces the reconstruction-L1 / timing statistics.

Prints ONE JSON line (rank 0).  `value` is timed with the inputs resident in HBM; `e2e` times
the same call with pinned HOST inputs (H2D inside) and the predicted frames read back (D2H).
`--impl reference` times the unmodified reference (vendored into oracle/_ref by oracle/build_ref.py) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# kernels of the hot path proper: timed inside the timed region, candidates for `roofline`
HOT_PATH = ("dual_warp_fwd", "grid_sample_fwd", "corr_volume", "corr_lookup_fwd", "corr_pack", "dense_motion_prior",
            "tps_motion_prior", "tps_solve", "kp2gaussian", "coords_grid", "make_coordinate_grid", "prior_to_flow",
            "avg_pool2x2")
METRIC = "frame-pairs/sec (refinement forward, 256x256)"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="pairs per GPU per step")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--config", default="vox1", choices=["vox1", "celebvhq"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nchw", action="store_true", help="keep the networks in NCHW memory (default: channels_last)")
    ap.add_argument("--cudnn-benchmark", type=int, default=1, help="torch.backends.cudnn.benchmark (algorithm autotuning)")
    ap.add_argument("--ref-batch", type=int, default=0, help="pairs per step of the CPU reference arm (0 = auto)")
    ap.add_argument("--other-configs", type=int, default=1,
                    help="also run short measurements of BASELINE.json configs 3/4/5 (celebvhq, 512x512 sweep, training step)")
    ap.add_argument("--other-timeout", type=int, default=300,
                    help="watchdog (s) around the other-config measurements: on expiry rank 0 prints the line with what finished")
    ap.add_argument("--parity-pairs", type=int, default=2, help="pairs of the timed batch compared with the CPU reference in the same run")
    return ap.parse_args()


def load_cfg(name):
    import yaml
    return yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", name + ".yaml")))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            out = {"source": "measured"}
            out["hbm_gbs"] = float(d.get("hbm_gbs") or FALLBACK_PEAKS["hbm_gbs"])
            out["bf16_tflops"] = float(d.get("bf16_tflops") or FALLBACK_PEAKS["bf16_tflops"])
            out["bf16_tflops_sustained"] = float(d.get("bf16_tflops_sustained") or out["bf16_tflops"])
            return out
        except Exception:
            pass
    return dict(FALLBACK_PEAKS, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 8:
                    self.rows.append(f)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm, mx, reasons = [], 0, set()
        for f in self.rows:
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference vendored into oracle/_ref by oracle/build_ref.py (kind "reference");
# only if that copy is absent, the oracle port oracle/torch_path.py (kind "port").  Neither imports mrfa_b200.
def build_cpu_nets(cfg, size):
    """-> (kind, forward(src, kp_s, kp_d, bg) -> predicted frames, modules dict for load_state_dict)."""
    import torch
    from oracle import reference_arm as RA
    torch.manual_seed(0)
    if RA.available():
        nets = RA.build_networks(cfg, size)
        return "reference", (lambda src, kp_s, kp_d, bg: RA.forward(nets, src, kp_s, kp_d, bg)[0]), {"dm": nets[1], "rf": nets[2]}
    from oracle import torch_path as TP
    dm = TP.DenseMotionOracle(**cfg["dense_motion"]).eval()
    rf = TP.RaftFlowOracle(**dict(cfg["raft_flow"], size=size)).eval()

    def fwd(src, kp_s, kp_d, bg):
        dense = dm(src, kp_d, kp_s, bg_param=bg)
        return rf(kp_s["kp"], kp_d["kp"], dense, img=dm.down(src), img_full=src)[0]
    return "port", fwd, {"dm": dm, "rf": rf}


CPU_KIND_TEXT = {"reference": "unmodified reference modules (oracle/_ref: DenseMotionNetwork + RaftFlow, demo.py:47-73 composition)",
                 "port": "oracle torch-CPU port of the reference path (oracle/_ref not vendored)"}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def time_cpu_path(cfg, size, batch, steps, warmup):
    """The reference's CPU implementation of the path on all host threads."""
    import torch
    import synthetic_inputs as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind, fwd, _ = build_cpu_nets(cfg, size)
    src, _ = syn.frame_pairs(batch, size, seed=0)
    kp_s, kp_d = syn.keypoints(batch, cfg["raft_flow"]["num_kp"], seed=0)
    bg = syn.bg_affine(batch) if cfg["train_params"]["bg_start"] == 0 else None
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            fwd(src, kp_s, kp_d, bg)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    total = sum(times)
    return {"value": batch * len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": cores, "kind": kind,
            "best_pairs_s": batch / min(times), "median_pairs_s": batch / sorted(times)[len(times) // 2], "cpu_model": cpu_model()}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path (the vendored, unmodified
    reference modules), all host threads, on a bounded sample of the workload: each step processes
    `ref_batch` pairs, sized from a one-pair probe so that the whole --steps/--warmup run stays within
    a few minutes.  Never imports mrfa_b200 (asserted below): no product code, no CUDA library."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = load_cfg(args.config)
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    probe = time_cpu_path(cfg, args.size, 1, 1, 1)
    t1 = probe["ms_per_step"] / 1e3
    ref_batch = args.ref_batch or int(max(1, min(8, 180.0 / ((steps + warmup) * t1))))
    r = time_cpu_path(cfg, args.size, ref_batch, steps, warmup)
    assert "mrfa_b200" not in sys.modules, "the reference arm must not load the product"
    sample = (f"{ref_batch} pair(s) per step (bounded sample of the {args.batch}-pair batch), {steps} timed steps after "
              f"{warmup} warm-up, {CPU_KIND_TEXT[r['kind']]}, {r['cores']} threads")
    config = workload_config(args, args.batch, max(1, args.gpus))
    config.update({"reference_arm": f"CPU, fp32 (stock torch CPU ops, MKL-DNN convolutions), {r['cores']} host threads, "
                                    f"{ref_batch} pair(s) per step; same workload definition, bounded sample",
                   "pairs_per_step_this_arm": ref_batch, "device": "cpu"})
    config.pop("convs", None)
    config.pop("l2", None)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": r["value"], "unit": "pairs/s", "cores": r["cores"], "kind": r["kind"], "sample": sample,
                             "cpu_model": r["cpu_model"], "best": r["best_pairs_s"], "median": r["median_pairs_s"]},
            "e2e": {"value": r["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, batch, n):
    return {"workload": f"{args.config}.yaml refinement forward (DenseMotionNetwork + RaftFlow), {args.size}x{args.size}, "
                        f"batch {batch} pairs per GPU, random-init weights, synthetic key-points",
            "pairs_per_gpu": batch, "global_pairs": batch * n, "size": args.size, "parallelism": f"dp{n} (batch-sharded, no data-path collective)",
            "l2": "inputs and intermediates (>3 GB per step) exceed the 126 MB L2; no explicit flush",
            "convs": "cuDNN, TF32 allowed (PyTorch default), " + ("NCHW" if args.nchw else "channels_last (NHWC) memory")}


def same_run_parity(cfg, size, host, parity_in, P, use_bg):
    """BASELINE.md section 3 "parity in the same run": the first P pairs of the batch the GPU just timed (bench
    settings: TF32 convolutions, channels_last, cuDNN autotuning, full batch) against the CPU reference
    (oracle/_ref, else the oracle port) carrying the SAME weights, on the same inputs."""
    import torch
    kind, fwd, mods = build_cpu_nets(cfg, size)
    mods["dm"].load_state_dict(parity_in["sd_dm"])
    mods["rf"].load_state_dict(parity_in["sd_rf"])
    torch.set_num_threads(os.cpu_count() or 1)
    kp_s = {"kp": host["kp_s"][:P].clone(), "jacobian": host["jac_s"][:P].clone()}
    kp_d = {"kp": host["kp_d"][:P].clone(), "jacobian": host["jac_d"][:P].clone()}
    bg = host["bg"][:P].clone() if use_bg else None
    with torch.no_grad():
        ref = fwd(host["src"][:P].clone(), kp_s, kp_d, bg).double()
    got = parity_in["out"].double()
    err = (got - ref).abs()
    rel_l2 = float((got - ref).norm() / ref.norm())
    tol = 2e-2
    return {"against": kind, "pairs": P, "max_abs_err": float(err.max()), "mean_abs_err": float(err.mean()), "rel_l2": rel_l2,
            "ref_rms": float(ref.pow(2).mean().sqrt()), "tolerance_rel": tol,
            "ok": bool(rel_l2 < tol and float(err.max()) < tol * (1.0 + float(ref.abs().max()))),
            "settings": "same process, same weights (state_dict copied to the CPU modules), product under the timed settings"}


def measure_other_configs(args, world, rank, local, dev, res=None):
    """Short measurements (3 timed steps after 2 warm-up, CUDA events, max over ranks) of the BASELINE.json configs the
    headline line does not cover, at the same N: config 3 celebvhq 256x256 B=64 (background-affine path on), config 4
    the vox1 architecture at 512x512 swept over B, config 5 the vox1 training step B=16/GPU (fwd + bwd through the
    correlation lookup and the warps, Adam; DistributedDataParallel + SyncBatchNorm + NCCL gradient all-reduce when
    N > 1, train.py:37-48,64-77)."""
    import torch
    import torch.distributed as dist
    import mrfa_b200
    import synthetic_inputs as syn
    steps, warm = 3, 2
    res = {} if res is None else res                         # filled in place: the watchdog of run_ours reports what is there
    res.update({"steps": steps, "warmup": warm, "note": "device-resident inputs, ms = max over ranks"})

    def tmax(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def fwd_bench(cfg_name, size, batch):
        cfg = load_cfg(cfg_name)
        torch.manual_seed(0)
        dm = mrfa_b200.DenseMotionNetwork(**cfg["dense_motion"]).to(dev).eval().channels_last_()
        rf = mrfa_b200.RaftFlow(**dict(cfg["raft_flow"], size=size)).to(dev).eval().channels_last_()
        src = syn.frame_pairs(batch, size, seed=rank)[0].to(dev)
        ks, kd = ({k: v.to(dev) for k, v in d.items()} for d in syn.keypoints(batch, cfg["raft_flow"]["num_kp"], seed=rank))
        bg = syn.bg_affine(batch, seed=rank).to(dev) if cfg["train_params"]["bg_start"] == 0 else None

        def f():
            dense = dm(src, kd, ks, bg_param=bg)
            return rf(ks["kp"], kd["kp"], dense, img=dm.down(src), img_full=src)[0]
        with torch.no_grad():
            for _ in range(warm):
                f()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                o = f()
            e1.record()
            torch.cuda.synchronize()
        ms = tmax(e0.elapsed_time(e1) / steps)
        finite = bool(torch.isfinite(o).all())
        del dm, rf, o
        torch.cuda.empty_cache()
        return {"pairs_per_gpu": batch, "size": size, "ms_per_step": round(ms, 3), "pairs_per_s": round(batch * world / ms * 1e3, 1),
                "finite": finite}

    try:
        res["celebvhq_256_b64"] = dict(fwd_bench("celebvhq", 256, 64), config="celebvhq.yaml refinement forward, bg_param path on (bg_start: 0)")
    except Exception as e:                                   # keep the headline line even if a side config fails
        res["celebvhq_256_b64"] = {"error": repr(e)[:200]}
    sweep = []
    for b in (4, 8, 16):
        try:
            sweep.append(fwd_bench("vox1", 512, b))
        except Exception as e:
            sweep.append({"pairs_per_gpu": b, "size": 512, "error": repr(e)[:200]})
    res["vox1_512_sweep"] = sweep
    try:
        res["train_step_b16"] = train_step_bench(world, rank, local, dev, 16, 256, steps, warm)
    except Exception as e:
        res["train_step_b16"] = {"error": repr(e)[:200]}
    return res


def train_step_bench(world, rank, local, dev, batch, size, steps, warm):
    """Config 5: vox1 training step through the drop-in modules (L1 reconstruction loss, Adam(0.5, 0.999) as train.py:21-25).
    The perceptual (pretrained VGG19, needs the network) and equivariance losses of model.py:219-254 are outside the hot path."""
    import torch
    import torch.distributed as dist
    import mrfa_b200
    import synthetic_inputs as syn
    cfg = load_cfg("vox1")
    torch.manual_seed(0)

    class Refiner(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.dense_motion = mrfa_b200.DenseMotionNetwork(**cfg["dense_motion"])
            self.decoder = mrfa_b200.RaftFlow(**dict(cfg["raft_flow"], size=size))

        def forward(self, src, kp_s, kp_d):
            dense = self.dense_motion(src, kp_d, kp_s)
            return self.decoder(kp_s["kp"], kp_d["kp"], dense, img=self.dense_motion.down(src), img_full=src)[0]

    model = Refiner().to(dev).train()
    model.dense_motion.channels_last_()
    model.decoder.channels_last_()
    use_graph = os.environ.get("MRFA_TRAIN_GRAPH", "1") != "0"
    side = torch.cuda.Stream()
    if world > 1:
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)                     # train.py:43
        # train.py:45-48 passes find_unused_parameters=True; static_graph lets the reducer record the (fixed) set of unused
        # parameters once instead of walking the autograd graph every step, gradient_as_bucket_view drops the bucket copy.
        # Constructed on a side stream: the whole step is captured in a CUDA graph below.
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True,
                                                              static_graph=True, gradient_as_bucket_view=True)
        torch.cuda.current_stream().wait_stream(side)
    opt = torch.optim.Adam(model.parameters(), lr=2e-4, betas=(0.5, 0.999), capturable=use_graph)
    src, drv = (t.to(dev) for t in syn.frame_pairs(batch, size, seed=rank))
    # the key-points carry gradients like the jointly trained detector's outputs do (model.py:196-201): the backward runs
    # through the fused prior-motion / heat-map backward kernels
    kp_s, kp_d = ({k: v.to(dev).requires_grad_(True) for k, v in d.items()} for d in syn.keypoints(batch, 10, seed=rank))

    def clear_grads():
        opt.zero_grad(set_to_none=True)
        for d in (kp_s, kp_d):
            for v in d.values():
                v.grad = None

    def eager_step():
        clear_grads()
        loss = (model(src, kp_s, kp_d) - drv).abs().mean()
        loss.backward()
        opt.step()
        return loss

    step, graph, execution = eager_step, None, "eager"
    if use_graph:
        # B200-first: forward + backward + Adam and every NCCL collective of DDP / SyncBatchNorm captured ONCE in a CUDA graph
        # and replayed.  Eager SyncBatchNorm forces a host synchronisation per layer (the count mask of
        # torch/nn/modules/_functions.py, skipped under capture) and ~112 separately enqueued small collectives per step:
        # at N = 2 the eager step takes 123 ms against 96.8 ms on one GPU, the replayed graph 97.5 ms.
        ok = 1.0
        try:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(11):                       # DDP wants >= 11 eager iterations before a whole-network capture
                    eager_step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            clear_grads()
            with torch.cuda.graph(graph):
                static_loss = (model(src, kp_s, kp_d) - drv).abs().mean()
                static_loss.backward()
                opt.step()
        except Exception as e:                             # capture unsupported somewhere: every rank falls back together
            ok, graph = 0.0, None
            sys.stderr.write(f"train-step graph capture failed on rank {rank}: {e!r}\n")
        torch.cuda.synchronize()
        if world > 1:
            flag = torch.tensor([ok], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = float(flag[0])
        if ok > 0:
            def step():                                    # noqa: F811
                graph.replay()
                return static_loss
            execution = "one CUDA graph per step (forward + backward + Adam + NCCL collectives), replayed"
        else:
            graph = None

    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lt = loss.detach().clone()
        dist.reduce(lt, dst=0)                                                            # train.py:74-77
    ms = float(t[0])
    out = {"pairs_per_gpu": batch, "size": size, "ms_per_step": round(ms, 3), "pairs_per_s": round(batch * world / ms * 1e3, 1),
           "loss": float(loss.detach()), "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),
           "parallelism": f"DDP dp{world} + SyncBatchNorm, NCCL gradient all-reduce" if world > 1 else "single GPU (no collective)",
           "execution": execution}
    # share of NCCL / this library's / cuDNN kernels in one step (torch.profiler, rank 0; an extra untimed step)
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        tot = nccl = ours = 0.0
        for ev in prof.key_averages():
            us = float(getattr(ev, "device_time_total", 0.0) or getattr(ev, "cuda_time_total", 0.0))
            tot += us
            name = ev.key.lower()
            if "nccl" in name:
                nccl += us
            elif "mrfa::" in name or name.startswith("void mrfa"):
                ours += us
        if tot > 0:
            out.update({"kernel_time_ms": round(tot / 1e3, 3), "nccl_kernel_ms": round(nccl / 1e3, 3),
                        "nccl_share_of_kernel_time": round(nccl / tot, 4), "mrfa_kernel_ms": round(ours / 1e3, 3),
                        "limiting_collective": "ncclAllReduce of DDP gradient buckets (457 MB fp32 per step) + SyncBatchNorm all-gathers"
                        if world > 1 else None})
    except Exception as e:
        out["profile_error"] = repr(e)[:120]
    # release the captured NCCL nodes before anything tears the communicator down (destroying it under a live graph blocks)
    del step, loss
    if graph is not None:
        del static_loss
        graph = None
    import gc
    gc.collect()
    torch.cuda.synchronize()
    del model, opt
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    import mrfa_b200
    from mrfa_b200 import ops
    import synthetic_inputs as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: NCCL (version banner, NCCL_DEBUG lines) and other libraries write to file
    # descriptor 1 directly, so fd 1 is pointed at stderr for the whole run and the line goes to a saved copy of the real stdout
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = load_cfg(args.config)
    B, S, K = args.batch, args.size, cfg["raft_flow"]["num_kp"]
    use_bg = cfg["train_params"]["bg_start"] == 0

    torch.manual_seed(0)
    dm = mrfa_b200.DenseMotionNetwork(**cfg["dense_motion"]).to(dev).eval()
    rf = mrfa_b200.RaftFlow(**dict(cfg["raft_flow"], size=S)).to(dev).eval()
    if not args.nchw:
        dm.channels_last_()
        rf.channels_last_()

    # synthetic inputs: generated on the host (pinned), seeded per rank so ranks hold different pairs
    src_h, drv_h = syn.frame_pairs(B, S, seed=rank)
    kp_s_h, kp_d_h = syn.keypoints(B, K, seed=rank)
    host = {"src": src_h, "drv": drv_h, "kp_s": kp_s_h["kp"], "kp_d": kp_d_h["kp"], "jac_s": kp_s_h["jacobian"],
            "jac_d": kp_d_h["jacobian"]}
    if use_bg:
        host["bg"] = syn.bg_affine(B, seed=rank)
    host = {k: v.pin_memory() for k, v in host.items()}
    out_h = torch.empty((B, 3, S, S), dtype=torch.float32).pin_memory()

    def forward(d):
        kp_s = {"kp": d["kp_s"], "jacobian": d["jac_s"]}
        kp_d = {"kp": d["kp_d"], "jacobian": d["jac_d"]}
        dense = dm(d["src"], kp_d, kp_s, bg_param=d.get("bg"))
        out, _, _ = rf(kp_s["kp"], kp_d["kp"], dense, img=dm.down(d["src"]), img_full=d["src"])
        return out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    resident = {k: v.to(dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for k, v in host.items() if k != "drv")
    d2h_bytes = out_h.numel() * out_h.element_size()

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            out = forward(resident)
        sync_all()

        # ---- device-resident timing (value).  Inside the timed region only the hot-path kernels (SURVEY.md
        #      8(a), ~30 launches per step) are bracketed by CUDA events -- that is what `roofline` reports;
        #      every other launch of this library is counted, not timed, so the instrumentation does not
        #      inflate the step.  A second, fully instrumented pass of the same K steps fills kernels[]. ----
        sampler = ClockSampler(local)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ops.KernelTimer(only=HOT_PATH) as timer:
            sync_all()
            e0.record()
            for _ in range(args.steps):
                out = forward(resident)
            e1.record()
            sync_all()
        dev_ms = e0.elapsed_time(e1)
        clocks = sampler.stop()
        hot_kernels = timer.summary()
        launch_counts = dict(timer.counts)
        l1 = (out - resident["drv"]).abs().sum().double()
        e0b, e1b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ops.KernelTimer() as timer_all:
            e0b.record()
            for _ in range(args.steps):
                forward(resident)
            e1b.record()
            sync_all()
        instrumented_ms = e0b.elapsed_time(e1b)
        kernels = timer_all.summary()
        kernels.update(hot_kernels)                        # hot-path rows: the numbers of the timed region

        # ---- end to end: pinned host inputs -> H2D -> forward -> D2H of the predicted frames, every step.
        #      Three streams: the next step's inputs are uploaded and the previous step's frames downloaded
        #      while the current step computes (double-buffered pinned output) ----
        cur = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        out_hh = [out_h, torch.empty_like(out_h).pin_memory()]

        def stage():
            with torch.cuda.stream(s_in):
                d = {k: v.to(dev, non_blocking=True) for k, v in host.items() if k != "drv"}
                ev = torch.cuda.Event()
                ev.record(s_in)
            return d, ev

        def e2e_loop(n):
            d, ev = stage()
            for i in range(n):
                cur.wait_event(ev)
                for t in d.values():
                    t.record_stream(cur)
                nxt = stage() if i + 1 < n else None        # overlaps with this step's compute
                o = forward(d)
                done = torch.cuda.Event()
                done.record(cur)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(done)
                    out_hh[i & 1].copy_(o, non_blocking=True)
                o.record_stream(s_out)
                if nxt is not None:
                    d, ev = nxt
            cur.wait_stream(s_out)                          # the last download is inside the timed region

        e2e_loop(2)
        sync_all()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        e2e_loop(args.steps)
        e3.record()
        sync_all()
        e2e_ms = e2.elapsed_time(e3)

    # reconstruction-L1 / timing statistics: SUM of the L1 terms, MAX over ranks of the two elapsed times
    from mrfa_b200.dist import reduce_stats
    red = reduce_stats(float(l1), float(out.numel()), dev_ms, float(B * args.steps), device=dev, extra_max=(e2e_ms,))
    dev_ms, e2e_ms = red["elapsed_s"], red["extra_max"][0]
    recon_l1_mean = red["l1_mean"]

    # material for the same-run parity check (rank 0, evaluated after the GPU work): the first pairs of the timed batch
    P = max(0, min(args.parity_pairs, B))
    parity_in = None
    if rank == 0 and P > 0:
        parity_in = {"out": out[:P].float().cpu(), "sd_dm": {k: v.detach().cpu() for k, v in dm.state_dict().items()},
                     "sd_rf": {k: v.detach().cpu() for k, v in rf.state_dict().items()}}
    del out, resident
    if args.other_configs:
        del dm, rf
        torch.cuda.empty_cache()
    line = None
    if rank == 0:
        pk = peaks()
        pairs = B * world * args.steps
        step_ms = dev_ms / args.steps
        klist = []
        for name, k in sorted(kernels.items(), key=lambda kv: -kv[1]["total_ms"]):
            sec = k["total_ms"] / 1e3
            e = {"kernel": name, "launches": k["launches"], "total_ms": round(k["total_ms"], 4),
                 "share_of_step": round(k["total_ms"] / dev_ms, 4), "avg_ms": round(k["total_ms"] / max(1, k["calls"]), 5)}
            e["algorithmic_bytes_per_launch"] = round(k["bytes"] / max(1, k["calls"]))
            if sec > 0:
                e["hbm_gbs"] = round(k["bytes"] / sec / 1e9, 1)
                e["hbm_frac"] = round(k["bytes"] / sec / 1e9 / pk["hbm_gbs"], 4)
                if k["flops"]:
                    e["tflops"] = round(k["flops"] / sec / 1e12, 1)
                    e["tensor_frac"] = round(k["flops"] / sec / 1e12 / pk["bf16_tflops_sustained"], 4)
            klist.append(e)
        ours_ms = sum(k["total_ms"] for k in kernels.values())
        hot_ms = sum(k["total_ms"] for k in hot_kernels.values())
        # roofline: the dominant kernel of the hot path proper; the kernels of the
        # "next" rows 8(f) (fused elementwise helpers, small-channel convolution, cat) are listed in kernels[] only
        hot = [k for k in klist if k["kernel"] in HOT_PATH]
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        # the ncu capture was taken at the default workload; other shapes report traffic = null
        traffic_tbl = json.load(open(tpath)) if (os.path.exists(tpath) and args.batch == 64 and args.size == 256) else {}

        def roof(entry):
            if entry is None:
                return None
            t = traffic_tbl.get(entry["kernel"])
            per_step = max(1, entry["launches"] // max(1, args.steps))
            base = {"kernel": entry["kernel"], "launches_per_step": per_step, "avg_launch_ms": entry["avg_ms"],
                    "algorithmic_bytes_per_launch": entry["algorithmic_bytes_per_launch"],
                    "traffic": round(t / per_step) if t else None,
                    "traffic_scope": "average per launch: ncu dram__bytes_read+write summed over this kernel's launches in one step "
                                     "(B=64, 256x256; profiles/r2_hot_kernels_ncu.md), divided by launches_per_step" if t else None}
            if entry["kernel"] == "corr_volume":
                base.update({"bound": "tensor", "achieved": entry["tflops"], "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                             "frac": entry["tensor_frac"], "peak_source": pk["source"] + " (sustained bf16)",
                             "hbm_frac_of_output_bytes": entry["hbm_frac"]})
            else:
                base.update({"bound": "hbm", "achieved": entry["hbm_gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                             "frac": entry["hbm_frac"], "peak_source": pk["source"]})
            return base

        roofline = roof(hot[0] if hot else None)
        roofline_corr = roof(next((k for k in klist if k["kernel"] == "corr_volume"), None))

        line = {"metric": METRIC, "value": pairs / (dev_ms / 1e3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "fp32 warps/lookups + bf16 tensor-core correlation", "data": "synthetic",
                "config": workload_config(args, B, world), "clocks": clocks,
                "e2e": {"value": pairs / (e2e_ms / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": d2h_bytes},
                "gpu_launches": sum(launch_counts.values()),
                "roofline": roofline, "roofline_corr": roofline_corr, "kernels": klist,
                "kernels_note": "hot-path rows (SURVEY 8a) are timed inside the timed region; the other rows come from a "
                                f"second, fully instrumented pass of the same {args.steps} steps ({instrumented_ms / args.steps:.2f} ms/step)",
                "hot_path_share_of_step": round(ours_ms / dev_ms, 4),
                # two throughput tiers -- the SURVEY 8(a) kernels alone (K1-K9: correlation pack + volume, lookups,
                # feature warps, prior-motion synthesis, grids), summed from the CUDA events of the timed region on rank 0 and
                # scaled by the world size, and the end-to-end refinement forward (`value`, cuDNN convolutions included)
                "tiers": {"hot_path_only": {"ms_per_step": round(hot_ms / args.steps, 4),
                                            "pairs_per_s": round(B * world / (hot_ms / args.steps) * 1e3, 1) if hot_ms > 0 else None,
                                            "kernels": sorted(hot_kernels)},
                          "refinement_forward": {"ms_per_step": round(step_ms, 4), "pairs_per_s": round(pairs / (dev_ms / 1e3), 1)}},
                "recon_l1_mean": recon_l1_mean, "peaks": pk}
        if parity_in is not None:
            line["parity"] = same_run_parity(cfg, S, host, parity_in, P, use_bg)
            line["parity_max_err"], line["parity_rel_l2"] = line["parity"]["max_abs_err"], line["parity"]["rel_l2"]
        if not args.no_cpu_baseline and world == 1:           # reported at N = 1 only (torchrun pins OMP threads per rank)
            r = time_cpu_path(cfg, S, 1, 3, 1)
            line["cpu_baseline"] = {"value": r["value"], "unit": "pairs/s", "cores": r["cores"], "kind": r["kind"],
                                    "cpu_model": r["cpu_model"], "best": r["best_pairs_s"], "median": r["median_pairs_s"],
                                    "sample": f"1 pair per step, 3 timed steps after 1 warm-up ({r['ms_per_step']:.0f} ms/step), "
                                              f"{CPU_KIND_TEXT[r['kind']]}, {r['cores']} threads"}

    # ---- the other BASELINE.json configs, under a watchdog: whatever happens there (a stuck collective at some N), rank 0
    #      still prints the headline line with the side measurements finished so far
    emitted = threading.Event()

    def emit():
        if rank == 0 and not emitted.is_set():
            emitted.set()
            for _ in range(5):                      # the watchdog thread may serialise while the main thread still adds results
                try:
                    text = json.dumps(line)
                    break
                except RuntimeError:
                    time.sleep(0.05)
            print(text, file=real_stdout, flush=True)

    if args.other_configs:
        others = {}
        if rank == 0:
            line["other_configs"] = others

        def on_timeout():
            others["watchdog"] = f"other_configs did not finish within {args.other_timeout} s; partial results"
            emit()
            os._exit(0)

        dog = threading.Timer(args.other_timeout, on_timeout)
        dog.daemon = True
        dog.start()
        measure_other_configs(args, world, rank, local, dev, others)
        dog.cancel()
    emit()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
