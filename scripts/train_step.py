"""Config 5: vox1 training step (fwd + bwd through the correlation lookup and the warps), batch
16 per GPU, Adam, L1 reconstruction loss; DistributedDataParallel + SyncBatchNorm when launched
with torchrun (train.py:37-48).  The perceptual (pretrained VGG19) and equivariance losses of
model.py:219-254 are outside the hot path and are not part of this step.

    python scripts/train_step.py [--batch 16] [--steps 5]
    python -m torch.distributed.run --nproc-per-node 8 ... scripts/train_step.py
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mrfa_b200                                    # noqa: E402
import synthetic_inputs as syn             # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--nchw", action="store_true", help="keep NCHW memory (default: channels_last)")
ap.add_argument("--cudnn-benchmark", type=int, default=1, help="torch.backends.cudnn.benchmark (algorithm autotuning)")
ap.add_argument("--profile", action="store_true", help="print the top CUDA kernels of one step (torch.profiler)")
ap.add_argument("--graph", action="store_true", help="capture forward + backward + optimizer step (incl. the NCCL collectives of DDP "
                "and SyncBatchNorm) in one CUDA graph and replay it")
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
torch.cuda.set_device(local)
torch.backends.cudnn.benchmark = bool(a.cudnn_benchmark)
dev = torch.device("cuda", local)
if world > 1:
    if a.graph:
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")      # whole-network capture with DDP (CUDA graphs notes)
    dist.init_process_group("nccl", device_id=dev)

cfg = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", "vox1.yaml")))
torch.manual_seed(0)


class Refiner(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.dense_motion = mrfa_b200.DenseMotionNetwork(**cfg["dense_motion"])
        self.decoder = mrfa_b200.RaftFlow(**dict(cfg["raft_flow"], size=a.size))

    def forward(self, src, kp_s, kp_d):
        dense = self.dense_motion(src, kp_d, kp_s)
        return self.decoder(kp_s["kp"], kp_d["kp"], dense, img=self.dense_motion.down(src), img_full=src)[0]


model = Refiner().to(dev).train()
if not a.nchw:
    model.dense_motion.channels_last_()
    model.decoder.channels_last_()
side = torch.cuda.Stream() if a.graph else None
if world > 1:
    model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    # whole-network capture wants DDP constructed (and warmed up) on a side stream
    with torch.cuda.stream(side) if a.graph else torch.cuda.stream(torch.cuda.current_stream()):
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True,
                                                          static_graph=True, gradient_as_bucket_view=True)
opt = torch.optim.Adam(model.parameters(), lr=2e-4, betas=(0.5, 0.999), capturable=a.graph)
src, drv = (t.to(dev) for t in syn.frame_pairs(a.batch, a.size, seed=rank))
kp_s, kp_d = syn.keypoints(a.batch, 10, seed=rank)
kp_s = {k: v.to(dev).requires_grad_(True) for k, v in kp_s.items()}     # live key-point gradients (model.py:196-201)
kp_d = {k: v.to(dev).requires_grad_(True) for k, v in kp_d.items()}


def step():
    opt.zero_grad(set_to_none=True)
    for d in (kp_s, kp_d):
        for v in d.values():
            v.grad = None
    out = model(src, kp_s, kp_d)
    loss = (out - drv).abs().mean()
    loss.backward()
    opt.step()
    return loss


if a.graph:
    # >= 11 eager DDP iterations on the side stream, then ONE capture of the whole step.  Under capture SyncBatchNorm skips
    # its count-mask host synchronisation (torch/nn/modules/_functions.py) and every collective becomes a graph node.
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(a.warmup, 11)):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    opt.zero_grad(set_to_none=True)
    for d in (kp_s, kp_d):
        for v in d.values():
            v.grad = None
    with torch.cuda.graph(graph):
        static_loss = (model(src, kp_s, kp_d) - drv).abs().mean()
        static_loss.backward()
        opt.step()

    def step():                                       # noqa: F811  (replay replaces the eager step)
        graph.replay()
        return static_loss
for _ in range(a.warmup):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if a.profile and rank == 0:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70), file=sys.stderr)
if rank == 0:
    print(json.dumps({"workload": "vox1 training step (L1 loss, Adam), fwd+bwd", "pairs_per_gpu": a.batch, "n_gpus": world,
                      "ms_per_step": float(ms), "pairs_per_s": a.batch * world / float(ms) * 1e3, "loss": float(loss.detach()),
                      "peak_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}))
if a.graph:
    # release the captured NCCL nodes before the communicator goes away (destroying it under a live graph blocks)
    import gc
    del graph, static_loss, step, loss
    gc.collect()
    torch.cuda.synchronize()
if world > 1:
    import faulthandler
    faulthandler.dump_traceback_later(60, exit=True)          # a stuck teardown must not hold the GPUs
    dist.barrier()
    dist.destroy_process_group()
