"""Does capturing the refinement forward in a CUDA graph help (launch gaps / CPU-bound stretches)?"""
import os, sys, time
import torch, yaml
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mrfa_b200
import synthetic_inputs as syn
cfg = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", "vox1.yaml")))
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dm = mrfa_b200.DenseMotionNetwork(**cfg["dense_motion"]).to(dev).eval().channels_last_()
rf = mrfa_b200.RaftFlow(**cfg["raft_flow"]).to(dev).eval().channels_last_()
src, _ = syn.frame_pairs(B, 256)
kp_s, kp_d = syn.keypoints(B, 10)
src = src.to(dev); kp_s = {k: v.to(dev) for k, v in kp_s.items()}; kp_d = {k: v.to(dev) for k, v in kp_d.items()}
def step():
    dense = dm(src, kp_d, kp_s)
    return rf(kp_s["kp"], kp_d["kp"], dense, img=dm.down(src), img_full=src)[0]
with torch.no_grad():
    for _ in range(3): out = step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): out = step()
    t_cpu = (time.perf_counter() - t0) / 5
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 5
    print(f"eager: CPU enqueue {1e3*t_cpu:.1f} ms/step, wall {1e3*t_all:.1f} ms/step")
    ref = out.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out_g = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2): g.replay()
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"graph replay: {e0.elapsed_time(e1)/10:.2f} ms/step; max|graph - eager| = {(out_g - ref).abs().max().item():.3e}")
