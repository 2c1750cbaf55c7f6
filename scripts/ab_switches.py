"""Same-box A/B of the inference fast-path switches: runs bench.py once with everything on, then once per
MRFA_* switch turned off (all others on) and prints the step time of each run.

    python scripts/ab_switches.py [--steps 10] > gpurun_out/ab_switches.jsonl
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SWITCHES = [
    ("MRFA_FUSED_CARRY", "fused flow/occlusion level hand-over (mrfa_flow_carry)"),
    ("MRFA_SMALL_CONV", "tcgen05 TF32 kernel for the 7x7 small-channel convolutions"),
    ("MRFA_S2D_FINAL", "final 7x7 convolution as 3x3 over a 4x4 space-to-depth layout"),
    ("MRFA_HG_SUBPIXEL", "hourglass up-blocks as sub-pixel convolutions + shuffle-cat kernel"),
    ("MRFA_SPLIT_K", "update-block cat+conv pairs as two accumulating convolutions (no cat pass)"),
    ("MRFA_CAT_SLICES", "coarse warp / blend written into the decoder's cat buffers"),
    ("MRFA_FAST_CONV", "all conv-block fusions (BN folding, fused bias+ReLU, blends, ...)"),
]

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()


def run(env_off=None):
    env = dict(os.environ)
    if env_off:
        env[env_off] = "0"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--steps", str(a.steps)],
                         env=env, capture_output=True, text=True, timeout=600)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    return line["ms_per_step"], line["value"], line["clocks"]["sm_mhz"], line["gpu_launches"]


base = run()
print(json.dumps({"switch": "all on", "ms_per_step": round(base[0], 2), "pairs_per_s": round(base[1], 1), "sm_mhz": base[2],
                  "gpu_launches": base[3]}), flush=True)
for name, what in SWITCHES:
    r = run(name)
    print(json.dumps({"switch": name + "=0", "what": what, "ms_per_step": round(r[0], 2), "pairs_per_s": round(r[1], 1),
                      "delta_ms": round(r[0] - base[0], 2), "sm_mhz": r[2], "gpu_launches": r[3]}), flush=True)
again = run()
print(json.dumps({"switch": "all on (repeat)", "ms_per_step": round(again[0], 2), "pairs_per_s": round(again[1], 1),
                  "sm_mhz": again[2]}), flush=True)
