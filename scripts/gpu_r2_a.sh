#!/bin/bash
# round 2, GPU call A: parity tests, isolated kernel timings (A/B of the warp variants), one full bench line
mkdir -p gpurun_out/r2a
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2a/pytest.txt
echo "pytest rc=$?" >> gpurun_out/r2a/pytest.txt
timeout 300 python scripts/bench_kernels.py --only corr --stock > gpurun_out/r2a/k_corr.jsonl 2> gpurun_out/r2a/k_corr.err
timeout 300 python scripts/bench_kernels.py --only warp > gpurun_out/r2a/k_warp_run8.jsonl 2> gpurun_out/r2a/k_warp.err
MRFA_WARP_VEC=4 timeout 300 python scripts/bench_kernels.py --only warp > gpurun_out/r2a/k_warp_run4.jsonl 2>> gpurun_out/r2a/k_warp.err
MRFA_WARP_RUN=0 timeout 300 python scripts/bench_kernels.py --only warp > gpurun_out/r2a/k_warp_plain.jsonl 2>> gpurun_out/r2a/k_warp.err
timeout 300 python scripts/bench_kernels.py --only warp --flow random > gpurun_out/r2a/k_warp_run8_randomflow.jsonl 2>> gpurun_out/r2a/k_warp.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err
echo "bench rc=$?" >> gpurun_out/r2a/bench.err
tail -c 600 gpurun_out/r2a/pytest.txt
head -c 1500 gpurun_out/r2a/bench.json
