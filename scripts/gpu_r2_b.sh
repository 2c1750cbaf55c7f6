#!/bin/bash
# round 2, GPU call B: parity tests (per-tap lookup fractions, tcgen05 correlation backward), lookup / warp timings
mkdir -p gpurun_out/r2b
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2b/pytest.txt
echo "pytest rc=$?" >> gpurun_out/r2b/pytest.txt
timeout 300 python scripts/bench_kernels.py --only corr > gpurun_out/r2b/k_corr.jsonl 2> gpurun_out/r2b/k_corr.err
timeout 300 python scripts/bench_kernels.py --only warp > gpurun_out/r2b/k_warp.jsonl 2> gpurun_out/r2b/k_warp.err
MRFA_WARP_PW=32 timeout 300 python scripts/bench_kernels.py --only warp > gpurun_out/r2b/k_warp_pw32.jsonl 2>> gpurun_out/r2b/k_warp.err
MRFA_WARP_PW=8 timeout 300 python scripts/bench_kernels.py --only warp > gpurun_out/r2b/k_warp_pw8.jsonl 2>> gpurun_out/r2b/k_warp.err
timeout 600 python bench.py --steps 10 --warmup 3 --other-configs 0 > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err
echo "bench rc=$?" >> gpurun_out/r2b/bench.err
tail -c 1500 gpurun_out/r2b/pytest.txt
