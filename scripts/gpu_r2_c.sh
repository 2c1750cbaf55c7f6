#!/bin/bash
# round 2, GPU call C: run-walk test diagnostics, store-mode A/B of the correlation GEMM, warp timings, ncu captures
mkdir -p gpurun_out/r2c
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "run_walk" 2>&1 | grep -E "run-walk|passed|failed|Error|assert" | head -40 > gpurun_out/r2c/pytest_runwalk.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2c/pytest.txt
for mode in 2 0; do
  MRFA_CORR_STORE=$mode timeout 300 python scripts/bench_kernels.py --only corr_volume > gpurun_out/r2c/k_corr_store$mode.jsonl 2> gpurun_out/r2c/k_corr.err
done
timeout 300 python scripts/bench_kernels.py --only warp > gpurun_out/r2c/k_warp.jsonl 2> gpurun_out/r2c/k_warp.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:corr_lookup_fwd_tiled -s 1 -c 1 -f -o gpurun_out/r2c/ncu_lookup python scripts/ncu_targets.py --only lookup > gpurun_out/r2c/ncu_lookup.log 2>&1
timeout 600 $NCU -k regex:dual_warp_fwd_nhwc_run -s 1 -c 1 -f -o gpurun_out/r2c/ncu_dual_warp python scripts/ncu_targets.py --only warp > gpurun_out/r2c/ncu_warp.log 2>&1
timeout 600 $NCU -k regex:corr_volume_tma -s 1 -c 1 -f -o gpurun_out/r2c/ncu_corr python scripts/ncu_targets.py --only corr > gpurun_out/r2c/ncu_corr.log 2>&1
ls -la gpurun_out/r2c
cat gpurun_out/r2c/pytest_runwalk.txt
tail -5 gpurun_out/r2c/pytest.txt
cat gpurun_out/r2c/k_corr_store*.jsonl
