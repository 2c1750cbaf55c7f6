#!/bin/bash
# round 2, GPU call G: warp-autonomous channels-last lookup kernel: parity, timing A/B against the block kernel, ncu
mkdir -p gpurun_out/r2g
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "lookup or corr" 2>&1 | tail -15 > gpurun_out/r2g/pytest.txt
timeout 300 python scripts/bench_kernels.py --only corr_lookup > gpurun_out/r2g/k_lookup_warp.jsonl 2> gpurun_out/r2g/k.err
MRFA_LOOKUP_WARP=0 timeout 300 python scripts/bench_kernels.py --only corr_lookup > gpurun_out/r2g/k_lookup_block.jsonl 2>> gpurun_out/r2g/k.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:corr_lookup_fwd_tiled_nhwc -s 1 -c 1 -f -o gpurun_out/r2g/ncu_lookup python scripts/ncu_targets.py --only lookup > gpurun_out/r2g/ncu_lookup.log 2>&1
cat gpurun_out/r2g/pytest.txt; cat gpurun_out/r2g/k_lookup_warp.jsonl; echo; cat gpurun_out/r2g/k_lookup_block.jsonl; tail -3 gpurun_out/r2g/k.err
