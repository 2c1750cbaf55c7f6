#!/bin/bash
# round 2, GPU call D: full parity suite, A/B of the epilogue conversion / occupancy cap / backward pre-reduction, bench line
mkdir -p gpurun_out/r2d
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2d/pytest.txt
for cvt in 1 0; do
  MRFA_CORR_CVT=$cvt timeout 300 python scripts/bench_kernels.py --only corr_volume > gpurun_out/r2d/k_corr_cvt$cvt.jsonl 2> gpurun_out/r2d/k_corr.err
done
MRFA_CORR_CVT=1 MRFA_CORR_STORE=0 timeout 300 python scripts/bench_kernels.py --only corr_volume > gpurun_out/r2d/k_corr_cvt1_store0.jsonl 2>> gpurun_out/r2d/k_corr.err
timeout 300 python scripts/bench_kernels.py --only corr > gpurun_out/r2d/k_corr_all.jsonl 2>> gpurun_out/r2d/k_corr.err
timeout 300 python scripts/bench_kernels.py --only warp > gpurun_out/r2d/k_warp_occ4.jsonl 2> gpurun_out/r2d/k_warp.err
MRFA_WARP_OCC4=0 timeout 300 python scripts/bench_kernels.py --only warp > gpurun_out/r2d/k_warp_occ3.jsonl 2>> gpurun_out/r2d/k_warp.err
timeout 300 python scripts/bench_kernels.py --only bwd --stock > gpurun_out/r2d/k_bwd_run1.jsonl 2> gpurun_out/r2d/k_bwd.err
MRFA_BWD_RUN=0 timeout 300 python scripts/bench_kernels.py --only bwd > gpurun_out/r2d/k_bwd_run0.jsonl 2>> gpurun_out/r2d/k_bwd.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d/bench.json 2> gpurun_out/r2d/bench.err
echo "bench rc=$?" >> gpurun_out/r2d/bench.err
tail -8 gpurun_out/r2d/pytest.txt
cat gpurun_out/r2d/k_corr_cvt*.jsonl | grep corr_volume
