timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "corr or warp" 2>&1 | tail -3
timeout 120 python scripts/bench_kernels.py --only corr_volume 2>/dev/null | grep corr_volume | sed "s/^/v3 /"
MRFA_CORR_VARIANT=4 timeout 120 python scripts/bench_kernels.py --only corr_volume 2>/dev/null | grep corr_volume | sed "s/^/v3-cluster2 /"
MRFA_CORR_VARIANT=2 timeout 120 python scripts/bench_kernels.py --only corr_volume 2>/dev/null | grep corr_volume | sed "s/^/v2 /"
timeout 120 python scripts/bench_kernels.py --only corr_volume --size 512 --batch 8 2>/dev/null | grep corr_volume | sed "s/^/v3-512 /"
timeout 120 python scripts/bench_kernels.py --only corr_volume --batch 1 2>/dev/null | grep corr_volume | sed "s/^/v3-b1 /"
timeout 300 python scripts/bench_kernels.py --only warp 2>/dev/null | grep -E "dual_warp_fwd.nhwc"
