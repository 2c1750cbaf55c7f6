timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "warp or channels_last" 2>&1 | tail -2
timeout 300 python scripts/bench_kernels.py --only warp 2>/dev/null | grep -E "dual_warp_fwd.nhwc"
