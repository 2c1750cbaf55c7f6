timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "warp or channels_last" 2>&1 | tail -2
for u in 1 2 4; do echo "unroll=$u"; MRFA_WARP_UNROLL=$u timeout 300 python scripts/bench_kernels.py --only warp 2>/dev/null | grep -E "nhwc C=(256|128|64) "; done
