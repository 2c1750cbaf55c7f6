timeout 300 python scripts/bench_kernels.py --only corr --stock 2>/dev/null | grep -E "corr_volume|cublas|einsum|pack"
