timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "corr" 2>&1 | tail -3
timeout 300 python scripts/bench_kernels.py --only corr 2>/dev/null | grep -E "pack|lookup"
