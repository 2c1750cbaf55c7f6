mkdir -p gpurun_out
timeout 120 python scripts/bench_kernels.py --only membw 2>/dev/null
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "corr" 2>&1 | tail -5
for dbg in 0 1 2 3; do
MRFA_CORR_DEBUG=$dbg timeout 120 python scripts/bench_kernels.py --only corr_volume 2>/dev/null | grep corr_volume
done
MRFA_CORR_VARIANT=1 timeout 120 python scripts/bench_kernels.py --only corr_volume 2>/dev/null | grep corr_volume
timeout 120 python scripts/bench_kernels.py --only corr_volume --size 512 --batch 8 2>/dev/null | grep corr_volume
timeout 120 python scripts/bench_kernels.py --only corr_volume --batch 1 2>/dev/null | grep corr_volume
