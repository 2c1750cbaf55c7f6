timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "corr" 2>&1 | tail -5
for st in 1 0; do for dbg in 0 1 3; do
MRFA_CORR_STORE=$st MRFA_CORR_DEBUG=$dbg timeout 120 python scripts/bench_kernels.py --only corr_volume 2>/dev/null | grep corr_volume | sed "s/^/cluster2 store=$st /"
done; done
MRFA_CORR_VARIANT=2 timeout 120 python scripts/bench_kernels.py --only corr_volume 2>/dev/null | grep corr_volume | sed "s/^/nocluster /"
timeout 120 python scripts/bench_kernels.py --only corr_volume --size 512 --batch 8 2>/dev/null | grep corr_volume
timeout 120 python scripts/bench_kernels.py --only corr_volume --batch 1 2>/dev/null | grep corr_volume
timeout 120 python scripts/bench_kernels.py --only corr_volume --batch 8 2>/dev/null | grep corr_volume
