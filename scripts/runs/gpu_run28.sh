MRFA_CORR_VARIANT=5 timeout -s KILL 120 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "corr" -x 2>&1 | tail -5
MRFA_CORR_VARIANT=5 timeout -s KILL 120 python scripts/bench_kernels.py --only corr_volume 2>/dev/null | grep corr_volume | sed "s/^/2sm /"
timeout -s KILL 120 python scripts/bench_kernels.py --only corr_volume 2>/dev/null | grep corr_volume | sed "s/^/1sm /"
MRFA_CORR_VARIANT=5 timeout -s KILL 120 python scripts/bench_kernels.py --only corr_volume --size 512 --batch 8 2>/dev/null | grep corr_volume | sed "s/^/2sm-512 /"
MRFA_CORR_VARIANT=5 timeout -s KILL 120 python scripts/bench_kernels.py --only corr_volume --batch 1 2>/dev/null | grep corr_volume | sed "s/^/2sm-b1 /"
timeout 300 python scripts/bench_kernels.py --only warp 2>/dev/null | grep -E "nhwc C=(128|64) "
