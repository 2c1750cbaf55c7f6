for dbg in 0 8 10 16 24 26; do
MRFA_CORR_DEBUG=$dbg timeout 120 python scripts/bench_kernels.py --only corr_volume 2>/dev/null | grep corr_volume
done
