for c in 64 128 256; do for dbg in 0 1 2 3; do
MRFA_CORR_DEBUG=$dbg timeout 120 python scripts/bench_kernels.py --only corr_volume --C $c 2>/dev/null | grep corr_volume | sed "s/^/C=$c /"
done; done
