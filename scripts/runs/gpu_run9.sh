for ns in 1 2 4 8 16; do
MRFA_CORR_NSPLIT=$ns timeout 120 python scripts/bench_kernels.py --only corr_volume 2>/dev/null | grep corr_volume | sed "s/^/cluster2 nsplit=$ns /"
MRFA_CORR_VARIANT=2 MRFA_CORR_NSPLIT=$ns timeout 120 python scripts/bench_kernels.py --only corr_volume 2>/dev/null | grep corr_volume | sed "s/^/nocluster nsplit=$ns /"
done
