mkdir -p gpurun_out
for dbg in 0 1 2 3 4 5 7; do
MRFA_CORR_DEBUG=$dbg timeout 120 python scripts/bench_kernels.py --only corr_volume 2>/dev/null | grep corr_volume
done
MRFA_CORR_EPILOGUE=direct timeout 120 python scripts/bench_kernels.py --only corr_volume 2>/dev/null | grep corr_volume
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_volume_tma -c 1 -o gpurun_out/prof_corr_tma python scripts/bench_kernels.py --only corr_volume --iters 1 > gpurun_out/ncu_corr.log 2>&1
tail -3 gpurun_out/ncu_corr.log
