mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'nhwc_kernel<|lookup_fwd|pack_nhwc' -c 40 -o gpurun_out/prof18 python scripts/bench_kernels.py --only "warp,corr" --iters 1 > gpurun_out/ncu18.log 2>&1
tail -2 gpurun_out/ncu18.log
