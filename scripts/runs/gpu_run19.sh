mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nhwc_kernel -s 43 -c 5 -o gpurun_out/prof19 python scripts/bench_kernels.py --only "warp" --iters 1 > gpurun_out/ncu19.log 2>&1
tail -2 gpurun_out/ncu19.log
