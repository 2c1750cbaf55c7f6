mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nhwc -s 35 -c 13 -o gpurun_out/prof_warp_nhwc python scripts/bench_kernels.py --only warp --iters 1 > gpurun_out/ncu_warp.log 2>&1
tail -3 gpurun_out/ncu_warp.log
