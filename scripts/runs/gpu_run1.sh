mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "not corr" 2>&1 | tail -80 > gpurun_out/t_nocorr.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "corr" 2>&1 | tail -120 > gpurun_out/t_corr.log
timeout 300 python -m pytest tests/test_gpu_refine.py -q -m gpu 2>&1 | tail -80 > gpurun_out/t_refine.log
timeout 200 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
tail -5 gpurun_out/t_nocorr.log gpurun_out/t_corr.log gpurun_out/t_refine.log gpurun_out/smoke.log
