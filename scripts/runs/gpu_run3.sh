mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "corr" 2>&1 | tail -30 > gpurun_out/t3_corr.log
tail -4 gpurun_out/t3_corr.log
timeout 300 python scripts/bench_kernels.py --stock > gpurun_out/kern_tma.jsonl 2> gpurun_out/kern_tma.err
MRFA_CORR_EPILOGUE=direct timeout 300 python scripts/bench_kernels.py --only corr_v > gpurun_out/kern_direct.jsonl 2>> gpurun_out/kern_tma.err
cat gpurun_out/kern_tma.jsonl gpurun_out/kern_direct.jsonl; tail -5 gpurun_out/kern_tma.err
