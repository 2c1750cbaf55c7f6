mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "warp or channels_last" 2>&1 | tail -3
timeout 300 python scripts/bench_kernels.py --only warp 2>/dev/null | grep nhwc
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench14_n2.json 2> gpurun_out/bench14_n2.err
python -c "
import json
d=json.load(open('gpurun_out/bench14_n2.json')); print('N=2 value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'n_gpus', d['n_gpus'], 'l1', d['recon_l1_mean'], d['clocks'])
" || tail -20 gpurun_out/bench14_n2.err
