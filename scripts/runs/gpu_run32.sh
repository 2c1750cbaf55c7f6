mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench32_n4.json 2> gpurun_out/bench32_n4.err
python -c "
import json
d=json.load(open('gpurun_out/bench32_n4.json')); print('N=4 value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['clocks'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 scripts/train_step.py --batch 16 --steps 3 2> gpurun_out/train32_n4.err | tail -1
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=1 value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))"
tail -3 gpurun_out/train32_n4.err
