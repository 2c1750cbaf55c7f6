#!/bin/bash
# round 2, GPU call J (2 GPUs): the driver's N>1 launch of bench.py incl. other_configs (DDP + SyncBatchNorm training step)
mkdir -p gpurun_out/r2j
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2j/bench_n2.json 2> gpurun_out/r2j/bench_n2.err
echo "rc=$?" >> gpurun_out/r2j/bench_n2.err
tail -c 600 gpurun_out/r2j/bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/r2j/bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])
        print(json.dumps(d.get('other_configs'))[:1500])
PY
