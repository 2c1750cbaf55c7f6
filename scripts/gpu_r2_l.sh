#!/bin/bash
# round 2, GPU call L: parity after the prior / heat-map / flow-update changes, quick bench line, launch list
mkdir -p gpurun_out/r2l
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2l/pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 --other-configs 0 --no-cpu-baseline > gpurun_out/r2l/bench_quick.json 2> gpurun_out/r2l/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2l/launches.csv python scripts/profile_step.py > gpurun_out/r2l/launches.log 2>&1
tail -4 gpurun_out/r2l/pytest.txt
python - <<'PY'
import json
for l in open('gpurun_out/r2l/bench_quick.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity_rel_l2'), d.get('parity_max_err'), d['gpu_launches'])
        for k in d['kernels']: print(k['kernel'], k['launches'], k['total_ms'], k.get('hbm_frac'), k.get('tensor_frac'))
PY
