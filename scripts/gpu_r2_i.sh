#!/bin/bash
# round 2, GPU call I: parity of the re-written streaming kernels, stock / same-output baselines, launch list of one step
mkdir -p gpurun_out/r2i
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_refine.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2i/pytest.txt
timeout 600 python scripts/bench_kernels.py --only corr --stock > gpurun_out/r2i/k_corr_stock.jsonl 2> gpurun_out/r2i/k.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2i/launches.csv python scripts/profile_step.py > gpurun_out/r2i/launches.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --other-configs 0 --no-cpu-baseline --parity-pairs 0 > gpurun_out/r2i/bench_quick.json 2> gpurun_out/r2i/bench.err
tail -4 gpurun_out/r2i/pytest.txt
grep -E "corr_volume|cublas|einsum|stock" gpurun_out/r2i/k_corr_stock.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r2i/bench_quick.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print(d['value'], d['ms_per_step'], d['e2e']['value'])
        for k in d['kernels']: print(k['kernel'], k['launches'], k['total_ms'], k.get('hbm_frac'), k.get('tensor_frac'))
PY
