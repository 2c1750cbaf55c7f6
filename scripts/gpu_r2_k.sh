#!/bin/bash
# round 2, GPU call K: parity after the bias fusions (resize epilogue, pack kernel), quick bench line
mkdir -p gpurun_out/r2k
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2k/pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 --other-configs 0 --no-cpu-baseline > gpurun_out/r2k/bench_quick.json 2> gpurun_out/r2k/bench.err
tail -4 gpurun_out/r2k/pytest.txt
python - <<'PY'
import json
for l in open('gpurun_out/r2k/bench_quick.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity_rel_l2'), d.get('parity_max_err'))
        for k in d['kernels']: print(k['kernel'], k['launches'], k['total_ms'], k.get('hbm_frac'), k.get('tensor_frac'))
PY
