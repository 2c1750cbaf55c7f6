#!/bin/bash
# round 2, GPU call H: parity suite + bench line + ncu of the "next"-row kernels after the index-math / tiling changes
mkdir -p gpurun_out/r2h
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2h/pytest.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h/bench.json 2> gpurun_out/r2h/bench.err
echo "bench rc=$?" >> gpurun_out/r2h/bench.err
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"resize_bilinear|conv7x7_small|subpixel_shuffle|antialias|occlusion_blend|flow_carry|dense_motion_prior|kp2gaussian|corr_lookup|corr_pack|cast_bf16|avg_pool|channel_affine" -f -o /tmp/ncu_next python scripts/profile_step.py > gpurun_out/r2h/ncu_next.log 2>&1
ncu -i /tmp/ncu_next.ncu-rep --page raw --csv > gpurun_out/r2h/ncu_next_raw.csv 2>/dev/null
tail -6 gpurun_out/r2h/pytest.txt
python - <<'PY'
import json
for l in open('gpurun_out/r2h/bench.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity',{}).get('rel_l2'))
        for k in d['kernels']: print(k['kernel'], k['launches'], k['total_ms'], k.get('hbm_frac'), k.get('tensor_frac'))
PY
