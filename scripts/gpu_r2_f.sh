#!/bin/bash
# round 2, GPU call F: store-order micro-benchmark, epilogue-variant check, ncu of the current lookup + "next"-row kernels (CSV only)
mkdir -p gpurun_out/r2f
timeout 120 scripts/micro/store_order > gpurun_out/r2f/store_order.txt 2>&1
timeout 300 python scripts/diag/epilogue_variants.py > gpurun_out/r2f/epilogue_variants.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:corr_lookup_fwd_tiled -s 1 -c 1 -f -o gpurun_out/r2f/ncu_lookup python scripts/ncu_targets.py --only lookup > gpurun_out/r2f/ncu_lookup.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"resize_bilinear|conv7x7_small|subpixel_shuffle|antialias|occlusion_blend|flow_carry|dense_motion_prior|kp2gaussian|corr_lookup|corr_pack|cast_bf16" -f -o /tmp/ncu_next python scripts/profile_step.py > gpurun_out/r2f/ncu_next.log 2>&1
ncu -i /tmp/ncu_next.ncu-rep --page raw --csv > gpurun_out/r2f/ncu_next_raw.csv 2>/dev/null
ls -la gpurun_out/r2f
cat gpurun_out/r2f/store_order.txt gpurun_out/r2f/epilogue_variants.txt
