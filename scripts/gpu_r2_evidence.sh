#!/bin/bash
# round 2 evidence run (final code): parity suite, smoke, both bench arms, isolated kernel timings (+ stock / library baselines),
# ncu summary of this library's kernels inside one step (+ DRAM traffic), ncu rows of the backward kernels, launch list,
# power trace of the correlation GEMM, store-order replay.  Everything lands in gpurun_out/r2final; what is judged is copied
# to profiles/ by hand afterwards.
O=gpurun_out/r2final
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > $O/tests.log; tail -2 $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | grep -E "smoke|Error|error" | tee $O/smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$? lines=$(wc -l < $O/bench_n1.json)"
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "ref rc=$?"
timeout 900 python scripts/bench_kernels.py --stock > $O/bench_kernels.jsonl 2> $O/bench_kernels.err
timeout 120 scripts/micro/store_order > $O/store_order.txt 2>&1
timeout 120 python scripts/clock_probe.py > $O/corr_power_trace.txt 2>&1
K='dual_warp|grid_sample|corr_|cast_bf16|conv7x7|resize_bilinear|flow_carry|flow_update|subpixel|occlusion_blend|channel_affine|avg_pool|antialias|dense_motion_prior|kp2gaussian|coords_grid|prior_to_flow'
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum"
timeout 1500 ncu --profile-from-start off --clock-control none $SEC -k regex:"$K" -f -o /tmp/r2_prof_hot python scripts/profile_step.py --batch 64 > $O/ncu.log 2>&1
ncu -i /tmp/r2_prof_hot.ncu-rep --page raw --csv > $O/hot_raw.csv 2>/dev/null
python scripts/ncu_summary.py $O/hot_raw.csv $O/hot_kernels_ncu.md $O/traffic.json
# backward kernels (training step, B = 16): correlation backward, warp backward, lookup backward, prior-motion backward
timeout 900 ncu --clock-control none $SEC -k regex:"bwd|transpose_bf16" -f -o /tmp/r2_prof_bwd python scripts/ncu_targets.py --only bwd --reps 1 > $O/ncu_bwd.log 2>&1
ncu -i /tmp/r2_prof_bwd.ncu-rep --page raw --csv > $O/bwd_raw.csv 2>/dev/null
python scripts/ncu_summary.py $O/bwd_raw.csv $O/bwd_kernels_ncu.md $O/bwd_traffic.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python scripts/profile_step.py --batch 64 > $O/ncu_launch.log 2>&1
python - <<'PY'
import json
O="gpurun_out/r2final/"
d=json.loads([l for l in open(O+"bench_n1.json") if l.startswith("{")][-1]); print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"]["value"], d["clocks"], "launches", d["gpu_launches"], "parity", d.get("parity_rel_l2"))
print(d["roofline"]); print(d["roofline_corr"]); print(d["tiers"])
r=json.loads([l for l in open(O+"bench_reference_arm.json") if l.startswith("{")][-1]); print("ref", r["value"], r["cpu_baseline"]["sample"])
PY
rm -f $O/ncu_corr_volume.ncu-rep
ls -la $O; du -sh $O
