#!/bin/bash
# round 2: ncu summary of this library's kernels inside one step (sections instead of --set full: same columns, fewer replay passes)
O=gpurun_out/r2final
mkdir -p $O
K='dual_warp|grid_sample|corr_|cast_bf16|conv7x7|resize_bilinear|flow_carry|flow_update|subpixel|occlusion_blend|channel_affine|avg_pool|antialias|dense_motion_prior|kp2gaussian|coords_grid|prior_to_flow'
timeout 1500 ncu --profile-from-start off --clock-control none --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:"$K" -f -o /tmp/r2_prof_hot python scripts/profile_step.py --batch 64 > $O/ncu.log 2>&1
ncu -i /tmp/r2_prof_hot.ncu-rep --page raw --csv > $O/hot_raw.csv 2>/dev/null
python scripts/ncu_summary.py $O/hot_raw.csv $O/hot_kernels_ncu.md $O/traffic.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_volume_tma -s 1 -c 1 -f -o $O/ncu_corr_volume python scripts/ncu_targets.py --only corr > $O/ncu_corr.log 2>&1
ls -la $O | tail -8; head -30 $O/hot_kernels_ncu.md
