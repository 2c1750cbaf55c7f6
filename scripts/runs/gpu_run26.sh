mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'dual_warp|grid_sample_fwd|corr_' -o gpurun_out/prof26_hot python scripts/profile_step.py --batch 64 > gpurun_out/ncu26.log 2>&1
tail -2 gpurun_out/ncu26.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches26.csv python scripts/profile_step.py --batch 64 > gpurun_out/ncu_launch26.log 2>&1
