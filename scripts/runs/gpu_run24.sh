mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches24.csv python scripts/profile_step.py --batch 64 > gpurun_out/ncu_launch24.log 2>&1
