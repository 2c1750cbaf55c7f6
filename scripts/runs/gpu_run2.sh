mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/t_all.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err
timeout 300 python bench.py --steps 2 --warmup 1 --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b64.csv python scripts/profile_step.py --batch 64 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'corr_|dual_warp|grid_sample|dense_motion|kp2gaussian' -o gpurun_out/prof_hot_b64 python scripts/profile_step.py --batch 64 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/t_all.log; cat gpurun_out/bench_b64.json | head -c 3000; tail -3 gpurun_out/bench_b64.err; cat gpurun_out/bench_ref.json; ls -la gpurun_out
