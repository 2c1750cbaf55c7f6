mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/t13.log; tail -3 gpurun_out/t13.log
timeout 300 python scripts/bench_kernels.py --only warp,corr_pack > gpurun_out/kern13.jsonl 2>gpurun_out/kern13.err; grep -E "nhwc|pack" gpurun_out/kern13.jsonl; tail -3 gpurun_out/kern13.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches13_cl.csv python scripts/profile_step.py --batch 64 > gpurun_out/ncu_launch13.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench13_cl.json 2> gpurun_out/bench13.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench13_cl.json")); print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "hot share", d["hot_path_share_of_step"])
for k in d["kernels"][:6]: print("   ", k)
PY
