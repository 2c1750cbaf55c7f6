mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/t16.log; tail -4 gpurun_out/t16.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench16_fast.json 2> gpurun_out/bench16.err
MRFA_FAST_CONV=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench16_plain.json 2>> gpurun_out/bench16.err
python - <<'PY'
import json
for f in ("bench16_fast","bench16_plain"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "hot share", d["hot_path_share_of_step"], "l1", d["recon_l1_mean"])
        for k in d["kernels"][:8]: print("   ", k)
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/bench16.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches16_fast.csv python scripts/profile_step.py --batch 64 > gpurun_out/ncu_launch16.log 2>&1
