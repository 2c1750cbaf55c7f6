mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/t11.log; tail -6 gpurun_out/t11.log
timeout 300 python scripts/bench_kernels.py > gpurun_out/kern11.jsonl 2>gpurun_out/kern11.err; cat gpurun_out/kern11.jsonl; tail -3 gpurun_out/kern11.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench11_cl.json 2> gpurun_out/bench11.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --nchw > gpurun_out/bench11_nchw.json 2>> gpurun_out/bench11.err
python - <<'PY'
import json
for f in ("bench11_cl","bench11_nchw"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "hot share", d["hot_path_share_of_step"])
        for k in d["kernels"][:6]: print("   ", k)
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/bench11.err
