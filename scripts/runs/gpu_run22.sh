mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/t22.log; tail -3 gpurun_out/t22.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench22_vox1.json 2> gpurun_out/bench22.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --config celebvhq > gpurun_out/bench22_celebvhq.json 2>> gpurun_out/bench22.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --size 512 --batch 8 > gpurun_out/bench22_512.json 2>> gpurun_out/bench22.err
timeout 600 python scripts/train_step.py --batch 16 --steps 3 > gpurun_out/train22.json 2>> gpurun_out/bench22.err
python - <<'PY'
import json
for f in ("bench22_vox1","bench22_celebvhq","bench22_512"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "hot", d["hot_path_share_of_step"], "cpu", d.get("cpu_baseline",{}).get("value"), d["clocks"])
        for k in d["kernels"][:4]: print("   ", k)
    except Exception as e: print(f, "ERR", e)
print(open("gpurun_out/train22.json").read())
PY
tail -5 gpurun_out/bench22.err
