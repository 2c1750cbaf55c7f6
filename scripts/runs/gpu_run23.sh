mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/t23.log; tail -3 gpurun_out/t23.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench23.json 2> gpurun_out/bench23.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench23.json")); print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "hot", d["hot_path_share_of_step"], d["clocks"])
for k in d["kernels"][:12]: print("   ", k)
PY
tail -3 gpurun_out/bench23.err
