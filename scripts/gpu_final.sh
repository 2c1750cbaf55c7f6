mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/final_tests.log; tail -2 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | grep -E "smoke|Error|error" 
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/final_bench_n1.json)"
timeout 900 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/final_ref_n1.json 2> gpurun_out/final_ref_n1.err; echo "ref rc=$? lines=$(wc -l < gpurun_out/final_ref_n1.json)"
timeout 600 python scripts/bench_kernels.py --stock > gpurun_out/final_kernels.jsonl 2>/dev/null
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'dual_warp|grid_sample_fwd|corr_|conv7x7|cat2|resize_bilinear_nhwc_kernel|flow_carry|subpixel' -o /tmp/final_prof_hot python scripts/profile_step.py --batch 64 > gpurun_out/final_ncu.log 2>&1
ncu -i /tmp/final_prof_hot.ncu-rep --page raw --csv > gpurun_out/final_hot_raw.csv 2>/dev/null; python scripts/ncu_summary.py gpurun_out/final_hot_raw.csv gpurun_out/final_hot_kernels.md gpurun_out/final_traffic.json
ls -la /tmp/final_prof_hot.ncu-rep | awk '{print "ncu-rep bytes", $5}'
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv python scripts/profile_step.py --batch 64 > gpurun_out/final_ncu_launch.log 2>&1
timeout 300 python scripts/graph_probe.py 1 2>&1 | tail -2 | tee gpurun_out/final_graph_probe.log
timeout 300 python bench.py --size 512 --batch 8 --no-cpu-baseline > gpurun_out/final_bench_512.json 2>/dev/null; echo "512 rc=$?"
timeout 300 python bench.py --config celebvhq --no-cpu-baseline > gpurun_out/final_bench_celebvhq.json 2>/dev/null; echo "celebvhq rc=$?"
timeout 300 python scripts/train_step.py --batch 16 --steps 3 2>/dev/null | tail -1 | tee gpurun_out/final_train_step.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/final_bench_n1.json")); print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"]["value"], d["clocks"], "launches", d["gpu_launches"])
print(d["roofline"]); print(d["roofline_corr"])
r=json.load(open("gpurun_out/final_ref_n1.json")); print("ref", r["value"], r["cpu_baseline"]["sample"])
PY
