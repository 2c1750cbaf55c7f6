for cb in 1 0; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --cudnn-benchmark $cb 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cudnn.benchmark=$cb value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))"
done
