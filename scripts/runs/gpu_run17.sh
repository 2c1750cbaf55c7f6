timeout 600 python -m pytest tests/test_gpu_refine.py -q -m gpu -k "training" 2>&1 | tail -30
