python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -8
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/bench_models.py eres eres_w24 2>&1 | tail -2
