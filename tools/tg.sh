python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -6
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/bench_models.py ecapa 2>&1 | tail -1
