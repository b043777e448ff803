python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/time_batch64.py 4096
SPK_GRAPH_MAX_BATCH=0 python tools/time_batch64.py 4096
python tools/bench_models.py eres eres_w24 ecapa 2>&1 | tail -3
