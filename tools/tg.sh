CHUNK=256 python tools/bench_models.py eres eres_w24 2>&1 | tail -2
BATCH_SCALE=2 CHUNK=512 python tools/bench_models.py eres 2>&1 | tail -1
BATCH_SCALE=2 CHUNK=256 python tools/bench_models.py eres_w24 ecapa 2>&1 | tail -2
BATCH_SCALE=4 CHUNK=1024 python tools/bench_models.py eres 2>&1 | tail -1
BATCH_SCALE=4 CHUNK=512 python tools/bench_models.py eres_w24 ecapa 2>&1 | tail -2
