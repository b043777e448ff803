B=163 H=20 W=75 C=512 CO=1024 KS=3 ST=2 python tools/time_conv.py
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/bench_models.py eres eres_w24 ecapa 2>&1 | tail -3
bash tools/quick_bench.sh 2>&1 | tail -1
