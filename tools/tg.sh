python tools/time_conv.py
B=163 H=10 W=38 C=208 python tools/time_conv.py
B=163 H=40 W=149 C=128 CO=256 KS=1 ST=2 python tools/time_conv.py
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/bench_models.py eres eres_w24 ecapa 2>&1 | tail -3
