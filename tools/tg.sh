python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for cfg in "64 128 0 0" "64 128 0 1" "128 64 0 0" "128 256 0 1" "32 112 0 0" "512 128 1 0" "256 128 1 0"; do
  set -- $cfg
  K=$1 N=$2 PRO=$3 RES=$4 B=2048 W=1900 python tools/time_gemm.py
done
python tools/bench_models.py eres eres_w24 ecapa 2>&1 | tail -3
