timeout 300 python -m pytest tests/test_gpu_campplus.py -x -q 2>&1 | tail -5
python tools/stem_timeline.py | sed -n 1,12p
NSEG=16384 timeout 120 python tools/sweep_chunks.py 8192/2048
SPK_NO_STEM_FUSE=1 NSEG=16384 timeout 120 python tools/sweep_chunks.py 8192/2048
