timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
PB_CHUNK_POINTS=40000000 timeout 300 python tools/profile_step.py --scenes 312 --steps 3 2>&1 | tail -2 | head -1
