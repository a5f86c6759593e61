python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
python tools/bench_small_calls.py 2>&1 | tail -3
PB_CHUNK_POINTS=40000000 python tools/profile_step.py --scenes 312 --steps 3 2>&1 | tail -3
