python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -x -q 2>&1 | tail -3
python tools/bench_small_calls.py 2>&1 | tail -8
python bench.py --no-cpu-baseline --steps 4 > gpurun_out/bench_t.json 2>gpurun_out/bench_t.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_t.json')); print(round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['stage_ms']['degree'], d['e2e_dropin'])
PY
