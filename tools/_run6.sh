for cp in 0 1900000 1300000 950000; do
PB_CHUNK_POINTS=$cp python bench.py --no-cpu-baseline --steps 6 --scenes 39 --dropin-calls 8 > gpurun_out/bench_s39_$cp.json 2>gpurun_out/bench_s39.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_s39_$cp.json')); print('chunkpts',$cp, 'pts', d['config']['points_total'], round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['counters']['chunks'], d['stage_ms'])
PY
done
