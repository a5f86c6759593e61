for prio in 0 1; do for cp in 0 10000000 7500000 5000000; do
PB_STREAM_PRIO=$prio PB_CHUNK_POINTS=$cp python bench.py --no-cpu-baseline --steps 4 > gpurun_out/bench_p${prio}_c${cp}.json 2>gpurun_out/bench_p.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_p${prio}_c${cp}.json')); print('prio',$prio,'chunkpts',$cp, round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['counters']['chunks'])
PY
done; done
