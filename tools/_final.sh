python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_r01_v9_n1.json 2> gpurun_out/bench_r01_v9_n1.err; tail -c 200 gpurun_out/bench_r01_v9_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_v9.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --dropin-calls 4 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
