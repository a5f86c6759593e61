python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_r01_final2_n1.json 2> gpurun_out/bench_r01_final2_n1.err; tail -c 200 gpurun_out/bench_r01_final2_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_v8.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --dropin-calls 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_degree|k_nn|k_union|k_label|k_hp_cells|k_centres' -s 6 -c 8 -o gpurun_out/prof_v8 -f python tools/profile_step.py --scenes 312 --steps 2 > gpurun_out/prof_v8.log 2>&1
tail -2 gpurun_out/prof_v8.log
