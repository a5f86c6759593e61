python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r01_s3_n1.json 2> gpurun_out/bench_r01_s3_n1.err; tail -c 1500 gpurun_out/bench_r01_s3_n1.json; tail -3 gpurun_out/bench_r01_s3_n1.err
