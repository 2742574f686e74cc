set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1l.json 2> gpurun_out/bench_r1l.err; tail -2 gpurun_out/bench_r1l.err
