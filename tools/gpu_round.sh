python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2a.json 2>gpurun_out/bench_r2a.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2a.json'));print('value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),d['stage_ms_per_step'])"
