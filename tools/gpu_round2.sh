python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu_r1b.json 2> gpurun_out/bench_2gpu_r1b.err
tail -1 gpurun_out/bench_2gpu_r1b.json | cut -c1-400
python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -2
python bench.py --impl reference --gpus 1 --steps 1 --warmup 0 | cut -c1-600
