N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/run_c4_sharded.py 4500 > gpurun_out/c4_sharded_$N.json 2> gpurun_out/c4_sharded_$N.err; tail -1 gpurun_out/c4_sharded_$N.json | cut -c1-700; tail -3 gpurun_out/c4_sharded_$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu_final.json 2> gpurun_out/bench_${N}gpu_final.err; python -c "
import json;d=json.load(open('gpurun_out/bench_${N}gpu_final.json'));print('bench n_gpus',d['n_gpus'],'value',round(d['value'],2),'e2e',round(d['e2e']['value'],2))"
