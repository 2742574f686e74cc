python tools/e2e_trace.py 600 2>&1 | tail -2
for c in 1 2 3; do python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-contexts $c > gpurun_out/bench_r1z_$c.json 2>gpurun_out/bench_r1z.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r1z_$c.json'));print($c,'value',round(d['value'],2),'e2e',round(d['e2e']['value'],2))"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_full.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1_full_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_analyze|k_parse_blocks" -s 6 -c 2 -o gpurun_out/prof_full_r1 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1_full_b.log 2>&1
tail -1 gpurun_out/ncu_r1_full_b.log | cut -c1-200
