set -x
for v in NOLEVELS NOWALK NOBIAS NOKSER NOPREP; do python tools/time_variant.py build/var/lib_$v.so; done > gpurun_out/variants_r1k.log 2>&1
python tools/time_variant.py lossless-audio-codec_b200/liblac_b200.so >> gpurun_out/variants_r1k.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1k.csv python bench.py --steps 2 --warmup 1 --seconds 120 --no-cpu-baseline > gpurun_out/ncu_r1k_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_analyze -s 2 -c 1 -o gpurun_out/prof_an_r1k -f python bench.py --steps 1 --warmup 1 --seconds 120 --no-cpu-baseline > gpurun_out/ncu_r1k_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_parse_blocks -s 1 -c 1 -o gpurun_out/prof_parse_r1k -f python bench.py --steps 1 --warmup 1 --seconds 120 --no-cpu-baseline > gpurun_out/ncu_r1k_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_restore_blocks|k_autocorr" -s 3 -c 3 -o gpurun_out/prof_misc_r1k -f python bench.py --steps 1 --warmup 1 --seconds 120 --no-cpu-baseline > gpurun_out/ncu_r1k_d.log 2>&1
cat gpurun_out/variants_r1k.log
