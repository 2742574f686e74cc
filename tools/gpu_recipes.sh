#!/usr/bin/env bash
# tools/gpu_recipes.sh -- the exact commands behind profiles/ (run ON the GPU box, from the repo root).
#
# From the build container:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_recipes.sh tests bench launches'
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_recipes.sh ncu-analyze'
# Every recipe writes into gpurun_out/<tag>_*; copy what should be judged into profiles/.
# Numbers printed by a run under ncu are never bench values: `bench` always runs un-profiled first.
set -u
TAG=${TAG:-r2}
OUT=gpurun_out
mkdir -p $OUT
BENCH_ARGS=${BENCH_ARGS:---steps 5 --warmup 3}

for recipe in "$@"; do
  case $recipe in
    tests)        # GPU parity suite (all through the C ABI)
      python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_gpu_tests.log ;;
    smoke)
      python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log ;;
    bench)        # the un-profiled bench line
      python bench.py $BENCH_ARGS > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench.json ;;
    bench-ref)    # the reference arm on the box's host cores
      python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "bench-ref rc=$?"; cat $OUT/${TAG}_bench_ref.json ;;
    launches)     # ncu launch list of the same command (per-launch times: cold cache, serialised)
      ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
          python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1; echo "launches rc=$?" ;;
    ncu-analyze)  # one full-size k_analyze launch, full set + source lines
      ncu --set full --import-source on --clock-control none -k regex:k_analyze -s 2 -c 1 -f -o $OUT/${TAG}_analyze \
          python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_analyze.log 2>&1; echo "ncu-analyze rc=$?" ;;
    ncu-parse)    # one full-size k_parse_blocks launch
      ncu --set full --import-source on --clock-control none -k regex:k_parse_blocks -s 2 -c 1 -f -o $OUT/${TAG}_parse \
          python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_parse.log 2>&1; echo "ncu-parse rc=$?" ;;
    ncu-all)      # one launch of every kernel of a step, full set (DRAM traffic per kernel)
      ncu --set full --clock-control none -s 39 -c 13 -f -o $OUT/${TAG}_step \
          python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_step.log 2>&1; echo "ncu-all rc=$?" ;;
    configs)      # all BASELINE configs at full size on one GPU
      python tools/run_configs.py > $OUT/${TAG}_all_configs.jsonl 2> $OUT/${TAG}_all_configs.err; echo "configs rc=$?"; cat $OUT/${TAG}_all_configs.jsonl ;;
    cli)          # lac_cli wall times beside the reference CLI
      python tools/cli_timing.py > $OUT/${TAG}_cli_timing.jsonl 2> $OUT/${TAG}_cli_timing.err; echo "cli rc=$?"; cat $OUT/${TAG}_cli_timing.jsonl ;;
    configs-host) # the BASELINE configs through the host-buffer C ABI (page-locked caller buffers)
      CUDA_DEVICE_MAX_CONNECTIONS=32 python tools/run_configs_host.py > $OUT/${TAG}_all_configs_host.jsonl 2> $OUT/${TAG}_all_configs_host.err; echo "configs-host rc=$?"; cat $OUT/${TAG}_all_configs_host.jsonl ;;
    trace)        # device timeline of every slice of one host encode + decode of the bench workload
      CUDA_DEVICE_MAX_CONNECTIONS=32 LACB_TRACE=1 python tools/e2e_trace.py > $OUT/${TAG}_e2e_trace.txt 2>&1; echo "trace rc=$?"; grep "slice .*:\|slice .* done\|^it" $OUT/${TAG}_e2e_trace.txt | tail -24 ;;
    ncu-probe)    # the stereo-probe analysis kernel (first k_analyze launch of an auto-stereo encode); LACB_PROBE_GANG=1 gives the one-warp-CTA form
      ncu --set full --import-source on --clock-control none -k regex:k_analyze -c 1 -f -o $OUT/${TAG}_probe \
          python tools/run_configs.py --only "24/192k" > $OUT/${TAG}_ncu_probe.log 2>&1; echo "ncu-probe rc=$?"
      ncu -i $OUT/${TAG}_probe.ncu-rep --page raw --csv > $OUT/${TAG}_probe_raw.csv ;;
    ncu-others)   # one full-size launch each of the parser, the emitter, the restore and the autocorrelation kernels
      ncu --set full --clock-control none -k regex:"k_parse_blocks|k_emit|k_restore_blocks|k_autocorr_stream" -c 4 -f -o $OUT/${TAG}_others \
          python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_others.log 2>&1; echo "ncu-others rc=$?"
      ncu -i $OUT/${TAG}_others.ncu-rep --page raw --csv > $OUT/${TAG}_others_raw.csv ;;
    phase)        # clock64 phase counters of k_analyze (-DLACB_PHASE_CLK build made by `make -C lossless-audio-codec_b200 phase`)
      python tools/phase_clk.py lossless-audio-codec_b200/build/liblac_b200_phase.so > $OUT/${TAG}_phase_clocks.txt 2>&1; echo "phase rc=$?"; cat $OUT/${TAG}_phase_clocks.txt ;;
    *) echo "unknown recipe $recipe"; exit 2 ;;
  esac
done
