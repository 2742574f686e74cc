#!/usr/bin/env bash
# k_analyze under ncu on ONE section type of the bench signal (tools/section_timing.py --only=S), for one or more
# builds of the library: instruction count, issue utilisation, ALU pipe, stall reasons.
# usage (on the GPU box): bash tools/ncu_sections.sh lib.so [lib2.so ...]      sections: 0 AR noise, 3 stepped noise
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
for lib in "${@:-lossless-audio-codec_b200/liblac_b200.so}"; do
 for s in 0 3; do
  echo "== $lib section $s"
  ncu --metrics $M --clock-control none -k regex:k_analyze -s 1 -c 1 --csv python tools/section_timing.py --only=$s $lib 2>/dev/null | grep -E "k_analyze" | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"'
 done
done
