#!/bin/bash
# usage: tools/ncu_quick.sh <kernel-regex> <out-csv> -- <command...>   (run under gpurun, 1 GPU)
# NCU_SKIP / NCU_COUNT (default 3 / 2) select the launches.
# A cheap ncu pass: duration, issue rate, top stall reasons, DRAM bytes for a few launches after warm-up.
K="$1"; OUT="$2"; shift 3
M=gpu__time_duration.sum,smsp__issue_active.avg.per_cycle_active,smsp__warps_active.avg.per_cycle_active,smsp__warps_eligible.avg.per_cycle_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem
for r in no_instruction long_scoreboard wait math_pipe_throttle not_selected dispatch_stall short_scoreboard mio_throttle lg_throttle barrier branch_resolving; do
  M="$M,smsp__average_warps_issue_stalled_${r}_per_issue_active.ratio"
done
ncu --metrics "$M" --clock-control none -k "regex:$K" -s ${NCU_SKIP:-3} -c ${NCU_COUNT:-2} --csv --log-file "$OUT" "$@"
