#!/bin/bash
# GPU call C: parity subset of the current default, bench of both pipelines, ncu launch list + full capture of k_cull.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
log() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/call_c.log; }
log "parity subset"
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "phase_parity or whole_step or pipelines_agree or sliced or config5" > gpurun_out/pytest_parity_c.log 2>&1
log "parity exit $? : $(tail -1 gpurun_out/pytest_parity_c.log)"
timeout 100 python tools/ab_bench.py cur:1 cur:0 2>&1 | tee -a gpurun_out/call_c.log
log "ncu launch list"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 150 ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/launches_c.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-spin > gpurun_out/ncu_list.log 2>&1
log "ncu list exit $?"
timeout 100 ncu --set full --clock-control none --import-source on -k regex:k_cull -s 1 -c 1 -o gpurun_out/k_cull_full python bench.py --steps 1 --warmup 3 --no-cpu --no-spin > gpurun_out/ncu_full.log 2>&1
log "ncu full exit $?"
log "done"
