#!/bin/bash
# The round's last GPU seconds: the default bench line of the final library, then a few quick GPU tests on it.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
log() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/shot2.log; }
log "start"
timeout 30 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err
log "bench exit $? : $(cut -c1-220 gpurun_out/final_bench_1gpu.json)"
timeout 25 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "empty_and_tiny or sliced_equals_whole or determinism or (whole_step and sheet_wall)" > gpurun_out/shot2_tests.log 2>&1
log "tests exit $? : $(tail -1 gpurun_out/shot2_tests.log)"
log "done"
