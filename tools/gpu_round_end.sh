#!/bin/bash
# One gpurun call that mirrors the driver's round-end checks on ONE GPU: smoke, the GPU test suite, the default bench.
# (The reference arm -- bench.py --impl reference -- times one full config-4 step of the reference: ~10 min of one host core;
#  run it outside gpurun, it needs no GPU.)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
log() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/round_end.log; }
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a gpurun_out/round_end.log
log "pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/pytest_gpu_final.log 2>&1
log "pytest exit $? : $(tail -1 gpurun_out/pytest_gpu_final.log)"
log "bench"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
log "bench exit $? $(cut -c1-200 gpurun_out/bench_final.json)"
timeout 100 python bench.py --workload config3 --no-cpu --no-api-default --steps 20 > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.err
log "config3 exit $? $(cut -c1-200 gpurun_out/bench_config3.json)"
log "done"
