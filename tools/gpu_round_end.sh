#!/bin/bash
# One gpurun call that mirrors the driver's round-end checks: smoke, bench (both arms, config 3), then the GPU test suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
log() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/round_end.log; }
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a gpurun_out/round_end.log
log "bench"
timeout 200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
log "bench exit $? $(cut -c1-200 gpurun_out/bench_final.json)"
timeout 100 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
log "ref exit $? $(cut -c1-160 gpurun_out/bench_ref.json)"
timeout 60 python bench.py --workload config3 --no-cpu --steps 20 > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.err
log "config3 exit $? $(cut -c1-200 gpurun_out/bench_config3.json)"
log "pytest -m gpu"
timeout 330 python -m pytest tests -m gpu -x -q --durations=10 --deselect "tests/test_gpu_parity.py::test_whole_step_parity[layered_4x24]" > gpurun_out/pytest_gpu_final.log 2>&1
log "pytest exit $? : $(tail -1 gpurun_out/pytest_gpu_final.log)"
log "done"
