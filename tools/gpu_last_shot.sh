#!/bin/bash
# One short gpurun call (the round's last GPU seconds): (A) the GPU tests that have not run since they were written or
# changed, on the in-tree library; (B) parity of the tuning variant `opt` (tools/build_variants.py) -- both at once, they
# are bound by the oracle on the host cores; then (C) A/B bench of the in-tree library against `opt` and against a run
# without phase marks; (D) smoke().  Every leg has its own timeout and log under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
log() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/shot.log; }
log "start"
( timeout 80 python -m pytest tests/test_gpu_vs_reference.py tests/test_gpu_host_cpp.py tests/test_gpu_fullsize.py -q -x -s -m gpu \
    -k "cloth_spheres or cpp or config3" > gpurun_out/shot_A.log 2>&1
  log "A exit $? : $(tail -1 gpurun_out/shot_A.log)" ) &
( CLSN_LIB=$PWD/collision_b200/variants/libclsn_opt.so timeout 80 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -q -x -s -m gpu \
    -k "(pinned_pass and C0) or (whole_step and (two_sheets or mixed or ball_plane or cloth_spheres)) or (phase_parity and (mixed or string_string)) or empty_and_tiny" \
    > gpurun_out/shot_B.log 2>&1
  log "B exit $? : $(tail -1 gpurun_out/shot_B.log)" ) &
wait
timeout 50 python tools/ab_bench.py --steps 10 cur opt > gpurun_out/shot_C.log 2>&1
log "C exit $? : $(cut -c1-150 gpurun_out/shot_C.log | tr '\n' '|')"
timeout 25 python bench.py --steps 10 --warmup 3 --no-cpu --no-api-default --no-phase-marks > gpurun_out/shot_C_nomarks.json 2> gpurun_out/shot_C_nomarks.err
log "C2 exit $? : $(cut -c1-200 gpurun_out/shot_C_nomarks.json)"
timeout 25 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/shot_D.log 2>&1
log "D exit $? : $(tail -1 gpurun_out/shot_D.log)"
log "done"
