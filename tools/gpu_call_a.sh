#!/bin/bash
# One GPU call: parity of the fused pipeline, A/B of the pipelines, full-size parity, tuning variants.
# Every step is bounded; steps are skipped once the call's time budget is used up.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
BUDGET=${BUDGET:-480}
left() { echo $(( BUDGET - ($(date +%s) - T0) )); }
log() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/call_a.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | tee -a gpurun_out/call_a.log
log "parity (fused pipeline)"
CLSN_PIPELINE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q --durations=12 > gpurun_out/pytest_parity_p1.log 2>&1
log "parity exit $? : $(tail -1 gpurun_out/pytest_parity_p1.log)"
log "bench staged / fused"
timeout 200 python tools/ab_bench.py cur:0 cur:1 2>&1 | tee -a gpurun_out/call_a.log
if [ $(left) -gt 200 ]; then
  log "full-size parity (fused pipeline)"
  CLSN_PIPELINE=1 timeout 240 python -m pytest tests/test_gpu_fullsize.py -x -q --durations=8 -k "sample_of_config4 or config3_full or deterministic" > gpurun_out/pytest_fullsize_p1.log 2>&1
  log "fullsize exit $? : $(tail -1 gpurun_out/pytest_fullsize_p1.log)"
fi
for v in f4:1 c5:1 c3:1 g32:1 t256:1 t64:1 f2:1 g8:1; do
  if [ $(left) -gt 45 ]; then timeout 100 python tools/ab_bench.py $v 2>&1 | tee -a gpurun_out/call_a.log; fi
done
log "done"
