#!/bin/bash
# GPU call B: parity of the default build (batched cull, aggregated contact counter) incl. all three pipelines,
# then A/B of pipelines and tuning variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
BUDGET=${BUDGET:-330}
left() { echo $(( BUDGET - ($(date +%s) - T0) )); }
log() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/call_b.log; }
log "parity subset"
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q --durations=8 -k "phase_parity or whole_step or pipelines_agree or sliced or config5" > gpurun_out/pytest_parity_b.log 2>&1
log "parity exit $? : $(tail -1 gpurun_out/pytest_parity_b.log)"
for v in cur:0 cur:1 cur:2 cb0:0 pm:0 s2:0 r2:0 r5:0 fa5:2 fa3:2 ex5:2 c5:0; do
  if [ $(left) -gt 30 ]; then timeout 100 python tools/ab_bench.py $v 2>&1 | tee -a gpurun_out/call_b.log; fi
done
log "done"
