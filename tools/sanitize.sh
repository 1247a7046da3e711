#!/bin/bash
# compute-sanitizer over the whole hot path: ONE tool per gpurun call (B200_PROFILING.md).
#   tools/sanitize.sh memcheck | racecheck | synccheck | initcheck
# Workload: two full MoE train steps per detector at E=3 / B=24 (every kernel on the path: router, TMA-fed tcgen05 GEMM pair
# kernels, strip / gather variants, norms, fused discriminator, fp32 convs, loss tails, Adam) checked against the golden
# vectors while the tool watches, plus batch inference.  Output: gpurun_out/sanitize_<tool>.log (+ a one-line summary).
cd "$(dirname "$0")/.." || exit 1
TOOL=${1:-memcheck}
O=gpurun_out
mkdir -p $O
# the same command must have exited 0 without the tool first
python -m pytest tests/test_step_gpu.py -q -p no:cacheprovider -k "E3_B24_golden or generate_matches or generate_neutron" > $O/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_step_gpu.py -q -p no:cacheprovider -k "E3_B24_golden or generate_matches or generate_neutron" > $O/sanitize_$TOOL.log 2>&1
rc=$?
echo "compute-sanitizer --tool $TOOL rc=$rc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/sanitize_$TOOL.log | tail -1) | $(grep -E 'passed|failed' $O/sanitize_$TOOL.log | tail -1)"
exit 0
