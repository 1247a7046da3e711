#!/bin/bash
# Runs the per-kernel GPU parity tests in separate processes (a faulting kernel poisons its CUDA context, so
# each family gets its own interpreter) and leaves logs under gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
nvidia-smi > gpurun_out/smi.txt 2>&1
(ls /root/reference; ls baseline/_ref) > gpurun_out/ref_probe.txt 2>&1
i=0
rc_all=0
for k in "router or gather" "hinge or loss_tails" "igemm_fwd and simt" "igemm_fwd and not simt and not fullsize" \
         "fullsize" "igemm_wgrad and simt" "igemm_wgrad and not simt" "dense_dgrad" \
         "pack or fc1 or gn_lrelu or ln_lrelu or gen_out" \
         "conv2d or groupnorm or layernorm or maxpool or linear or spectral or elementwise or expm1"; do
  i=$((i+1))
  timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$k" -p no:cacheprovider > gpurun_out/kt_$i.log 2>&1
  rc=$?
  echo "[$i] '$k' -> rc=$rc : $(tail -1 gpurun_out/kt_$i.log)"
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
