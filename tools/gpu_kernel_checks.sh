#!/bin/bash
# Runs every per-kernel parity case in its own process (a faulting kernel cannot poison the others),
# logging to gpurun_out/kernels_<case>.log.  Usage: tools/gpu_kernel_checks.sh [case ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
CASES=${@:-"misc groupnorm conv_basic conv_epilogue conv_n256 conv_small_hw conv_1x1 conv_shortcut conv_stride2 conv_lastconv conv_up2 conv_tproj attention sampler gemm wgrad groupnorm_bwd backward_misc optimizer ode_samplers"}
rc=0
for c in $CASES; do
  timeout 300 python tests/kernel_cases.py $c > gpurun_out/kernels_$c.log 2>&1
  r=$?
  echo "case $c exit $r"
  tail -n 3 gpurun_out/kernels_$c.log | cut -c1-400
  [ $r -ne 0 ] && rc=1
done
exit $rc
