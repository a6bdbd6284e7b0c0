#!/bin/bash
# Runs the end-to-end parity cases, one process per case, logging to gpurun_out/e2e_<case>.log.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CASES=${@:-"unet_forward ddim50 ddpm_noise timing"}
rc=0
for c in $CASES; do
  timeout 600 python tests/e2e_cases.py $c > gpurun_out/e2e_$c.log 2>&1
  r=$?
  echo "case $c exit $r"
  grep -E '^\{|^===' gpurun_out/e2e_$c.log | cut -c1-300 | tail -n 12
  [ $r -ne 0 ] && { rc=1; tail -n 15 gpurun_out/e2e_$c.log | cut -c1-300; }
done
exit $rc
