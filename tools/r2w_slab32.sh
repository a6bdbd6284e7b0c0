#!/bin/bash
# round 2: GroupNorm adjoint experiments: parity, then same-box A/B of the training step over the env settings in $AB (default:
# reverse walk of the second pass on / off)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tests/kernel_cases.py groupnorm_bwd > gpurun_out/r2w_k_gnbwd.log 2>&1; echo "kernel case groupnorm_bwd rc=$?"
grep -E '"ok": false|mismatch": [1-9]|exception' gpurun_out/r2w_k_gnbwd.log | cut -c1-300 | head -n 6
for c in ${E_CASES-train_step}; do
  timeout 900 python tests/e2e_cases.py $c > gpurun_out/r2w_e2e_$c.log 2>&1; echo "e2e case $c rc=$?"
  grep -E '^===' gpurun_out/r2w_e2e_$c.log | cut -c1-330 | tail -n 6
done
for rep in 1 2; do
for cfg in ${AB:-B200_GNB_REVERSE=0 B200_NOP=1}; do
  env $(echo $cfg | tr ',' ' ') python tools/bench_train.py cfg 128 20 > gpurun_out/r2w_train.json 2> gpurun_out/r2w_train.err
  python - $cfg <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2w_train.json').read().strip().splitlines()[-1])
print(sys.argv[1].ljust(28), round(d['ms_per_step'],3), 'ms', d['kernels_per_step'], 'launches', {k:round(v['ms'],3) for k,v in d['kernels'].items() if k in ('groupnorm_bwd','groupnorm_apply','conv_gemm','conv_wgrad','grad_cast')})
PY
done
done
