#!/bin/bash
# round 2: one-launch GroupNorm adjoint for small images: parity, training e2e, A/B of the training step; then the
# block-output fusion policy A/B of the sampling path
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tests/kernel_cases.py groupnorm_bwd > gpurun_out/r2s_k_gnbwd.log 2>&1; echo "kernel case exit $?"
grep -E '"ok": false|PASS|FAIL|rror' gpurun_out/r2s_k_gnbwd.log | cut -c1-400 | tail -n 8
for c in train_step train_step_pesser train_step_adm train_multi_step; do
  timeout 900 python tests/e2e_cases.py $c > gpurun_out/r2s_e2e_$c.log 2>&1; echo "$c exit $?"
  grep -E '^\{|^===' gpurun_out/r2s_e2e_$c.log | cut -c1-200 | tail -n 3
done
for v in 0 1 0 1; do
  B200_GNB_SLAB=$v timeout 600 python tools/bench_train.py cfg 128 10 > gpurun_out/r2s_train_$v.json 2> gpurun_out/r2s_train_$v.err
  python - $v <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/r2s_train_{sys.argv[1]}.json'))
print('GNB_SLAB', sys.argv[1], round(d['ms_per_step'],3), d['kernels_per_step'], d['kernels']['groupnorm_bwd'])
PY
done
bash tools/r2r_policy.sh
