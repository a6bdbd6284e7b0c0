#!/bin/bash
# round 2: one-launch GroupNorm adjoint for 32x32 images: parity, training parity (B = 4 and the benchmarked B = 128), same-box A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tests/kernel_cases.py groupnorm_bwd > gpurun_out/r2w_k_gnbwd.log 2>&1; echo "kernel case groupnorm_bwd rc=$?"
grep -E '"ok": false|mismatch": [1-9]|exception' gpurun_out/r2w_k_gnbwd.log | cut -c1-300 | head -n 6
for c in train_step train_step_b128; do
  timeout 900 python tests/e2e_cases.py $c > gpurun_out/r2w_e2e_$c.log 2>&1; echo "e2e case $c rc=$?"
  grep -E '^\{|^===' gpurun_out/r2w_e2e_$c.log | cut -c1-330 | tail -n 6
done
for rep in 1 2; do
for hw in 256 1024; do
  B200_GNB_SLAB_HW=$hw python tools/bench_train.py cfg 128 20 > gpurun_out/r2w_train_$hw.json 2> gpurun_out/r2w_train_$hw.err
  python - $hw <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r2w_train_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('slab_hw', sys.argv[1], round(d['ms_per_step'],3), 'ms', d['kernels_per_step'], 'launches', {k:round(v['ms'],3) for k,v in d['kernels'].items() if 'groupnorm' in k})
PY
done
done
