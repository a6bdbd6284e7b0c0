#!/bin/bash
# round 2: block-output GroupNorm fusion in training forwards: training parity (all families), same-box A/B of the step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in train_step train_step_b128 train_multi_step train_step_pesser train_step_adm engine_hygiene; do
  timeout 900 python tests/e2e_cases.py $c > gpurun_out/r2y_e2e_$c.log 2>&1; echo "e2e case $c rc=$?"
  grep -E '^\{|^===' gpurun_out/r2y_e2e_$c.log | cut -c1-260 | tail -n 5
done
for rep in 1 2; do
for cfg in B200_FUSE_GN1_TRAIN=0 B200_NOP=1; do
  env $cfg python tools/bench_train.py cfg 128 20 > gpurun_out/r2y_train.json 2> gpurun_out/r2y_train.err
  python - $cfg <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2y_train.json').read().strip().splitlines()[-1])
print(sys.argv[1].ljust(24), round(d['ms_per_step'],3), 'ms', d['kernels_per_step'], 'launches', {k:(v['n'],round(v['ms'],3)) for k,v in d['kernels'].items() if k in ('groupnorm_apply','conv_gemm')}, 'loss', round(d['loss_last'],4), 'mem', round(d['peak_mem_gib'],2))
PY
done
done
