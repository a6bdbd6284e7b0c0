#!/bin/bash
# round 2: block-output form of the fused conv + GroupNorm (conv2 -> next norm1 / head norm): parity, then A/B of DDIM-50
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in conv_gnfuse_out conv_gnfuse; do
  timeout 600 python tests/kernel_cases.py $c > gpurun_out/r2o_k_$c.log 2>&1; echo "kernel case $c exit $?"
  grep -E '"ok": false|PASS|FAIL|Error|error' gpurun_out/r2o_k_$c.log | cut -c1-400 | tail -n 8
done
for c in unet_forward ddim50 cfg engine_hygiene; do
  timeout 900 python tests/e2e_cases.py $c > gpurun_out/r2o_e2e_$c.log 2>&1; echo "$c exit $?"
  grep -E '^\{|^===' gpurun_out/r2o_e2e_$c.log | cut -c1-260 | tail -n 5
done
for v in 0 1; do
  B200_FUSE_GN1_CAT=$v python bench.py --no-extras --no-cpu-baseline --no-parity > gpurun_out/r2o_bench_$v.json 2> gpurun_out/r2o_bench_$v.err
  python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r2o_bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('FUSE_GN1_CAT', sys.argv[1], round(d['value'],1), {k:(v['n_per_forward'],round(v['ms_per_forward'],3)) for k,v in d['kernels'].items() if k in ('conv_gemm','groupnorm_apply')})
PY
done
