#!/bin/bash
# round 2: tensor-core first convolution + dead fp32 block outputs + late output stores in the multi-tile fused epilogue: kernel / e2e parity, then same-box A/B (DDIM-50, batch 256)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in ${K_CASES-first_conv_tc conv_gnfuse conv_gnfuse_out}; do
  timeout 600 python tests/kernel_cases.py $c > gpurun_out/r2u_k_$c.log 2>&1; echo "kernel case $c rc=$?"
  grep -E '"ok": false|mismatch": [1-9]|exception|^===' gpurun_out/r2u_k_$c.log | cut -c1-400 | head -n 12
done
for c in ${E_CASES-unet_forward ddim50 cfg}; do
  timeout 900 python tests/e2e_cases.py $c > gpurun_out/r2u_e2e_$c.log 2>&1; echo "e2e case $c rc=$?"
  grep -E '^\{|^===' gpurun_out/r2u_e2e_$c.log | cut -c1-300 | tail -n 8
done
AB_CONFIGS=${AB_CONFIGS:-"B200_FIRST_TC=0,B200_SKIP_DEAD_OUT=0,B200_GN_LATE_OUT=0 B200_GN_LATE_OUT=0 B200_FIRST_TC=0 B200_SKIP_DEAD_OUT=0 default"}
for rep in 1 2; do
for cfg in $AB_CONFIGS; do
  env $(echo $cfg | tr ',' ' ' | sed 's/^default$/B200_NOP=1/') python bench.py --no-extras --no-cpu-baseline --no-parity > gpurun_out/r2u_b.json 2> gpurun_out/r2u_b.err
  python - "$cfg" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2u_b.json').read().strip().splitlines()[-1])
print((sys.argv[1] or 'default').ljust(40), round(d['value'],1), {k:(v['n_per_forward'],round(v['ms_per_forward'],3)) for k,v in d['kernels'].items()}, d['clocks']['sm_mhz'])
PY
done
done
