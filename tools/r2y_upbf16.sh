#!/bin/bash
# round 2: bf16 up-sampling conv outputs + GroupNorm over (bf16 first source, fp32 skip): parity, e2e, same-box A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in groupnorm conv_up2; do
  timeout 600 python tests/kernel_cases.py $c > gpurun_out/r2y_k_$c.log 2>&1; echo "kernel case $c rc=$?"
  grep -E '"ok": false|mismatch": [1-9]|exception' gpurun_out/r2y_k_$c.log | cut -c1-300 | head -n 6
done
for c in unet_forward ddim50 cfg engine_hygiene; do
  timeout 900 python tests/e2e_cases.py $c > gpurun_out/r2y_e2e_$c.log 2>&1; echo "e2e case $c rc=$?"
  grep -E '^\{|^===' gpurun_out/r2y_e2e_$c.log | cut -c1-220 | tail -n 5
done
for rep in 1 2; do
for cfg in B200_UP_BF16=0 B200_NOP=1; do
  env $cfg python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2y_b.json 2> gpurun_out/r2y_b.err
  python - $cfg <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2y_b.json').read().strip().splitlines()[-1])
print(sys.argv[1].ljust(18), round(d['value'],1), 'img/s', {k:(v['n_per_forward'],round(v['ms_per_forward'],3)) for k,v in d['kernels'].items() if k in ('conv_gemm','groupnorm_apply')}, 'eps', round(d['parity']['eps_rel_l2'],5), 'psnr', round(d['parity']['ddim_psnr_db'],2))
PY
done
done
