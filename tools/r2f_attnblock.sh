#!/bin/bash
# round 2: validation + timing of the one-launch attention block (b200_attn_block_fwd); optional ncu capture (arg "ncu")
mkdir -p gpurun_out
timeout 300 python tests/kernel_cases.py attn_block > gpurun_out/r2f_k_attn_block.log 2>&1; echo "attn_block rc=$?"
timeout 300 python tests/kernel_cases.py conv_up2 > gpurun_out/r2f_k_conv_up2.log 2>&1; echo "conv_up2 rc=$?"; cut -c1-200 gpurun_out/r2f_k_conv_up2.log | tail -n 16
cut -c1-250 gpurun_out/r2f_k_attn_block.log | tail -n 14
if grep -q "=== attn_block: PASS" gpurun_out/r2f_k_attn_block.log; then
  timeout 600 python tests/e2e_cases.py ddim50 > gpurun_out/r2f_e2e.log 2>&1; echo "e2e rc=$?"
  cut -c1-300 gpurun_out/r2f_e2e.log | tail -n 7
  timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2f_bench_on.json 2> gpurun_out/r2f_bench_on.err
  python - <<'PY'
import json
for n in ('on',):
    try:
        d=json.loads(open(f'gpurun_out/r2f_bench_{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:(v['n_per_forward'],round(v['ms_per_forward'],3)) for k,v in d['kernels'].items()}, d.get('parity',{}).get('eps_rel_l2'), d.get('parity',{}).get('ddim_psnr_db'), d.get('parity',{}).get('bitwise_reproducible'))
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/r2f_bench_{n}.err').read()[-1500:])
PY
  if [ "$1" = "ncu" ]; then
    ncu --set full --clock-control none --import-source on -k regex:attn_block --launch-skip 5 -c 1 -f \
      -o gpurun_out/r2g_prof_attnblock python tools/profile_forward.py 256 2 > gpurun_out/r2g_prof_attnblock.log 2>&1; echo "attn_block full rc=$?"
  fi
fi
