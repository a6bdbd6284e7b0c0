#!/bin/bash
# round 2: validation + A/B of the 4-CTA-cluster fused conv1 -> GroupNorm epilogue at 32x32
mkdir -p gpurun_out
timeout 300 python tests/kernel_cases.py conv_gnfuse > gpurun_out/r2i_k_gnfuse.log 2>&1; echo "conv_gnfuse rc=$?"
cut -c1-330 gpurun_out/r2i_k_gnfuse.log | tail -n 10
if grep -q "=== conv_gnfuse: PASS" gpurun_out/r2i_k_gnfuse.log; then
  timeout 600 python tests/e2e_cases.py unet_forward ddim50 > gpurun_out/r2i_e2e.log 2>&1; echo "e2e rc=$?"
  cut -c1-300 gpurun_out/r2i_e2e.log | grep -E "PASS|FAIL|false"
  B200_FUSE_GN2_CLUSTER=0 timeout 600 python bench.py --no-extras --no-cpu-baseline --no-parity > gpurun_out/r2i_bench_off.json 2> gpurun_out/r2i_bench_off.err
  timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2i_bench_on.json 2> gpurun_out/r2i_bench_on.err
  python - <<'PY'
import json
for n in ('off','on'):
    try:
        d=json.loads(open(f'gpurun_out/r2i_bench_{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:(v['n_per_forward'],round(v['ms_per_forward'],3)) for k,v in d['kernels'].items()}, d.get('parity',{}).get('eps_rel_l2'), d.get('parity',{}).get('ddim_psnr_db'), d.get('parity',{}).get('bitwise_reproducible'))
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/r2i_bench_{n}.err').read()[-1500:])
PY
fi
