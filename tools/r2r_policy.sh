#!/bin/bash
# round 2: same-box A/B of the block-output GroupNorm fusion policy (off / without the concat form / default = every eligible layer)
cd "$(dirname "$0")/.."
for rep in 1 2; do
for cfg in "B200_FUSE_GN1=0" "B200_FUSE_GN1_CAT=0" ""; do
  env $cfg python bench.py --no-extras --no-cpu-baseline --no-parity > gpurun_out/r2r_b.json 2> gpurun_out/r2r_b.err
  python - "$cfg" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2r_b.json').read().strip().splitlines()[-1])
print((sys.argv[1] or 'default').ljust(22), round(d['value'],1), {k:(v['n_per_forward'],round(v['ms_per_forward'],3)) for k,v in d['kernels'].items() if k in ('conv_gemm','groupnorm_apply')}, d['clocks']['sm_mhz'])
PY
done
done
