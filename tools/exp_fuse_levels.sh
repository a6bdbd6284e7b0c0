#!/bin/bash
# round 2: which resolutions profit from the fused conv1 -> GroupNorm entry?  (B200_FUSE_GN2_HW = allowed image sizes)
cd "$(dirname "$0")/.."
for hw in "16,64,256,1024" "16,256,1024" "64,256,1024" "16,64,1024" "256,1024" "1024" ""; do
  B200_FUSE_GN2_HW="$hw" python bench.py --no-extras --no-cpu-baseline --no-parity > gpurun_out/r2k_b.json 2> gpurun_out/r2k_b.err
  python - "$hw" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2k_b.json').read().strip().splitlines()[-1])
print('HW', sys.argv[1] or '(none)', round(d['value'],1), {k:(v['n_per_forward'],round(v['ms_per_forward'],3)) for k,v in d['kernels'].items() if k in ('conv_gemm','groupnorm_apply')})
PY
done
