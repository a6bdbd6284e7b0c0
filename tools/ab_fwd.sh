#!/bin/bash
# A/B of forward-path experiment knobs: tools/ab_fwd.sh "ENV=.. ENV=.." "ENV=.." ...  (one bench.py run per argument)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg python bench.py --no-cpu-baseline --no-extras --steps 3 --warmup 3 > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err
  echo "[$cfg] $(python -c "import json,sys; d=json.loads(open('gpurun_out/ab_$i.json').read().strip().splitlines()[-1]); print(round(d['value'],1),'img/s', round(d['ms_per_step'],2),'ms', 'frac', round(d.get('unet_fwd_frac_of_bf16_peak',0),4))" 2>&1 | tail -1)"
done
