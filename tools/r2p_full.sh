#!/bin/bash
# round 2: full GPU suite + default bench + launch list of the current forward
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r2p_gputest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2p_gputest.log | cut -c1-300
python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"
python tools/profile_forward.py 256 3 > gpurun_out/r2p_pf_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2p_launches_fwd_b256.csv \
    python tools/profile_forward.py 256 3 > gpurun_out/r2p_pf_ncu.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['unet_fwd_frac_of_bf16_peak'], d['roofline']['frac'], d['parity'])
print({k:(v.get('ms_per_step') or v.get('ms_per_forward') or v.get('ms_per_run')) for k,v in d['extras'].items()})
PY
