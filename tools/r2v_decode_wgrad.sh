#!/bin/bash
# round 2: shift-based tile decode (all conv kernels) + division-free K-block stepping in the weight-gradient producers:
# conv / wgrad / gemm parity cases, wgrad micro-benchmark, training step, default DDIM-50 bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in conv_basic conv_epilogue conv_n256 conv_small_hw conv_1x1 conv_shortcut conv_stride2 conv_lastconv conv_up2 conv_tproj conv_gnfuse_out wgrad gemm attention_bwd; do
  timeout 600 python tests/kernel_cases.py $c > gpurun_out/r2v_k_$c.log 2>&1; echo "kernel case $c rc=$?"
  grep -E '"ok": false|mismatch": [1-9]|exception' gpurun_out/r2v_k_$c.log | cut -c1-300 | head -n 4
done
python tools/bench_wgrad.py > gpurun_out/r2v_wgrad.txt 2>&1; cat gpurun_out/r2v_wgrad.txt | tail -n 16
python tools/bench_train.py cfg 128 20 > gpurun_out/r2v_train.json 2> gpurun_out/r2v_train.err; tail -n 2 gpurun_out/r2v_train.json | cut -c1-600
python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2v_b.json 2> gpurun_out/r2v_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_b.json').read().strip().splitlines()[-1])
print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['unet_fwd_frac_of_bf16_peak'], d['roofline']['frac'], d['parity'], d['clocks'])
PY
