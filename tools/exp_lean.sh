#!/bin/bash
# Round-2 first experiment: the lean conv epilogue (B200_EPI_LEAN=1, conv_epilogue.cuh) was written after round 1's GPU
# budget was spent -- it builds and its SASS was inspected, but it has NOT run on hardware.  This script (1) runs every
# conv parity case and the forward / sampling / training e2e cases under the knob, (2) times the default bench with and
# without it.  Flip the default in conv_gemm.cu / conv_gemm2.cu only if (1) is green and (2) is faster.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rc=0
for c in conv_basic conv_epilogue conv_n256 conv_small_hw conv_1x1 conv_shortcut conv_stride2 conv_up2 conv_tproj; do
  B200_EPI_LEAN=1 timeout 200 python tests/kernel_cases.py $c > gpurun_out/lean_$c.log 2>&1; r=$?
  echo "lean case $c exit $r"; [ $r -ne 0 ] && { rc=1; grep -v '"ok": true' gpurun_out/lean_$c.log | tail -n 5 | cut -c1-300; }
done
for c in unet_forward ddim50 cfg train_step; do
  B200_EPI_LEAN=1 timeout 300 python tests/e2e_cases.py $c > gpurun_out/lean_e2e_$c.log 2>&1; r=$?
  echo "lean e2e $c exit $r"; tail -n 2 gpurun_out/lean_e2e_$c.log | cut -c1-300; [ $r -ne 0 ] && rc=1
done
python bench.py > gpurun_out/bench_lean_off.json 2>/dev/null
B200_EPI_LEAN=1 python bench.py > gpurun_out/bench_lean_on.json 2>/dev/null
# second hypothesis: 128-pixel tiles for the short-K layers, alone and with the lean epilogue
B200_SHORTK_NP=128 timeout 100 python tests/kernel_cases.py conv_1x1 2>&1 | tail -n 1
B200_SHORTK_NP=128 python bench.py > gpurun_out/bench_lean_np128.json 2>/dev/null
B200_SHORTK_NP=128 B200_EPI_LEAN=1 python bench.py > gpurun_out/bench_lean_on_np128.json 2>/dev/null
python - <<'PY'
import json
for tag in ('off', 'on', 'np128', 'on_np128'):
    l = json.load(open(f'gpurun_out/bench_lean_{tag}.json'))
    print(tag, 'ddim50 images/s', round(l['value'], 1), 'conv ms/forward', round(l['kernels']['conv_gemm']['ms_per_forward'], 3),
          'adm256 ms', round(l['extras']['adm256_forward']['ms_per_forward'], 2), 'train ms', round(l['extras']['cfg_train_step']['ms_per_step'], 2))
PY
# third experiment: conv1 -> norm2 fused in the conv epilogue (b200_conv2d_gn_fwd), kernel-level parity only so far
timeout 200 python tests/kernel_cases.py conv_gnfuse > gpurun_out/gnfuse.log 2>&1; echo "conv_gnfuse exit $?"; tail -n 6 gpurun_out/gnfuse.log | cut -c1-300
# ... and through the engine (B200_FUSE_GN2=1): forward / sampling parity, then the bench A/B
for c in unet_forward ddim50 cfg; do
  B200_FUSE_GN2=1 timeout 300 python tests/e2e_cases.py $c > gpurun_out/fuse_e2e_$c.log 2>&1; echo "fuse e2e $c exit $?"
  tail -n 2 gpurun_out/fuse_e2e_$c.log | cut -c1-300
done
B200_FUSE_GN2=1 python bench.py --no-extras 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); print('fuse_gn2 ddim50 images/s', round(l['value'],1), {k: round(v['ms_per_forward'],3) for k, v in l['kernels'].items()})"
exit $rc
