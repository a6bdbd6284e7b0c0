#!/bin/bash
# Round-2 ncu evidence (one GPU; every ncu run only after the same command exited 0 without ncu):
#   1. launch list of three eager CIFAR-10 forwards at batch 256 (device time per launch; cold-cache, serialised)
#   2. ncu --set full of conv_gemm (3x3 @32x32, @16x16, the fused conv+GroupNorm instantiation), groupnorm_apply and
#      attention launches of the steady-state forward
# The .ncu-rep files come back in gpurun_out/ and are summarised HERE by tools/summarize_ncu.py into profiles/*.details.txt
# (tensor-pipe utilisation, DRAM throughput / bytes, occupancy, registers).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/profile_forward.py 256 3 > gpurun_out/r02_pf_plain.log 2>&1 || { echo "plain run failed"; tail -n 5 gpurun_out/r02_pf_plain.log; exit 1; }
cat gpurun_out/r02_pf_plain.log | tail -n 4
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_fwd_b256.csv \
    python tools/profile_forward.py 256 3 > gpurun_out/r02_pf_ncu.log 2>&1; echo "launch list rc=$?"
# second forward = launches 120.. (119 per forward): conv layers of the 32x32 / 16x16 levels and one fused conv+GN
ncu --set full --clock-control none --import-source on -k regex:conv_gemm --launch-skip 74 -c 14 -f \
    -o gpurun_out/r02_prof_conv python tools/profile_forward.py 256 2 > gpurun_out/r02_prof_conv.log 2>&1; echo "conv full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:groupnorm_apply --launch-skip 34 -c 6 -f \
    -o gpurun_out/r02_prof_gn python tools/profile_forward.py 256 2 > gpurun_out/r02_prof_gn.log 2>&1; echo "gn full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention --launch-skip 6 -c 2 -f \
    -o gpurun_out/r02_prof_attn python tools/profile_forward.py 256 2 > gpurun_out/r02_prof_attn.log 2>&1; echo "attn full rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tail -5
