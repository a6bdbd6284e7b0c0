#!/bin/bash
# round 2: launch list (cold-cache per-launch times) of the current forward at batch 256 + ncu --set full with source
# counters of the 32x32 / 16x16 fused-epilogue conv launches of the steady-state forward
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/profile_forward.py 256 3 > gpurun_out/r2x_pf_plain.log 2>&1 || { echo "plain run failed"; tail -n 5 gpurun_out/r2x_pf_plain.log; exit 1; }
tail -n 3 gpurun_out/r2x_pf_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2x_launches_fwd_b256.csv \
    python tools/profile_forward.py 256 3 > gpurun_out/r2x_pf_ncu.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_gemm --launch-skip 56 -c 12 -f \
    -o gpurun_out/r2x_prof_conv python tools/profile_forward.py 256 2 > gpurun_out/r2x_prof_conv.log 2>&1; echo "conv full rc=$?"
ls -la gpurun_out/r2x_*
