#!/bin/bash
# round 2: ncu --set full of the one-launch attention block inside a steady-state forward (+ launch list of the new forward)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/profile_forward.py 256 3 > gpurun_out/r2g_pf_plain.log 2>&1 || { echo "plain run failed"; tail -n 5 gpurun_out/r2g_pf_plain.log; exit 1; }
tail -n 4 gpurun_out/r2g_pf_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2g_launches_fwd_b256.csv \
    python tools/profile_forward.py 256 3 > gpurun_out/r2g_pf_ncu.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_block --launch-skip 5 -c 2 -f \
    -o gpurun_out/r2g_prof_attnblock python tools/profile_forward.py 256 2 > gpurun_out/r2g_prof_attnblock.log 2>&1; echo "attn_block full rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tail -3
