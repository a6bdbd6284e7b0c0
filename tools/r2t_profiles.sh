#!/bin/bash
# Round-2 final ncu evidence of the sampling forward (one GPU; every ncu run only after the same command exited 0 without
# ncu): ncu --set full of EVERY conv launch and every GroupNorm-apply launch of the steady-state (second) eager forward at
# batch 256, two attention-block launches, and the DRAM-traffic capture behind bench.py's roofline.traffic.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/profile_forward.py 256 2 > gpurun_out/r2t_pf_plain.log 2>&1 || { echo "plain run failed"; tail -n 5 gpurun_out/r2t_pf_plain.log; exit 1; }
tail -n 3 gpurun_out/r2t_pf_plain.log
ncu --set full --clock-control none --import-source on -k regex:conv_gemm --launch-skip 56 -c 55 -f \
    -o gpurun_out/r2t_prof_conv python tools/profile_forward.py 256 2 > gpurun_out/r2t_prof_conv.log 2>&1; echo "conv full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:groupnorm_apply --launch-skip 17 -c 17 -f \
    -o gpurun_out/r2t_prof_gn python tools/profile_forward.py 256 2 > gpurun_out/r2t_prof_gn.log 2>&1; echo "gn full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_block --launch-skip 5 -c 2 -f \
    -o gpurun_out/r2t_prof_attnblock python tools/profile_forward.py 256 2 > gpurun_out/r2t_prof_attnblock.log 2>&1; echo "attn_block full rc=$?"
bash tools/capture_traffic.sh; echo "traffic rc=$?"
ls -la gpurun_out/r2t_*.ncu-rep gpurun_out/traffic_r01.csv
