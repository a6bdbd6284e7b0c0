#!/bin/bash
# Round-2 final ncu evidence of the sampling forward (one GPU; every ncu run only after the same command exited 0 without
# ncu).  gpurun brings back at most 64 MiB, so:
#   A. every conv and GroupNorm-apply launch of the steady-state (second) eager forward at batch 256 with the metric subset
#      tools/summarize_ncu.py prints (tensor-pipe utilisation, DRAM bytes / throughput, occupancy ...): small reports;
#   B. ncu --set full --import-source on for nine conv launches of the 32x32 / 16x16 levels (fused conv1 -> norm2 multi-tile,
#      block-output form, lean, one-tile fused) and two attention-block launches.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/profile_forward.py 256 2 > gpurun_out/r2t_pf_plain.log 2>&1 || { echo "plain run failed"; tail -n 5 gpurun_out/r2t_pf_plain.log; exit 1; }
tail -n 3 gpurun_out/r2t_pf_plain.log
M=gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__cycles_active.avg,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,gpc__cycles_elapsed.avg.per_second
ncu --metrics $M --clock-control none -k regex:conv_gemm --launch-skip 56 -c 55 -f \
    -o gpurun_out/r2t_conv_all python tools/profile_forward.py 256 2 > gpurun_out/r2t_conv_all.log 2>&1; echo "conv metrics rc=$?"
ncu --metrics $M --clock-control none -k regex:groupnorm_apply --launch-skip 17 -c 17 -f \
    -o gpurun_out/r2t_gn_all python tools/profile_forward.py 256 2 > gpurun_out/r2t_gn_all.log 2>&1; echo "gn metrics rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_gemm --launch-skip 57 -c 9 -f \
    -o gpurun_out/r2t_prof_conv python tools/profile_forward.py 256 2 > gpurun_out/r2t_prof_conv.log 2>&1; echo "conv full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_block --launch-skip 5 -c 2 -f \
    -o gpurun_out/r2t_prof_attnblock python tools/profile_forward.py 256 2 > gpurun_out/r2t_prof_attnblock.log 2>&1; echo "attn_block full rc=$?"
ls -la gpurun_out/r2t_*.ncu-rep; du -sh gpurun_out
