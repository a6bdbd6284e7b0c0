#!/bin/bash
# DRAM traffic per launch of the dominant kernels (for bench.py's roofline.traffic): ncu over two eager CIFAR-10 forwards
# at batch 256, metrics dram__bytes_read/write + duration, conv_gemm* and groupnorm_apply launches only; the steady-state
# (second) forward is averaged by tools/summarize_traffic.py into profiles/conv_gemm_traffic.json.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/profile_forward.py 256 2 > gpurun_out/traffic_plain.log 2>&1 || exit 1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'conv_gemm|groupnorm_apply' --csv --log-file gpurun_out/traffic_r01.csv \
    python tools/profile_forward.py 256 2 > gpurun_out/traffic_ncu.log 2>&1
python tools/summarize_traffic.py gpurun_out/traffic_r01.csv
