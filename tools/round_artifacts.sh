#!/bin/bash
# End-of-round evidence: default bench line (+ wall time), reference arm, launch lists and ncu --set full captures of
# the dominant kernels (each only after the same command exited 0 without ncu).  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 ))s"
T0=$(date +%s)
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$? wall=$(( $(date +%s) - T0 ))s"
python tools/profile_forward.py 256 3 > gpurun_out/pf_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_fwd_b256_v7.csv \
    python tools/profile_forward.py 256 3 > gpurun_out/pf_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel --launch-skip 40 -c 2 -f \
    -o gpurun_out/prof_conv_v7 python tools/profile_forward.py 256 2 > gpurun_out/prof_conv_v7.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:groupnorm_apply --launch-skip 60 -c 2 -f \
    -o gpurun_out/prof_gn_v7 python tools/profile_forward.py 256 2 > gpurun_out/prof_gn_v7.log 2>&1
ls -la gpurun_out | tail -8
