#!/bin/bash
# round 2: attention-adjoint micro-benchmark + ncu capture, launch list of the training step (eager steps under ncu)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/bench_attn_bwd.py 128 256 4 20 | tee gpurun_out/r2m_attnbwd_bench.txt
python tools/bench_attn_bwd.py 128 64 4 20 | tee -a gpurun_out/r2m_attnbwd_bench.txt
python tools/bench_attn_bwd.py 128 16 4 20 | tee -a gpurun_out/r2m_attnbwd_bench.txt
ncu --set full --clock-control none --import-source on -k regex:attn_bwd --launch-skip 3 -c 1 -f \
    -o gpurun_out/r2m_prof_attnbwd python tools/bench_attn_bwd.py 128 256 4 3 > gpurun_out/r2m_prof_attnbwd.log 2>&1; echo "attn_bwd full rc=$?"
python tools/profile_train.py 128 3 > gpurun_out/r2m_pt_plain.log 2>&1 || { echo "plain train run failed"; tail -n 5 gpurun_out/r2m_pt_plain.log; }
tail -n 3 gpurun_out/r2m_pt_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2m_launches_train_cfg_b128.csv \
    python tools/profile_train.py 128 3 > gpurun_out/r2m_pt_ncu.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/r2m_*
