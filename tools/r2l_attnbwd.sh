#!/bin/bash
# round 2: fused attention adjoint (b200_attention_bwd): kernel parity, training-gradient parity, A/B of the training step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tests/kernel_cases.py attention_bwd > gpurun_out/r2l_k_attnbwd.log 2>&1; echo "kernel case exit $?"
tail -n 30 gpurun_out/r2l_k_attnbwd.log | cut -c1-400
timeout 600 python tests/kernel_cases.py attention > gpurun_out/r2l_k_attn.log 2>&1; echo "attention fwd case exit $?"
timeout 900 python tests/e2e_cases.py train_step > gpurun_out/r2l_e2e_train.log 2>&1; echo "train_step exit $?"
grep -E '^\{|^===|rel|grad' gpurun_out/r2l_e2e_train.log | cut -c1-300 | tail -n 20
timeout 900 python tests/e2e_cases.py train_step_adm > gpurun_out/r2l_e2e_train_adm.log 2>&1; echo "train_step_adm exit $?"
tail -n 5 gpurun_out/r2l_e2e_train_adm.log | cut -c1-300
B200_ATTN_BWD_FUSED=0 timeout 600 python tools/bench_train.py cfg 128 10 > gpurun_out/r2l_train_off.json 2> gpurun_out/r2l_train_off.err; echo "bench off exit $?"
timeout 600 python tools/bench_train.py cfg 128 10 > gpurun_out/r2l_train_on.json 2> gpurun_out/r2l_train_on.err; echo "bench on exit $?"
tail -c 1500 gpurun_out/r2l_train_off.json; echo; tail -c 1500 gpurun_out/r2l_train_on.json
