#!/bin/bash
# round 2: fused attention adjoint + batched embedding / qkv weight gradients: parity, then the training step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tests/kernel_cases.py attention_bwd > gpurun_out/r2l_k_attnbwd.log 2>&1; echo "kernel case exit $?"
grep -E "PASS|FAIL|rel-L2" gpurun_out/r2l_k_attnbwd.log | tail -n 20
for c in train_step train_step_pesser train_step_adm train_multi_step; do
  timeout 900 python tests/e2e_cases.py $c > gpurun_out/r2l_e2e_$c.log 2>&1; echo "$c exit $?"
  grep -E '^\{|^===' gpurun_out/r2l_e2e_$c.log | cut -c1-260 | tail -n 6
done
timeout 600 python tools/bench_train.py cfg 128 10 > gpurun_out/r2l_train_on.json 2> gpurun_out/r2l_train_on.err; echo "bench exit $?"
tail -c 1600 gpurun_out/r2l_train_on.json
