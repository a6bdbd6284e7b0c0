#!/bin/bash
# End-of-session evidence for the kernels changed last (K4 streaming kernel, 8-channel cast, attention row split, tiled
# weight pack): e2e parity subset, forward launch list, ncu --set full of K4 (each ncu pass only after the same command
# exited 0 without ncu).  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rc=0
for c in unet_forward ddim50 cfg; do
  timeout 240 python tests/e2e_cases.py $c > gpurun_out/e2e_$c.log 2>&1; r=$?
  echo "e2e $c exit $r"; tail -n 3 gpurun_out/e2e_$c.log | cut -c1-300
  [ $r -ne 0 ] && rc=1
done
python tools/profile_forward.py 256 3 > gpurun_out/pf_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_fwd_b256_v9.csv \
    python tools/profile_forward.py 256 3 > gpurun_out/pf_ncu.log 2>&1
echo "launch list exit $?"
timeout 120 python tools/bench_sampler.py --one > gpurun_out/bench_sampler_one.log 2>&1 && \
timeout 240 ncu --set full --clock-control none --import-source on -k regex:sampler_step_vec4 --launch-skip 2 -c 1 -f \
    -o gpurun_out/r01_sampler_step_full python tools/bench_sampler.py --one > gpurun_out/prof_sampler.log 2>&1
echo "ncu sampler exit $?"
exit $rc
