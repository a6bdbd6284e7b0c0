#!/bin/bash
# K4 (sampler update) evidence in one short GPU call: parity cases, smoke(), the HBM-roofline micro-benchmark, one
# ncu --set full capture of the streaming kernel (after the same command exited 0 without ncu), then the default bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rc=0
for c in sampler sampler_cfg sampler_large; do
  timeout 200 python tests/kernel_cases.py $c > gpurun_out/kernels_$c.log 2>&1; r=$?
  echo "case $c exit $r"; tail -n 2 gpurun_out/kernels_$c.log | cut -c1-300
  [ $r -ne 0 ] && rc=1
done
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log | cut -c1-300
timeout 200 python tools/bench_sampler.py --json gpurun_out/bench_sampler.json > gpurun_out/bench_sampler.log 2>&1; echo "bench_sampler exit $?"
cat gpurun_out/bench_sampler.log
timeout 120 python tools/bench_sampler.py --one > gpurun_out/bench_sampler_one.log 2>&1 && \
timeout 240 ncu --set full --clock-control none --import-source on -k regex:sampler_step_vec4 --launch-skip 2 -c 1 -f \
    -o gpurun_out/r01_sampler_step_full python tools/bench_sampler.py --one > gpurun_out/prof_sampler.log 2>&1
echo "ncu exit $?"
T0=$(date +%s)
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 ))s"
cut -c1-1500 gpurun_out/bench_default.json
exit $rc
