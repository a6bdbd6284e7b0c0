#!/bin/bash
# Round-2 end state, one GPU: default bench line (+ extras), reference arm (short), launch lists of the forward and the
# training step (part a), ncu --set full (source counters; part b = `bash tools/r2z_final.sh b`) of the 32x32 / 16x16 conv launches, the GroupNorm launches and the
# attention block, DRAM traffic of the dominant kernels.  Every ncu run follows a plain run of the same command.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PART=${1:-a}
if [ "$PART" = "b" ]; then
python tools/profile_forward.py 256 2 > gpurun_out/r2z_pf_plain_b.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm --launch-skip 56 -c 8 -f \
    -o gpurun_out/r2z_prof_conv python tools/profile_forward.py 256 2 > gpurun_out/r2z_prof_conv.log 2>&1; echo "conv full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_block --launch-skip 5 -c 1 -f \
    -o gpurun_out/r2z_prof_attnblock python tools/profile_forward.py 256 2 > gpurun_out/r2z_prof_attnblock.log 2>&1; echo "attn_block full rc=$?"
du -sh gpurun_out; exit 0
fi
python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; echo "reference arm rc=$?"
python tools/profile_forward.py 256 3 > gpurun_out/r2z_pf_plain.log 2>&1 || { echo "plain run failed"; tail -n 5 gpurun_out/r2z_pf_plain.log; exit 1; }
tail -n 2 gpurun_out/r2z_pf_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2z_launches_fwd_b256.csv \
    python tools/profile_forward.py 256 3 > gpurun_out/r2z_pf_ncu.log 2>&1; echo "forward launch list rc=$?"
M=gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__cycles_active.avg,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,gpc__cycles_elapsed.avg.per_second
ncu --metrics $M --clock-control none -k regex:'conv_gemm|groupnorm_apply|attn_block|conv3x3_first|cast_bf16' --launch-skip 86 -c 86 -f \
    -o gpurun_out/r2z_fwd_all python tools/profile_forward.py 256 2 > gpurun_out/r2z_fwd_all.log 2>&1; echo "forward metrics rc=$?"
bash tools/capture_traffic.sh > gpurun_out/r2z_traffic.json; echo "traffic rc=$?"; tail -c 600 gpurun_out/r2z_traffic.json
python tools/profile_train.py 128 3 > gpurun_out/r2z_pt_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2z_launches_train_cfg_b128.csv \
    python tools/profile_train.py 128 3 > gpurun_out/r2z_pt_ncu.log 2>&1; echo "training launch list rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['unet_fwd_frac_of_bf16_peak'], d['roofline']['frac'], d['clocks'])
print({k:(v.get('ms_per_step') or v.get('ms_per_forward') or v.get('ms_per_run') or v.get('images_per_s')) for k,v in d['extras'].items()})
PY
du -sh gpurun_out; ls -la gpurun_out/r2z_*
