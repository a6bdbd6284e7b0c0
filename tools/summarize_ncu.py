"""Summarises an `ncu --set full` report into a text file for profiles/: per profiled launch the figures BASELINE.json's
north_star asks for -- tensor-pipe utilisation (sm__pipe_tensor_cycles_active, % of elapsed cycles), DRAM throughput and
bytes, duration, SM clock, occupancy, registers, shared memory.  Runs in the build container (no GPU needed):
    python tools/summarize_ncu.py gpurun_out/r02_prof_conv.ncu-rep profiles/r02_prof_conv.details.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    ('gpu__time_duration.sum', 'duration'),
    ('sm__cycles_elapsed.max', 'elapsed cycles (max over SMs)'),
    ('smsp__cycles_active.avg', 'SMSP active cycles (avg)'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor pipe active, % of elapsed'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active, % of active'),
    ('sm__inst_executed_pipe_tensor.sum', 'tensor-pipe instructions'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
    ('dram__bytes_read.sum', 'DRAM bytes read'),
    ('dram__bytes_write.sum', 'DRAM bytes written'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('launch__registers_per_thread', 'registers / thread'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem / block'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__cycles_active.avg', 'SM active cycles'),
    ('gpc__cycles_elapsed.avg.per_second', 'GPC clock'),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    lines = [l for l in raw.splitlines() if not l.startswith('==')]
    rows = list(csv.reader(io.StringIO('\n'.join(lines))))
    header, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(header)}
    with open(out, 'w') as f:
        f.write(f'# {rep}: ncu --set full --clock-control none (per-launch, cold cache, ~40 replays); summarised by tools/summarize_ncu.py\n')
        for r in data:
            name = r[idx['Kernel Name']] if 'Kernel Name' in idx else '?'
            f.write(f"\n{name.split('(')[0][:110]}  id={r[idx['ID']]}\n")
            for key, label in WANT:
                cands = [h for h in header if h == key or h.startswith(key)]
                if not cands:
                    continue
                i = idx[cands[0]]
                f.write(f'    {label:38s} {r[i]:>18s} {units[i]}\n')
    print(f'wrote {out}: {len(data)} launches')


if __name__ == '__main__':
    main()
