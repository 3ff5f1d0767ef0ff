#!/usr/bin/env python3
"""Summarise an .ncu-rep (captured on the B200 box with `ncu --set full --clock-control none --import-source on`)
into a small text table that can be committed: per kernel launch the duration, DRAM bytes, throughput
percentages, occupancy, registers and the top warp-stall reasons.

    python profiles/summarize_ncu.py gpurun_out/x.ncu-rep > profiles/rNN_x.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'dram_read'),
    ('dram__bytes_write.sum', 'dram_write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'),
    ('l1tex__t_sector_hit_rate.pct', 'l1_hit_pct'),
    ('lts__t_sector_hit_rate.pct', 'l2_hit_pct'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy_pct'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__waves_per_multiprocessor', 'waves'),
    ('smsp__inst_executed.sum', 'warp_insts'),
]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
    print(f'# {path}')
    for r in rows[2:]:
        print(f"\n== {r[idx['Kernel Name']][:110]}")
        for key, name in WANT:
            if key in idx:
                print(f'   {name:14s} {r[idx[key]]:>16s} {units[idx[key]]}')
        st = sorted(((float(r[idx[h]].replace(',', '')), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))
                     for h in stall), reverse=True)[:5]
        print('   top stalls (warps per issue): ' + ', '.join(f'{n}={v:.2f}' for v, n in st))


if __name__ == '__main__':
    main(sys.argv[1])
