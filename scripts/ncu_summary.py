"""Summarise an ncu report (read here, no GPU): per kernel the metrics the roofline / DESIGN numbers come from.
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/summary.md
"""
import csv
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'duration'), ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM %'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %'),
        ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1/TEX %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %'),
        ('sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'tensor(hmma) inst %'),
        ('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor pipe active %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
        ('launch__registers_per_thread', 'regs/thread'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('lts__t_sector_hit_rate.pct', 'L2 hit %')]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {}
    for i, h in enumerate(hdr):
        idx.setdefault(h, i)
        idx.setdefault(h.split('.', 1)[-1] if h.startswith(('TPC.', 'SM_')) else h, i)
    print(f'# ncu summary of `{path}` (--set full, --clock-control none)\n')
    for r in rows[2:]:
        name = r[idx['Kernel Name']]
        print(f'## {name[:110]}\n')
        print('| metric | value | unit |\n|---|---|---|')
        for key, label in WANT:
            j = idx.get(key)
            if j is None:
                cands = [i for i, h in enumerate(hdr) if h.endswith(key)]
                j = cands[0] if cands else None
            if j is not None:
                print(f'| {label} (`{key}`) | {r[j]} | {units[j]} |')
        print()


if __name__ == '__main__':
    main(sys.argv[1])
