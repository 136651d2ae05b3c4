"""Key counters of `ncu --set full` captures exported with `--page raw --csv`:  python profiles/summarize_raw.py profiles/r02r_*_raw.csv"""
import csv
import sys

KEYS = [('gpu__time_duration.sum', 'duration'), ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM %'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor pipe % (elapsed)'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe % (active)'),
        ('sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'issue %'), ('sm__warps_active.avg.per_cycle_active', 'warps / SM'),
        ('launch__registers_per_thread', 'regs'), ('l1tex__t_sector_hit_rate.pct', 'L1 hit %'), ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block')]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        print(path, '(empty)')
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    name = vals[hdr.index('Kernel Name')][:60]
    out = []
    for k, lab in KEYS:
        if k in hdr:
            i = hdr.index(k)
            v = vals[i]
            try:
                v = f'{float(v.replace(",", "")):.4g}'
            except ValueError:
                pass
            out.append(f'{lab} {v} {units[i]}'.strip())
    print(f'{path.split("/")[-1]}: {name}\n    ' + ' | '.join(out))
