"""Summarise an `ncu --page raw --csv` export: one line per launch with the metrics the roofline discussion uses.

    ncu -i X.ncu-rep --page raw --csv > X_raw.csv ; python tools/ncu_summary.py X_raw.csv
"""
import csv
import sys

r = list(csv.reader(open(sys.argv[1])))
hdr, units, rows = r[0], r[1], r[2:]
idx = {h: i for i, h in enumerate(hdr)}


def f(row, h):
    try:
        return float(row[idx[h]].replace(',', ''))
    except (ValueError, KeyError):
        return float('nan')


cols = [('gpu__time_duration.sum', 'us'), ('launch__registers_per_thread', 'regs'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('smsp__issue_active.avg.pct', 'issue%'),
        ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma%'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
        ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu%'),
        ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu%'),
        ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'alu%'),
        ('dram__bytes_read.sum', 'rd'), ('dram__bytes_write.sum', 'wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('lts__t_bytes.sum', 'l2')]
print(' '.join(f"{n:>8s}" for _, n in cols), 'kernel   [units: rd/wr %s, l2 %s]' % (
    units[idx['dram__bytes_read.sum']], units[idx.get('lts__t_bytes.sum', 0)]))
tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(units[idx['gpu__time_duration.sum']], 1.0)
for row in rows:
    vals = [f(row, c) * (tscale if c == 'gpu__time_duration.sum' else 1.0) for c, _ in cols]
    print(' '.join(f"{v:8.1f}" for v in vals), row[idx['Kernel Name']][:44])
