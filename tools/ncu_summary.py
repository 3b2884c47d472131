"""Markdown table from an ncu report (raw page): one row per captured launch with the metrics the roofline uses.

  python tools/ncu_summary.py REPORT.ncu-rep|RAW.csv > profiles/<name>.md
"""
import csv
import io
import subprocess
import sys

COLS = [('gpu__time_duration.sum', 'ms', 1e-6), ('launch__grid_size', 'grid', 1), ('launch__block_size', 'block', 1),
        ('launch__registers_per_thread', 'regs', 1), ('dram__bytes_read.sum', 'dram_rd_MB', None), ('dram__bytes_write.sum', 'dram_wr_MB', None),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_%', 1), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_act_%', 1),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pipe_act_%', 1),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor_pipe_elapsed_%', 1),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram_%', 1), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_%', 1)]


def to_mb(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(unit, 1.0)


def main():
    if sys.argv[1].endswith('.csv'):                 # already exported with `ncu -i X.ncu-rep --page raw --csv`
        out = open(sys.argv[1]).read()
    else:
        out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print('| # | kernel | ' + ' | '.join(c[1] for c in COLS) + ' | dram_GB/s |')
    print('|---|---|' + '---|' * (len(COLS) + 1))
    tot_ms = tot_mb = 0.0
    for n, r in enumerate(body):
        cells = []
        row_ms = row_mb = 0.0
        for name, label, scale in COLS:
            if name not in idx:
                cells.append('')
                continue
            v, u = r[idx[name]], units[idx[name]]
            if scale is None:
                x = to_mb(v, u)
                tot_mb += x
                row_mb += x
                cells.append('%.2f' % x)
            elif label == 'ms':
                x = float(v.replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}.get(u, 1e-6)
                tot_ms += x
                row_ms = x
                cells.append('%.4f' % x)
            elif label in ('grid', 'block', 'regs'):
                cells.append(v.replace(',', '').split('.')[0])
            else:
                cells.append('%.1f' % float(v.replace(',', '')))
        cells.append('%.0f' % (row_mb / row_ms) if row_ms > 0 else '')      # MB / ms = GB/s achieved DRAM traffic
        print('| %d | %s | ' % (n, r[idx['Kernel Name']][:48]) + ' | '.join(cells) + ' |')
    print()
    print('%d launches, %.3f ms, DRAM read+write %.1f MB (%.1f MB per launch)' % (len(body), tot_ms, tot_mb, tot_mb / max(len(body), 1)))


if __name__ == '__main__':
    main()
