"""profiles/r2_ncu_traffic.json from an `ncu --set full` capture of one bench step (raw page csv or .ncu-rep): mean
dram__bytes_read.sum + dram__bytes_write.sum per launch (MB) for the bench's stage keys.  bench.py reads the file at run
time for `roofline.traffic`, so the number in the JSON line comes from a committed ncu capture, not from a literal.

  python tools/ncu_traffic.py RAW.csv|REPORT.ncu-rep > profiles/r2_ncu_traffic.json
"""
import csv
import io
import json
import subprocess
import sys

FAMILIES = {'net_2d/conv3x3': 'tc_conv3x3_', 'net_2d/conv_general': 'tc_convg_kernel', 'fused_sa_fa_tc2': 'tc2_kernel', 'fused_mlp_tc': 'tc_fused_mlp_kernel'}


def to_mb(v, unit):
    return float(v.replace(',', '')) * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(unit, 1.0)


def main():
    src = sys.argv[1]
    out = open(src).read() if src.endswith('.csv') else subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    res = {'_source': src, '_metric': 'dram__bytes_read.sum + dram__bytes_write.sum, MB per launch (mean over the captured launches)'}
    for key, pat in FAMILIES.items():
        mbs = [to_mb(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']]) + to_mb(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
               for r in body if pat in r[idx['Kernel Name']]]
        if mbs:
            res[key] = round(sum(mbs) / len(mbs), 2)
            res[key + '_launches'] = len(mbs)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
