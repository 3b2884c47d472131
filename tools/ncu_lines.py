"""Per-source-line summary of an ncu report's source page (needs -lineinfo and --import-source on).

  python tools/ncu_lines.py REPORT.ncu-rep KERNEL_ID [TOP]   (KERNEL_ID = 0-based ID column of the raw page)

Prints the lines of the kernel's CUDA source with the most warp-stall samples, the dominant stall reasons of
each, and the executed warp-instruction count — the view used to decide what to optimise next.
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kid = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-id', ':::%d' % (int(kid) + 1)],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    files = {}
    cur = None
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path':
            cur = r[1]
            continue
        if len(r) >= 2 and r[0] == 'Function Name':
            print('#', r[1][:120])
            continue
        if r and r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or not r or r[0] == '':
            continue
        files.setdefault(cur, []).append(r)
    col = {n: i for i, n in enumerate(hdr)}
    stall_cols = [(n, i) for n, i in col.items() if n.startswith('stall_') and 'Not Issued' not in n]
    samp = col['Warp Stall Sampling (All Samples)']
    inst = col['Instructions Executed']
    lines = []
    total = 0
    total_inst = 0
    for f, rs in files.items():
        for r in rs:
            try:
                s = int(r[samp])
                n = int(r[inst])
            except ValueError:
                continue
            total += s
            total_inst += n
            reasons = sorted(((int(r[i]) if r[i].isdigit() else 0, n_) for n_, i in stall_cols), reverse=True)[:3]
            lines.append((s, n, f, r[0], r[1].strip()[:110], reasons))
    lines.sort(reverse=True)
    print('total samples %d, warp instructions %d' % (total, total_inst))
    for s, n, f, ln, src, reasons in lines[:top]:
        why = ' '.join('%s=%d' % (n_.replace('stall_', ''), c) for c, n_ in reasons if c)
        print('%5.1f%% %8d inst  %s:%s  %s\n         [%s]' % (100.0 * s / max(total, 1), n, f.split('/')[-1], ln, src, why))


if __name__ == '__main__':
    main()
