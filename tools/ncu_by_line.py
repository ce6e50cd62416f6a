"""Per-CUDA-source-line view of an ncu capture: warp instructions executed and stall samples aggregated over the SASS of each
line (ncu's own source page only lists them per SASS instruction).  Needs a report captured with `--import-source on` from a
library built with -lineinfo, and the sources at the same revision on disk (line numbers are matched against the files here).

usage: python tools/ncu_by_line.py REPORT.ncu-rep UNITS [min_instr] [file-substring ...]
       UNITS = number of work items of the launch (frames, columns ...): counts are printed per unit.

This is the view that found the integer divisions in the mixed-radix passes and in the inverse kernel's gather-sum, the 64-bit
index multiplies per store and the multi-channel loader's per-sample bounds test (profiles/r2_small_experiments.md)."""
import csv
import io
import os
import subprocess
import sys


def main():
    rep, units = sys.argv[1], float(sys.argv[2])
    min_instr = float(sys.argv[3]) if len(sys.argv) > 3 else 5.0
    wanted = sys.argv[4:]
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    for block in txt.split('"File Path",')[1:]:
        lines = block.split('\n')
        fname = lines[0].strip().strip('"')
        rows = list(csv.reader(io.StringIO('\n'.join(lines[2:]))))
        if not rows or '# Samples' not in rows[0]:
            continue
        i_s, i_e = rows[0].index('# Samples'), rows[0].index('Instructions Executed')
        agg, cur = {}, None
        for r in rows[1:]:
            if len(r) <= i_e:
                continue
            if r[0] not in ('', '-'):
                try:
                    cur = int(r[0])
                except ValueError:
                    pass
            if r[2] in ('-', ''):          # the CUDA source row itself; its SASS rows follow
                continue
            try:
                s, e = int(r[i_s]), int(r[i_e])
            except ValueError:
                continue
            a = agg.setdefault(cur, [0, 0])
            a[0] += s
            a[1] += e
        total = sum(v[1] for v in agg.values())
        if not total or not agg:
            continue
        # ncu labels the blocks by file, but the label order has been seen to lag: identify the file by its length instead
        guess = fname
        if not os.path.exists(fname) or max(agg) > sum(1 for _ in open(fname, errors='replace')):
            guess = fname + ' (label may be off: check the line numbers)'
        if wanted and not any(w in fname for w in wanted):
            continue
        print('== %s: %.0f instructions per unit, %d samples' % (guess, total / units, sum(v[0] for v in agg.values())))
        src = open(fname, errors='replace').read().split('\n') if os.path.exists(fname) else []
        for ln in sorted(agg):
            s, e = agg[ln]
            if e / units >= min_instr or s >= 100:
                text = src[ln - 1].strip()[:110] if 0 < ln <= len(src) else ''
                print('  %5d %6d samples %8.1f  %s' % (ln, s, e / units, text))


if __name__ == '__main__':
    main()
