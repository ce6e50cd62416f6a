"""Summarise an ncu '--page source --csv' dump: opcode mix, stall totals, hottest SASS lines."""
import csv
import sys
from collections import Counter

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
op_inst, op_samp = Counter(), Counter()
stall_tot = Counter()
tot_inst = 0
lines = []
for r in body:
    if len(r) < len(hdr):
        continue
    src = r[ix['Source']].strip()
    toks = src.split()
    op = toks[0] if toks else '?'
    if op.startswith('@') and len(toks) > 1:
        op = toks[1]
    parts = op.split('.')
    op = parts[0] + ('.' + parts[1] if parts[0] in ('LDS', 'STS', 'LDG', 'STG') and len(parts) > 1 else '')
    n = int(r[ix['Instructions Executed']] or 0)
    s = int(r[ix['# Samples']] or 0)
    op_inst[op] += n
    op_samp[op] += s
    tot_inst += n
    for c in stall_cols:
        stall_tot[c] += int(r[ix[c]] or 0)
    lines.append((s, n, src, r))
tot_samp = sum(op_samp.values())
print('total warp instructions %d, samples %d' % (tot_inst, tot_samp))
print('--- opcode mix (inst%, sample%)')
for op, n in op_inst.most_common(28):
    print('%-12s %6.2f%% %6.2f%%' % (op, 100.0 * n / tot_inst, 100.0 * op_samp[op] / max(tot_samp, 1)))
print('--- stall reasons (all samples)')
ts = sum(stall_tot.values())
for c, v in stall_tot.most_common(12):
    print('%-28s %6.2f%%' % (c, 100.0 * v / max(ts, 1)))
print('--- hottest lines')
for s, n, src, r in sorted(lines, key=lambda t: -t[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    top = sorted(((int(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    extra = ''
    if r[ix['L1 Wavefronts Shared']] not in ('', '0'):
        extra = ' smem wf %s ideal %s' % (r[ix['L1 Wavefronts Shared']], r[ix['L1 Wavefronts Shared Ideal']])
    print('%6d %9d  %-60s %s%s' % (s, n, src[:60], ','.join('%s=%d' % (c[6:], v) for v, c in top), extra))
