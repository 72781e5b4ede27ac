"""Per-instruction stall samples of the hottest SASS region of an ncu source page (csv).
usage: ncu_hot.py src.csv [first_instr last_instr]   (without a range: lists the hot segments)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isamp, isrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
cols = [c for c in hdr if c.startswith('stall_') and 'Not Issued' not in c]
ci = [hdr.index(c) for c in cols]
ins = rows[2:]
tot_s = sum(int(r[isamp]) for r in ins)
if len(sys.argv) < 4:
    segs, cur = [], None
    for i, r in enumerate(ins):
        n, s = int(r[ia]), int(r[isamp])
        if cur and cur[0] == n: cur[2] += 1; cur[3] += s
        else:
            cur = [n, i, 1, s]; segs.append(cur)
    for c in sorted(sorted(segs, key=lambda c: -c[3])[:30], key=lambda c: c[1]):
        print(f'start {c[1]:6d} len {c[2]:4d} exec {c[0]:9d} samples {c[3]:6d} ({100*c[3]/tot_s:4.1f}%)')
else:
    a, b = int(sys.argv[2]), int(sys.argv[3])
    unit = None
    for i in range(a, b):
        r = ins[i]
        st = ' '.join(f'{c[6:]}={r[j]}' for c, j in zip(cols, ci) if r[j] not in ('0', ''))
        print(f'{i:6d} {r[isrc].strip()[:58]:58s} {int(r[ia]):8d} {int(r[isamp]):5d}  {st}')
    print('segment samples', sum(int(ins[i][isamp]) for i in range(a, b)), 'of', tot_s)
