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
def line_map(dis, ksub):
    import re
    lines = open(dis).read().split('\n')
    start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and ksub in l and l.endswith(':'))
    cur, out = None, []
    for l in lines[start + 1:]:
        if l.startswith('//--------------------- .') and out: break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l): out.append(cur)
    return out
if len(sys.argv) == 4 and not sys.argv[2].isdigit():
    lm = line_map(sys.argv[2], sys.argv[3])
    import collections
    segs, cur = [], None
    for i, r in enumerate(ins):
        n, s = int(r[ia]), int(r[isamp])
        if cur and cur[0] == n: cur[2] += 1; cur[3] += s
        else:
            cur = [n, i, 1, s]; segs.append(cur)
    for c in sorted(sorted(segs, key=lambda c: -c[3])[:40], key=lambda c: c[1]):
        cnt = collections.Counter(lm[i] for i in range(c[1], c[1] + c[2]) if i < len(lm))
        top = ', '.join(f'{f}:{ln}x{n}' for (f, ln), n in cnt.most_common(3) if f)
        print(f'start {c[1]:6d} len {c[2]:4d} exec {c[0]:9d} samples {c[3]:6d} ({100*c[3]/tot_s:4.1f}%)  {top}')
elif len(sys.argv) < 4:
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
