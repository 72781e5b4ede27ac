"""Join an ncu SASS source page (csv) with nvdisasm line info and aggregate by source function.
usage: ncu_by_line.py src.csv dis.txt kernel_mangled_substr"""
import csv, re, sys, collections
src_csv, dis, ksub = sys.argv[1:4]
# --- nvdisasm: list of (file,line) per instruction of the kernel, tracking inlined-at chains
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and ksub in l and l.endswith(':'))
cur = None; per_instr = []
for l in lines[start + 1:]:
    if l.startswith('//--------------------- .text.') or l.startswith('//--------------------- .'):
        if per_instr: break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        if 'inlined at' in m.group(3) and cur is not None and False:
            pass
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l):
        per_instr.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); isrc = hdr.index('Source')
ins = rows[2:]
print('sass instrs: ncu', len(ins), 'nvdisasm', len(per_instr))
# function ranges in core.cuh by line number
def func_of(f, ln):
    if f is None: return 'unknown'
    return f'{f}:{ln}'
agg = collections.defaultdict(lambda: [0, 0, 0])
for r, li in zip(ins, per_instr):
    k = li
    agg[k][0] += int(r[ia]); agg[k][1] += int(r[isamp]); agg[k][2] += 1
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print('total warp-instr', tot_i, 'samples', tot_s)
# bucket into functions using line ranges from the source file
import os
def load_funcs(path):
    out = []
    for i, l in enumerate(open(path), 1):
        m = re.match(r'\s*static LB_HD \w[\w<>:, ]* (\w+)\(', l) or re.match(r'\s*(?:template.*)?__device__ __forceinline__ \w+ (\w+)\(', l)
        if m: out.append((i, m.group(1)))
    return out
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'learning-based-mpc_b200', 'csrc')
funcs = {'lbmpc_core.cuh': load_funcs(os.path.join(root, 'lbmpc_core.cuh')), 'lbmpc_kernels.cuh': load_funcs(os.path.join(root, 'lbmpc_kernels.cuh'))}
byf = collections.defaultdict(lambda: [0, 0, 0])
for (k, v) in agg.items():
    if k is None: name = 'unknown'
    else:
        f, ln = k; name = f
        for s, n in funcs.get(f, []):
            if ln >= s: name = f'{f}:{n}'
    for j in range(3): byf[name][j] += v[j]
for name, v in sorted(byf.items(), key=lambda kv: -kv[1][1]):
    print(f'{name:45s} instr {v[0]:>11d} ({100*v[0]/tot_i:5.1f}%)  samples {v[1]:>7d} ({100*v[1]/max(tot_s,1):5.1f}%)  sass {v[2]}')
