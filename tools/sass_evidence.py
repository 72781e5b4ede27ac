"""profiles/r2_ptxas_v.txt and profiles/r2_sass_counts.txt from the in-tree build: registers / spills per kernel (nvcc -Xptxas -v)
and SASS opcode counts per kernel (cuobjdump -sass).  usage: python tools/sass_evidence.py   (no GPU needed; ~2 min)"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = os.path.join(ROOT, "learning-based-mpc_b200")
out = subprocess.run(["make", "-C", pkg, "ptxas-info"], capture_output=True, text=True)
txt = out.stdout + out.stderr
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
lines, cur = ["nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xptxas -v (make -C learning-based-mpc_b200 ptxas-info)"], None
rows = {}
for l in txt.split("\n"):
    m = re.search(r"Function properties for (\S+)", l)
    if m:
        cur = re.sub(r"\(.*", "", demangle(m.group(1)))
        continue
    if cur and "bytes stack frame" in l:
        rows.setdefault(cur, {})["stack"] = l.strip()
    m = re.search(r"Used (\d+) registers, used (\d+) barriers(.*)", l)
    if m and cur:
        rows[cur]["regs"] = f"{m.group(1)} registers"; rows[cur]["bar"] = f"used {m.group(2)} barriers{m.group(3)}"; cur = None
for k, v in rows.items():
    lines.append(f"{k}: {v.get('regs')}; {v.get('stack')}; {v.get('bar')}")
open(os.path.join(ROOT, "profiles", "r2_ptxas_v.txt"), "w").write("\n".join(lines) + "\n")
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(pkg, "liblbmpc_b200.so")], capture_output=True, text=True).stdout
want = ["DFMA", "DMUL", "DADD", "MUFU", "FSEL", "SHFL", "LDS", "STS", "LD", "LDG", "STG", "LDL", "STL", "LDC", "UBLKCP", "SYNCS", "BAR", "FENCE", "ATOMG", "F2F", "BRA"]
lines = ["cuobjdump -sass liblbmpc_b200.so (sm_100a cubin built by learning-based-mpc_b200/Makefile), opcode counts per kernel"]
for blk in sass.split("Function : ")[1:]:
    name = re.sub(r"\(.*", "", demangle(blk.split("\n")[0].strip()))
    ops = collections.Counter()
    n = 0
    for l in blk.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            n += 1
            ops[m.group(1)] += 1
    lines.append(f"{name}: {n} SASS instructions; " + ", ".join(f"{o} {ops[o]}" for o in want if ops[o]))
open(os.path.join(ROOT, "profiles", "r2_sass_counts.txt"), "w").write("\n".join(lines) + "\n")
print(len(rows), "kernels")
