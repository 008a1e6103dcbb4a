"""Opcode histogram (warp instructions executed) + sample share of one kernel from an .ncu-rep (SASS source page).
   python scripts/ncu_sass_hist.py rep kernel_regex [top]"""
import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]   # rx: substring that the demangled kernel name must contain
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
heads = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')] + [len(lines)]
sel = [j for j in range(len(heads) - 1) if all(t in lines[heads[j]] for t in rx.split("&"))]
start = [heads[sel[0]], heads[sel[0] + 1]]
blk = lines[start[0] + 1:start[1]]
rows = list(csv.reader(io.StringIO("\n".join(blk))))
hdr = rows[0]
iS, iN, iX = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
hist, samp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[1:]:
    if len(r) <= iX: continue
    op = r[iS].split()
    if not op: continue
    name = op[1] if op[0].startswith("@") else op[0]
    name = name.split(".")[0] + ("." + name.split(".")[1] if name.startswith(("LDS", "STS", "LDG", "STG", "SHFL")) and "." in name else "")
    n = int(r[iX] or 0)
    hist[name] += n
    samp[name] += int(r[iN] or 0)
    tot += n
ts = sum(samp.values())
print(lines[start[0]][:120])
print(f"total warp instructions {tot}, samples {ts}, static instructions {len(rows)-1}")
for k, v in hist.most_common(top):
    print(f"  {k:14s} {v:10d} {100*v/tot:5.1f}%   samples {100*samp[k]/max(ts,1):5.1f}%")
