"""Top CUDA source lines by warp instructions executed for one kernel of an .ncu-rep (needs -lineinfo, --import-source on).
   python scripts/ncu_lines.py rep "substr&substr" [top]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
res, cur_file, cur_fn, hdr = [], None, None, None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0] or not r[0].isdigit(): continue
    if not all(t in (cur_fn or "") for t in rx.split("&")): continue
    iX = hdr.index("Instructions Executed"); iN = hdr.index("# Samples")
    num = lambda v: int(v) if v.lstrip("-").isdigit() else 0
    res.append((num(r[iX]), num(r[iN]), cur_file, int(r[0]), r[1].strip()[:90]))
tot = sum(x[0] for x in res); ts = sum(x[1] for x in res)
print(f"total warp instructions {tot}, samples {ts}")
for n, sm, f, ln, src in sorted(res, reverse=True)[:top]:
    print(f"{100*n/tot:5.1f}% inst {100*sm/max(ts,1):5.1f}% smp  {f}:{ln}  {src}")
