"""Per-kernel warp-stall breakdown (ratio per issue) from an .ncu-rep: python scripts/ncu_stalls.py rep [regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
ki = hdr.index("Kernel Name")
cols = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    if pat and not pat.search(r[ki]):
        continue
    vals = sorted([(float(r[i].replace(",", "")), hdr[i][len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for i in cols if r[i]], reverse=True)[:18]
    print(r[ki][:70])
    print("    " + ", ".join(f"{n}={v:.2f}" for v, n in vals))
