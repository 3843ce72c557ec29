"""Top source lines by warp-stall samples from an `ncu --set full --import-source on` report.
usage: python tools/ncu_lines.py report.ncu-rep [kernel-regex] [top N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]
if len(sys.argv) > 2 and sys.argv[2] != "-":
    cmd += ["-k", "regex:" + sys.argv[2]]
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname = func = None
hdr = None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name":
        func = r[1][:40]; continue
    if r[0] == "Line No":
        hdr = r; idx = {}
        for i, k in enumerate(hdr):
            idx.setdefault(k, i)
        stall = [(k, i) for i, k in enumerate(hdr) if k.startswith("stall_") and "Not Issued" not in k]
        continue
    if hdr is None or r[2] != "-":          # keep the per-CUDA-line summary rows (Address == "-")
        continue
    try:
        n = int(r[idx["# Samples"]])
    except ValueError:
        continue
    key = (func, fname, r[0], r[1].strip()[:100])
    a = agg.setdefault(key, [0, {}, 0])
    a[0] += n
    a[2] += int(r[idx["Instructions Executed"]] or 0)
    for k, i in stall:
        a[1][k] = a[1].get(k, 0) + int(r[i] or 0)
tot = sum(a[0] for a in agg.values())
print(f"total samples {tot}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    st = sorted(a[1].items(), key=lambda kv: -kv[1])[:3]
    print(f"{a[0]:7d} {100*a[0]/max(tot,1):5.1f}%  {key[1]}:{key[2]:>4}  inst {a[2]:9d}  {key[3][:80]:80s} {[(k[6:], v) for k, v in st if v]}")
