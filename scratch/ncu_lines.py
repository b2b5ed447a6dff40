"""Aggregate ncu warp-stall samples per CUDA source line: python scratch/ncu_lines.py report.ncu-rep [top]"""
import collections, csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True).stdout.decode(errors="replace")
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None; agg = collections.Counter(); txt = {}; stall = collections.defaultdict(collections.Counter)
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; ci = {}; [ci.setdefault(n, i) for i, n in enumerate(hdr)]; continue
    if r and r[0] == "Function Name": continue
    if hdr and len(r) > ci["# Samples"]:
        try: n = int(r[ci["# Samples"]] or 0)
        except ValueError: continue
        key = (cur, r[0]); agg[key] += n; txt[key] = r[1].strip()
tot = sum(agg.values()); print("total samples", tot)
for key, n in agg.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{n:7d} {100*n/tot:5.1f}%  {key[0]}:{key[1]:>4}  {txt[key][:110]}")
