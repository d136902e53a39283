"""Per-source-line executed-instruction / stall-sample table from an ncu report captured with --import-source on.
  python scripts/ncu_source_hot.py <report.ncu-rep> [top_n]
"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur = None; hdr = None; agg = []
per_file = collections.Counter()
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
        try: n = int(r[ie]); s = int(r[isamp])
        except ValueError: continue
        agg.append((n, s, cur, r[0], r[1].strip()[:110])); per_file[cur] += n
tot = sum(a[0] for a in agg); tots = sum(a[1] for a in agg)
print(f"total warp-instructions {tot}, samples {tots}")
for f, n in per_file.most_common(): print(f"  {f:24s} {n/tot*100:5.1f}%")
for n, s, f, ln, src in sorted(agg, reverse=True)[:top]:
    print(f"{n/tot*100:5.1f}% inst {s/max(tots,1)*100:5.1f}% samp  {f}:{ln:>4s}  {src}")
