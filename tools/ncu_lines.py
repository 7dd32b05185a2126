"""Aggregate executed instructions / stall samples per CUDA source line from an .ncu-rep (needs -lineinfo)."""
import csv, io, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; agg = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 7 and r[0].isdigit():
        try:
            agg.append((int(r[7]), int(r[4]) if r[4].isdigit() else 0, cur, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(a[0] for a in agg); samp = sum(a[1] for a in agg)
print("total warp-instructions", tot, " stall samples", samp)
for a in sorted(agg, reverse=True)[:top]:
    print("%10d %5.1f%%  samp %5.1f%%  %s:%d  %s" % (a[0], 100 * a[0] / tot, 100 * a[1] / max(samp, 1), a[2], a[3], a[4]))
