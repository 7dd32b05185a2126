"""Executed warp-instructions per source-line RANGE of an .ncu-rep (ranges given as file:lo-hi[=label] ...)."""
import csv, io, subprocess, sys
path = sys.argv[1]
walkers = float(sys.argv[2])
ranges = []
for a in sys.argv[3:]:
    spec, _, label = a.partition("=")
    f, _, r = spec.partition(":")
    lo, _, hi = r.partition("-")
    ranges.append((f, int(lo), int(hi), label or spec))
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None
agg = {}
tot = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 7 and r[0].isdigit():
        try:
            n = int(r[7]); s = int(r[4]) if r[4].isdigit() else 0
        except ValueError:
            continue
        tot += n
        for f, lo, hi, label in ranges:
            if cur == f and lo <= int(r[0]) <= hi:
                a = agg.setdefault(label, [0, 0]); a[0] += n; a[1] += s
                break
        else:
            a = agg.setdefault("other:" + str(cur), [0, 0]); a[0] += n; a[1] += s
print("attributed total %.1f K/walker" % (tot / walkers / 1e3))
for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-40s %8.2f K/walker  %5.1f%%   samples %d" % (k, n / walkers / 1e3, 100 * n / tot, s))
