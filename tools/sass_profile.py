"""Per-function and per-opcode breakdown of an .ncu-rep SASS page (ncu -i X --page source --csv --print-source sass).

    python tools/sass_profile.py gpurun_out/X.ncu-rep [walkers]

Groups SASS instructions by the function they sit in (the kernel and every __noinline__ device function appear as
separate address ranges, split at RET/EXIT boundaries is not reliable, so functions are told apart by address gaps)
and prints executed warp-instructions, stall samples and the top opcodes of each range."""
import csv, io, subprocess, sys, re
from collections import defaultdict

path = sys.argv[1]
walkers = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
insts = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].startswith("0x"):
        insts.append((int(r[0], 16), r[1].strip(), int(r[hdr.index("# Samples")] or 0), int(r[hdr.index("Instructions Executed")] or 0),
                      int(r[hdr.index("Thread Instructions Executed")] or 0), int(r[hdr.index("L1 Wavefronts Shared")] or 0),
                      int(r[hdr.index("L1 Wavefronts Shared Ideal")] or 0)))
tot = sum(i[3] for i in insts)
samp = sum(i[2] for i in insts)
print("instructions (SASS)", len(insts), " executed warp-instructions", tot, " samples", samp)
if walkers:
    print("per walker: %.1f K warp-instructions" % (tot / walkers / 1e3))
# opcode histogram
ops = defaultdict(lambda: [0, 0])
for a, s, sm, ex, th, w, wi in insts:
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", s)
    op = m.group(2).split(".")[0] if m else "?"
    ops[op][0] += ex
    ops[op][1] += sm
print("\ntop opcodes: op, executed, share, stall-sample share")
for op, (ex, sm) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:28]:
    print("  %-10s %14d %5.1f%%  %5.1f%%" % (op, ex, 100 * ex / tot, 100 * sm / max(samp, 1)))
# shared-memory wavefront excess
wv = sum(i[5] for i in insts); wvi = sum(i[6] for i in insts)
print("\nshared wavefronts %d  ideal %d  excess %.1f%%" % (wv, wvi, 100 * (wv - wvi) / max(wv, 1)))
print("worst shared-memory instructions (excess wavefronts):")
for i in sorted(insts, key=lambda i: -(i[5] - i[6]))[:12]:
    print("  %x  %-60s exec %12d  wavefronts %12d ideal %12d" % (i[0], i[1][:60], i[3], i[5], i[6]))
