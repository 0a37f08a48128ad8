"""dev aid: aggregate an ncu source page (cuda,sass view) of aw_align_kernel by code region and by line.
usage: python tools/ncu_regions.py report.ncu-rep [top_lines]"""
import collections, csv, io, subprocess, sys, os

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows[hi + 1:]:
    try:
        data.append((int(r[0]), int(r[iS] or 0), int(r[iI] or 0), r[1]))
    except (ValueError, IndexError):
        pass
ti, ts = sum(d[2] for d in data) or 1, sum(d[1] for d in data) or 1
print("warp instructions", ti, "samples", ts)
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "allwave_b200", "csrc", "aw_wfa.cuh")).read().split("\n")
marks = [(1, "top")]
strip = "--old" in sys.argv  # report taken before the markers were added: count lines without them
nm = 0
for i, l in enumerate(src):
    if "//@region " in l:
        marks.append((i + 1 - (nm if strip else 0), l.split("//@region ")[1].strip()))
        nm += 1
agg = collections.OrderedDict((m[1], [0, 0]) for m in marks)
for ln, s, n, _ in data:
    name = ([m[1] for m in marks if m[0] <= ln] or ["top"])[-1]
    agg[name][0] += s
    agg[name][1] += n
for k, (s, n) in agg.items():
    print(f"{k:28s} inst {100*n/ti:5.1f}%  samples {100*s/ts:5.1f}%")
byline = collections.defaultdict(lambda: [0, 0, ""])
for ln, s, n, t in data:
    byline[ln][0] += s
    byline[ln][1] += n
    byline[ln][2] = t or byline[ln][2]
print("-- top lines by instructions")
for ln, (s, n, t) in sorted(byline.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{ln:>6} inst {100*n/ti:5.1f}% samp {100*s/ts:5.1f}%  {t.strip()[:100]}")
print("-- top lines by samples")
for ln, (s, n, t) in sorted(byline.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ln:>6} inst {100*n/ti:5.1f}% samp {100*s/ts:5.1f}%  {t.strip()[:100]}")
