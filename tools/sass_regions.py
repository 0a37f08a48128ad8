"""dev aid: static SASS size (KB) of one kernel by //@region of aw_wfa.cuh.  usage: python tools/sass_regions.py lib.so <kernel-substring>"""
import re, collections, subprocess, sys, os, tempfile, glob
lib, fn = os.path.abspath(sys.argv[1]), sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
cub = glob.glob(os.path.join(d, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "allwave_b200", "csrc", "aw_wfa.cuh")).read().split("\n")
marks = [(1, "top")]
for i, l in enumerate(src):
    if "//@region " in l:
        marks.append((i + 1, l.split("//@region ")[1].strip()))
on, cur = False, None
cnt = collections.Counter()
for l in dis.split("\n"):
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m:
        on = fn in m.group(1)
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        f, ln = m.group(1).split("/")[-1], int(m.group(2))
        cur = ([x[1] for x in marks if x[0] <= ln] or ["top"])[-1] if f == "aw_wfa.cuh" else "hdr:" + f
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
        cnt[cur] += 1
for k, v in cnt.most_common():
    print(f"{v*16/1024:8.1f} KB {k}")
print(f"{sum(cnt.values())*16/1024:8.1f} KB total")
