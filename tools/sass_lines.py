"""dev aid: static SASS instruction count per source line of one device function.
usage: python tools/sass_lines.py lib.so <function-substring> [min_count]"""
import re, collections, subprocess, sys, os, tempfile, glob
lib, fn = os.path.abspath(sys.argv[1]), sys.argv[2]
minc = int(sys.argv[3]) if len(sys.argv) > 3 else 4
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
cub = glob.glob(os.path.join(d, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "allwave_b200", "csrc", "aw_wfa.cuh")).read().split("\n")
cur, on, seen = None, False, False
cnt = collections.Counter()
ops = collections.Counter()
for l in dis.split("\n"):
    m = re.match(r"\s*\.type\s+(\S+),@function", l)
    if m:
        on = (fn in m.group(1)) and not seen
        if on: seen = True
        continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if m:
        cnt[cur] += 1
        ops[m.group(2).split(".")[0]] += 1
print("total", sum(cnt.values()))
for (f, ln), v in sorted(cnt.items(), key=lambda kv: (kv[0][0] != "aw_wfa.cuh", kv[0][1])):
    if v >= minc:
        print(f"{ln:5d} {v:4d}  {src[ln-1].strip()[:120] if f == 'aw_wfa.cuh' else f}")
print(ops.most_common(25))
