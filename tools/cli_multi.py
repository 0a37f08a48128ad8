"""dev probe: allwave CLI on C2-shaped input with --gpus 1 / N; compares the sorted PAF."""
import os, subprocess, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from allwave_b200 import synth
n, G = int(sys.argv[1]), int(sys.argv[2])
c, ids, seqs, rc = synth.config("C2", n=n)
d = tempfile.mkdtemp()
fa = os.path.join(d, "in.fa")
open(fa, "w").write("".join(f">{i}\n{s.decode()}\n" for i, s in zip(ids, seqs)))
exe = os.path.join(ROOT, "allwave_b200", "allwave")
res = {}
for g in (1, G):
    out = os.path.join(d, f"o{g}.paf")
    t0 = time.time()
    subprocess.check_call([exe, "-i", fa, "-o", out, "-p", "none", "--gpus", str(g)])
    res[g] = (time.time() - t0, sorted(open(out).read().splitlines()))
    print(f"--gpus {g}: {len(res[g][1])} lines, {res[g][0]:.2f}s wall")
print("identical:", res[1][1] == res[G][1])
