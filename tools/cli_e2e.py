"""dev probe: the allwave CLI end to end (FASTA in, PAF out) on BASELINE-shaped inputs; checks a sample against the oracle."""
import os, subprocess, sys, time, tempfile, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from allwave_b200 import synth
import oracle_lib as O
cfg, n = sys.argv[1], int(sys.argv[2])
length = int(sys.argv[3]) if len(sys.argv) > 3 else None
c, ids, seqs, rc = synth.config(cfg, n=n, length=length)
d = tempfile.mkdtemp()
fa, paf = os.path.join(d, "in.fa"), os.path.join(d, "out.paf")
with open(fa, "w") as f:
    for i, s in zip(ids, seqs):
        f.write(f">{i}\n{s.decode()}\n")
exe = os.path.join(ROOT, "allwave_b200", "allwave")
t0 = time.time()
subprocess.check_call([exe, "-i", fa, "-o", paf, "-p", c["sparsification"], "-s", c["scores"]])
dt = time.time() - t0
lines = open(paf).read().splitlines()
print(f"{cfg} n={n}: {len(lines)} PAF lines in {dt:.2f}s wall ({len(lines)/dt:.0f} pairs/s incl. FASTA parse, sketch, pair selection, PAF write), {os.path.getsize(paf)/1e6:.1f} MB")
sc = [int(x) for x in c["scores"].split(",")]
p = O.params(sc[0], sc[1], sc[2], sc[3], sc[4] if len(sc) > 4 else None, sc[5] if len(sc) > 5 else None)
idx = {name: i for i, name in enumerate(ids)}
rnd = random.Random(1)
bad = 0
sample = rnd.sample(lines, min(int(os.environ.get("AW_CHECK", "24")), len(lines)))
for ln in sample:
    f = ln.split("\t")
    q, t = idx[f[0]], idx[f[5]]
    o = O.align_pair(seqs[q], seqs[t], q, t, p, use_mash=True, qname=ids[q], tname=ids[t])
    bad += o["paf"] != ln
print(f"oracle check on {len(sample)} random lines: {len(sample)-bad} identical")
