"""dev probe: the CLI on a BASELINE-shaped input with the PAF going to a file and to /dev/null (is the writer the limit?)
usage: python tools/cli_devnull.py C3 1415"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from allwave_b200 import synth
cfg, n = sys.argv[1], int(sys.argv[2])
c, ids, seqs, rc = synth.config(cfg, n=n)
d = tempfile.mkdtemp()
fa = os.path.join(d, "in.fa")
with open(fa, "w") as f:
    for i, s in zip(ids, seqs):
        f.write(f">{i}\n{s.decode()}\n")
exe = os.path.join(ROOT, "allwave_b200", "allwave")
for out in (os.path.join(d, "out.paf"), "/dev/null", "/dev/shm/allwave_probe.paf"):
    print("== -o", out, flush=True)
    subprocess.call([exe, "-i", fa, "-o", out, "-p", c["sparsification"], "-s", c["scores"]], env=dict(os.environ, ALLWAVE_TIMING="1"))
if os.path.exists("/dev/shm/allwave_probe.paf"):
    os.remove("/dev/shm/allwave_probe.paf")
