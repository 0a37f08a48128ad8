"""dev probe: Mb-scale pairs (BASELINE config 4 shape) through the device-resident path.
usage: python tools/c4_probe.py <nseq> <length> <npairs> [key=value ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import allwave_b200 as aw
from allwave_b200 import synth

nseq, length, npairs = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
opts = dict(kv.split("=") for kv in sys.argv[4:])
c, ids, seqs, _ = synth.config("C4", n=nseq, length=length)
ctx = aw.Context(0)
for k, v in opts.items():
    ctx.set_option(k, int(v))
ctx.load_sequences(ids, seqs)
pairs = [(i, j) for i in range(nseq) for j in range(nseq) if i != j][:npairs]
est = ctx.estimate_divergence(pairs)
p = aw.make_params(0, 5, 8, 2, 24, 1)
b = aw.Batch(ctx, p, pairs, flags=0)
t0 = time.time(); b.launch(); res = b.fetch(collect=True); t = time.time() - t0
st = b.stats(); kms = b.kernel_ms()
print(f"C4 n={nseq} L={length} pairs={len(pairs)} opts={opts} wall={t:.2f}s kernel={kms:.0f}ms pairs/s={len(pairs)/(kms/1e3):.3f} cells/s={st['cells']/(kms/1e3):.3e} "
      f"steps={st['steps']} retried={st['pairs_retried']} failed={st['failed_pairs']} paf_bytes={st['paf_bytes']}")
print("scores", [r["score"] for r in res][:12], "est_div", [round(e, 4) for e in est][:12])
import hashlib
print("paf sha1", hashlib.sha1("\n".join(r["paf"] for r in res).encode()).hexdigest())
b.close(); ctx.close()
