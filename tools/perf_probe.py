"""dev probe: time the device-resident path on C2-shaped batches (not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import allwave_b200 as aw
from allwave_b200 import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
nl = (sys.argv[2] if len(sys.argv) > 2 else "40").split(":")
nseq = int(nl[0])
length = int(nl[1]) if len(nl) > 1 else None
sizes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [148, 296, 592]
opts = dict(kv.split("=") for kv in sys.argv[4:] if kv != "check")
check = "check" in sys.argv[4:]
c, ids, seqs, rc = synth.config(cfg, n=nseq, length=length)
ctx = aw.Context(0)
for k, v in opts.items():
    ctx.set_option(k, int(v))
t0 = time.time(); ctx.load_sequences(ids, seqs); print("load %.3fs" % (time.time() - t0))
pairs = [(i, j) for i in range(nseq) for j in range(nseq) if i != j]
sc = [int(x) for x in c["scores"].split(",")]
p = aw.make_params(sc[0], sc[1], sc[2], sc[3], sc[4] if len(sc) > 4 else None, sc[5] if len(sc) > 5 else None)
for B in sizes:
    b = aw.Batch(ctx, p, pairs[:B], flags=0)
    for it in range(2):
        t0 = time.time(); b.launch(); res = b.fetch(collect=False); t = time.time() - t0
        st = b.stats(); kms = b.kernel_ms()
        dc = b.debug_cycles(); tot = max(1, sum(dc.values()))
        print('   cycles%:', {k: round(100.0 * v / tot, 1) for k, v in dc.items()}, 'Mcyc/pair', round(tot / B / 1e6, 1))
        print(f"{cfg} B={B} it={it} wall={t:.3f}s kernel={kms:.1f}ms pairs/s={B/(kms/1e3):.1f} cells/s={st['cells']/(kms/1e3):.3e} steps={st['steps']} retried={st['pairs_retried']} failed={st['failed_pairs']} paf_bytes={st['paf_bytes']}")
    b.close()

if check:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import oracle_lib as O
    sample = pairs[: min(len(pairs), 16)]
    res = ctx.align_pairs(p, sample)
    op = O.params(sc[0], sc[1], sc[2], sc[3], sc[4] if len(sc) > 4 else None, sc[5] if len(sc) > 5 else None)
    t0 = time.time()
    r = O.run_pairs(ids, seqs, sample, op, use_mash=True, threads=os.cpu_count())
    same = sum(1 for g, e in zip(res, r["paf"]) if g["paf"] == e)
    print(f"oracle check: {same}/{len(sample)} PAF identical, oracle {time.time()-t0:.1f}s, scores {[g['score'] for g in res][:6]}")
