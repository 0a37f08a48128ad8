"""Generates tests/golden/c4_pairs.json.gz: oracle results (score, strand, whole PAF line) for a few pairs of BASELINE config 4
exactly as generated (200 x 1 Mb haplotypes, seed 4, divergence 0.1-2 % + SVs; the first pairs of the `giant:0.99` list).
The oracle needs many minutes per 1 Mb pair, so the GPU suite compares against this frozen fixture instead of re-running it
(the fixture is produced by the oracle in its plain checker mode unless --fast is given; both modes are tested to agree).
usage: python tests/make_golden_c4.py [npairs] [--fast]"""
import gzip
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import oracle_lib as O  # noqa: E402
from allwave_b200 import synth  # noqa: E402

npairs = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 3
fast = "--fast" in sys.argv
c, ids, seqs, _ = synth.config("C4")
full = O.pair_list(ids, seqs, kind=O.SPARS_GIANT, fraction=0.99)
pairs = [tuple(p) for p in full[:npairs]]
print("pairs", pairs, "of", len(full), flush=True)
t0 = time.time()
r = O.run_pairs(ids, seqs, pairs, O.params(0, 5, 8, 2, 24, 1), use_mash=True, threads=min(len(pairs), os.cpu_count() or 1), fast=fast)
print("oracle seconds", time.time() - t0, r["scores"], r["work"], flush=True)
out = {"config": "C4", "seed": c["seed"], "n": c["n"], "length": c["length"], "scores": "0,5,8,2,24,1", "sparsification": "giant:0.99",
       "n_pairs_in_list": len(full), "pair_list_sha1": hashlib.sha1(json.dumps(full).encode()).hexdigest(),
       "oracle_mode": "fast" if fast else "plain", "pairs": [{"q": q, "t": t, "score": s, "paf": paf} for (q, t), s, paf in zip(pairs, r["scores"], r["paf"])]}
with gzip.open(os.path.join(HERE, "golden", "c4_pairs.json.gz"), "wt") as f:
    json.dump(out, f)
print("written", flush=True)
