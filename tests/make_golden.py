"""Regenerates tests/golden/*.json from the CPU oracle.

These fixtures are SELF-GENERATED: the reference (Rust + un-vendored WFA2-lib) cannot be run
in this image and its tests hold no exact CIGAR/score/hash vector, so parity with
WFA2-lib's tie-breaks stays "unpinned".  What the fixtures pin is (a) the published / hand-derived
anchors listed in `anchors` and (b) the oracle's own behaviour, so that neither the oracle nor
the CUDA path can drift silently.     usage: python tests/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
from allwave_b200 import synth  # noqa: E402

PENS = {"affine2p": (0, 5, 8, 2, 24, 1), "edit": (0, 1, 1, 1, None, None), "affine": (0, 4, 6, 2, None, None)}


def main():
    out = {"anchors": [], "alignments": [], "hashes": [], "pair_lists": []}
    # published / hand-derived anchors
    out["anchors"] = [
        dict(src="WFA2-lib README example (SURVEY A.9)", pen="affine", q="TCTTTACTCGCGCGTTGGAGAAATACAATAGT", t="TCTATACTGCGCGTTTGGAGAAATAAAATAGT",
             score=-24, ops="MMMXMMMMDMMMMMMMIMMMMMMMMMXMMMMMM", cg="3=1X4=1I7=1D9=1X6="),
        dict(src="tests/debug/check_wfa_ops.rs:15-16", pen="affine2p", q="ACGTACGTACGT", t="ACGTACGTTCGT", score=-5, ops="MMMMMMMMXMMM", cg="8=1X3="),
        dict(src="tests/debug/debug_cigar.rs:15-16 (12 vs 10 bases)", pen="affine2p", q="ACGTACGTACGT", t="ACGTACGTAC", score=-12, ops="MMMMMMMMMMDD", cg="10=2I"),
        dict(src="tests/debug/debug_cigar.rs (swapped)", pen="affine2p", q="ACGTACGTAC", t="ACGTACGTACGT", score=-12, ops="MMMMMMMMMMII", cg="10=2D"),
    ]
    out["hashes"] = [
        dict(src="SipHash-2-4 paper vector key=00..0f msg=00..0e", kind="siphash24", value="a129ca6149be45e5"),
        dict(src="Rust DefaultHasher::new().finish()", kind="empty13", value="d1fba762150c532c"),
        dict(src="SURVEY Appendix B (self-derived)", kind="bytes", msg="ACGTACGTACGTACG", value="b80df090b2f60358"),
        dict(src="SURVEY Appendix B (self-derived)", kind="bytes", msg="AAAAAAAAAAAAAAA", value="3a6b161f7b57a325"),
        dict(src="SURVEY Appendix B (self-derived)", kind="bytes", msg="ATCG", value="8d5c1dea1fb91a61"),
        dict(src="SURVEY Appendix B (self-derived)", kind="str", msg="seq1:seq2", value="dda85666d94b8609"),
        dict(src="SURVEY Appendix B (self-derived)", kind="str", msg="seq2:seq1", value="13dc9564ab31a805"),
        dict(src="SURVEY Appendix B (self-derived)", kind="str", msg="s000000:s000001", value="6358f4dc9d837732"),
    ]
    # oracle behaviour on seeded inputs (shapes of the five configs, reduced)
    cases = [("C1", 1, 3, 2500, 0.01, 0.0), ("C2", 2, 3, 1200, 0.05, 0.0), ("C3", 3, 6, 150, 0.02, 0.0), ("C5", 5, 4, 1500, 0.03, 0.5), ("short", 11, 4, 90, 0.08, 0.0)]
    for name, seed, n, length, d, rcp in cases:
        ids, seqs, rc = synth.generate(seed, n, length, d, rc_prob=rcp)
        for pen_name in ("affine2p", "edit"):
            p = O.params(*PENS[pen_name])
            for q in range(n):
                for t in range(n):
                    if q == t:
                        continue
                    r = O.align_pair(seqs[q], seqs[t], q, t, p, use_mash=True, qname=ids[q], tname=ids[t])
                    out["alignments"].append(dict(case=name, seed=seed, n=n, length=length, d=d, rc_prob=rcp, pen=pen_name, q=q, t=t,
                                                  score=r["score"], is_reverse=int(r["is_reverse"]), paf=r["paf"]))
    # pair lists
    ids = ["s%06d" % i for i in range(40)]
    for kind, frac in ((O.SPARS_RANDOM, 0.3), (O.SPARS_GIANT, 0.99), (O.SPARS_AUTO, 0.0)):
        pl = O.pair_list(ids, None, kind=kind, fraction=frac)
        out["pair_lists"].append(dict(n=40, kind=kind, fraction=frac, pairs=pl))
    _, seqs, _ = synth.generate(21, 12, 1200, 0.04)
    pl = O.pair_list(ids[:12], seqs, kind=O.SPARS_TREE, fraction=0.1, k_nearest=2, k_farthest=1)
    out["pair_lists"].append(dict(n=12, kind=O.SPARS_TREE, fraction=0.1, k_nearest=2, k_farthest=1, seed=21, length=1200, d=0.04, pairs=pl))
    with open(os.path.join(HERE, "golden", "oracle_golden.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote", len(out["alignments"]), "alignments,", len(out["pair_lists"]), "pair lists")


if __name__ == "__main__":
    main()
