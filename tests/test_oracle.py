"""CPU tests of the oracle: published / hand-derived anchors, the committed golden fixtures,
every property the reference's own tests assert at this boundary, and score
optimality against the independent Gotoh DP."""
import json
import os
import random

import pytest

from allwave_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "oracle_golden.json")))
PENS = {"affine2p": (0, 5, 8, 2, 24, 1), "edit": (0, 1, 1, 1, None, None), "affine": (0, 4, 6, 2, None, None)}


def walk_cigar(ops, q, t):
    """validate_alignment_simple (src/validation_simple.rs:73-161): every M matches, every X differs,
    the CIGAR consumes both sequences fully"""
    i = j = 0
    for op in ops:
        c = chr(op)
        if c == "M":
            assert q[i] == t[j]
            i += 1
            j += 1
        elif c == "X":
            assert q[i] != t[j]
            i += 1
            j += 1
        elif c == "I":  # WFA2: consumes text/target
            j += 1
        elif c == "D":  # WFA2: consumes pattern/query
            i += 1
        else:
            raise AssertionError(c)
    assert (i, j) == (len(q), len(t))


def test_siphash_vectors(oracle):
    L = oracle.lib()
    key = bytes(range(16))
    k0, k1 = int.from_bytes(key[:8], "little"), int.from_bytes(key[8:], "little")
    for h in GOLD["hashes"]:
        if h["kind"] == "siphash24":
            got = L.awo_siphash(2, 4, k0, k1, bytes(range(15)), 15)
        elif h["kind"] == "empty13":
            got = L.awo_siphash(1, 3, 0, 0, b"", 0)
        elif h["kind"] == "bytes":
            got = oracle.hash_bytes(h["msg"].encode())
        else:
            got = oracle.hash_str(h["msg"])
        assert "%016x" % got == h["value"], h


def test_anchor_alignments(oracle):
    for a in GOLD["anchors"]:
        p = oracle.params(*PENS[a["pen"]])
        st, sc, ops, _ = oracle.wfa_align(p, a["q"].encode(), a["t"].encode())
        assert st == 0 and sc == a["score"] and ops.decode() == a["ops"] and oracle.cigar_string(ops) == a["cg"], a["src"]
    # Appendix C worked example
    r = oracle.align_pair(b"ACGTACGTACGT", b"ACGTACGTTCGT", 0, 1, oracle.params(), qname="seq1", tname="seq2")
    assert r["paf"] == "seq1\t12\t0\t12\t+\tseq2\t12\t0\t12\t11\t12\t60\tgi:f:0.916667\tcg:Z:8=1X3="
    # README example through allwave's glue: num_matches 29, block 32, gi 0.935484
    r = oracle.align_pair(b"TCTTTACTCGCGCGTTGGAGAAATACAATAGT", b"TCTATACTGCGCGTTTGGAGAAATAAAATAGT", 0, 1, oracle.params(0, 4, 6, 2, None, None), qname="p", tname="t")
    assert r["num_matches"] == 29 and "\t32\t60\tgi:f:0.935484\tcg:Z:3=1X4=1I7=1D9=1X6=" in r["paf"]


def test_golden_alignments(oracle):
    cache = {}
    for g in GOLD["alignments"]:
        key = (g["seed"], g["n"], g["length"], g["d"], g["rc_prob"])
        if key not in cache:
            cache[key] = synth.generate(g["seed"], g["n"], g["length"], g["d"], rc_prob=g["rc_prob"])
        ids, seqs, _ = cache[key]
        r = oracle.align_pair(seqs[g["q"]], seqs[g["t"]], g["q"], g["t"], oracle.params(*PENS[g["pen"]]), qname=ids[g["q"]], tname=ids[g["t"]])
        assert (r["score"], int(r["is_reverse"]), r["paf"]) == (g["score"], g["is_reverse"], g["paf"])


def test_golden_pair_lists(oracle):
    for g in GOLD["pair_lists"]:
        ids = ["s%06d" % i for i in range(g["n"])]
        if g["kind"] == oracle.SPARS_TREE:
            _, seqs, _ = synth.generate(g["seed"], g["n"], g["length"], g["d"])
            pl = oracle.pair_list(ids, seqs, kind=g["kind"], fraction=g["fraction"], k_nearest=g["k_nearest"], k_farthest=g["k_farthest"])
        else:
            pl = oracle.pair_list(ids, None, kind=g["kind"], fraction=g["fraction"])
        assert [list(p) for p in pl] == g["pairs"]


def test_parse_scores_and_mode(oracle):
    # src/lib.rs:159-192
    p = oracle.parse_scores("0,1,1,1")
    assert (p.match_score, p.mismatch_penalty, p.gap_open, p.gap_extend, p.has_gap2_open) == (0, 1, 1, 1, 0)
    assert oracle.mode(p) == 0  # EditDistance
    p = oracle.parse_scores(" 0, 5,8 ,2,24,1")
    assert (p.gap2_open, p.gap2_extend, p.has_gap2_open, p.has_gap2_extend) == (24, 1, 1, 1)
    assert oracle.mode(p) == 2  # TwoPieceAffine
    assert oracle.mode(oracle.parse_scores("0,4,6,2")) == 1
    for bad in ("0,1,1", "0,1,1,1,1", "a,b,c,d", "", "0,,1,1"):
        with pytest.raises(ValueError):
            oracle.parse_scores(bad)


def test_score_optimal_and_cigar_valid(oracle):
    rnd = random.Random(3)
    for it in range(120):
        n = rnd.randint(1, 900)
        ids, seqs, _ = synth.generate(500 + it, 2, n, rnd.choice([0.0, 0.01, 0.05, 0.15, 0.4]))
        q, t = seqs
        for pen in PENS.values():
            p = oracle.params(*pen)
            st, sc, ops, w = oracle.wfa_align(p, q, t)
            assert st == 0
            walk_cigar(ops, q, t)
            assert -sc == oracle.cigar_penalty(p, ops) == oracle.gotoh_penalty(p, q, t)
            # unidirectional (MemoryMode::High) WFA reaches the same score
            st2, sc2, ops2, _ = oracle.wfa_align(p, q, t, unidirectional=True)
            assert st2 == 0 and sc2 == sc
            walk_cigar(ops2, q, t)


def test_unrelated_and_degenerate(oracle):
    rnd = random.Random(9)
    a = bytes(rnd.choice(b"ACGT") for _ in range(400))
    b = bytes(rnd.choice(b"ACGT") for _ in range(380))
    for q, t in ((a, b), (a, b""), (b"", a), (b"", b""), (a, a[:1]), (b"A" * 300, b"A" * 250), (a, a[::-1])):
        for pen in PENS.values():
            p = oracle.params(*pen)
            st, sc, ops, _ = oracle.wfa_align(p, q, t)
            assert st == 0
            walk_cigar(ops, q, t)
            assert -sc == oracle.gotoh_penalty(p, q, t)


def test_reference_integration_properties(oracle):
    p = oracle.params()
    # test_identical_sequences (tests/integration_tests.rs:216-260)
    _, seqs, _ = synth.generate(4, 1, 5000, 0.0)
    r = oracle.align_pair(seqs[0], seqs[0], 0, 1, p, qname="a", tname="b")
    assert r["cigar_bytes"] == b"M" * 5000 and "gi:f:1.000000\tcg:Z:5000=" in r["paf"] and not r["is_reverse"]
    # test_alignment_correctness (:599-672): (ATCG)x25 with 2 SNPs, 1 base removed, 1 inserted -> X=2, I=1, D=1 (standard letters)
    ref = bytearray(b"ATCG" * 25)
    qry = bytearray(ref)
    qry[10] = ord("G") if qry[10] != ord("G") else ord("A")
    qry[50] = ord("C") if qry[50] != ord("C") else ord("T")
    del qry[30]
    qry.insert(70, ord("T"))
    r = oracle.align_pair(bytes(ref), bytes(qry), 0, 1, p, qname="r", tname="q")
    ops = r["cigar_bytes"]
    assert ops.count(b"X") == 2 and ops.count(b"I") == 1 and ops.count(b"D") == 1
    # test_reverse_complement_alignment (:443-555)
    _, seqs, _ = synth.generate(8, 1, 3000, 0.0)
    ref = seqs[0]
    _, mut, _ = synth.generate(8, 1, 3000, 0.02)
    fwd, rev = mut[0], oracle.reverse_complement(mut[0])
    a = oracle.align_pair(ref, fwd, 0, 1, p, qname="ref", tname="f")
    b = oracle.align_pair(ref, rev, 0, 2, p, qname="ref", tname="r")
    assert not a["is_reverse"] and b["is_reverse"]
    ia = a["num_matches"] / a["alignment_length"]
    ib = b["num_matches"] / b["alignment_length"]
    assert abs(ia - ib) < 0.01


def test_pair_counts_and_strategies(oracle):
    ids = ["seq%d" % i for i in range(6)]
    assert len(oracle.pair_list(ids[:3], None)) == 6  # tests/integration_tests.rs:809-817
    assert len(oracle.pair_list(ids, None)) == 30     # :1662-1668
    assert len(oracle.pair_list(ids[:3], None, exclude_self=False)) == 9
    # compute_connectivity_probability table (src/iterator.rs:309-316) and clamps
    L = oracle.lib()
    assert [L.awo_connectivity_probability(n, 0.99) for n in (1, 2, 3, 4, 5, 6, 10)] == [1.0, 1.0, 0.8, 0.7, 0.6, 0.5, 0.5]
    assert abs(L.awo_connectivity_probability(200, 0.99) - 0.0494923) < 1e-6
    assert L.awo_connectivity_probability(10**9, 0.5) == 0.001
    # giant:0.99 with n=3 -> p=0.8 -> non-empty (tests/integration_tests.rs:1676-1804)
    assert len(oracle.pair_list(ids[:3], None, kind=oracle.SPARS_GIANT, fraction=0.99)) > 0
    # directed hash: (a,b) and (b,a) are decided independently
    pl = set(oracle.pair_list(["s%06d" % i for i in range(60)], None, kind=oracle.SPARS_RANDOM, fraction=0.5))
    assert any((b, a) not in pl for a, b in pl)
    assert 0.4 < len(pl) / (60 * 59) < 0.6


def test_knn_graph_literal_matrices(oracle):
    # src/knn_graph.rs:226-260, 318-338
    m = [[0.0, 0.1, 0.5], [0.1, 0.0, 0.3], [0.5, 0.3, 0.0]]
    near = oracle.build_knn_graph(m, 1, False)
    assert (0, 1) in near and (1, 0) in near and (2, 1) in near and len(near) == 3
    far = oracle.build_knn_graph(m, 1, True)
    assert (0, 2) in far and (1, 2) in far and len(far) == 3
    assert len(oracle.build_knn_graph(m, 2, False)) == 6
    assert oracle.build_knn_graph([], 2, False) == [] and oracle.build_knn_graph([[0.0]], 2, False) == []
    # ties resolve by ascending index (stable sort)
    t = [[0.0, 0.2, 0.2, 0.2], [0.2, 0.0, 0.2, 0.2], [0.2, 0.2, 0.0, 0.2], [0.2, 0.2, 0.2, 0.0]]
    assert oracle.build_knn_graph(t, 1, False) == [(0, 1), (1, 0), (2, 0), (3, 0)]
    assert oracle.build_knn_graph(t, 1, True) == [(0, 1), (1, 0), (2, 0), (3, 0)]


def test_mash_properties(oracle):
    # src/mash.rs:186-260
    _, seqs, _ = synth.generate(12, 3, 2000, 0.05)
    s0 = oracle.sketch(seqs[0], canonical=True)
    assert len(s0) == 1000 and s0 == sorted(s0)
    assert oracle.jaccard_counts(s0, s0)[0] == oracle.jaccard_counts(s0, s0)[1]
    assert oracle.mash_distance(s0, s0) < 1e-10
    dm = oracle.distance_matrix(seqs)
    assert all(dm[i][i] == 0.0 for i in range(3)) and dm[0][1] == dm[1][0] > 0
    assert oracle.sketch(b"ACGT", k=15) == [] and oracle.jaccard_counts([], []) == (0, 0)
    assert oracle.reverse_complement(b"ATCG") == b"CGAT" and oracle.reverse_complement(b"acgtnNx") == b"NNNACGT"
    # k-mers with non-ACGT bytes are skipped; lower case counts as DNA but hashes as raw bytes
    assert len(oracle.sketch(b"ACGTACGTACGTACGNACGTACGTACGTACG")) == 2
    assert oracle.sketch(b"acgtacgtacgtacg") != oracle.sketch(b"ACGTACGTACGTACG")


def test_run_pairs_threads_agree(oracle):
    ids, seqs, _ = synth.generate(14, 5, 700, 0.04, rc_prob=0.5)
    pairs = [(i, j) for i in range(5) for j in range(5) if i != j]
    a = oracle.run_pairs(ids, seqs, pairs, oracle.params(), threads=1)
    b = oracle.run_pairs(ids, seqs, pairs, oracle.params(), threads=4)
    assert a["paf"] == b["paf"] and a["scores"] == b["scores"] and a["sum_block_len"] == b["sum_block_len"]


def test_fast_mode_agrees(oracle):
    """the opt-in fast mode of the restatement (bench baseline legs) gives byte-identical PAF and scores"""
    O = oracle

    rnd = random.Random(11)
    for (n, length, d, pen) in [(10, 600, 0.05, (0, 5, 8, 2, 24, 1)), (8, 300, 0.2, (0, 1, 1, 1, None, None)), (6, 2500, 0.03, (0, 4, 6, 2, None, None)),
                                (12, 90, 0.1, (0, 5, 8, 2, 24, 1)), (4, 7000, 0.05, (0, 5, 8, 2, 24, 1)), (6, 1000, 0.02, (0, 4, 12, 1, 6, 3))]:
        ids, seqs, _ = synth.generate(rnd.randrange(1 << 30), n, length, d, rc_prob=0.3)
        seqs[0] = seqs[1][: length // 2]  # a long end gap
        pairs = [(i, j) for i in range(n) for j in range(n) if i != j]
        p = O.params(*pen)
        a = O.run_pairs(ids, seqs, pairs, p, use_mash=True, threads=4)
        b = O.run_pairs(ids, seqs, pairs, p, use_mash=True, threads=4, fast=True)
        assert a["paf"] == b["paf"] and a["scores"] == b["scores"] and a["work"]["cells"] == b["work"]["cells"]


def test_match_score_bonus(oracle):
    """match_score < 0 (WFA2 penalty shifting, include/aw_wfa2_compat.h): the alignment is optimal for the shifted penalties
    (independent Gotoh DP) and the reported score is the CIGAR's score under the user's own penalties, matches included"""
    O = oracle
    rnd = random.Random(41)
    for (m, x, o, e, o2, e2) in [(-1, 4, 6, 2, None, None), (-2, 5, 8, 2, 24, 1), (-3, 2, 1, 1, None, None), (-1, 1, 1, 1, None, None)]:
        p = O.params(m, x, o, e, o2, e2)
        shifted = O.params(0, 2 * x - 2 * m, 2 * o, 2 * e - m, None if o2 is None else 2 * o2, None if e2 is None else 2 * e2 - m)
        for _ in range(25):
            n = rnd.randint(1, 300)
            a = bytes(rnd.choice(b"ACGT") for _ in range(n))
            b = bytearray(a)
            for _ in range(rnd.randint(0, 12)):
                pos = rnd.randrange(len(b) + 1)
                k = rnd.random()
                if k < 0.5 and pos < len(b):
                    b[pos] = rnd.choice(b"ACGT")
                elif k < 0.75:
                    b[pos:pos] = bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(1, 30)))
                else:
                    del b[pos:pos + rnd.randint(1, 30)]
            b = bytes(b)
            st, sc, ops, _ = O.wfa_align(p, a, b)
            assert st == 0
            n_m, n_x = ops.count(b"M"), ops.count(b"X")
            assert n_m + n_x + ops.count(b"D") == len(a) and n_m + n_x + ops.count(b"I") == len(b)
            assert sc == -O.cigar_penalty(p, ops)                      # user's penalties, match bonus included
            s_shift = O.cigar_penalty(shifted, ops)                    # the penalty the wavefronts minimised
            assert s_shift == O.gotoh_penalty(shifted, a, b)           # ... optimally
            assert 2 * sc == -(s_shift + m * (len(a) + len(b)))       # WFA2's score conversion
    with pytest.raises(Exception):
        st, sc, ops, _ = O.wfa_align(O.params(1, 4, 6, 2, None, None), b"ACGT", b"ACGT")
        assert st == 0


# ---- the hook that pins the oracle against a real WFA2-lib build (PARITY UNPINNED until the file exists) ----
WFA2_INPUTS = os.path.join(HERE, "golden", "wfa2_inputs.tsv.gz")
WFA2_VECTORS = os.environ.get("AW_WFA2_VECTORS") or os.path.join(HERE, "golden", "wfa2_lib_vectors.tsv.gz")


def read_wfa2_inputs():
    import gzip
    rows = {}
    with gzip.open(WFA2_INPUTS, "rt") as f:
        for line in f:
            v = line.rstrip("\n").split("\t")
            rows[v[0]] = (tuple(int(x) for x in v[1:7]), v[7].encode(), v[8].encode())
    return rows


def test_wfa2_vector_inputs_run_through_oracle(oracle):
    """the committed input set of tools/wfa2_vectors/ is well formed and the oracle aligns it: valid CIGARs, and optimal scores
    (independent Gotoh DP on the shifted penalties) for a seeded sample of the short ones"""
    rows = read_wfa2_inputs()
    assert len(rows) > 3000
    rnd = random.Random(3)
    short = [k for k, (_, a, b) in rows.items() if len(a) <= 800 and len(b) <= 1600]
    for k in rnd.sample(short, 160):
        (m, x, o1, e1, o2, e2), a, b = rows[k]
        two = o2 >= 0
        p = oracle.params(m, x, o1, e1, o2 if two else None, e2 if two else None)
        st, sc, ops, _ = oracle.wfa_align(p, a, b)
        assert st == 0, k
        walk_cigar(ops, a, b)
        assert sc == -oracle.cigar_penalty(p, ops), k
        shifted = oracle.params(0, 2 * x - 2 * m, 2 * o1, 2 * e1 - m, 2 * o2 if two else None, 2 * e2 - m if two else None) if m else p
        assert oracle.cigar_penalty(shifted, ops) == oracle.gotoh_penalty(shifted, a, b), k


def test_real_wfa2_vectors(oracle):
    """score and every CIGAR operation of the oracle against the output of a real WFA2-lib build on the committed inputs
    (tools/wfa2_vectors/dump_wfa2_vectors.c, run by a maintainer who has WFA2-lib at the pinned commit)"""
    if not os.path.exists(WFA2_VECTORS):
        pytest.skip("PARITY UNPINNED: tests/golden/wfa2_lib_vectors.tsv.gz absent (no WFA2-lib in this build container); "
                    "see tools/wfa2_vectors/dump_wfa2_vectors.c")
    import gzip
    rows = read_wfa2_inputs()
    n = 0
    with gzip.open(WFA2_VECTORS, "rt") as f:
        for line in f:
            k, score, ops_ref = (line.rstrip("\n").split("\t") + [""])[:3]
            (m, x, o1, e1, o2, e2), a, b = rows[k]
            two = o2 >= 0
            st, sc, ops, _ = oracle.wfa_align(oracle.params(m, x, o1, e1, o2 if two else None, e2 if two else None), a, b)
            assert st == 0 and sc == int(score) and ops.decode() == ops_ref, k
            n += 1
    assert n == len(rows)
