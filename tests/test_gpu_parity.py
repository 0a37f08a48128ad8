"""GPU parity tests proper: the CUDA path, called through the C ABI, must equal the CPU oracle
bit for bit (score, WFA2 op string, PAF line) on the same seeded inputs."""
import os
import random

import pytest

import allwave_b200 as aw
from allwave_b200 import synth

pytestmark = pytest.mark.gpu

DEFAULT = dict(mismatch=5, gap_open=8, gap_extend=2, gap2_open=24, gap2_extend=1)
EDIT = dict(mismatch=1, gap_open=1, gap_extend=1, gap2_open=None, gap2_extend=None)
AFFINE = dict(mismatch=4, gap_open=6, gap_extend=2, gap2_open=None, gap2_extend=None)


def _all_pairs(n):
    return [(i, j) for i in range(n) for j in range(n) if i != j]


def _compare(oracle, ctx, ids, seqs, pairs, pen, orientation=aw.AW_ORIENT_FORWARD, flags=aw.AW_FLAG_CIGAR_BYTES):
    ctx.load_sequences(ids, seqs)
    res = ctx.align_pairs(aw.make_params(**pen), pairs, orientation=orientation, flags=flags)
    assert len(res) == len(pairs)
    op = oracle.params(**pen)
    bad = []
    for r, (qi, ti) in zip(res, pairs):
        assert (r["query_idx"], r["target_idx"]) == (qi, ti)
        if orientation == aw.AW_ORIENT_FORWARD:
            st, sc, ops, _ = oracle.wfa_align(op, seqs[qi], seqs[ti])
            exp = dict(score=sc, cigar_bytes=ops, is_reverse=False)
            exp_paf = None
        else:
            o = oracle.align_pair(seqs[qi], seqs[ti], qi, ti, op, use_mash=True, qname=ids[qi], tname=ids[ti])
            exp = dict(score=o["score"], cigar_bytes=o["cigar_bytes"], is_reverse=bool(o["is_reverse"]))
            exp_paf = o["paf"]
        ok = r["status"] == 0 and r["score"] == exp["score"] and r["is_reverse"] == exp["is_reverse"]
        if flags & aw.AW_FLAG_CIGAR_BYTES:
            ok = ok and r["cigar_bytes"] == exp["cigar_bytes"]
        ok = ok and r["cg"] == oracle.cigar_string(exp["cigar_bytes"])
        if exp_paf is not None:
            ok = ok and r["paf"] == exp_paf
        if not ok:
            bad.append((qi, ti, r["status"], r["score"], exp["score"], r["cg"][:80], oracle.cigar_string(exp["cigar_bytes"])[:80]))
    assert not bad, f"{len(bad)}/{len(pairs)} pairs differ, first: {bad[:3]}"
    return res


def test_known_answers(oracle, gpu_ctx):
    # WFA2-lib README example and the hand-derived cases of tests/debug/*.rs
    ids = ["p", "t", "seq1", "seq2", "a12", "a10"]
    seqs = [b"TCTTTACTCGCGCGTTGGAGAAATACAATAGT", b"TCTATACTGCGCGTTTGGAGAAATAAAATAGT", b"ACGTACGTACGT", b"ACGTACGTTCGT", b"ACGTACGTACGT", b"ACGTACGTAC"]
    gpu_ctx.load_sequences(ids, seqs)
    r = gpu_ctx.align_pairs(aw.make_params(**AFFINE), [(0, 1)], orientation=aw.AW_ORIENT_FORWARD, flags=aw.AW_FLAG_CIGAR_BYTES)[0]
    assert r["score"] == -24 and r["cigar_bytes"] == b"MMMXMMMMDMMMMMMMIMMMMMMMMMXMMMMMM" and r["cg"] == "3=1X4=1I7=1D9=1X6="
    assert r["num_matches"] == 29 and r["paf"].split("\t")[10] == "32" and "gi:f:0.935484" in r["paf"]
    r = gpu_ctx.align_pairs(aw.make_params(**DEFAULT), [(2, 3)], orientation=aw.AW_ORIENT_FORWARD)[0]
    assert r["paf"] == "seq1\t12\t0\t12\t+\tseq2\t12\t0\t12\t11\t12\t60\tgi:f:0.916667\tcg:Z:8=1X3="
    r = gpu_ctx.align_pairs(aw.make_params(**DEFAULT), [(4, 5), (5, 4)], orientation=aw.AW_ORIENT_FORWARD)
    assert r[0]["cg"] == "10=2I" and r[1]["cg"] == "10=2D"


@pytest.mark.parametrize("pen", [DEFAULT, EDIT, AFFINE], ids=["affine2p", "edit", "affine"])
@pytest.mark.parametrize("length,d,n", [(60, 0.05, 8), (150, 0.02, 12), (400, 0.1, 8), (1500, 0.03, 6), (3000, 0.08, 4)])
def test_forward_parity_small(oracle, gpu_ctx, pen, length, d, n):
    ids, seqs, _ = synth.generate(1000 + length, n, length, d)
    _compare(oracle, gpu_ctx, ids, seqs, _all_pairs(n), pen)


def test_edge_cases(oracle, gpu_ctx):
    rnd = random.Random(5)
    base = bytes(rnd.choice(b"ACGT") for _ in range(300))
    seqs = [b"", b"A", b"ACGT", base, base, base[:150], base[150:], base[::-1], b"A" * 200, b"A" * 180 + b"C" * 20, b"ACGT" * 40, b"ACGT" * 38 + b"AC"]
    ids = ["e%d" % i for i in range(len(seqs))]
    for pen in (DEFAULT, EDIT):
        _compare(oracle, gpu_ctx, ids, seqs, _all_pairs(len(seqs)), pen)


def test_non_acgt_bytes_use_byte_path(oracle, gpu_ctx):
    ids, seqs, _ = synth.generate(77, 4, 500, 0.04)
    seqs[1] = seqs[1][:100] + b"NNNNnnacgt" + seqs[1][110:]
    seqs[2] = seqs[2].lower()
    _compare(oracle, gpu_ctx, ids, seqs, _all_pairs(4), DEFAULT)
    _compare(oracle, gpu_ctx, ids, seqs, _all_pairs(4), DEFAULT, orientation=aw.AW_ORIENT_MASH)


def test_c1_shape_full_paf(oracle, gpu_ctx):
    # BASELINE config 1 shape at reduced n: 6 x 10 kb, d=1%, default scores, mash orientation
    c, ids, seqs, _ = synth.config("C1", n=6)
    _compare(oracle, gpu_ctx, ids, seqs, _all_pairs(6), DEFAULT, orientation=aw.AW_ORIENT_MASH)


def test_c2_shape_sample(oracle, gpu_ctx):
    # BASELINE config 2 shape: 10 kb at d=5% per haplotype; a few pairs (oracle ~1.6 s per pair)
    c, ids, seqs, _ = synth.config("C2", n=4)
    _compare(oracle, gpu_ctx, ids, seqs, [(0, 1), (1, 2), (3, 0), (2, 3)], DEFAULT, orientation=aw.AW_ORIENT_MASH)


def test_c3_shape(oracle, gpu_ctx):
    c, ids, seqs, _ = synth.config("C3", n=40)
    _compare(oracle, gpu_ctx, ids, seqs, _all_pairs(40), EDIT, orientation=aw.AW_ORIENT_MASH)


def test_c5_shape_mixed_orientation(oracle, gpu_ctx):
    c, ids, seqs, rc = synth.config("C5", n=8)
    assert 0 < sum(rc) < 8
    res = _compare(oracle, gpu_ctx, ids, seqs, _all_pairs(8), DEFAULT, orientation=aw.AW_ORIENT_MASH)
    for r in res:
        assert r["is_reverse"] == (rc[r["query_idx"]] != rc[r["target_idx"]])


def test_identical_sequences(gpu_ctx):
    # reference tests/integration_tests.rs:216-260: identity 1.0, CIGAR 5000=
    ids, seqs, _ = synth.generate(9, 1, 5000, 0.0)
    gpu_ctx.load_sequences(["a", "b"], [seqs[0], seqs[0]])
    r = gpu_ctx.align_pairs(aw.make_params(**DEFAULT), [(0, 1)])[0]
    assert r["cg"] == "5000=" and r["score"] == 0 and "gi:f:1.000000" in r["paf"] and not r["is_reverse"]


def test_sketches_and_orientation(oracle, gpu_ctx):
    c, ids, seqs, rc = synth.config("C5", n=6, length=3000)
    gpu_ctx.load_sequences(ids, seqs)
    for i in range(6):
        assert gpu_ctx.get_sketch(i) == oracle.sketch(seqs[i])
        assert gpu_ctx.get_sketch(i, reverse_complement=True) == oracle.sketch(oracle.reverse_complement(seqs[i]))
        assert gpu_ctx.get_sketch(i, canonical=True) == oracle.sketch(seqs[i], canonical=True)
    pairs = _all_pairs(6)
    assert gpu_ctx.orient_pairs(pairs) == [oracle.orientation_mash(seqs[q], seqs[t]) for q, t in pairs]
    inter, uni = gpu_ctx.mash_jaccard_counts()
    for i in range(6):
        for j in range(i + 1, 6):
            e = oracle.jaccard_counts(oracle.sketch(seqs[i], canonical=True), oracle.sketch(seqs[j], canonical=True))
            assert (int(inter[i, j]), int(uni[i, j])) == e == (int(inter[j, i]), int(uni[j, i]))


def test_aligner_api(oracle, gpu_ctx):
    al = aw.Aligner(gpu_ctx, 5, 8, 2, 24, 1)
    ids, seqs, _ = synth.generate(31, 2, 800, 0.05)
    assert al.align(seqs[0], seqs[1]) == 0
    st, sc, ops, _ = oracle.wfa_align(oracle.params(), seqs[0], seqs[1])
    assert al.score() == sc and al.cigar() == ops
    al.close()


def test_host_tree_pair_list_and_cli(oracle, gpu_ctx, tmp_path):
    """tree:2:1:0.1 pair list from the C++ host mirror (GPU sketches + Jaccard counts) and the CLI's PAF"""
    import subprocess

    from allwave_b200 import hostlib as H

    c, ids, seqs, rc = synth.config("C5", n=14, length=1500)
    gpu_ctx.load_sequences(ids, seqs)
    got = H.pair_list(ids, H.KIND_TREE, k_nearest=2, k_farthest=1, random_fraction=0.1, ctx=gpu_ctx)
    exp = oracle.pair_list(ids, seqs, kind=oracle.SPARS_TREE, fraction=0.1, k_nearest=2, k_farthest=1)
    assert got == exp and len(got) >= 14 * 2
    fa = tmp_path / "in.fa"
    fa.write_text("".join(f">{i} extra words\n{s.decode()}\n" for i, s in zip(ids, seqs)))
    out = tmp_path / "out.paf"
    exe = os.path.join(os.path.dirname(aw._cabi.so_path()), "allwave")
    subprocess.check_call([exe, "-i", str(fa), "-o", str(out), "-p", "tree:2:1:0.1", "-s", "0,5,8,2,24,1", "--no-progress"])
    lines = out.read_text().splitlines()
    p = oracle.params(**DEFAULT)
    want = [oracle.align_pair(seqs[q], seqs[t], q, t, p, use_mash=True, qname=ids[q], tname=ids[t])["paf"] for q, t in exp]
    assert lines == want


def test_wfa_orientation(oracle, gpu_ctx):
    """--wfa-orientation (src/alignment.rs:157-175): two edit-distance alignments decide the strand"""
    c, ids, seqs, rc = synth.config("C5", n=6, length=1200)
    gpu_ctx.load_sequences(ids, seqs)
    pairs = _all_pairs(6)
    res = gpu_ctx.align_pairs(aw.make_params(**DEFAULT), pairs, orientation=aw.AW_ORIENT_WFA)
    p = oracle.params(**DEFAULT)
    for r, (q, t) in zip(res, pairs):
        o = oracle.align_pair(seqs[q], seqs[t], q, t, p, use_mash=False, qname=ids[q], tname=ids[t])
        assert r["is_reverse"] == bool(o["is_reverse"]) == (rc[q] != rc[t]) and r["paf"] == o["paf"] and r["score"] == o["score"]


def test_c4_shape_long_pairs(oracle, gpu_ctx):
    """BASELINE config 4 shape at reduced length: haplotypes with SVs, too long for int16 offsets -> the int32 chunked
    path with sequences staged in global memory, biWFA recursion over many levels"""
    c, ids, seqs, _ = synth.config("C4", n=3, length=24000)
    assert max(len(s) for s in seqs) * 3 > 32000
    _compare(oracle, gpu_ctx, ids, seqs, [(0, 1), (1, 2), (2, 0)], DEFAULT)


@pytest.mark.parametrize("pen", [DEFAULT, AFFINE], ids=["affine2p", "affine"])
def test_int16_boundary_lengths(oracle, gpu_ctx, pen):
    """pairs just below / above the int16 storage limit (2*tlen+plen < 32000) take different kernels"""
    ids, seqs, _ = synth.generate(4242, 2, 10800, 0.02)
    a, b = seqs
    cases = [(a[:10600], b[:10600]), (a[:10700], b[:10700]), (a[:10660], b[:10500]), (a[:9000], b[:10800])]
    for x, y in cases:
        _compare(oracle, gpu_ctx, ["x", "y"], [x, y], [(0, 1), (1, 0)], pen)


def test_kernel_variants_agree(gpu_ctx):
    """the same pairs through every storage / CTA-size variant of the alignment kernel give identical PAF"""
    c, ids, seqs, _ = synth.config("C2", n=6, length=6000)
    pairs = _all_pairs(6)
    p = aw.make_params(**DEFAULT)
    ctx = aw.Context(0)
    try:
        ctx.load_sequences(ids, seqs)
        base = [r["paf"] for r in ctx.align_pairs(p, pairs, orientation=aw.AW_ORIENT_MASH)]
        for opts in (dict(threads_per_cta=256), dict(threads_per_cta=256, ws16=0), dict(threads_per_cta=128, ctas_per_sm=1), dict(threads_per_cta=32)):
            c2 = aw.Context(0)
            for k, v in opts.items():
                c2.set_option(k, v)
            c2.load_sequences(ids, seqs)
            got = [r["paf"] for r in c2.align_pairs(p, pairs, orientation=aw.AW_ORIENT_MASH)]
            c2.close()
            assert got == base, opts
    finally:
        ctx.close()


def _check_cigar(q, t, cg, pen):
    """applies a cg:Z: string (= match, X mismatch, I consumes query, D consumes target) and returns the penalty"""
    import re

    i = j = 0
    penalty = 0
    for n, op in re.findall(r"(\d+)([=XID])", cg):
        n = int(n)
        if op == "=":
            assert q[i:i + n] == t[j:j + n]
            i += n
            j += n
        elif op == "X":
            assert all(q[i + k] != t[j + k] for k in range(n))
            penalty += n * pen["mismatch"]
            i += n
            j += n
        else:
            g = pen["gap_open"] + n * pen["gap_extend"]
            if pen.get("gap2_open") is not None:
                g = min(g, pen["gap2_open"] + n * pen["gap2_extend"])
            penalty += g
            if op == "I":
                i += n
            else:
                j += n
    assert (i, j) == (len(q), len(t))
    return penalty


def test_c2_full_size_properties(gpu_ctx):
    """BASELINE config 2 at full sequence size, 600 pairs: every CIGAR is a valid global alignment of its two
    sequences whose penalty equals -score, forward/backward pairs score the same, and the PAF columns agree"""
    c, ids, seqs, _ = synth.config("C2", n=40)
    pairs = [(i, (i + s) % 40) for s in range(1, 16) for i in range(40)]
    gpu_ctx.load_sequences(ids, seqs)
    res = gpu_ctx.align_pairs(aw.make_params(**DEFAULT), pairs, orientation=aw.AW_ORIENT_MASH)
    score = {}
    for r, (q, t) in zip(res, pairs):
        assert r["status"] == 0 and not r["is_reverse"]
        assert _check_cigar(seqs[q], seqs[t], r["cg"], DEFAULT) == -r["score"]
        f = r["paf"].split("\t")
        assert (f[0], int(f[1]), int(f[3]), f[5], int(f[6]), int(f[8])) == (ids[q], len(seqs[q]), len(seqs[q]), ids[t], len(seqs[t]), len(seqs[t]))
        assert int(f[10]) == max(len(seqs[q]), len(seqs[t])) and int(f[9]) == r["num_matches"]
        score[(q, t)] = r["score"]
    for (q, t), s in score.items():
        if (t, q) in score:
            assert score[(t, q)] == s


def test_cli_multi_gpu_matches_single(tmp_path):
    """allwave --gpus N (pairs sharded by predicted cost, one host thread per GPU, no collective) writes the same set of
    PAF lines as one GPU; needs >= 2 devices"""
    import subprocess

    if aw._cabi.lib().aw_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    c, ids, seqs, _ = synth.config("C5", n=24, length=3000)
    fa = tmp_path / "in.fa"
    fa.write_text("".join(f">{i}\n{s.decode()}\n" for i, s in zip(ids, seqs)))
    exe = os.path.join(os.path.dirname(aw._cabi.so_path()), "allwave")
    outs = []
    for g in (1, 2):
        out = tmp_path / f"out{g}.paf"
        subprocess.check_call([exe, "-i", str(fa), "-o", str(out), "-p", "none", "--no-progress", "--gpus", str(g)])
        outs.append(sorted(out.read_text().splitlines()))
    assert len(outs[0]) == 24 * 23 and outs[0] == outs[1]


def test_cli_mash_matrix_and_preset(oracle, gpu_ctx, tmp_path):
    """--mash-matrix prints the reference's tab-separated distance table (src/mash.rs:168-184) from the GPU sketches;
    -x 90% aligns with the preset's scores; gzipped input"""
    import gzip
    import math
    import subprocess

    c, ids, seqs, rc = synth.config("C5", n=5, length=2500)
    fa = tmp_path / "in.fa.gz"
    with gzip.open(fa, "wt") as f:
        f.write("".join(f">{i}\n{s.decode()}\n" for i, s in zip(ids, seqs)))
    exe = os.path.join(os.path.dirname(aw._cabi.so_path()), "allwave")
    out = subprocess.run([exe, "-i", str(fa), "--mash-matrix"], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0].split("\t") == ["sequence"] + ids and len(out) == 6
    for i in range(5):
        row = out[1 + i].split("\t")
        assert row[0] == ids[i]
        for j in range(5):
            if i == j:
                exp = 0.0
            else:
                inter, uni = oracle.jaccard_counts(oracle.sketch(seqs[i], canonical=True), oracle.sketch(seqs[j], canonical=True))
                jac = inter / uni if uni else 0.0
                exp = 1.0 if jac <= 0.0 else (-1.0 / 15.0) * math.log(2.0 * jac / (1.0 + jac))
            assert row[1 + j] == f"{exp:.6f}"
    paf = tmp_path / "o.paf"
    subprocess.check_call([exe, "-i", str(fa), "-o", str(paf), "-p", "none", "-x", "90%", "--no-progress", "-e", ids[4]])
    lines = paf.read_text().splitlines()
    p = oracle.params(**DEFAULT)
    want = [oracle.align_pair(seqs[q], seqs[t], q, t, p, use_mash=True, qname=ids[q], tname=ids[t])["paf"] for q, t in _all_pairs(4)]
    assert lines == want


def test_golden_fixtures_gpu(gpu_ctx):
    """the CUDA path against the committed golden fixtures (tests/golden/oracle_golden.json): the four published /
    hand-derived anchors and the 132 frozen alignments (score, strand, whole PAF line)"""
    import json

    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.json")))
    pens = {"affine2p": DEFAULT, "edit": EDIT, "affine": AFFINE}
    for a in gold["anchors"]:
        gpu_ctx.load_sequences(["q", "t"], [a["q"].encode(), a["t"].encode()])
        r = gpu_ctx.align_pairs(aw.make_params(**pens[a["pen"]]), [(0, 1)], orientation=aw.AW_ORIENT_FORWARD, flags=aw.AW_FLAG_CIGAR_BYTES)[0]
        assert (r["score"], r["cigar_bytes"].decode(), r["cg"]) == (a["score"], a["ops"], a["cg"]), a["src"]
    groups = {}
    for g in gold["alignments"]:
        groups.setdefault((g["seed"], g["n"], g["length"], g["d"], g["rc_prob"], g["pen"]), []).append(g)
    checked = 0
    for (seed, n, length, d, rcp, pen), items in groups.items():
        ids, seqs, _ = synth.generate(seed, n, length, d, rc_prob=rcp)
        gpu_ctx.load_sequences(ids, seqs)
        res = gpu_ctx.align_pairs(aw.make_params(**pens[pen]), [(g["q"], g["t"]) for g in items], orientation=aw.AW_ORIENT_MASH)
        for r, g in zip(res, items):
            assert (r["score"], int(r["is_reverse"]), r["paf"]) == (g["score"], g["is_reverse"], g["paf"]), (g["case"], pen, g["q"], g["t"])
            checked += 1
    assert checked == len(gold["alignments"]) == 132


def test_retry_ladder_small_workspace(oracle):
    """pairs whose wavefronts outgrow the first-try workspace (capped width, tiny history arena) report AW_EWORKSPACE on the
    device and are re-run by the host with larger workspaces; the results must not change"""
    ids, seqs, _ = synth.generate(77, 4, 3000, 0.08)
    pairs = _all_pairs(4)
    for opts in (dict(max_wavefront_width=256), dict(hist_mb=1, max_wavefront_width=512), dict(max_wavefront_width=128, threads_per_cta=256, ws16=0)):
        ctx = aw.Context(0)
        try:
            for k, v in opts.items():
                ctx.set_option(k, v)
            _compare(oracle, ctx, ids, seqs, pairs, DEFAULT)
            b = aw.Batch(ctx, aw.make_params(**DEFAULT), pairs, flags=0)
            b.launch()
            b.fetch(collect=False)
            assert b.stats()["pairs_retried"] > 0 and b.stats()["failed_pairs"] == 0, opts
            b.close()
        finally:
            ctx.close()
