"""GPU parity tests, second batch: wide randomized sweeps against the CPU oracle, unusual penalty sets, the streaming
C ABI (aw_align_stream, PAF blocks), callback cancellation, the failure sentinel, the orientation retry ladder and the
full-size C2 / C5 samples.  Everything goes through the C ABI (ctypes) or the C++ host mirror above it."""
import os
import random

import pytest

import allwave_b200 as aw
from allwave_b200 import synth

pytestmark = pytest.mark.gpu

DEFAULT = dict(mismatch=5, gap_open=8, gap_extend=2, gap2_open=24, gap2_extend=1)
EDIT = dict(mismatch=1, gap_open=1, gap_extend=1, gap2_open=None, gap2_extend=None)
CORES = os.cpu_count() or 1

# penalty sets of the sweep: the CLI presets (src/main.rs:115-121) plus sets that break the usual ordering
# o1 < o2, e1 > e2, x < o+e (the phase-2 stop rule and the scope arithmetic must not depend on it)
PENALTY_SETS = [
    dict(mismatch=5, gap_open=8, gap_extend=2, gap2_open=24, gap2_extend=1),   # default / -x 85
    dict(mismatch=7, gap_open=12, gap_extend=2, gap2_open=36, gap2_extend=1),  # -x 95
    dict(mismatch=4, gap_open=6, gap_extend=2, gap2_open=18, gap2_extend=1),   # -x 75
    dict(mismatch=3, gap_open=4, gap_extend=1, gap2_open=None, gap2_extend=None),  # -x 65
    dict(mismatch=1, gap_open=1, gap_extend=1, gap2_open=None, gap2_extend=None),  # edit
    dict(mismatch=4, gap_open=12, gap_extend=1, gap2_open=6, gap2_extend=3),   # o1 > o2, e2 > e1
    dict(mismatch=2, gap_open=3, gap_extend=2, gap2_open=5, gap2_extend=2),    # e2 == e1
    dict(mismatch=9, gap_open=2, gap_extend=1, gap2_open=5, gap2_extend=1),    # x > o1 + e1
    dict(mismatch=6, gap_open=0, gap_extend=3, gap2_open=None, gap2_extend=None),  # o = 0
    dict(mismatch=2, gap_open=7, gap_extend=3, gap2_open=2, gap2_extend=4),    # second piece dominated for every length > 1
]


def _mutate(rnd, s, d):
    """substitutions, indel bursts, homopolymer and tandem-repeat expansions / contractions at rate ~d per base"""
    out = bytearray()
    i = 0
    while i < len(s):
        if rnd.random() < d:
            r = rnd.random()
            if r < 0.5:
                out.append(rnd.choice(b"ACGT"))
                i += 1
            elif r < 0.65:
                i += 1 + int(rnd.expovariate(0.5))          # deletion burst
            elif r < 0.8:
                out.extend(rnd.choice(b"ACGT") for _ in range(1 + int(rnd.expovariate(0.5))))  # insertion burst
            elif r < 0.9:
                out.extend(bytes([s[i]]) * rnd.randint(2, 12))  # homopolymer expansion
                i += 1
            else:
                unit = s[i:i + rnd.randint(2, 5)]
                out.extend(unit * rnd.randint(2, 6))            # tandem repeat
                i += len(unit)
        else:
            out.append(s[i])
            i += 1
    return bytes(out)


def _sweep_group(seed, nfam=40, members=5, maxlen=400, npairs=10000):
    rnd = random.Random(seed)
    seqs = []
    for f in range(nfam):
        L = rnd.choice([1, 2, 3, 7, 14, 15, 16, 31, 64, 99, 100, 101, 150, 250, 399, 400]) if f % 3 == 0 else rnd.randint(1, maxlen)
        kind = f % 5
        if kind == 0:
            root = bytes(rnd.choice(b"ACGT") for _ in range(L))
        elif kind == 1:
            unit = bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(1, 6)))
            root = (unit * (L // len(unit) + 1))[:L]        # tandem repeat
        elif kind == 2:
            root = b"".join(bytes([rnd.choice(b"ACGT")]) * rnd.randint(1, 20) for _ in range(L))[:L]  # homopolymer runs
        else:
            root = bytes(rnd.choice(b"AC") for _ in range(L)) if kind == 3 else bytes(rnd.choice(b"ACGT") for _ in range(L))
        seqs.append(root)
        for _ in range(members - 1):
            seqs.append(_mutate(rnd, root, rnd.choice([0.0, 0.01, 0.05, 0.1, 0.2, 0.3])))
    ids = ["w%04d" % i for i in range(len(seqs))]
    pairs = []
    for f in range(nfam):
        b = f * members
        pairs += [(b + i, b + j) for i in range(members) for j in range(members) if i != j]
    n = len(seqs)
    while len(pairs) < npairs:
        a, b = rnd.randrange(n), rnd.randrange(n)
        if a != b:
            pairs.append((a, b))
    return ids, seqs, pairs[:npairs]


def _paf_parity(oracle, ctx, ids, seqs, pairs, pen, orientation=aw.AW_ORIENT_MASH):
    ctx.load_sequences(ids, seqs)
    res = ctx.align_pairs(aw.make_params(**pen), pairs, orientation=orientation)
    exp = oracle.run_pairs(ids, seqs, pairs, oracle.params(**pen), use_mash=(orientation == aw.AW_ORIENT_MASH), threads=CORES)
    bad = [(q, t, r["status"], r["score"], s, r["paf"][-60:], e[-60:]) for r, (q, t), e, s in zip(res, pairs, exp["paf"], exp["scores"])
           if r["status"] != 0 or r["paf"] != e or r["score"] != s]
    assert not bad, f"{len(bad)}/{len(pairs)} pairs differ, first: {bad[:2]}"
    return res


@pytest.mark.parametrize("group", range(10))
def test_random_sweep_100k(oracle, gpu_ctx, group):
    """10 groups x 10,000 pairs = 10^5 pairs, lengths 1..400 (+ insert/repeat growth), divergence 0..30 % with indel bursts,
    homopolymers and tandem repeats, unrelated pairs, one penalty set per group: PAF line and score vs the oracle for every
    pair, Gotoh optimality (independent O(nm) DP) on a sample"""
    pen = PENALTY_SETS[group]
    ids, seqs, pairs = _sweep_group(9000 + group)
    res = _paf_parity(oracle, gpu_ctx, ids, seqs, pairs, pen)
    rnd = random.Random(group)
    p = oracle.params(**pen)
    for r in rnd.sample(res, 300):
        q, t = seqs[r["query_idx"]], seqs[r["target_idx"]]
        if r["is_reverse"]:
            q = oracle.reverse_complement(q)
        assert -r["score"] == oracle.gotoh_penalty(p, q, t)


def test_large_penalties_shared_memory_optin(oracle, gpu_ctx):
    """max_score_scope of 80..150 pushes static + dynamic shared memory over 48 KB: the launch must opt in"""
    ids, seqs, _ = synth.generate(31, 6, 2000, 0.03)
    pairs = [(i, j) for i in range(6) for j in range(6) if i != j]
    for pen in (dict(mismatch=5, gap_open=8, gap_extend=2, gap2_open=80, gap2_extend=1), dict(mismatch=3, gap_open=6, gap_extend=2, gap2_open=140, gap2_extend=1),
                dict(mismatch=90, gap_open=20, gap_extend=5, gap2_open=None, gap2_extend=None)):
        _paf_parity(oracle, gpu_ctx, ids, seqs, pairs, pen)
    short = [s[:300] for s in seqs]
    _paf_parity(oracle, gpu_ctx, ids, short, pairs, dict(mismatch=3, gap_open=6, gap_extend=2, gap2_open=140, gap2_extend=1))


def test_hypothesis_pairs(oracle, gpu_ctx):
    """property test: any two strings over ACGT (one derived from the other by random edits) and any valid penalty set
    give the oracle's score, op string and cg"""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    dna = st.text(alphabet="ACGT", min_size=0, max_size=260)
    edits = st.lists(st.tuples(st.integers(0, 259), st.sampled_from("SID"), st.text(alphabet="ACGT", min_size=1, max_size=12)), max_size=25)
    pens = st.tuples(st.integers(1, 9), st.integers(0, 12), st.integers(1, 4), st.one_of(st.none(), st.tuples(st.integers(0, 40), st.integers(1, 4))))

    @settings(max_examples=120, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(dna, edits, pens, st.booleans())
    def prop(a, ed, pen, unrelated):
        b = list(a)
        for pos, kind, ins in ed:
            pos = min(pos, len(b))
            if kind == "S" and pos < len(b):
                b[pos] = ins[0]
            elif kind == "I":
                b[pos:pos] = list(ins)
            elif kind == "D":
                del b[pos:pos + len(ins)]
        b = "".join(b)
        if unrelated:
            b = b[::-1]
        x, o, e, two = pen
        kw = dict(mismatch=x, gap_open=o, gap_extend=e, gap2_open=two[0] if two else None, gap2_extend=two[1] if two else None)
        gpu_ctx.load_sequences(["a", "b"], [a.encode(), b.encode()])
        r = gpu_ctx.align_pairs(aw.make_params(**kw), [(0, 1)], orientation=aw.AW_ORIENT_FORWARD, flags=aw.AW_FLAG_CIGAR_BYTES)[0]
        st_, sc, ops, _ = oracle.wfa_align(oracle.params(**kw), a.encode(), b.encode())
        assert r["status"] == 0 and r["score"] == sc and r["cigar_bytes"] == ops and r["cg"] == oracle.cigar_string(ops)

    prop()


def test_c2_full_size_256_pairs(oracle, gpu_ctx):
    """BASELINE config 2 at full size: 256 pairs of the 1000 x 10 kb set, PAF + score vs the oracle"""
    c, ids, seqs, _ = synth.config("C2", n=1000)
    rnd = random.Random(22)
    pairs = []
    while len(pairs) < 256:
        a, b = rnd.randrange(1000), rnd.randrange(1000)
        if a != b:
            pairs.append((a, b))
    _paf_parity(oracle, gpu_ctx, ids, seqs, pairs, DEFAULT)


def test_c5_full_size_256_pairs(oracle, gpu_ctx):
    """BASELINE config 5 at full size: 256 pairs of the 5000 x 5 kb mixed-orientation set (mash orientation)"""
    c, ids, seqs, rc = synth.config("C5", n=5000)
    rnd = random.Random(55)
    pairs = []
    while len(pairs) < 256:
        a, b = rnd.randrange(5000), rnd.randrange(5000)
        if a != b:
            pairs.append((a, b))
    res = _paf_parity(oracle, gpu_ctx, ids, seqs, pairs, DEFAULT)
    assert any(r["is_reverse"] for r in res) and not all(r["is_reverse"] for r in res)
    for r in res:
        assert r["is_reverse"] == (rc[r["query_idx"]] != rc[r["target_idx"]])


def test_c1_full_config(oracle, gpu_ctx):
    """BASELINE config 1 exactly: 16 x 10 kb at 1 %, -p none = 240 directed pairs"""
    c, ids, seqs, _ = synth.config("C1")
    pairs = [(i, j) for i in range(16) for j in range(16) if i != j]
    assert len(pairs) == 240
    _paf_parity(oracle, gpu_ctx, ids, seqs, pairs, DEFAULT)


# ---- streaming C ABI ---------------------------------------------------------------------------

def test_stream_blocks_equal_per_pair_lines(oracle):
    """aw_align_stream: chunks pulled from a source, two batches in flight, PAF blocks laid out in pair order on the device.
    The concatenated blocks must equal the per-pair lines of aw_align_pairs (and so the oracle's) line for line."""
    c, ids, seqs, rc = synth.config("C5", n=30, length=1200)
    pairs = [(i, j) for i in range(30) for j in range(30) if i != j]
    ctx = aw.Context(0)
    try:
        ctx.load_sequences(ids, seqs)
        p = aw.make_params(**DEFAULT)
        lines = [r["paf"] for r in ctx.align_pairs(p, pairs)]
        exp = oracle.run_pairs(ids, seqs, pairs[:200], oracle.params(**DEFAULT), use_mash=True, threads=CORES)
        assert lines[:200] == exp["paf"]
        chunks = [pairs[i:i + 97] for i in range(0, len(pairs), 97)]
        got, nlines, seen = [], [], []
        ctx.align_stream(p, chunks, flags=aw.AW_FLAG_PAF_BLOCKS, block_callback=lambda b, n: (got.append(b), nlines.append(n)) and False,
                         callback=lambda d: seen.append(d["paf"]) and False)
        assert sum(nlines) == len(pairs) and len(got) == len(chunks)
        assert b"".join(got).decode() == "".join(l + "\n" for l in lines)
        assert seen == lines
        # a narrow first-try workspace forces the retry ladder on the divergent pairs: the per-line fallback of the block path
        c2 = aw.Context(0)
        c2.set_option("max_wavefront_width", 48)
        c2.load_sequences(ids, seqs)
        got2 = []
        c2.align_stream(p, chunks, flags=aw.AW_FLAG_PAF_BLOCKS, block_callback=lambda b, n: got2.append(b) and False)
        c2.close()
        assert sorted(b"".join(got2).decode().splitlines()) == sorted(lines)
    finally:
        ctx.close()


def test_callback_cancel_c_abi(gpu_ctx):
    """a non-zero callback return cancels the run with AW_ECALLBACK (src/iterator.rs:235-251) and leaves the context usable"""
    c, ids, seqs, _ = synth.config("C5", n=12, length=800)
    pairs = [(i, j) for i in range(12) for j in range(12) if i != j]
    ctx = aw.Context(0)
    try:
        ctx.set_option("chunk_pairs", 16)  # several batches: the cancel arrives while the next batch is already in flight
        ctx.load_sequences(ids, seqs)
        p = aw.make_params(**DEFAULT)
        seen = []
        with pytest.raises(aw._cabi.AllwaveError) as ei:
            ctx.align_pairs(p, pairs, callback=lambda d: seen.append(d["paf"]) or len(seen) >= 21)
        assert ei.value.status == aw._cabi.AW_ECALLBACK and len(seen) == 21
        full = [r["paf"] for r in ctx.align_pairs(p, pairs)]
        assert len(full) == len(pairs) and full[:21] == seen
    finally:
        ctx.close()


def test_callback_cancel_cpp_wrapper(gpu_ctx):
    """the C++ mirror of for_each_with_callback: the first exception thrown by the callback stops the run and is rethrown"""
    from allwave_b200 import hostlib as H

    c, ids, seqs, _ = synth.config("C5", n=10, length=600)
    gpu_ctx.load_sequences(ids, seqs)
    seen, msg = H.cancel_probe(gpu_ctx, ids, [len(s) for s in seqs], fail_at=7, chunk_pairs=8)
    assert (seen, msg) == (7, "callback failed on purpose")
    seen, msg = H.cancel_probe(gpu_ctx, ids, [len(s) for s in seqs], fail_at=10 ** 9)
    assert seen == -1


def test_failure_sentinel_line(gpu_ctx):
    """a pair that cannot be aligned still becomes a PAF line (src/alignment.rs:49-64): forced here with a
    48-diagonal workspace and the retry ladder switched off"""
    ids, seqs, _ = synth.generate(808, 4, 900, 0.08)
    seqs[3] = seqs[0]  # an identical pair needs one diagonal and still succeeds
    pairs = [(0, 1), (0, 3), (2, 1)]
    ctx = aw.Context(0)
    try:
        ctx.set_option("max_wavefront_width", 48)
        ctx.set_option("max_retry_attempts", 0)
        ctx.load_sequences(ids, seqs)
        res = ctx.align_pairs(aw.make_params(**DEFAULT), pairs, orientation=aw.AW_ORIENT_FORWARD)
        for r, (q, t) in zip(res, pairs):
            if (q, t) == (0, 3):
                assert r["status"] == 0 and r["cg"] == f"{len(seqs[0])}="
                continue
            assert r["status"] == aw._cabi.AW_EALIGN and r["score"] == 2 ** 31 - 1 and r["cg"] == ""
            assert (r["query_end"], r["target_end"], r["num_matches"], r["alignment_length"]) == (0, 0, 0, 0)
            assert r["paf"] == f"{ids[q]}\t{len(seqs[q])}\t0\t0\t+\t{ids[t]}\t{len(seqs[t])}\t0\t0\t0\t0\t60\tgi:f:0.000000\tcg:Z:"
        # with the ladder on, the same context settings give real alignments
        ctx.set_option("max_retry_attempts", 3)
        res = ctx.align_pairs(aw.make_params(**DEFAULT), pairs, orientation=aw.AW_ORIENT_FORWARD)
        assert all(r["status"] == 0 and r["cg"] for r in res)
    finally:
        ctx.close()


def test_wfa_orientation_through_retry_ladder(oracle):
    """--wfa-orientation with a first-try workspace that is too narrow for the wrong-strand alignment: both
    count-only passes must go through the retry ladder before the strand is picked"""
    c, ids, seqs, rc = synth.config("C5", n=6, length=1500)
    pairs = [(i, j) for i in range(6) for j in range(6) if i != j]
    ctx = aw.Context(0)
    try:
        ctx.set_option("max_wavefront_width", 64)
        ctx.load_sequences(ids, seqs)
        res = ctx.align_pairs(aw.make_params(**DEFAULT), pairs, orientation=aw.AW_ORIENT_WFA)
        exp = oracle.run_pairs(ids, seqs, pairs, oracle.params(**DEFAULT), use_mash=False, threads=CORES)
        for r, (q, t), e in zip(res, pairs, exp["paf"]):
            assert r["status"] == 0 and r["is_reverse"] == (rc[q] != rc[t]) and r["paf"] == e
    finally:
        ctx.close()


def test_divergence_estimate_orders_pairs(gpu_ctx):
    """aw_estimate_divergence (scheduling aid): monotone in the true divergence of C4-shaped haplotypes"""
    rnd = random.Random(3)
    root = bytes(rnd.choice(b"ACGT") for _ in range(20000))
    seqs = [root] + [_mutate(rnd, root, d) for d in (0.002, 0.01, 0.03, 0.08)]
    ids = ["d%d" % i for i in range(len(seqs))]
    gpu_ctx.load_sequences(ids, seqs)
    est = gpu_ctx.estimate_divergence([(0, i) for i in range(1, 5)])
    assert est == sorted(est) and est[0] < 0.01 and est[-1] > 0.03


def test_run_job_matches_cli_semantics(oracle, tmp_path):
    """awh_run_job (what bench.py's strong-scaling mode drives) = the CLI path: same lines as the oracle, digest stable"""
    from allwave_b200 import hostlib as H

    c, ids, seqs, rc = synth.config("C5", n=16, length=1000)
    out = tmp_path / "job.paf"
    r = H.run_job(ids, seqs, sparsification="none", out_path=out, checksum=True)
    pairs = [(i, j) for i in range(16) for j in range(16) if i != j]
    exp = oracle.run_pairs(ids, seqs, pairs, oracle.params(**DEFAULT), use_mash=True, threads=CORES)
    assert r["pairs"] == len(pairs) and out.read_text().splitlines() == exp["paf"]
    r2 = H.run_job(ids, seqs, sparsification="none", checksum=True)
    assert r2["digest"] == r["digest"] and r2["paf_bytes"] == r["paf_bytes"] == sum(len(l) + 1 for l in exp["paf"])
    if aw._cabi.lib().aw_device_count() >= 2:
        r3 = H.run_job(ids, seqs, sparsification="none", n_gpus=2, checksum=True)
        assert r3["digest"] == r["digest"] and r3["pairs"] == r["pairs"]


# ---- Mb-scale regime: one pair per thread-block cluster -----------------------------------------

def _cluster_ctx(cluster_min_len, solo_len):
    ctx = aw.Context(0)
    ctx.set_option("cluster_min_len", cluster_min_len)
    ctx.set_option("cluster_always", 1)
    ctx.set_option("solo_len", solo_len)
    return ctx


@pytest.mark.parametrize("solo_len", [0, 3000, 40000])
def test_cluster_kernel_parity(oracle, solo_len):
    """the cluster kernel (8 CTAs per pair, wavefront cells split over the cluster, reductions folded through distributed
    shared memory) on C4-shaped pairs short enough for the oracle: every sub-problem cluster-wide (solo_len 0), a mix, and
    almost everything on CTA 0 -- each must give the oracle's PAF and the single-CTA kernel's"""
    c, ids, seqs, _ = synth.config("C4", n=4, length=30000)
    pairs = [(0, 1), (1, 2), (2, 3), (3, 0), (0, 2)]
    exp = oracle.run_pairs(ids, seqs, pairs, oracle.params(**DEFAULT), use_mash=True, threads=CORES, fast=True)
    ctx = _cluster_ctx(1000, solo_len)
    try:
        ctx.load_sequences(ids, seqs)
        res = ctx.align_pairs(aw.make_params(**DEFAULT), pairs)
        assert [r["status"] for r in res] == [0] * len(pairs)
        assert [r["paf"] for r in res] == exp["paf"] and [r["score"] for r in res] == exp["scores"]
    finally:
        ctx.close()


def test_cluster_kernel_edge_cases(oracle):
    """pairs that end in the base case at once (identical, one long gap), affine single-piece penalties, wfa orientation"""
    rnd = random.Random(77)
    a = bytes(rnd.choice(b"ACGT") for _ in range(40000))
    seqs = [a, a, a[:25000], a[5000:], _mutate(rnd, a, 0.01), oracle.reverse_complement(_mutate(rnd, a, 0.02))]
    ids = ["k%d" % i for i in range(len(seqs))]
    pairs = [(0, 1), (0, 2), (2, 0), (3, 0), (0, 4), (4, 5), (5, 0)]
    ctx = _cluster_ctx(1000, 2000)
    try:
        ctx.load_sequences(ids, seqs)
        for pen in (DEFAULT, dict(mismatch=4, gap_open=6, gap_extend=2, gap2_open=None, gap2_extend=None)):
            exp = oracle.run_pairs(ids, seqs, pairs, oracle.params(**pen), use_mash=True, threads=CORES, fast=True)
            res = ctx.align_pairs(aw.make_params(**pen), pairs)
            assert [r["paf"] for r in res] == exp["paf"]
    finally:
        ctx.close()


def test_c4_full_size_pairs_vs_golden():
    """BASELINE config 4 as generated (200 x 1 Mb haplotypes, seed 4, SVs, giant:0.99): the first pairs of the list against the
    frozen oracle results in tests/golden/c4_pairs.json.gz (tests/make_golden_c4.py; the oracle needs ~10 minutes per pair)"""
    import gzip
    import hashlib
    import json

    from allwave_b200 import hostlib as H

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c4_pairs.json.gz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/c4_pairs.json.gz has not been generated")
    gold = json.load(gzip.open(path, "rt"))
    c, ids, seqs, _ = synth.config("C4")
    assert (c["seed"], c["n"], c["length"]) == (gold["seed"], gold["n"], gold["length"])
    full = H.pair_list(ids, H.KIND_CONNECTIVITY, value=0.99)
    assert len(full) == gold["n_pairs_in_list"] and hashlib.sha1(json.dumps([list(p) for p in full]).encode()).hexdigest() == gold["pair_list_sha1"]
    pairs = [(g["q"], g["t"]) for g in gold["pairs"]]
    assert pairs == [tuple(p) for p in full[: len(pairs)]]
    ctx = aw.Context(0)
    try:
        ctx.load_sequences(ids, seqs)
        res = ctx.align_pairs(aw.make_params(**DEFAULT), pairs)
        for r, g in zip(res, gold["pairs"]):
            assert r["status"] == 0 and r["score"] == g["score"] and r["paf"] == g["paf"], (g["q"], g["t"], r["score"], g["score"])
    finally:
        ctx.close()


def test_match_score_bonus(oracle, gpu_ctx):
    """match_score < 0: WFA2 shifts the penalties (x' = 2x - 2M, o' = 2o, e' = 2e - M; include/aw_wfa2_compat.h) and reports
    the CIGAR's score under the user's own penalties.  Same op strings, cg and scores as the oracle; match > 0 is refused."""
    ids, seqs, pairs = _sweep_group(977, nfam=24, members=4, maxlen=320, npairs=2500)
    for pen in (dict(match=-1, mismatch=4, gap_open=6, gap_extend=2, gap2_open=None, gap2_extend=None),
                dict(match=-2, mismatch=5, gap_open=8, gap_extend=2, gap2_open=24, gap2_extend=1),
                dict(match=-3, mismatch=2, gap_open=1, gap_extend=1, gap2_open=None, gap2_extend=None)):
        res = _paf_parity(oracle, gpu_ctx, ids, seqs, pairs, pen)
        assert any(r["score"] > 0 for r in res)        # a bonus makes near-identical pairs score positive
    ids2, seqs2, _ = synth.generate(12, 6, 9000, 0.04)
    pairs2 = [(i, j) for i in range(6) for j in range(6) if i != j]
    _paf_parity(oracle, gpu_ctx, ids2, seqs2, pairs2, dict(match=-1, mismatch=5, gap_open=8, gap_extend=2, gap2_open=24, gap2_extend=1))
    gpu_ctx.load_sequences(ids, seqs)
    with pytest.raises(aw.AllwaveError) as ei:
        gpu_ctx.align_pairs(aw.make_params(match=1, mismatch=4, gap_open=6, gap_extend=2, gap2_open=None, gap2_extend=None), pairs[:4])
    assert ei.value.status == aw.AW_EUNSUPPORTED


def test_real_wfa2_vectors_gpu(gpu_ctx):
    """the CUDA path against the output of a real WFA2-lib build (tests/golden/wfa2_lib_vectors.tsv.gz, produced by
    tools/wfa2_vectors/dump_wfa2_vectors.c); skipped, saying so, until a maintainer with WFA2-lib commits that file"""
    import gzip
    here = os.path.dirname(os.path.abspath(__file__))
    vec = os.environ.get("AW_WFA2_VECTORS") or os.path.join(here, "golden", "wfa2_lib_vectors.tsv.gz")
    if not os.path.exists(vec):
        pytest.skip("PARITY UNPINNED: no WFA2-lib vectors in tests/golden/ (see tools/wfa2_vectors/dump_wfa2_vectors.c)")
    inputs = {}
    with gzip.open(os.path.join(here, "golden", "wfa2_inputs.tsv.gz"), "rt") as f:
        for line in f:
            v = line.rstrip("\n").split("\t")
            inputs[v[0]] = (tuple(int(x) for x in v[1:7]), v[7].encode(), v[8].encode())
    by_pen = {}
    with gzip.open(vec, "rt") as f:
        for line in f:
            k, score, ops = (line.rstrip("\n").split("\t") + [""])[:3]
            by_pen.setdefault(inputs[k][0], []).append((k, int(score), ops))
    for (m, x, o1, e1, o2, e2), items in by_pen.items():
        seqs, ids = [], []
        for k, _, _ in items:
            seqs += [inputs[k][1], inputs[k][2]]
            ids += [k + "q", k + "t"]
        gpu_ctx.load_sequences(ids, seqs)
        pen = dict(match=m, mismatch=x, gap_open=o1, gap_extend=e1, gap2_open=o2 if o2 >= 0 else None, gap2_extend=e2 if o2 >= 0 else None)
        res = gpu_ctx.align_pairs(aw.make_params(**pen), [(2 * i, 2 * i + 1) for i in range(len(items))], orientation=aw.AW_ORIENT_FORWARD,
                                  flags=aw.AW_FLAG_CIGAR_BYTES)
        for r, (k, score, ops) in zip(res, items):
            assert r["status"] == 0 and r["score"] == score and r["cigar_bytes"].decode() == ops, k


def test_legacy_align_sequences(oracle, gpu_ctx):
    """wfa::align_sequences of the C++ mirror (src/wfa.rs:178-258): a fresh aligner per call in each of the three modes; score,
    standard-convention CIGAR and the five counters against the oracle; the README example of WFA2-lib"""
    from allwave_b200 import hostlib as H

    r = H.align_sequences(gpu_ctx, b"TCTTTACTCGCGCGTTGGAGAAATACAATAGT", b"TCTATACTGCGCGTTTGGAGAAATAAAATAGT", 4, 6, 2, mode=H.MODE_AFFINE)
    assert (r["score"], r["cigar"]) == (-24, "3=1X4=1I7=1D9=1X6=")
    rnd = random.Random(5)
    for mode, pen in ((H.MODE_EDIT, (3, 0, 0, 0, 0)), (H.MODE_AFFINE, (4, 6, 2, 0, 0)), (H.MODE_AFFINE2P, (5, 8, 2, 24, 1))):
        for _ in range(6):
            a = bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(1, 700)))
            b = _mutate(rnd, a, 0.08) or b"A"
            r = H.align_sequences(gpu_ctx, a, b, *pen, mode=mode)
            if mode == H.MODE_EDIT:
                p = oracle.params(0, pen[0], pen[0], pen[0], None, None)
            elif mode == H.MODE_AFFINE:
                p = oracle.params(0, pen[0], pen[1], pen[2], None, None)
            else:
                p = oracle.params(0, *pen)
            st, sc, ops, _ = oracle.wfa_align(p, a, b)
            assert st == 0 and r["score"] == sc and r["cigar"] == oracle.cigar_string(ops)
            assert (r["matches"], r["mismatches"], r["insertions"], r["deletions"]) == (ops.count(b"M"), ops.count(b"X"), ops.count(b"D"), ops.count(b"I"))
            assert r["alignment_length"] == r["matches"] + r["mismatches"]
