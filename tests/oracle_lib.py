"""ctypes binding of the CPU oracle (oracle/liballwave_oracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never import this from allwave_b200/.
"""
import ctypes as C
import os
import subprocess

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")
_SO = os.path.join(_ORACLE_DIR, "liballwave_oracle.so")


class Params(C.Structure):
    _fields_ = [
        ("match_score", C.c_int32),
        ("mismatch_penalty", C.c_int32),
        ("gap_open", C.c_int32),
        ("gap_extend", C.c_int32),
        ("gap2_open", C.c_int32),
        ("gap2_extend", C.c_int32),
        ("has_gap2_open", C.c_uint8),
        ("has_gap2_extend", C.c_uint8),
    ]


class Work(C.Structure):
    _fields_ = [
        ("cells", C.c_uint64),
        ("steps", C.c_uint64),
        ("ext_chars", C.c_uint64),
        ("n_breakpoints", C.c_uint32),
        ("n_base", C.c_uint32),
        ("max_base_score", C.c_uint32),
        ("max_level", C.c_uint32),
        ("n_end_reached_fallbacks", C.c_uint32),
        ("max_width", C.c_uint64),
    ]


class Alignment(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("score", C.c_int32),
        ("cigar", C.POINTER(C.c_uint8)),
        ("cigar_len", C.c_size_t),
        ("work", Work),
    ]


class Result(C.Structure):
    _fields_ = [
        ("query_idx", C.c_uint64),
        ("target_idx", C.c_uint64),
        ("query_start", C.c_uint64),
        ("query_end", C.c_uint64),
        ("target_start", C.c_uint64),
        ("target_end", C.c_uint64),
        ("is_reverse", C.c_uint8),
        ("score", C.c_int32),
        ("num_matches", C.c_uint64),
        ("alignment_length", C.c_uint64),
        ("cigar_bytes", C.POINTER(C.c_uint8)),
        ("cigar_len", C.c_size_t),
        ("work", Work),
    ]


class Sparsification(C.Structure):
    _fields_ = [
        ("kind", C.c_int),
        ("fraction", C.c_double),
        ("k_nearest", C.c_uint64),
        ("k_farthest", C.c_uint64),
        ("kmer_size", C.c_int),
    ]


class Job(C.Structure):
    _fields_ = [
        ("npairs", C.c_uint64),
        ("paf", C.POINTER(C.c_char_p)),
        ("scores", C.POINTER(C.c_int32)),
        ("sum_block_len", C.c_uint64),
        ("work", Work),
        ("seconds", C.c_double),
    ]


SPARS_NONE, SPARS_RANDOM, SPARS_AUTO, SPARS_GIANT, SPARS_TREE = range(5)

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _ORACLE_DIR])


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    u8p = C.POINTER(C.c_uint8)
    L.awo_wfa_align.argtypes = [C.POINTER(Params), C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(Alignment)]
    L.awo_wfa_align.restype = C.c_int
    L.awo_wfa_align_unidirectional.argtypes = L.awo_wfa_align.argtypes
    L.awo_wfa_align_unidirectional.restype = C.c_int
    L.awo_alignment_free.argtypes = [C.POINTER(Alignment)]
    L.awo_cigar_penalty.argtypes = [C.POINTER(Params), C.c_char_p, C.c_size_t]
    L.awo_cigar_penalty.restype = C.c_int64
    L.awo_gotoh_penalty.argtypes = [C.POINTER(Params), C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    L.awo_gotoh_penalty.restype = C.c_int64
    L.awo_siphash.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_char_p, C.c_size_t]
    L.awo_siphash.restype = C.c_uint64
    L.awo_hash_bytes.argtypes = [C.c_char_p, C.c_size_t]
    L.awo_hash_bytes.restype = C.c_uint64
    L.awo_hash_str.argtypes = [C.c_char_p, C.c_size_t]
    L.awo_hash_str.restype = C.c_uint64
    for f in (L.awo_sketch_stranded, L.awo_sketch_canonical):
        f.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_size_t, C.POINTER(C.c_uint64)]
        f.restype = C.c_size_t
    L.awo_jaccard_counts.argtypes = [C.POINTER(C.c_uint64), C.c_size_t, C.POINTER(C.c_uint64), C.c_size_t,
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.awo_jaccard.argtypes = L.awo_jaccard_counts.argtypes[:4]
    L.awo_jaccard.restype = C.c_double
    L.awo_mash_distance.argtypes = L.awo_jaccard_counts.argtypes[:4] + [C.c_int]
    L.awo_mash_distance.restype = C.c_double
    L.awo_reverse_complement.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
    L.awo_orientation_mash.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    L.awo_orientation_mash.restype = C.c_int
    L.awo_cigar_bytes_to_string.argtypes = [C.c_char_p, C.c_size_t]
    L.awo_cigar_bytes_to_string.restype = C.c_void_p
    L.awo_align_pair.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_uint64, C.c_uint64,
                                 C.POINTER(Params), C.POINTER(Params), C.c_int, C.POINTER(Result)]
    L.awo_result_free.argtypes = [C.POINTER(Result)]
    L.awo_alignment_to_paf.argtypes = [C.POINTER(Result), C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    L.awo_alignment_to_paf.restype = C.c_void_p
    L.awo_parse_scores.argtypes = [C.c_char_p, C.POINTER(Params)]
    L.awo_parse_scores.restype = C.c_int
    L.awo_connectivity_probability.argtypes = [C.c_uint64, C.c_double]
    L.awo_connectivity_probability.restype = C.c_double
    L.awo_pair_list.argtypes = [C.c_uint64, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_size_t),
                                C.c_int, C.POINTER(Sparsification), C.POINTER(C.c_uint64)]
    L.awo_pair_list.restype = C.POINTER(C.c_uint64)
    L.awo_distance_matrix.argtypes = [C.c_uint64, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, C.c_size_t]
    L.awo_distance_matrix.restype = C.POINTER(C.c_double)
    L.awo_build_knn_graph.argtypes = [C.POINTER(C.c_double), C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
    L.awo_build_knn_graph.restype = C.POINTER(C.c_uint64)
    L.awo_run_pairs.argtypes = [C.c_uint64, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_size_t),
                                C.POINTER(C.c_uint64), C.c_uint64, C.POINTER(Params), C.c_int, C.c_int, C.POINTER(Job)]
    L.awo_run_pairs.restype = C.c_int
    L.awo_job_free.argtypes = [C.POINTER(Job)]
    L.free = C.CDLL(None).free
    L.free.argtypes = [C.c_void_p]
    _lib = L
    return L


def mode(p):
    """AlignmentMode::from_params: 0 EditDistance, 1 SinglePieceAffine, 2 TwoPieceAffine"""
    f = lib().awo_mode_from_params
    f.argtypes = [C.POINTER(Params)]
    return f(C.byref(p))


def parse_scores(s):
    p = Params()
    rc = lib().awo_parse_scores(s.encode(), C.byref(p))
    if rc != 0:
        raise ValueError("Failed to parse scores" if rc == -1 else "Invalid number of scores")
    return p


def params(match=0, mismatch=5, gap_open=8, gap_extend=2, gap2_open=24, gap2_extend=1):
    p = Params(match, mismatch, gap_open, gap_extend, 0, 0, 0, 0)
    if gap2_open is not None:
        p.gap2_open, p.has_gap2_open = gap2_open, 1
    if gap2_extend is not None:
        p.gap2_extend, p.has_gap2_extend = gap2_extend, 1
    return p


def _take_str(ptr):
    s = C.string_at(ptr).decode()
    lib().free(ptr)
    return s


def wfa_align(p, pattern: bytes, text: bytes, unidirectional=False):
    """-> (status, score, ops bytes, work dict)"""
    al = Alignment()
    f = lib().awo_wfa_align_unidirectional if unidirectional else lib().awo_wfa_align
    rc = f(C.byref(p), pattern, len(pattern), text, len(text), C.byref(al))
    if rc != 0:
        return al.status, al.score, b"", {}
    ops = C.string_at(al.cigar, al.cigar_len)
    work = {k: getattr(al.work, k) for k, _ in Work._fields_}
    lib().awo_alignment_free(C.byref(al))
    return 0, al.score, ops, work


def cigar_penalty(p, ops: bytes):
    return lib().awo_cigar_penalty(C.byref(p), ops, len(ops))


def gotoh_penalty(p, pattern: bytes, text: bytes):
    return lib().awo_gotoh_penalty(C.byref(p), pattern, len(pattern), text, len(text))


def hash_bytes(b: bytes):
    return lib().awo_hash_bytes(b, len(b))


def hash_str(s: str):
    b = s.encode()
    return lib().awo_hash_str(b, len(b))


def sketch(seq: bytes, k=15, size=1000, canonical=False):
    out = (C.c_uint64 * size)()
    f = lib().awo_sketch_canonical if canonical else lib().awo_sketch_stranded
    n = f(seq, len(seq), k, size, out)
    return list(out[:n])


def jaccard_counts(a, b):
    A = (C.c_uint64 * max(1, len(a)))(*a)
    B = (C.c_uint64 * max(1, len(b)))(*b)
    i, u = C.c_uint64(), C.c_uint64()
    lib().awo_jaccard_counts(A, len(a), B, len(b), C.byref(i), C.byref(u))
    return i.value, u.value


def mash_distance(a, b, k=15):
    A = (C.c_uint64 * max(1, len(a)))(*a)
    B = (C.c_uint64 * max(1, len(b)))(*b)
    return lib().awo_mash_distance(A, len(a), B, len(b), k)


def reverse_complement(seq: bytes):
    out = C.create_string_buffer(len(seq) + 1)
    lib().awo_reverse_complement(seq, len(seq), out)
    return out.raw[: len(seq)]


def orientation_mash(query: bytes, target: bytes):
    return lib().awo_orientation_mash(query, len(query), target, len(target))


def cigar_string(ops: bytes):
    return _take_str(lib().awo_cigar_bytes_to_string(ops, len(ops)))


def align_pair(query: bytes, target: bytes, qi, ti, p, use_mash=True, orientation_params=None, qname=None, tname=None):
    """-> dict mirroring AlignmentResult (+ 'paf' when names are given)"""
    op = orientation_params or Params(0, 1, 1, 1, 0, 0, 0, 0)
    r = Result()
    lib().awo_align_pair(query, len(query), target, len(target), qi, ti, C.byref(p), C.byref(op), 1 if use_mash else 0,
                         C.byref(r))
    d = {k: getattr(r, k) for k, _ in Result._fields_ if k not in ("cigar_bytes", "work")}
    d["cigar_bytes"] = C.string_at(r.cigar_bytes, r.cigar_len) if r.cigar_len else b""
    d["work"] = {k: getattr(r.work, k) for k, _ in Work._fields_}
    if qname is not None:
        d["paf"] = _take_str(lib().awo_alignment_to_paf(C.byref(r), qname.encode(), len(query), tname.encode(), len(target)))
    lib().awo_result_free(C.byref(r))
    return d


def _seq_arrays(ids, seqs):
    n = len(ids)
    idarr = (C.c_char_p * max(1, n))(*[i.encode() for i in ids])
    seqarr = (C.c_char_p * max(1, n))(*seqs) if seqs is not None else None
    lens = (C.c_size_t * max(1, n))(*[len(s) for s in seqs]) if seqs is not None else None
    return idarr, seqarr, lens


def pair_list(ids, seqs, kind=SPARS_NONE, fraction=0.0, k_nearest=0, k_farthest=0, kmer_size=0, exclude_self=True):
    n = len(ids)
    idarr, seqarr, lens = _seq_arrays(ids, seqs)
    sp = Sparsification(kind, fraction, k_nearest, k_farthest, kmer_size)
    cnt = C.c_uint64()
    ptr = lib().awo_pair_list(n, idarr, seqarr, lens, 1 if exclude_self else 0, C.byref(sp), C.byref(cnt))
    out = [(ptr[2 * i], ptr[2 * i + 1]) for i in range(cnt.value)]
    lib().free(ptr)
    return out


def distance_matrix(seqs, k=15, size=1000):
    n = len(seqs)
    _, seqarr, lens = _seq_arrays(["x"] * n, seqs)
    ptr = lib().awo_distance_matrix(n, seqarr, lens, k, size)
    out = [[ptr[i * n + j] for j in range(n)] for i in range(n)]
    lib().free(ptr)
    return out


def build_knn_graph(matrix, k, farthest):
    n = len(matrix)
    flat = (C.c_double * max(1, n * n))(*[v for row in matrix for v in row])
    cnt = C.c_uint64()
    ptr = lib().awo_build_knn_graph(flat, n, k, 1 if farthest else 0, C.byref(cnt))
    out = [(ptr[2 * i], ptr[2 * i + 1]) for i in range(cnt.value)]
    lib().free(ptr)
    return out


def set_fast(on):
    """opt-in fast mode of the biWFA restatement (free lists, unchecked interior loop, 8-byte extend): for CPU baseline legs"""
    lib().awo_set_fast(1 if on else 0)


def run_pairs(ids, seqs, pairs, p, use_mash=True, threads=1, fast=False):
    """whole-job CPU driver -> dict(paf=[...], scores=[...], seconds=..., sum_block_len=..., work={...})"""
    if fast:
        set_fast(True)
        try:
            return run_pairs(ids, seqs, pairs, p, use_mash, threads, fast=False)
        finally:
            set_fast(False)
    n = len(ids)
    idarr, seqarr, lens = _seq_arrays(ids, seqs)
    flat = (C.c_uint64 * max(1, 2 * len(pairs)))()
    for i, (a, b) in enumerate(pairs):
        flat[2 * i], flat[2 * i + 1] = a, b
    job = Job()
    lib().awo_run_pairs(n, idarr, seqarr, lens, flat, len(pairs), C.byref(p), 1 if use_mash else 0, threads, C.byref(job))
    out = {
        "paf": [job.paf[i].decode() for i in range(job.npairs)],
        "scores": [job.scores[i] for i in range(job.npairs)],
        "seconds": job.seconds,
        "sum_block_len": job.sum_block_len,
        "work": {k: getattr(job.work, k) for k, _ in Work._fields_},
    }
    lib().awo_job_free(C.byref(job))
    return out
