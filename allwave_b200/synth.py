"""Synthetic FASTA generator binding (allwave_b200/host/synth.c) and the five BASELINE configs."""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_PKG, "liballwave_synth.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise ImportError(f"{_SO} is missing: run __graft_entry__.build()")
        L = C.CDLL(_SO)
        L.aws_generate.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.c_int, C.c_double,
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)]
        L.aws_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def generate(seed, n, length, d_lo, d_hi=None, n_sv=0, rc_prob=0.0):
    """-> (ids, seqs[bytes], is_rc[list])  names s000000..."""
    L = _load()
    d_hi = d_lo if d_hi is None else d_hi
    ptrs = (C.c_void_p * max(1, n))()
    lens = (C.c_uint64 * max(1, n))()
    rc = (C.c_uint8 * max(1, n))()
    L.aws_generate(seed, n, length, d_lo, d_hi, n_sv, rc_prob, ptrs, lens, rc)
    seqs = []
    for i in range(n):
        seqs.append(C.string_at(ptrs[i], lens[i]))
        L.aws_free(ptrs[i])
    ids = ["s%06d" % i for i in range(n)]
    return ids, seqs, [int(rc[i]) for i in range(n)]


# the five BASELINE.json configs.  `scale` < 1 shrinks n for tests.
CONFIGS = {
    "C1": dict(seed=1, n=16, length=10000, d_lo=0.01, scores="0,5,8,2,24,1", sparsification="none"),
    "C2": dict(seed=2, n=1000, length=10000, d_lo=0.05, scores="0,5,8,2,24,1", sparsification="none"),
    "C3": dict(seed=3, n=1415, length=150, d_lo=0.02, scores="0,1,1,1", sparsification="none"),
    "C4": dict(seed=4, n=200, length=1000000, d_lo=0.001, d_hi=0.02, n_sv=5, scores="0,5,8,2,24,1", sparsification="giant:0.99"),
    "C5": dict(seed=5, n=5000, length=5000, d_lo=0.03, rc_prob=0.5, scores="0,5,8,2,24,1", sparsification="tree:2:1:0.1"),
}


def config(name, n=None, length=None):
    c = dict(CONFIGS[name])
    if n is not None:
        c["n"] = n
    if length is not None:
        c["length"] = length
    ids, seqs, rc = generate(c["seed"], c["n"], c["length"], c["d_lo"], c.get("d_hi"), c.get("n_sv", 0), c.get("rc_prob", 0.0))
    return c, ids, seqs, rc
