"""The oracle's answers on tests/golden/wfa2_inputs.tsv.gz, in the output format of dump_wfa2_vectors.c -- to diff against a real
WFA2-lib run by hand, and to exercise the vector tests' plumbing (AW_WFA2_VECTORS=<this file> python -m pytest tests -k wfa2_vectors).
The result is the ORACLE's opinion: never commit it as tests/golden/wfa2_lib_vectors.tsv.gz.
usage: python tools/wfa2_vectors/dump_oracle_vectors.py out.tsv.gz [max_len]"""
import gzip, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O

out, max_len = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 30
O.set_fast(True)
n = 0
with gzip.open(os.path.join(ROOT, "tests", "golden", "wfa2_inputs.tsv.gz"), "rt") as f, gzip.open(out, "wt") as g:
    for line in f:
        v = line.rstrip("\n").split("\t")
        m, x, o1, e1, o2, e2 = (int(t) for t in v[1:7])
        a, b = v[7].encode(), v[8].encode()
        if max(len(a), len(b)) > max_len:
            continue
        st, sc, ops, _ = O.wfa_align(O.params(m, x, o1, e1, o2 if o2 >= 0 else None, e2 if o2 >= 0 else None), a, b)
        g.write(f"{v[0]}\t{sc if st == 0 else 'ERR'}\t{ops.decode()}\n")
        n += 1
print(n, "vectors ->", out)
