"""Writes tests/golden/wfa2_inputs.tsv.gz: the seeded input set for tools/wfa2_vectors/dump_wfa2_vectors.c (see there).
Covers what the co-optimal CIGAR choice depends on: every CLI penalty preset and the odd orderings (o1 > o2, e2 >= e1, x > o+e,
o = 0, match < 0), lengths on both sides of the biWFA thresholds (max(len) <= 100 goes straight to the base case, remaining
score <= 250 ends the recursion), repeats and homopolymers (many co-optimal paths), unrelated pairs, empty-ish sequences.
usage: python tools/wfa2_vectors/make_inputs.py"""
import gzip, os, random

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PENS = [(0, 5, 8, 2, 24, 1), (0, 7, 12, 2, 36, 1), (0, 4, 6, 2, 18, 1), (0, 3, 4, 1, -1, -1), (0, 1, 1, 1, -1, -1), (0, 4, 6, 2, -1, -1),
        (0, 4, 12, 1, 6, 3), (0, 2, 3, 2, 5, 2), (0, 9, 2, 1, 5, 1), (0, 6, 0, 3, -1, -1), (0, 2, 7, 3, 2, 4),
        (-1, 4, 6, 2, -1, -1), (-2, 5, 8, 2, 24, 1)]


def mutate(rnd, s, d):
    out, i = [], 0
    while i < len(s):
        if rnd.random() < d:
            r = rnd.random()
            if r < 0.5:
                out.append(rnd.choice("ACGT")); i += 1
            elif r < 0.65:
                i += 1 + int(rnd.expovariate(0.4))
            elif r < 0.8:
                out.extend(rnd.choice("ACGT") for _ in range(1 + int(rnd.expovariate(0.4))))
            elif r < 0.9:
                out.append(s[i] * rnd.randint(2, 12)); i += 1
            else:
                u = s[i:i + rnd.randint(2, 5)]
                out.append(u * rnd.randint(2, 6)); i += len(u)
        else:
            out.append(s[i]); i += 1
    return "".join(out)


def main():
    rnd = random.Random(20260101)
    rows = []
    lengths = [1, 2, 5, 17, 64, 99, 100, 101, 150, 250, 400, 800, 1500, 3000, 6000, 12000]
    for pi, pen in enumerate(PENS):
        for L in lengths:
            if L > 3000 and pi not in (0, 3, 4, 6):
                continue
            for kind in range(4):
                if kind == 0:
                    root = "".join(rnd.choice("ACGT") for _ in range(L))
                elif kind == 1:
                    u = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(1, 6)))
                    root = (u * (L // len(u) + 1))[:L]
                elif kind == 2:
                    root = "".join(rnd.choice("ACGT") * rnd.randint(1, 20) for _ in range(L))[:L]
                else:
                    root = "".join(rnd.choice("AC") for _ in range(L))
                for d in (0.0, 0.01, 0.05, 0.15, 0.3):
                    if L > 3000 and d > 0.05:
                        continue
                    rows.append((pen, root, mutate(rnd, root, d)))
            rows.append((pen, "".join(rnd.choice("ACGT") for _ in range(L)), "".join(rnd.choice("ACGT") for _ in range(max(1, L // 2)))))  # unrelated
    out = os.path.join(ROOT, "tests", "golden", "wfa2_inputs.tsv.gz")
    with gzip.GzipFile(out, "wb", mtime=0) as f:
        for i, (pen, a, b) in enumerate(rows):
            if not a:
                a = "A"
            if not b:
                b = "A"   # WFA2-lib's own tools never align an empty sequence; allwave handles those before the aligner
            f.write(("v%05d\t%s\t%s\t%s\n" % (i, "\t".join(str(v) for v in pen), a, b)).encode())
    print(len(rows), "vectors ->", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
