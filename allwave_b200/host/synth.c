/*
 * synth.c -- deterministic synthetic FASTA generator for the five BASELINE configs
 * (SURVEY.md section 8d): xoshiro256** seeded through splitmix64; an ancestor of L uniform
 * ACGT bases; haplotypes that mutate it per base with probability d (80% substitution,
 * 10% insertion, 10% deletion, indel length geometric(0.5) capped at 10); optional
 * structural variants (C4) and random reverse-complementing (C5).  Inspired by
 * /root/reference/src/test_framework.rs:78-95,131-317 (whose ChaCha stream cannot be
 * reproduced without the `rand` crate); shares nothing with the oracle.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s[4]; } aws_rng;
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t splitmix64(uint64_t* x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static void rng_seed(aws_rng* r, uint64_t seed) { for (int i = 0; i < 4; ++i) r->s[i] = splitmix64(&seed); }
static uint64_t rng_next(aws_rng* r) {
    uint64_t* s = r->s;
    const uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
static double rng_f(aws_rng* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static uint64_t rng_below(aws_rng* r, uint64_t n) { return n ? rng_next(r) % n : 0; }

typedef struct { uint8_t* p; size_t n, cap; } buf_t;
static void push(buf_t* b, uint8_t c) {
    if (b->n == b->cap) { b->cap = b->cap ? b->cap * 2 : 1024; b->p = (uint8_t*)realloc(b->p, b->cap); }
    b->p[b->n++] = c;
}
static void push_n(buf_t* b, const uint8_t* s, size_t n) { for (size_t i = 0; i < n; ++i) push(b, s[i]); }
static const char ACGT[4] = {'A', 'C', 'G', 'T'};
static int indel_len(aws_rng* r) { int L = 1; while (L < 10 && rng_f(r) < 0.5) ++L; return L; }

/* Generates n haplotypes of an ancestor of length L.  d_lo..d_hi: per-haplotype divergence
 * drawn uniformly; n_sv>0 adds 2 deletions + 2 tandem duplications (1..20 kb) + 1 novel
 * insertion (1..10 kb) per haplotype (scaled down when L is small); rc_prob: probability
 * that a haplotype is reverse-complemented.  seqs_out[i] is malloc'ed; is_rc_out optional. */
int aws_generate(uint64_t seed, uint32_t n, uint32_t L, double d_lo, double d_hi, int n_sv, double rc_prob,
                 uint8_t** seqs_out, uint64_t* lens_out, uint8_t* is_rc_out) {
    aws_rng r;
    rng_seed(&r, seed);
    uint8_t* anc = (uint8_t*)malloc(L ? L : 1);
    for (uint32_t i = 0; i < L; ++i) anc[i] = (uint8_t)ACGT[rng_next(&r) & 3];
    for (uint32_t hi = 0; hi < n; ++hi) {
        const double d = d_lo + (d_hi - d_lo) * rng_f(&r);
        buf_t b = {NULL, 0, 0};
        uint32_t i = 0;
        while (i < L) {
            if (rng_f(&r) < d) {
                const double t = rng_f(&r);
                if (t < 0.8) {
                    uint8_t c;
                    do { c = (uint8_t)ACGT[rng_next(&r) & 3]; } while (c == anc[i]);
                    push(&b, c);
                    ++i;
                } else if (t < 0.9) {
                    int len = indel_len(&r);
                    for (int j = 0; j < len; ++j) push(&b, (uint8_t)ACGT[rng_next(&r) & 3]);
                } else {
                    i += (uint32_t)indel_len(&r);
                }
            } else {
                push(&b, anc[i]);
                ++i;
            }
        }
        if (n_sv > 0) {
            const uint64_t big = L >= 200000 ? 20000 : (L / 10 ? L / 10 : 1), ins_max = L >= 200000 ? 10000 : (L / 20 ? L / 20 : 1);
            for (int sv = 0; sv < 5; ++sv) {
                buf_t o = {NULL, 0, 0};
                const uint64_t len = 1 + rng_below(&r, sv < 4 ? big : ins_max);
                const uint64_t pos = rng_below(&r, b.n > len ? b.n - len : 1);
                if (sv < 2) { /* deletion */
                    push_n(&o, b.p, pos);
                    if (pos + len < b.n) push_n(&o, b.p + pos + len, b.n - pos - len);
                } else if (sv < 4) { /* tandem duplication */
                    const uint64_t e = pos + len < b.n ? pos + len : b.n;
                    push_n(&o, b.p, e);
                    push_n(&o, b.p + pos, e - pos);
                    push_n(&o, b.p + e, b.n - e);
                } else { /* novel insertion */
                    push_n(&o, b.p, pos);
                    for (uint64_t j = 0; j < len; ++j) push(&o, (uint8_t)ACGT[rng_next(&r) & 3]);
                    push_n(&o, b.p + pos, b.n - pos);
                }
                free(b.p);
                b = o;
            }
        }
        uint8_t rc = 0;
        if (rc_prob > 0.0 && rng_f(&r) < rc_prob) {
            rc = 1;
            for (size_t a = 0, z = b.n; a < z; ) {
                --z;
                uint8_t x = b.p[a], y = b.p[z];
                #define COMP(c) ((c) == 'A' ? 'T' : (c) == 'T' ? 'A' : (c) == 'C' ? 'G' : 'C')
                if (a == z) { b.p[a] = COMP(x); break; }
                b.p[a] = COMP(y);
                b.p[z] = COMP(x);
                ++a;
            }
        }
        if (!b.p) b.p = (uint8_t*)malloc(1);
        seqs_out[hi] = b.p;
        lens_out[hi] = b.n;
        if (is_rc_out) is_rc_out[hi] = rc;
    }
    free(anc);
    return 0;
}
void aws_free(void* p) { free(p); }
