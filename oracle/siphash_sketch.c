/*
 * siphash_sketch.c -- oracle restatement of Rust's DefaultHasher (SipHash-1-3, zero keys) with
 * the std `Hash` framing, and of allwave's min-hash sketches / Jaccard / orientation.
 * TEST INFRASTRUCTURE ONLY (see aw_oracle.h).
 *
 * Follows: /root/reference/src/alignment.rs:69-154 (determine_orientation_mash,
 * sketch_sequence_stranded, jaccard_similarity, hash_kmer, is_dna_base), :178-190
 * (reverse_complement); /root/reference/src/mash.rs:40-135 (KmerSketch::jaccard, mash_distance,
 * sketch_sequence, reverse_complement_kmer).  Hash framing: SURVEY.md Appendix B.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../include/aw_wfa2_compat.h"
#include "aw_oracle.h"

#define ROTL(x, b) (uint64_t)(((x) << (b)) | ((x) >> (64 - (b))))
#define SIPROUND           \
    do {                   \
        v0 += v1;          \
        v1 = ROTL(v1, 13); \
        v1 ^= v0;          \
        v0 = ROTL(v0, 32); \
        v2 += v3;          \
        v3 = ROTL(v3, 16); \
        v3 ^= v2;          \
        v0 += v3;          \
        v3 = ROTL(v3, 21); \
        v3 ^= v0;          \
        v2 += v1;          \
        v1 = ROTL(v1, 17); \
        v1 ^= v2;          \
        v2 = ROTL(v2, 32); \
    } while (0)

uint64_t awo_siphash(int c_rounds, int d_rounds, uint64_t k0, uint64_t k1, const uint8_t* msg, size_t len) {
    uint64_t v0 = k0 ^ 0x736f6d6570736575ULL, v1 = k1 ^ 0x646f72616e646f6dULL;
    uint64_t v2 = k0 ^ 0x6c7967656e657261ULL, v3 = k1 ^ 0x7465646279746573ULL;
    size_t i = 0;
    for (; i + 8 <= len; i += 8) {
        uint64_t m = 0;
        for (int j = 0; j < 8; ++j) m |= (uint64_t)msg[i + j] << (8 * j);
        v3 ^= m;
        for (int r = 0; r < c_rounds; ++r) SIPROUND;
        v0 ^= m;
    }
    uint64_t b = (uint64_t)(len & 0xff) << 56;
    for (int j = 0; i + j < len; ++j) b |= (uint64_t)msg[i + j] << (8 * j);
    v3 ^= b;
    for (int r = 0; r < c_rounds; ++r) SIPROUND;
    v0 ^= b;
    v2 ^= 0xff;
    for (int r = 0; r < d_rounds; ++r) SIPROUND;
    return v0 ^ v1 ^ v2 ^ v3;
}

/* <[u8] as Hash>::hash: write_length_prefix(len as usize LE) then the bytes (hash_kmer,
 * src/alignment.rs:142-149, src/mash.rs:110-114) */
uint64_t awo_hash_bytes(const uint8_t* bytes, size_t len) {
    uint8_t stack[64];
    uint8_t* buf = len + 8 <= sizeof(stack) ? stack : (uint8_t*)malloc(len + 8);
    uint64_t l = (uint64_t)len;
    for (int j = 0; j < 8; ++j) buf[j] = (uint8_t)(l >> (8 * j));
    memcpy(buf + 8, bytes, len);
    uint64_t h = awo_siphash(1, 3, 0, 0, buf, len + 8);
    if (buf != stack) free(buf);
    return h;
}
/* <str as Hash>::hash: bytes then 0xFF (src/iterator.rs:272-274, src/knn_graph.rs:168-169) */
uint64_t awo_hash_str(const uint8_t* bytes, size_t len) {
    uint8_t stack[256];
    uint8_t* buf = len + 1 <= sizeof(stack) ? stack : (uint8_t*)malloc(len + 1);
    memcpy(buf, bytes, len);
    buf[len] = 0xFF;
    uint64_t h = awo_siphash(1, 3, 0, 0, buf, len + 1);
    if (buf != stack) free(buf);
    return h;
}

static int is_dna_base(uint8_t b) {
    if (b >= 'a' && b <= 'z') b = (uint8_t)(b - 32);
    return b == 'A' || b == 'C' || b == 'G' || b == 'T';
}
static int cmp_u64(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

static size_t sketch_impl(const uint8_t* seq, size_t len, int k, size_t sketch_size, uint64_t* out, int canonical) {
    if (len < (size_t)k) return 0;
    size_t nk = len - (size_t)k + 1, n = 0;
    uint64_t* hashes = (uint64_t*)malloc(sizeof(uint64_t) * nk);
    uint8_t rc[64];
    for (size_t i = 0; i < nk; ++i) {
        const uint8_t* kmer = seq + i;
        int ok = 1;
        for (int j = 0; j < k; ++j)
            if (!is_dna_base(kmer[j])) {
                ok = 0;
                break;
            }
        if (!ok) continue;
        uint64_t h = awo_hash_bytes(kmer, (size_t)k);
        if (canonical) {
            /* reverse_complement_kmer (src/mash.rs:122-135): upper-cased complement */
            for (int j = 0; j < k; ++j) {
                uint8_t b = kmer[k - 1 - j];
                uint8_t u = (b >= 'a' && b <= 'z') ? (uint8_t)(b - 32) : b;
                rc[j] = u == 'A' ? 'T' : u == 'T' ? 'A' : u == 'C' ? 'G' : u == 'G' ? 'C' : b;
            }
            uint64_t hr = awo_hash_bytes(rc, (size_t)k);
            if (hr < h) h = hr;
        }
        hashes[n++] = h;
    }
    qsort(hashes, n, sizeof(uint64_t), cmp_u64); /* sort_unstable; duplicates kept */
    if (n > sketch_size) n = sketch_size;        /* truncate */
    memcpy(out, hashes, n * sizeof(uint64_t));
    free(hashes);
    return n;
}
size_t awo_sketch_stranded(const uint8_t* seq, size_t len, int k, size_t sketch_size, uint64_t* out) {
    return sketch_impl(seq, len, k, sketch_size, out, 0);
}
size_t awo_sketch_canonical(const uint8_t* seq, size_t len, int k, size_t sketch_size, uint64_t* out) {
    return sketch_impl(seq, len, k, sketch_size, out, 1);
}

/* HashSet semantics on two ascending lists that may hold duplicates */
void awo_jaccard_counts(const uint64_t* a, size_t na, const uint64_t* b, size_t nb, uint64_t* inter, uint64_t* uni) {
    size_t i = 0, j = 0;
    uint64_t in = 0, un = 0;
    while (i < na || j < nb) {
        uint64_t x;
        if (j >= nb || (i < na && a[i] <= b[j])) x = a[i];
        else x = b[j];
        int ina = 0, inb = 0;
        while (i < na && a[i] == x) { ++i; ina = 1; }
        while (j < nb && b[j] == x) { ++j; inb = 1; }
        ++un;
        if (ina && inb) ++in;
    }
    *inter = in;
    *uni = un;
}
double awo_jaccard(const uint64_t* a, size_t na, const uint64_t* b, size_t nb) {
    uint64_t in, un;
    awo_jaccard_counts(a, na, b, nb, &in, &un);
    return un == 0 ? 0.0 : (double)in / (double)un;
}
double awo_mash_distance(const uint64_t* a, size_t na, const uint64_t* b, size_t nb, int k) {
    double j = awo_jaccard(a, na, b, nb);
    if (j <= 0.0) return 1.0;
    double ratio = (2.0 * j) / (1.0 + j);
    if (ratio <= 0.0) return 1.0;
    return (-1.0 / (double)k) * log(ratio);
}

void awo_reverse_complement(const uint8_t* seq, size_t len, uint8_t* out) {
    for (size_t i = 0; i < len; ++i) {
        uint8_t b = seq[len - 1 - i], c;
        switch (b) {
            case 'A': case 'a': c = 'T'; break;
            case 'T': case 't': c = 'A'; break;
            case 'C': case 'c': c = 'G'; break;
            case 'G': case 'g': c = 'C'; break;
            default: c = 'N'; break;
        }
        out[i] = c;
    }
}

int awo_orientation_mash(const uint8_t* query, size_t qlen, const uint8_t* target, size_t tlen) {
    uint64_t* ts = (uint64_t*)malloc(sizeof(uint64_t) * AW_SKETCH_SIZE * 3);
    uint64_t* fs = ts + AW_SKETCH_SIZE;
    uint64_t* rs = fs + AW_SKETCH_SIZE;
    uint8_t* rc = (uint8_t*)malloc(qlen ? qlen : 1);
    size_t nt = awo_sketch_stranded(target, tlen, AW_ORIENT_K, AW_SKETCH_SIZE, ts);
    size_t nf = awo_sketch_stranded(query, qlen, AW_ORIENT_K, AW_SKETCH_SIZE, fs);
    awo_reverse_complement(query, qlen, rc);
    size_t nr = awo_sketch_stranded(rc, qlen, AW_ORIENT_K, AW_SKETCH_SIZE, rs);
    double fj = awo_jaccard(fs, nf, ts, nt);
    double rj = awo_jaccard(rs, nr, ts, nt);
    free(rc);
    free(ts);
    return fj >= rj ? 0 : 1;
}
