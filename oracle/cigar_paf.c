/*
 * cigar_paf.c -- oracle restatement of allwave's per-pair glue: align_pair, CIGAR passes,
 * PAF record, parse_scores.  TEST INFRASTRUCTURE ONLY (see aw_oracle.h).
 *
 * Follows /root/reference/src/alignment.rs:25-66 (align_pair), :157-175
 * (determine_orientation_wfa), :292-376 (count_cigar_operations, edit_distance_from_cigar,
 * parse_cigar_lengths, cigar_bytes_to_string); /root/reference/src/lib.rs:71-153
 * (alignment_to_paf, parse_scores).
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "aw_oracle.h"

char* awo_cigar_bytes_to_string(const uint8_t* ops, size_t n) {
    size_t cap = 64, len = 0;
    char* s = (char*)malloc(cap);
    size_t i = 0;
    while (i < n) {
        uint8_t op = ops[i];
        size_t j = i + 1;
        while (j < n && ops[j] == op) ++j;
        char c = op == 'M' ? '=' : op == 'X' ? 'X' : op == 'I' ? 'D' : op == 'D' ? 'I' : '?';
        if (len + 32 > cap) {
            cap *= 2;
            s = (char*)realloc(s, cap);
        }
        len += (size_t)sprintf(s + len, "%zu%c", j - i, c);
        i = j;
    }
    s[len] = 0;
    return s;
}

static void count_ops(const uint8_t* ops, size_t n, uint64_t* m, uint64_t* x, uint64_t* i_, uint64_t* d) {
    *m = *x = *i_ = *d = 0;
    for (size_t k = 0; k < n; ++k) {
        switch (ops[k]) {
            case 'M': ++*m; break;
            case 'X': ++*x; break;
            case 'I': ++*i_; break;
            case 'D': ++*d; break;
            default: break;
        }
    }
}

/* perform_wfa_alignment_with_cache (src/alignment.rs:201-261) minus the aligner cache */
static int perform_wfa(const uint8_t* q, size_t qlen, const uint8_t* t, size_t tlen, const awo_params_t* params,
                       awo_result_t* out) {
    awo_alignment_t al;
    if (qlen > (size_t)INT_MAX || tlen > (size_t)INT_MAX || awo_wfa_align(params, q, (int)qlen, t, (int)tlen, &al) != 0)
        return -1;
    uint64_t m, x, i_, d;
    count_ops(al.cigar, al.cigar_len, &m, &x, &i_, &d);
    out->query_start = 0;
    out->target_start = 0;
    out->query_end = m + x + d;  /* WFA2 'D' consumes the query  */
    out->target_end = m + x + i_; /* WFA2 'I' consumes the target */
    out->num_matches = m;
    out->alignment_length = m + x;
    out->score = al.score;
    out->cigar_bytes = al.cigar;
    out->cigar_len = al.cigar_len;
    out->work = al.work;
    return 0;
}

static uint64_t wfa_edit_ops(const uint8_t* q, size_t qlen, const uint8_t* t, size_t tlen, const awo_params_t* params) {
    awo_result_t r;
    memset(&r, 0, sizeof(r));
    if (perform_wfa(q, qlen, t, tlen, params, &r) != 0) return UINT64_MAX;
    uint64_t n = 0;
    for (size_t k = 0; k < r.cigar_len; ++k) n += (r.cigar_bytes[k] == 'X' || r.cigar_bytes[k] == 'I' || r.cigar_bytes[k] == 'D');
    free(r.cigar_bytes);
    return n;
}

void awo_align_pair(const uint8_t* query, size_t qlen, const uint8_t* target, size_t tlen, uint64_t query_idx,
                    uint64_t target_idx, const awo_params_t* params, const awo_params_t* orientation_params,
                    int use_mash, awo_result_t* out) {
    memset(out, 0, sizeof(*out));
    uint8_t* rc = (uint8_t*)malloc(qlen ? qlen : 1);
    int is_reverse;
    if (use_mash) {
        is_reverse = awo_orientation_mash(query, qlen, target, tlen);
        if (is_reverse) awo_reverse_complement(query, qlen, rc);
    } else {
        awo_reverse_complement(query, qlen, rc);
        uint64_t fd = wfa_edit_ops(query, qlen, target, tlen, orientation_params);
        uint64_t rd = wfa_edit_ops(rc, qlen, target, tlen, orientation_params);
        is_reverse = fd <= rd ? 0 : 1;
    }
    const uint8_t* q = is_reverse ? rc : query;
    if (perform_wfa(q, qlen, target, tlen, params, out) != 0) {
        /* failure sentinel, src/alignment.rs:49-64 */
        memset(out, 0, sizeof(*out));
        out->score = INT32_MAX;
    }
    out->query_idx = query_idx;
    out->target_idx = target_idx;
    out->is_reverse = (uint8_t)is_reverse;
    free(rc);
}

void awo_result_free(awo_result_t* r) {
    if (r) {
        free(r->cigar_bytes);
        r->cigar_bytes = NULL;
        r->cigar_len = 0;
    }
}

char* awo_alignment_to_paf(const awo_result_t* r, const char* qname, size_t qlen, const char* tname, size_t tlen) {
    uint64_t qal = r->query_end - r->query_start, tal = r->target_end - r->target_start;
    uint64_t block_len = tal > qal ? tal : qal;
    double identity = r->alignment_length > 0 ? (double)r->num_matches / (double)r->alignment_length : 0.0;
    char* cg = awo_cigar_bytes_to_string(r->cigar_bytes, r->cigar_len);
    size_t cap = strlen(qname) + strlen(tname) + strlen(cg) + 256;
    char* line = (char*)malloc(cap);
    snprintf(line, cap, "%s\t%zu\t%llu\t%llu\t%c\t%s\t%zu\t%llu\t%llu\t%llu\t%llu\t%d\tgi:f:%.6f\tcg:Z:%s", qname, qlen,
             (unsigned long long)r->query_start, (unsigned long long)r->query_end, r->is_reverse ? '-' : '+', tname, tlen,
             (unsigned long long)r->target_start, (unsigned long long)r->target_end, (unsigned long long)r->num_matches,
             (unsigned long long)block_len, 60, identity, cg);
    free(cg);
    return line;
}

/* Rust: s.split(',').map(|s| s.trim().parse::<i32>()) ; 4 or 6 values */
static int parse_i32_rust(const char* b, const char* e, int32_t* out) {
    while (b < e && (*b == ' ' || *b == '\t' || *b == '\n' || *b == '\r' || *b == '\f' || *b == '\v')) ++b;
    while (e > b && (e[-1] == ' ' || e[-1] == '\t' || e[-1] == '\n' || e[-1] == '\r' || e[-1] == '\f' || e[-1] == '\v')) --e;
    if (b == e) return -1;
    int neg = 0;
    if (*b == '+' || *b == '-') {
        neg = (*b == '-');
        ++b;
        if (b == e) return -1;
    }
    int64_t v = 0;
    for (; b < e; ++b) {
        if (*b < '0' || *b > '9') return -1;
        v = v * 10 + (*b - '0');
        if (v > (int64_t)INT32_MAX + 1) return -1;
    }
    if (neg) v = -v;
    if (v > INT32_MAX || v < INT32_MIN) return -1;
    *out = (int32_t)v;
    return 0;
}
int awo_parse_scores(const char* s, awo_params_t* out) {
    int32_t vals[8];
    int n = 0;
    const char* b = s;
    for (;;) {
        const char* e = strchr(b, ',');
        const char* end = e ? e : b + strlen(b);
        int32_t v;
        if (parse_i32_rust(b, end, &v)) return -1;
        if (n < 8) vals[n] = v;
        ++n;
        if (!e) break;
        b = e + 1;
    }
    if (n != 4 && n != 6) return -2;
    memset(out, 0, sizeof(*out));
    out->match_score = vals[0];
    out->mismatch_penalty = vals[1];
    out->gap_open = vals[2];
    out->gap_extend = vals[3];
    if (n == 6) {
        out->gap2_open = vals[4];
        out->gap2_extend = vals[5];
        out->has_gap2_open = out->has_gap2_extend = 1;
    }
    return 0;
}
