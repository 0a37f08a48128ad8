/*
 * pairs.c -- oracle restatement of allwave's pair scheduling.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src/iterator.rs:30-92 (AllPairIterator::with_options), :256-284
 * (apply_random_sparsification), :300-334 (compute_connectivity_probability);
 * /root/reference/src/knn_graph.rs:12-52 (extract_tree_pairs), :112-174 (build_knn_graph,
 * generate_random_pairs, should_include_pair); /root/reference/src/mash.rs:141-165
 * (compute_distance_matrix_with_params).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/aw_wfa2_compat.h"
#include "aw_oracle.h"

double awo_connectivity_probability(uint64_t n, double connectivity_prob) {
    if (n <= 1) return 1.0;
    double x = connectivity_prob;
    if (x < 0.001) x = 0.001;
    if (x > 0.999) x = 0.999;
    if (n <= 10) {
        switch (n) {
            case 2: return 1.0;
            case 3: return 0.8;
            case 4: return 0.7;
            case 5: return 0.6;
            default: return 0.5;
        }
    }
    double n_f = (double)n;
    double log_n = log(n_f);
    double c = -log(-log(x));
    double p = (log_n + c) / n_f;
    if (p < 0.001) p = 0.001;
    if (p > 1.0) p = 1.0;
    return p;
}

/* hash("idA:idB") as f64 / u64::MAX as f64 < fraction */
static int keep_pair(const char* a, const char* b, double fraction) {
    size_t la = strlen(a), lb = strlen(b);
    char stack[256];
    char* buf = la + lb + 1 <= sizeof(stack) ? stack : (char*)malloc(la + lb + 1);
    memcpy(buf, a, la);
    buf[la] = ':';
    memcpy(buf + la + 1, b, lb);
    uint64_t h = awo_hash_str((const uint8_t*)buf, la + lb + 1);
    if (buf != stack) free(buf);
    double normalized = (double)h / 18446744073709551616.0; /* u64::MAX as f64 == 2^64 */
    return normalized < fraction;
}

double* awo_distance_matrix(uint64_t n, const uint8_t* const* seqs, const size_t* lens, int k, size_t sketch_size) {
    double* m = (double*)calloc((size_t)(n > 0 ? n * n : 1), sizeof(double));
    uint64_t* sk = (uint64_t*)malloc(sizeof(uint64_t) * sketch_size * (size_t)(n ? n : 1));
    size_t* cnt = (size_t*)malloc(sizeof(size_t) * (size_t)(n ? n : 1));
    for (uint64_t i = 0; i < n; ++i) cnt[i] = awo_sketch_canonical(seqs[i], lens[i], k, sketch_size, sk + i * sketch_size);
    for (uint64_t i = 0; i < n; ++i)
        for (uint64_t j = i + 1; j < n; ++j) {
            double d = awo_mash_distance(sk + i * sketch_size, cnt[i], sk + j * sketch_size, cnt[j], k);
            m[i * n + j] = d;
            m[j * n + i] = d;
        }
    free(sk);
    free(cnt);
    return m;
}

typedef struct {
    double d;
    uint64_t j;
} nb_t;
/* Rust's sort_by is stable and the list starts in ascending j, so ties break by ascending j */
static int cmp_near(const void* a, const void* b) {
    const nb_t *x = (const nb_t*)a, *y = (const nb_t*)b;
    if (x->d < y->d) return -1;
    if (x->d > y->d) return 1;
    return x->j < y->j ? -1 : (x->j > y->j ? 1 : 0);
}
static int cmp_far(const void* a, const void* b) {
    const nb_t *x = (const nb_t*)a, *y = (const nb_t*)b;
    if (x->d > y->d) return -1;
    if (x->d < y->d) return 1;
    return x->j < y->j ? -1 : (x->j > y->j ? 1 : 0);
}

uint64_t* awo_build_knn_graph(const double* matrix, uint64_t n, uint64_t k_neighbors, int farthest, uint64_t* npairs) {
    uint64_t cap = n * (k_neighbors < n ? k_neighbors : n) + 1, cnt = 0;
    uint64_t* out = (uint64_t*)malloc(sizeof(uint64_t) * 2 * cap);
    nb_t* nb = (nb_t*)malloc(sizeof(nb_t) * (size_t)(n ? n : 1));
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t m = 0;
        for (uint64_t j = 0; j < n; ++j)
            if (i != j) {
                nb[m].d = matrix[i * n + j];
                nb[m].j = j;
                ++m;
            }
        qsort(nb, m, sizeof(nb_t), farthest ? cmp_far : cmp_near);
        uint64_t ka = k_neighbors < m ? k_neighbors : m;
        for (uint64_t q = 0; q < ka; ++q) {
            out[2 * cnt] = i;
            out[2 * cnt + 1] = nb[q].j;
            ++cnt;
        }
    }
    free(nb);
    *npairs = cnt;
    return out;
}

static int cmp_pair(const void* a, const void* b) {
    const uint64_t *x = (const uint64_t*)a, *y = (const uint64_t*)b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    if (x[1] != y[1]) return x[1] < y[1] ? -1 : 1;
    return 0;
}

uint64_t* awo_pair_list(uint64_t n, const char* const* ids, const uint8_t* const* seqs, const size_t* lens,
                        int exclude_self, const awo_sparsification_t* sp, uint64_t* npairs) {
    uint64_t cnt = 0;
    if (sp->kind == AWO_SPARS_TREE) {
        /* extract_tree_pairs replaces the all-pairs list entirely */
        if (n < 2) {
            *npairs = 0;
            return (uint64_t*)malloc(16);
        }
        int k = sp->kmer_size ? sp->kmer_size : AW_ORIENT_K;
        double* dm = awo_distance_matrix(n, seqs, lens, k, AW_SKETCH_SIZE);
        uint64_t nn = 0, nf = 0;
        uint64_t* near = sp->k_nearest > 0 ? awo_build_knn_graph(dm, n, sp->k_nearest, 0, &nn) : NULL;
        uint64_t* far = sp->k_farthest > 0 ? awo_build_knn_graph(dm, n, sp->k_farthest, 1, &nf) : NULL;
        free(dm);
        uint64_t cap = nn + nf + 1024;
        uint64_t* out = (uint64_t*)malloc(sizeof(uint64_t) * 2 * cap);
        if (near) memcpy(out, near, sizeof(uint64_t) * 2 * nn);
        if (far) memcpy(out + 2 * nn, far, sizeof(uint64_t) * 2 * nf);
        cnt = nn + nf;
        free(near);
        free(far);
        if (sp->fraction > 0.0) {
            for (uint64_t i = 0; i < n; ++i)
                for (uint64_t j = 0; j < n; ++j)
                    if (i != j && keep_pair(ids[i], ids[j], sp->fraction)) {
                        if (cnt == cap) {
                            cap *= 2;
                            out = (uint64_t*)realloc(out, sizeof(uint64_t) * 2 * cap);
                        }
                        out[2 * cnt] = i;
                        out[2 * cnt + 1] = j;
                        ++cnt;
                    }
        }
        qsort(out, cnt, 2 * sizeof(uint64_t), cmp_pair);
        uint64_t w = 0;
        for (uint64_t r = 0; r < cnt; ++r)
            if (w == 0 || out[2 * r] != out[2 * w - 2] || out[2 * r + 1] != out[2 * w - 1]) {
                out[2 * w] = out[2 * r];
                out[2 * w + 1] = out[2 * r + 1];
                ++w;
            }
        *npairs = w;
        return out;
    }
    double keep = 1.0;
    int filter = 0;
    switch (sp->kind) {
        case AWO_SPARS_NONE: break;
        case AWO_SPARS_RANDOM: keep = sp->fraction; filter = 1; break;
        case AWO_SPARS_AUTO: keep = awo_connectivity_probability(n, 0.95); filter = 1; break;
        case AWO_SPARS_GIANT: keep = awo_connectivity_probability(n, sp->fraction); filter = 1; break;
        default: break;
    }
    uint64_t cap = 1024;
    uint64_t* out = (uint64_t*)malloc(sizeof(uint64_t) * 2 * cap);
    for (uint64_t i = 0; i < n; ++i)
        for (uint64_t j = 0; j < n; ++j) {
            if (exclude_self && i == j) continue;
            if (filter && !keep_pair(ids[i], ids[j], keep)) continue;
            if (cnt == cap) {
                cap *= 2;
                out = (uint64_t*)realloc(out, sizeof(uint64_t) * 2 * cap);
            }
            out[2 * cnt] = i;
            out[2 * cnt + 1] = j;
            ++cnt;
        }
    *npairs = cnt;
    return out;
}
