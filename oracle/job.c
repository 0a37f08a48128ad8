/*
 * job.c -- whole-job driver of the oracle: data-parallel over pairs with T threads, mirroring the
 * rayon loop of /root/reference/src/iterator.rs:208-252 (align_pair per pair, then
 * alignment_to_paf as src/main.rs:370-371).  Used as the CPU baseline leg of bench.py and by
 * tests.  TEST INFRASTRUCTURE ONLY (see aw_oracle.h).
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "aw_oracle.h"

typedef struct {
    uint64_t n;
    const char* const* ids;
    const uint8_t* const* seqs;
    const size_t* lens;
    const uint64_t* pairs;
    uint64_t npairs;
    const awo_params_t* params;
    int use_mash;
    awo_job_t* out;
    uint64_t next; /* atomic work counter */
    pthread_mutex_t mu;
} job_ctx_t;

static void work_add(awo_work_t* a, const awo_work_t* b) {
    a->cells += b->cells;
    a->steps += b->steps;
    a->ext_chars += b->ext_chars;
    a->n_breakpoints += b->n_breakpoints;
    a->n_base += b->n_base;
    a->n_end_reached_fallbacks += b->n_end_reached_fallbacks;
    if (b->max_base_score > a->max_base_score) a->max_base_score = b->max_base_score;
    if (b->max_level > a->max_level) a->max_level = b->max_level;
    if (b->max_width > a->max_width) a->max_width = b->max_width;
}

static void* worker(void* arg) {
    job_ctx_t* c = (job_ctx_t*)arg;
    awo_params_t orient;
    memset(&orient, 0, sizeof(orient)); /* AlignmentParams::edit_distance(), src/iterator.rs:85 */
    orient.mismatch_penalty = orient.gap_open = orient.gap_extend = 1;
    awo_work_t local;
    memset(&local, 0, sizeof(local));
    uint64_t block = 0;
    for (;;) {
        uint64_t i = __atomic_fetch_add(&c->next, 1, __ATOMIC_RELAXED);
        if (i >= c->npairs) break;
        uint64_t qi = c->pairs[2 * i], ti = c->pairs[2 * i + 1];
        awo_result_t r;
        awo_align_pair(c->seqs[qi], c->lens[qi], c->seqs[ti], c->lens[ti], qi, ti, c->params, &orient, c->use_mash, &r);
        c->out->paf[i] = awo_alignment_to_paf(&r, c->ids[qi], c->lens[qi], c->ids[ti], c->lens[ti]);
        c->out->scores[i] = r.score;
        uint64_t qa = r.query_end - r.query_start, ta = r.target_end - r.target_start;
        block += ta > qa ? ta : qa;
        work_add(&local, &r.work);
        awo_result_free(&r);
    }
    awo_pool_release();
    pthread_mutex_lock(&c->mu);
    work_add(&c->out->work, &local);
    c->out->sum_block_len += block;
    pthread_mutex_unlock(&c->mu);
    return NULL;
}

int awo_run_pairs(uint64_t n, const char* const* ids, const uint8_t* const* seqs, const size_t* lens,
                  const uint64_t* pairs, uint64_t npairs, const awo_params_t* params, int use_mash, int threads,
                  awo_job_t* out) {
    memset(out, 0, sizeof(*out));
    out->npairs = npairs;
    out->paf = (char**)calloc((size_t)(npairs ? npairs : 1), sizeof(char*));
    out->scores = (int32_t*)calloc((size_t)(npairs ? npairs : 1), sizeof(int32_t));
    job_ctx_t c;
    memset(&c, 0, sizeof(c));
    c.n = n; c.ids = ids; c.seqs = seqs; c.lens = lens; c.pairs = pairs; c.npairs = npairs;
    c.params = params; c.use_mash = use_mash; c.out = out;
    pthread_mutex_init(&c.mu, NULL);
    if (threads < 1) threads = 1;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, worker, &c);
    for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    pthread_mutex_destroy(&c.mu);
    out->seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    return 0;
}

void awo_job_free(awo_job_t* j) {
    if (!j) return;
    if (j->paf) for (uint64_t i = 0; i < j->npairs; ++i) free(j->paf[i]);
    free(j->paf);
    free(j->scores);
    j->paf = NULL;
    j->scores = NULL;
}
