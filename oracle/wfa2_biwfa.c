/*
 * wfa2_biwfa.c -- oracle restatement of WFA2-lib's gap-affine / gap-affine-2p wavefront aligner
 * in MemoryMode::Ultralow (biWFA), End2End span, no heuristic: exactly the configuration
 * allwave builds in /root/reference/src/alignment.rs:226-228,263-289 and calls at :231.
 *
 * TEST INFRASTRUCTURE ONLY (see aw_oracle.h).  PARITY UNPINNED: WFA2-lib is not on disk; the
 * upstream function each block restates is named in its comment.
 *
 * Conventions: pattern = query (v), text = target (h), diagonal k = h - v, wavefront
 * offset = h.  Ops: M match, X mismatch, I consumes text, D consumes pattern.
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/aw_wfa2_compat.h"
#include "aw_oracle.h"

#define OFFSET_NULL AW_OFFSET_NULL
#define MAXI(a, b) ((a) > (b) ? (a) : (b))
#define MINI(a, b) ((a) < (b) ? (a) : (b))

/* status values of wavefront_align_status_t */
enum { ST_OK = 0, ST_END_REACHED = 1, ST_END_UNREACHABLE = 2, ST_COMPLETED = 3, ST_ERROR = -1 };

typedef struct {
    int lo, hi; /* effective range after trimming */
    int null;   /* lo > hi */
    int alo;    /* off[k - alo] */
    int32_t* off;
    size_t cap; /* allocated elements */
} wf_t;

typedef struct {
    int x, o1, e1, o2, e2;
    int two_piece;
} pen_t;

typedef struct {
    pen_t pen;
    int modular; /* wavefronts kept modulo max_score_scope (score-only fwd/rev aligners) */
    int scope;   /* max_score_scope = max(x, o1+e1, o2+e2) + 1 */
    /* sequence views (wavefront_sequences_set_bounds): reverse aligners read back to front */
    const uint8_t* p;
    int plen;
    const uint8_t* t;
    int tlen;
    int reverse;
    int comp_begin, comp_end;
    /* fast mode: forward-readable copies for the reverse aligner (p[v] == rp[v]) */
    uint8_t *rp, *rt;
    size_t rp_cap, rt_cap;
    /* components, indexed by score (mod scope when modular) */
    wf_t** comp[5];
    int nslots;
    int num_null_steps;
    int status, end_score, end_k, end_off;
    awo_work_t* work;
} aligner_t;

static wf_t g_wf_null = {1, -1, 1, 0, NULL, 0}; /* wavefront_null: lo=1, hi=-1 */
static FILE* g_trace = NULL;

void awo_set_trace_file(void* f) { g_trace = (FILE*)f; }

static inline int pchar(const aligner_t* a, int v) { return a->reverse ? a->p[a->plen - 1 - v] : a->p[v]; }
static inline int tchar(const aligner_t* a, int h) { return a->reverse ? a->t[a->tlen - 1 - h] : a->t[h]; }

/*
 * FAST MODE (awo_set_fast(1)) -- used by the CPU *baseline* legs of bench.py only; the checker keeps the simple path, and
 * tests/test_oracle.py asserts that both give identical results.  It does what a WFA2-lib build does for speed and nothing
 * else: (a) wavefronts come from a per-thread free list instead of malloc/free per score (WFA2's slab), (b) every wavefront is
 * allocated with WF_PAD NULL cells either side so that the compute loop reads its inputs without range checks wherever all
 * reads fall inside the padded allocations (wavefront_compute_init_ends), (c) the match extension compares 8 bytes at a time
 * (XOR + ctz, wavefront_extend_matches_packed), the reverse aligner on reversed copies of the sequences.
 */
#define WF_PAD 64
static int g_fast = 0;
void awo_set_fast(int on) { g_fast = on ? 1 : 0; }
int awo_get_fast(void) { return g_fast; }

#define WF_POOL_MAX 512
static __thread wf_t* t_pool[WF_POOL_MAX];
static __thread int t_pool_n = 0;
static __thread size_t t_pool_cap[WF_POOL_MAX];

void awo_pool_release(void) { /* worker threads call this before they exit */
    for (int i = 0; i < t_pool_n; ++i) {
        free(t_pool[i]->off);
        free(t_pool[i]);
    }
    t_pool_n = 0;
}

static wf_t* wf_new(int lo, int hi) {
    int n = hi >= lo ? hi - lo + 1 : 0;
    const int pad = g_fast ? WF_PAD : 0;
    const size_t need = (size_t)(n > 0 ? n : 1) + 2 * (size_t)pad;
    wf_t* w = NULL;
    if (g_fast) {
        for (int i = t_pool_n - 1; i >= 0 && i >= t_pool_n - 8; --i) /* recently freed wavefronts have the right size */
            if (t_pool_cap[i] >= need) {
                w = t_pool[i];
                w->cap = t_pool_cap[i];
                t_pool[i] = t_pool[t_pool_n - 1];
                t_pool_cap[i] = t_pool_cap[t_pool_n - 1];
                --t_pool_n;
                break;
            }
    }
    if (w == NULL) {
        w = (wf_t*)malloc(sizeof(wf_t));
        w->cap = g_fast ? need + need / 4 + 64 : need;
        w->off = (int32_t*)malloc(sizeof(int32_t) * w->cap);
    }
    w->lo = lo;
    w->hi = hi;
    w->null = lo > hi;
    w->alo = lo - pad;
    if (pad) {
        for (int i = 0; i < pad; ++i) w->off[i] = OFFSET_NULL;
        for (size_t i = (size_t)pad + (size_t)(n > 0 ? n : 1); i < need; ++i) w->off[i] = OFFSET_NULL;
    }
    return w;
}
static void wf_free(wf_t* w) {
    if (w && w != &g_wf_null) {
        if (g_fast && t_pool_n < WF_POOL_MAX) {
            t_pool[t_pool_n] = w;
            t_pool_cap[t_pool_n] = w->cap;
            ++t_pool_n;
            return;
        }
        free(w->off);
        free(w);
    }
}
/* reads outside the (trimmed) range see NULL: restates wavefront_compute_init_ends */
static inline int32_t wf_at(const wf_t* w, int k) {
    return (k >= w->lo && k <= w->hi) ? w->off[k - w->alo] : OFFSET_NULL;
}

static void aligner_clear(aligner_t* a) {
    for (int c = 0; c < 5; ++c) {
        if (!a->comp[c]) continue;
        for (int s = 0; s < a->nslots; ++s) {
            wf_free(a->comp[c][s]);
            a->comp[c][s] = NULL;
        }
    }
}
static void aligner_reserve(aligner_t* a, int nslots) {
    if (nslots <= a->nslots) return;
    int cap = a->nslots ? a->nslots : 64;
    while (cap < nslots) cap *= 2;
    for (int c = 0; c < 5; ++c) {
        a->comp[c] = (wf_t**)realloc(a->comp[c], sizeof(wf_t*) * (size_t)cap);
        for (int s = a->nslots; s < cap; ++s) a->comp[c][s] = NULL;
    }
    a->nslots = cap;
}
static void aligner_init(aligner_t* a, const pen_t* pen, int modular, awo_work_t* work) {
    memset(a, 0, sizeof(*a));
    a->pen = *pen;
    a->modular = modular;
    int sc = pen->x;
    sc = MAXI(sc, pen->o1 + pen->e1);
    if (pen->two_piece) sc = MAXI(sc, pen->o2 + pen->e2);
    a->scope = sc + 1;
    a->work = work;
    aligner_reserve(a, modular ? a->scope : 64);
}
static void aligner_destroy(aligner_t* a) {
    aligner_clear(a);
    free(a->rp);
    free(a->rt);
    for (int c = 0; c < 5; ++c) free(a->comp[c]);
}
static inline int slot_of(const aligner_t* a, int score) { return a->modular ? score % a->scope : score; }
/* wavefront_compute_get_*wavefront: negative score or NULL pointer -> wavefront_null */
static inline wf_t* wf_get(const aligner_t* a, int c, int score) {
    if (score < 0) return &g_wf_null;
    int s = slot_of(a, score);
    if (s >= a->nslots || a->comp[c][s] == NULL) return &g_wf_null;
    return a->comp[c][s];
}
static inline wf_t* wf_ptr(const aligner_t* a, int c, int score) { /* raw pointer, may be NULL */
    if (score < 0) return NULL;
    int s = slot_of(a, score);
    if (s >= a->nslots) return NULL;
    return a->comp[c][s];
}
static void wf_set(aligner_t* a, int c, int score, wf_t* w) {
    int s = slot_of(a, score);
    aligner_reserve(a, s + 1);
    wf_free(a->comp[c][s]);
    a->comp[c][s] = w;
}

/* wavefront_unialign_init + wavefront_unialign_init_end2end: score-0 wavefront in component_begin */
static void aligner_start(aligner_t* a, const uint8_t* p, int plen, const uint8_t* t, int tlen, int reverse,
                          int comp_begin, int comp_end) {
    aligner_clear(a);
    a->p = p;
    a->plen = plen;
    a->t = t;
    a->tlen = tlen;
    a->reverse = reverse;
    a->comp_begin = comp_begin;
    a->comp_end = comp_end;
    a->num_null_steps = 0;
    a->status = ST_OK;
    a->end_score = -1;
    if (g_fast && reverse) { /* wavefront_sequences_init keeps reversed copies too */
        if ((size_t)plen + 8 > a->rp_cap) {
            a->rp_cap = (size_t)plen + 8 + (size_t)plen / 4;
            a->rp = (uint8_t*)realloc(a->rp, a->rp_cap);
        }
        if ((size_t)tlen + 8 > a->rt_cap) {
            a->rt_cap = (size_t)tlen + 8 + (size_t)tlen / 4;
            a->rt = (uint8_t*)realloc(a->rt, a->rt_cap);
        }
        for (int i = 0; i < plen; ++i) a->rp[i] = p[plen - 1 - i];
        for (int i = 0; i < tlen; ++i) a->rt[i] = t[tlen - 1 - i];
    }
    wf_t* w = wf_new(0, 0);
    w->off[0 - w->alo] = 0;
    wf_set(a, comp_begin, 0, w);
}

/* wavefront_compute_trim_ends */
static void wf_trim(const aligner_t* a, wf_t* w) {
    int k;
    const int lo = w->lo, old_hi = w->hi;
    for (k = w->hi; k >= lo; --k) {
        int32_t off = w->off[k - w->alo];
        uint32_t h = (uint32_t)off, v = (uint32_t)(off - k);
        if (h <= (uint32_t)a->tlen && v <= (uint32_t)a->plen) break;
    }
    w->hi = k;
    const int hi = w->hi;
    for (k = w->lo; k <= hi; ++k) {
        int32_t off = w->off[k - w->alo];
        uint32_t h = (uint32_t)off, v = (uint32_t)(off - k);
        if (h <= (uint32_t)a->tlen && v <= (uint32_t)a->plen) break;
    }
    w->lo = k;
    w->null = (w->lo > w->hi);
    if (g_fast) { /* cells trimmed away must read as NULL without a range check */
        for (k = lo; k < w->lo && k <= old_hi; ++k) w->off[k - w->alo] = OFFSET_NULL;
        for (k = (w->hi >= lo ? w->hi + 1 : lo); k <= old_hi; ++k) w->off[k - w->alo] = OFFSET_NULL;
    }
}


/* fast mode, interior of a wavefront (every read inside the NULL-padded allocations): the elementwise recurrences of
 * wavefront_compute_affine(2p)_idm over disjoint rows -- restrict-qualified so that the compiler vectorises them, cloned for AVX2
 * (resolved at load time; the baseline x86-64 clone has no packed 32-bit max) as a WFA2-lib build compiled for the host would be */
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define AW_FAST_CLONES __attribute__((target_clones("avx2", "default")))
#else
#define AW_FAST_CLONES
#endif
AW_FAST_CLONES
static void fast_interior_2p(int f_lo, int f_hi, const int32_t* restrict pmx, const int32_t* restrict pmo1, const int32_t* restrict pi1,
                             const int32_t* restrict pd1, const int32_t* restrict pmo2, const int32_t* restrict pi2, const int32_t* restrict pd2,
                             int32_t* restrict om, int32_t* restrict oi1, int32_t* restrict od1, int32_t* restrict oi2, int32_t* restrict od2, uint32_t tlen,
                             uint32_t plen) {
    for (int q = f_lo; q <= f_hi; ++q) {
        const int32_t ins1 = MAXI(pmo1[q - 1], pi1[q - 1]) + 1, del1 = MAXI(pmo1[q + 1], pd1[q + 1]);
        const int32_t ins2 = MAXI(pmo2[q - 1], pi2[q - 1]) + 1, del2 = MAXI(pmo2[q + 1], pd2[q + 1]);
        oi1[q] = ins1;
        od1[q] = del1;
        oi2[q] = ins2;
        od2[q] = del2;
        int32_t mx = MAXI(MAXI(del1, del2), MAXI(pmx[q] + 1, MAXI(ins1, ins2)));
        if ((uint32_t)mx > tlen || (uint32_t)(mx - q) > plen) mx = OFFSET_NULL;
        om[q] = mx;
    }
}
AW_FAST_CLONES
static void fast_interior_1p(int f_lo, int f_hi, const int32_t* restrict pmx, const int32_t* restrict pmo1, const int32_t* restrict pi1,
                             const int32_t* restrict pd1, int32_t* restrict om, int32_t* restrict oi1, int32_t* restrict od1, uint32_t tlen, uint32_t plen) {
    for (int q = f_lo; q <= f_hi; ++q) {
        const int32_t ins1 = MAXI(pmo1[q - 1], pi1[q - 1]) + 1, del1 = MAXI(pmo1[q + 1], pd1[q + 1]);
        oi1[q] = ins1;
        od1[q] = del1;
        int32_t mx = MAXI(del1, MAXI(pmx[q] + 1, ins1));
        if ((uint32_t)mx > tlen || (uint32_t)(mx - q) > plen) mx = OFFSET_NULL;
        om[q] = mx;
    }
}

/* wavefront_compute_affine / wavefront_compute_affine2p (+ _idm kernels, limits_input,
 * allocate_output, process_ends) for one score */
static void wf_compute(aligner_t* a, int s) {
    const pen_t* pn = &a->pen;
    const int two = pn->two_piece;
    wf_t* m_x = wf_get(a, AW_COMP_M, s - pn->x);
    wf_t* m_o1 = wf_get(a, AW_COMP_M, s - pn->o1 - pn->e1);
    wf_t* i1_e = wf_get(a, AW_COMP_I1, s - pn->e1);
    wf_t* d1_e = wf_get(a, AW_COMP_D1, s - pn->e1);
    wf_t* m_o2 = two ? wf_get(a, AW_COMP_M, s - pn->o2 - pn->e2) : &g_wf_null;
    wf_t* i2_e = two ? wf_get(a, AW_COMP_I2, s - pn->e2) : &g_wf_null;
    wf_t* d2_e = two ? wf_get(a, AW_COMP_D2, s - pn->e2) : &g_wf_null;
    a->work->steps++;
    if (m_x->null && m_o1->null && i1_e->null && d1_e->null && m_o2->null && i2_e->null && d2_e->null) {
        a->num_null_steps++;
        for (int c = 0; c < 5; ++c) wf_set(a, c, s, NULL); /* allocate_output_null */
        return;
    }
    a->num_null_steps = 0;
    /* wavefront_compute_limits_input (null wavefronts contribute lo=1,hi=-1 like upstream) */
    int lo = m_x->lo, hi = m_x->hi;
    lo = MINI(lo, m_o1->lo - 1);
    hi = MAXI(hi, m_o1->hi + 1);
    lo = MINI(lo, i1_e->lo + 1);
    hi = MAXI(hi, i1_e->hi + 1);
    lo = MINI(lo, d1_e->lo - 1);
    hi = MAXI(hi, d1_e->hi - 1);
    if (two) {
        lo = MINI(lo, m_o2->lo - 1);
        hi = MAXI(hi, m_o2->hi + 1);
        lo = MINI(lo, i2_e->lo + 1);
        hi = MAXI(hi, i2_e->hi + 1);
        lo = MINI(lo, d2_e->lo - 1);
        hi = MAXI(hi, d2_e->hi - 1);
    }
    /* wavefront_compute_allocate_output: I/D outputs exist only if one of their inputs does */
    wf_t* out_m = wf_new(lo, hi);
    wf_t* out_i1 = (!m_o1->null || !i1_e->null) ? wf_new(lo, hi) : NULL;
    wf_t* out_d1 = (!m_o1->null || !d1_e->null) ? wf_new(lo, hi) : NULL;
    wf_t* out_i2 = (two && (!m_o2->null || !i2_e->null)) ? wf_new(lo, hi) : NULL;
    wf_t* out_d2 = (two && (!m_o2->null || !d2_e->null)) ? wf_new(lo, hi) : NULL;
    const uint32_t tlen = (uint32_t)a->tlen, plen = (uint32_t)a->plen;
    int32_t* const om = out_m->off - out_m->alo;
    int32_t* const oi1 = out_i1 ? out_i1->off - out_i1->alo : NULL;
    int32_t* const od1 = out_d1 ? out_d1->off - out_d1->alo : NULL;
    int32_t* const oi2 = out_i2 ? out_i2->off - out_i2->alo : NULL;
    int32_t* const od2 = out_d2 ? out_d2->off - out_d2->alo : NULL;
    /* fast mode: [f_lo, f_hi] = diagonals whose nine reads all fall inside the (NULL-padded) allocations of non-null inputs */
    int f_lo = 1, f_hi = 0;
    if (g_fast && !m_x->null && !m_o1->null && !i1_e->null && !d1_e->null && out_i1 && out_d1 &&
        (!two || (!m_o2->null && !i2_e->null && !d2_e->null && out_i2 && out_d2))) {
        f_lo = lo;
        f_hi = hi;
        const wf_t* ins[7] = {m_x, m_o1, i1_e, d1_e, m_o2, i2_e, d2_e};
        for (int i = 0; i < (two ? 7 : 4); ++i) {
            f_lo = MAXI(f_lo, ins[i]->lo - WF_PAD + 1);
            f_hi = MINI(f_hi, ins[i]->hi + WF_PAD - 1);
        }
    }
    for (int k = lo; k <= hi; ++k) {
        if (k == f_lo && f_lo <= f_hi) {
            const int32_t* pmx = m_x->off - m_x->alo;
            const int32_t* pmo1 = m_o1->off - m_o1->alo;
            const int32_t* pi1 = i1_e->off - i1_e->alo;
            const int32_t* pd1 = d1_e->off - d1_e->alo;
            if (two)
                fast_interior_2p(f_lo, f_hi, pmx, pmo1, pi1, pd1, m_o2->off - m_o2->alo, i2_e->off - i2_e->alo, d2_e->off - d2_e->alo, om, oi1, od1, oi2, od2, tlen, plen);
            else
                fast_interior_1p(f_lo, f_hi, pmx, pmo1, pi1, pd1, om, oi1, od1, tlen, plen);
            k = f_hi;
            continue;
        }
        int32_t ins1 = MAXI(wf_at(m_o1, k - 1), wf_at(i1_e, k - 1)) + 1;
        int32_t del1 = MAXI(wf_at(m_o1, k + 1), wf_at(d1_e, k + 1));
        int32_t ins = ins1, del = del1;
        if (out_i1) oi1[k] = ins1;
        if (out_d1) od1[k] = del1;
        if (two) {
            int32_t ins2 = MAXI(wf_at(m_o2, k - 1), wf_at(i2_e, k - 1)) + 1;
            int32_t del2 = MAXI(wf_at(m_o2, k + 1), wf_at(d2_e, k + 1));
            if (out_i2) oi2[k] = ins2;
            if (out_d2) od2[k] = del2;
            ins = MAXI(ins, ins2);
            del = MAXI(del, del2);
        }
        int32_t misms = wf_at(m_x, k) + 1;
        int32_t mx = MAXI(del, MAXI(misms, ins));
        uint32_t h = (uint32_t)mx, v = (uint32_t)(mx - k);
        if (h > tlen) mx = OFFSET_NULL;
        if (v > plen) mx = OFFSET_NULL;
        om[k] = mx;
    }
    a->work->cells += (uint64_t)(hi - lo + 1) * (two ? 5 : 3);
    if ((uint64_t)(hi - lo + 1) > a->work->max_width) a->work->max_width = (uint64_t)(hi - lo + 1);
    /* wavefront_compute_process_ends */
    wf_trim(a, out_m);
#if AW_TRIM_ALL_COMPONENTS
    if (out_i1) wf_trim(a, out_i1);
    if (out_d1) wf_trim(a, out_d1);
    if (out_i2) wf_trim(a, out_i2);
    if (out_d2) wf_trim(a, out_d2);
#endif
    wf_set(a, AW_COMP_M, s, out_m);
    wf_set(a, AW_COMP_I1, s, out_i1);
    wf_set(a, AW_COMP_D1, s, out_d1);
    wf_set(a, AW_COMP_I2, s, out_i2);
    wf_set(a, AW_COMP_D2, s, out_d2);
}

/* wavefront_termination_end2end */
static int wf_termination(aligner_t* a, const wf_t* mwf, int score) {
    const int ak = a->tlen - a->plen;
    const int32_t aoff = a->tlen;
    const wf_t* w = mwf;
    if (a->comp_end != AW_COMP_M) {
        w = wf_ptr(a, a->comp_end, score);
        if (w == NULL) return 0;
    }
    if (w->lo > ak || ak > w->hi) return 0;
    if (w->off[ak - w->alo] < aoff) return 0;
    a->end_score = score;
    a->end_k = ak;
    a->end_off = aoff;
    return 1;
}

/* wavefront_extend_end2end / wavefront_extend_end2end_max: returns 1 when finished */
static int wf_extend(aligner_t* a, int score, int* max_ak) {
    if (max_ak) *max_ak = 0;
    wf_t* mwf = wf_ptr(a, AW_COMP_M, score);
    if (mwf == NULL) {
        if (a->num_null_steps > a->scope) {
            a->status = ST_END_UNREACHABLE;
            a->end_score = score;
            return 1;
        }
        return 0;
    }
    int best = 0;
    for (int k = mwf->lo; k <= mwf->hi; ++k) {
        int32_t off = mwf->off[k - mwf->alo];
        if (off == OFFSET_NULL) continue;
        int v = off - k, h = off;
        if (g_fast) { /* 8 bytes per step on forward-readable sequences (wavefront_extend_matches_packed) */
            const uint8_t* ps = a->reverse ? a->rp : a->p;
            const uint8_t* ts = a->reverse ? a->rt : a->t;
            while (v + 8 <= a->plen && h + 8 <= a->tlen) {
                uint64_t x, y;
                memcpy(&x, ps + v, 8);
                memcpy(&y, ts + h, 8);
                const uint64_t d = x ^ y;
                if (d) {
                    const int n = __builtin_ctzll(d) >> 3;
                    v += n;
                    h += n;
                    goto extended;
                }
                v += 8;
                h += 8;
            }
            while (v < a->plen && h < a->tlen && ps[v] == ts[h]) {
                ++v;
                ++h;
            }
        extended:;
        } else
        while (v < a->plen && h < a->tlen && pchar(a, v) == tchar(a, h)) {
            ++v;
            ++h;
        }
        a->work->ext_chars += (uint64_t)(h - off) + 1;
        mwf->off[k - mwf->alo] = h;
        int ak = 2 * h - k; /* WAVEFRONT_ANTIDIAGONAL */
        if (best < ak) best = ak;
    }
    if (wf_termination(a, mwf, score)) {
        a->status = ST_END_REACHED;
        return 1;
    }
    if (max_ak) *max_ak = best;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * cigar buffer
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint8_t* ops;
    size_t n, cap;
} cigar_t;
static void cigar_push_n(cigar_t* c, uint8_t op, size_t n) {
    if (c->n + n > c->cap) {
        size_t cap = c->cap ? c->cap : 256;
        while (cap < c->n + n) cap *= 2;
        c->ops = (uint8_t*)realloc(c->ops, cap);
        c->cap = cap;
    }
    memset(c->ops + c->n, op, n);
    c->n += n;
}

/* ------------------------------------------------------------------------------------------
 * wavefront_backtrace_affine: candidates (offset<<4)|type, max wins
 * ---------------------------------------------------------------------------------------- */
static inline int64_t bt_cand(const aligner_t* a, int c, int score, int k, int add, int type) {
    if (score < 0) return OFFSET_NULL;
    const wf_t* w = wf_ptr(a, c, score);
    if (w != NULL && w->lo <= k && k <= w->hi)
        return (int64_t)(((uint64_t)(int64_t)(w->off[k - w->alo] + add)) << AW_BT_TYPE_BITS) | type;
    return OFFSET_NULL;
}
#define MAX64(a, b) ((a) > (b) ? (a) : (b))

/* writes ops back to front into rev (caller reverses) */
static void wf_backtrace(const aligner_t* a, int end_score, int end_k, int32_t end_off, cigar_t* rev) {
    const pen_t* pn = &a->pen;
    int type = a->comp_end;
    int score = end_score;
    int k = end_k;
    int32_t offset = end_off;
    int h = offset, v = offset - k;
    /* (end2end: v==plen, h==tlen here, so no trailing indels to account for) */
    while (v > 0 && h > 0 && score > 0) {
        const int s_x = score - pn->x;
        const int s_o1 = score - pn->o1 - pn->e1, s_e1 = score - pn->e1;
        const int s_o2 = score - pn->o2 - pn->e2, s_e2 = score - pn->e2;
        int64_t max_all = OFFSET_NULL;
        switch (type) {
            case AW_COMP_M: {
                int64_t misms = bt_cand(a, AW_COMP_M, s_x, k, 1, AW_BT_M);
                int64_t i1o = bt_cand(a, AW_COMP_M, s_o1, k - 1, 1, AW_BT_I1_OPEN);
                int64_t i1e = bt_cand(a, AW_COMP_I1, s_e1, k - 1, 1, AW_BT_I1_EXT);
                int64_t d1o = bt_cand(a, AW_COMP_M, s_o1, k + 1, 0, AW_BT_D1_OPEN);
                int64_t d1e = bt_cand(a, AW_COMP_D1, s_e1, k + 1, 0, AW_BT_D1_EXT);
                max_all = MAX64(misms, MAX64(MAX64(i1o, i1e), MAX64(d1o, d1e)));
                if (pn->two_piece) {
                    int64_t i2o = bt_cand(a, AW_COMP_M, s_o2, k - 1, 1, AW_BT_I2_OPEN);
                    int64_t i2e = bt_cand(a, AW_COMP_I2, s_e2, k - 1, 1, AW_BT_I2_EXT);
                    int64_t d2o = bt_cand(a, AW_COMP_M, s_o2, k + 1, 0, AW_BT_D2_OPEN);
                    int64_t d2e = bt_cand(a, AW_COMP_D2, s_e2, k + 1, 0, AW_BT_D2_EXT);
                    max_all = MAX64(max_all, MAX64(MAX64(i2o, i2e), MAX64(d2o, d2e)));
                }
                break;
            }
            case AW_COMP_I1:
                max_all = MAX64(bt_cand(a, AW_COMP_M, s_o1, k - 1, 1, AW_BT_I1_OPEN),
                                bt_cand(a, AW_COMP_I1, s_e1, k - 1, 1, AW_BT_I1_EXT));
                break;
            case AW_COMP_I2:
                max_all = MAX64(bt_cand(a, AW_COMP_M, s_o2, k - 1, 1, AW_BT_I2_OPEN),
                                bt_cand(a, AW_COMP_I2, s_e2, k - 1, 1, AW_BT_I2_EXT));
                break;
            case AW_COMP_D1:
                max_all = MAX64(bt_cand(a, AW_COMP_M, s_o1, k + 1, 0, AW_BT_D1_OPEN),
                                bt_cand(a, AW_COMP_D1, s_e1, k + 1, 0, AW_BT_D1_EXT));
                break;
            case AW_COMP_D2:
                max_all = MAX64(bt_cand(a, AW_COMP_M, s_o2, k + 1, 0, AW_BT_D2_OPEN),
                                bt_cand(a, AW_COMP_D2, s_e2, k + 1, 0, AW_BT_D2_EXT));
                break;
        }
        if (type == AW_COMP_M) {
            const int32_t max_offset = (int32_t)(max_all >> AW_BT_TYPE_BITS);
            const int num_matches = offset - max_offset;
            if (num_matches > 0) cigar_push_n(rev, 'M', (size_t)num_matches);
            offset = max_offset;
            v = offset - k;
            h = offset;
            if (v <= 0 || h <= 0) break;
        }
        const int bt = (int)(max_all & 0xF);
        switch (bt) {
            case AW_BT_M: score = s_x; type = AW_COMP_M; break;
            case AW_BT_I1_OPEN: score = s_o1; type = AW_COMP_M; break;
            case AW_BT_I1_EXT: score = s_e1; type = AW_COMP_I1; break;
            case AW_BT_I2_OPEN: score = s_o2; type = AW_COMP_M; break;
            case AW_BT_I2_EXT: score = s_e2; type = AW_COMP_I2; break;
            case AW_BT_D1_OPEN: score = s_o1; type = AW_COMP_M; break;
            case AW_BT_D1_EXT: score = s_e1; type = AW_COMP_D1; break;
            case AW_BT_D2_OPEN: score = s_o2; type = AW_COMP_M; break;
            case AW_BT_D2_EXT: score = s_e2; type = AW_COMP_D2; break;
            default:
                fprintf(stderr, "[aw_oracle] backtrace: invalid type %d (score=%d k=%d)\n", bt, score, k);
                abort();
        }
        switch (bt) {
            case AW_BT_M:
                cigar_push_n(rev, 'X', 1);
                --offset;
                break;
            case AW_BT_I1_OPEN: case AW_BT_I1_EXT: case AW_BT_I2_OPEN: case AW_BT_I2_EXT:
                cigar_push_n(rev, 'I', 1);
                --k;
                --offset;
                break;
            default:
                cigar_push_n(rev, 'D', 1);
                ++k;
                break;
        }
        v = offset - k;
        h = offset;
    }
    if (v > 0 && h > 0) { /* score == 0: leading run of matches */
        const int nm = MINI(v, h);
        cigar_push_n(rev, 'M', (size_t)nm);
        v -= nm;
        h -= nm;
    }
    if (v > 0) cigar_push_n(rev, 'D', (size_t)v);
    if (h > 0) cigar_push_n(rev, 'I', (size_t)h);
}

/* ------------------------------------------------------------------------------------------
 * wavefront_unialign on the subsidiary aligner (memory_high, full history) + backtrace
 * ---------------------------------------------------------------------------------------- */
static int unialign_and_backtrace(aligner_t* sub, const uint8_t* p, int plen, const uint8_t* t, int tlen,
                                  int comp_begin, int comp_end, cigar_t* cigar) {
    aligner_start(sub, p, plen, t, tlen, 0, comp_begin, comp_end);
    int score = 0;
    for (;;) {
        if (wf_extend(sub, score, NULL)) break;
        ++score;
        wf_compute(sub, score);
        if (score > 4 * (plen + tlen + 64) * MAXI(sub->pen.x, sub->pen.o1 + sub->pen.e1) + 1024) {
            sub->status = ST_ERROR;
            break;
        }
    }
    if (sub->status != ST_END_REACHED) return ST_ERROR;
    if ((uint32_t)sub->end_score > sub->work->max_base_score) sub->work->max_base_score = (uint32_t)sub->end_score;
    cigar_t rev = {NULL, 0, 0};
    wf_backtrace(sub, sub->end_score, sub->end_k, sub->end_off, &rev);
    size_t base = cigar->n;
    cigar_push_n(cigar, 0, rev.n);
    for (size_t i = 0; i < rev.n; ++i) cigar->ops[base + i] = rev.ops[rev.n - 1 - i];
    free(rev.ops);
    return ST_OK;
}

/* ------------------------------------------------------------------------------------------
 * biWFA (wavefront_bialign.c)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int score, score_forward, score_reverse;
    int k_forward, k_reverse;
    int32_t offset_forward, offset_reverse;
    int component;
} breakpoint_t;

typedef struct {
    aligner_t fwd, rev, sub;
    awo_work_t* work;
} bialigner_t;

/* wavefront_bialign_breakpoint_indel2indel / _m2m */
static void bp_test(const aligner_t* a0, int bp_forward, int score_0, int score_1, const wf_t* w0, const wf_t* w1,
                    int component, int credit, breakpoint_t* bp) {
    const int tlen = a0->tlen, plen = a0->plen;
    const int lo_0 = w0->lo, hi_0 = w0->hi;
    const int lo_1 = (tlen - plen) - w1->hi, hi_1 = (tlen - plen) - w1->lo;
    if (hi_1 < lo_0 || hi_0 < lo_1) return;
    const int min_hi = MINI(hi_0, hi_1), max_lo = MAXI(lo_0, lo_1);
    for (int k_0 = max_lo; k_0 <= min_hi; ++k_0) {
        const int k_1 = (tlen - plen) - k_0;
        const int32_t h_0 = w0->off[k_0 - w0->alo];
        const int32_t h_1 = w1->off[k_1 - w1->alo];
        if (h_0 + h_1 >= tlen && score_0 + score_1 - credit < bp->score) {
            if (component != AW_COMP_M) { /* only indel2indel checks bounds of the forward cell */
                const int kk = bp_forward ? k_0 : k_1;
                const int32_t hh = bp_forward ? h_0 : h_1;
                if (hh - kk > plen || hh > tlen) continue;
            }
            if (bp_forward) {
                bp->score_forward = score_0;
                bp->score_reverse = score_1;
                bp->k_forward = k_0;
                bp->k_reverse = k_1;
                bp->offset_forward = h_0;
                bp->offset_reverse = h_1;
            } else {
                bp->score_forward = score_1;
                bp->score_reverse = score_0;
                bp->k_forward = k_1;
                bp->k_reverse = k_0;
                bp->offset_forward = h_1;
                bp->offset_reverse = h_0;
            }
            bp->score = score_0 + score_1 - credit;
            bp->component = component;
            return;
        }
    }
}

/* wavefront_bialign_overlap */
static void bialign_overlap(const aligner_t* a0, const aligner_t* a1, int score_0, int score_1, int bp_forward,
                            breakpoint_t* bp) {
    const pen_t* pn = &a0->pen;
    const wf_t* m0 = wf_ptr(a0, AW_COMP_M, score_0);
    if (m0 == NULL) return;
    const wf_t* d1_0 = wf_ptr(a0, AW_COMP_D1, score_0);
    const wf_t* i1_0 = wf_ptr(a0, AW_COMP_I1, score_0);
    const wf_t* d2_0 = pn->two_piece ? wf_ptr(a0, AW_COMP_D2, score_0) : NULL;
    const wf_t* i2_0 = pn->two_piece ? wf_ptr(a0, AW_COMP_I2, score_0) : NULL;
    for (int i = 0; i < a0->scope; ++i) {
        const int score_i = score_1 - i;
        if (score_i < 0) break;
        if (pn->two_piece && score_0 + score_i - pn->o2 < bp->score) {
            const wf_t* d2_1 = wf_ptr(a1, AW_COMP_D2, score_i);
            if (d2_0 && d2_1) bp_test(a0, bp_forward, score_0, score_i, d2_0, d2_1, AW_COMP_D2, pn->o2, bp);
            const wf_t* i2_1 = wf_ptr(a1, AW_COMP_I2, score_i);
            if (i2_0 && i2_1) bp_test(a0, bp_forward, score_0, score_i, i2_0, i2_1, AW_COMP_I2, pn->o2, bp);
        }
        if (score_0 + score_i - pn->o1 < bp->score) {
            const wf_t* d1_1 = wf_ptr(a1, AW_COMP_D1, score_i);
            if (d1_0 && d1_1) bp_test(a0, bp_forward, score_0, score_i, d1_0, d1_1, AW_COMP_D1, pn->o1, bp);
            const wf_t* i1_1 = wf_ptr(a1, AW_COMP_I1, score_i);
            if (i1_0 && i1_1) bp_test(a0, bp_forward, score_0, score_i, i1_0, i1_1, AW_COMP_I1, pn->o1, bp);
        }
        if (score_0 + score_i >= bp->score) continue;
        const wf_t* m1 = wf_ptr(a1, AW_COMP_M, score_i);
        if (m1) bp_test(a0, bp_forward, score_0, score_i, m0, m1, AW_COMP_M, 0, bp);
    }
}

/* wavefront_bialign_find_breakpoint */
static int bialign_find_breakpoint(bialigner_t* b, const uint8_t* p, int plen, const uint8_t* t, int tlen,
                                   int comp_begin, int comp_end, breakpoint_t* bp) {
    aligner_t* f = &b->fwd;
    aligner_t* r = &b->rev;
    aligner_start(f, p, plen, t, tlen, 0, comp_begin, comp_end);
    aligner_start(r, p, plen, t, tlen, 1, comp_end, comp_begin);
    const int max_antidiagonal = plen + tlen - 1;
    int score_f = 0, score_r = 0, f_max_ak = 0, r_max_ak = 0, max_ak = 0;
    bp->score = INT_MAX;
    if (wf_extend(f, score_f, &f_max_ak)) return f->status;
    if (wf_extend(r, score_r, &r_max_ak)) return r->status;
    int last_forward = 0;
    const int64_t step_limit = 4LL * ((int64_t)plen + tlen + 64) * MAXI(f->pen.x, f->pen.o1 + f->pen.e1) + 1024;
    for (;;) {
        if (f_max_ak + r_max_ak >= max_antidiagonal) break;
        ++score_f;
        wf_compute(f, score_f);
        int done = wf_extend(f, score_f, &max_ak);
        if (f_max_ak < max_ak) f_max_ak = max_ak;
        last_forward = 1;
#if AW_BIALIGN_PHASE1_END_REACHED_RETURNS
        if (done) return f->status;
#endif
        if (f_max_ak + r_max_ak >= max_antidiagonal) break;
        ++score_r;
        wf_compute(r, score_r);
        done = wf_extend(r, score_r, &max_ak);
        if (r_max_ak < max_ak) r_max_ak = max_ak;
        last_forward = 0;
#if AW_BIALIGN_PHASE1_END_REACHED_RETURNS
        if (done) return r->status;
#endif
        (void)done;
        if ((int64_t)score_f + score_r > step_limit) return ST_ERROR;
    }
    const int scope = f->scope;
    const int gap_opening = AW_BIALIGN_GAP_OPENING(f->pen.two_piece, f->pen.o1, f->pen.o2);
    for (;;) {
        if (last_forward) {
            const int min_r = (score_r > scope - 1) ? score_r - (scope - 1) : 0;
            if (score_f + min_r - gap_opening >= bp->score) break;
            bialign_overlap(f, r, score_f, score_r, 1, bp);
            ++score_r;
            wf_compute(r, score_r);
            int done = wf_extend(r, score_r, NULL);
#if AW_BIALIGN_PHASE2_END_REACHED_RETURNS
            if (done) return r->status;
#endif
            (void)done;
        }
        const int min_f = (score_f > scope - 1) ? score_f - (scope - 1) : 0;
        if (min_f + score_r - gap_opening >= bp->score) break;
        bialign_overlap(r, f, score_r, score_f, 0, bp);
        ++score_f;
        wf_compute(f, score_f);
        int done = wf_extend(f, score_f, NULL);
#if AW_BIALIGN_PHASE2_END_REACHED_RETURNS
        if (done) return f->status;
#endif
        (void)done;
        last_forward = 1;
        if ((int64_t)score_f + score_r > step_limit) return ST_ERROR;
    }
    return ST_OK;
}

/* wavefront_bialign_alignment (recursive) */
static int bialign_alignment(bialigner_t* b, const uint8_t* p, int plen, const uint8_t* t, int tlen, int comp_begin,
                             int comp_end, int score_remaining, cigar_t* cigar, int level) {
    if ((uint32_t)level > b->work->max_level) b->work->max_level = (uint32_t)level;
    if (tlen == 0) {
        cigar_push_n(cigar, 'D', (size_t)plen);
        return ST_OK;
    } else if (plen == 0) {
        cigar_push_n(cigar, 'I', (size_t)tlen);
        return ST_OK;
    }
    if (score_remaining <= AW_BIALIGN_FALLBACK_MIN_SCORE) {
        b->work->n_base++;
        if (g_trace) fprintf(g_trace, "base L%d plen=%d tlen=%d cb=%d ce=%d rem=%d\n", level, plen, tlen, comp_begin, comp_end, score_remaining);
        return unialign_and_backtrace(&b->sub, p, plen, t, tlen, comp_begin, comp_end, cigar);
    }
    breakpoint_t bp;
    int st = bialign_find_breakpoint(b, p, plen, t, tlen, comp_begin, comp_end, &bp);
    if (st == ST_END_REACHED || st == ST_END_UNREACHABLE) {
        b->work->n_base++;
        b->work->n_end_reached_fallbacks++;
        if (g_trace) fprintf(g_trace, "fallback L%d plen=%d tlen=%d cb=%d ce=%d\n", level, plen, tlen, comp_begin, comp_end);
        return unialign_and_backtrace(&b->sub, p, plen, t, tlen, comp_begin, comp_end, cigar);
    }
    if (st != ST_OK) return st;
    b->work->n_breakpoints++;
    const int bh = bp.offset_forward, bv = bp.offset_forward - bp.k_forward;
    if (g_trace)
        fprintf(g_trace, "bp L%d plen=%d tlen=%d cb=%d ce=%d -> score=%d sf=%d sr=%d kf=%d of=%d comp=%d\n", level, plen, tlen,
                comp_begin, comp_end, bp.score, bp.score_forward, bp.score_reverse, bp.k_forward, bp.offset_forward, bp.component);
    if (bv < 0 || bv > plen || bh < 0 || bh > tlen) return ST_ERROR;
    st = bialign_alignment(b, p, bv, t, bh, comp_begin, bp.component, bp.score_forward, cigar, level + 1);
    if (st != ST_OK) return st;
    return bialign_alignment(b, p + bv, plen - bv, t + bh, tlen - bh, bp.component, comp_end, bp.score_reverse, cigar,
                             level + 1);
}

/* ------------------------------------------------------------------------------------------
 * public entry points
 * ---------------------------------------------------------------------------------------- */
int awo_mode_from_params(const awo_params_t* p) {
    if (p->has_gap2_open && p->has_gap2_extend) return AWO_MODE_AFFINE2P;
    if (p->gap_open == p->gap_extend && p->gap_open == p->mismatch_penalty) return AWO_MODE_EDIT;
    return AWO_MODE_AFFINE;
}

/* create_wfa_aligner (src/alignment.rs:263-289) -> WFA2 penalties */
/* `shifted` = the penalties the wavefronts run on (AW_SHIFT_* when match < 0); otherwise the user's own, for scoring the CIGAR */
static int pen_from_params_ex(const awo_params_t* p, pen_t* pen, int shifted) {
    memset(pen, 0, sizeof(*pen));
    const int mode = awo_mode_from_params(p);
    if (p->match_score > 0) return -1; /* wavefront_penalties_set_*: the match score must be negative or zero */
    pen->x = p->mismatch_penalty;
    if (mode == AWO_MODE_EDIT) {
        pen->o1 = p->mismatch_penalty;
        pen->e1 = p->mismatch_penalty;
    } else {
        pen->o1 = p->gap_open;
        pen->e1 = p->gap_extend;
    }
    if (mode == AWO_MODE_AFFINE2P) {
        pen->two_piece = 1;
        pen->o2 = p->gap2_open;
        pen->e2 = p->gap2_extend;
        if (pen->o2 < 0 || pen->e2 <= 0) return -1;
    }
    if (pen->x <= 0 || pen->o1 < 0 || pen->e1 <= 0) return -1;
    if (shifted && p->match_score < 0) {
        const int m = p->match_score;
        pen->x = AW_SHIFT_MISMATCH(pen->x, m);
        pen->o1 = AW_SHIFT_GAP_OPEN(pen->o1, m);
        pen->e1 = AW_SHIFT_GAP_EXTEND(pen->e1, m);
        if (pen->two_piece) {
            pen->o2 = AW_SHIFT_GAP_OPEN(pen->o2, m);
            pen->e2 = AW_SHIFT_GAP_EXTEND(pen->e2, m);
        }
    }
    return 0;
}
static int pen_from_params(const awo_params_t* p, pen_t* pen) { return pen_from_params_ex(p, pen, 1); }

/* penalty of a CIGAR under the user's own penalties, matches included (cigar_score_gap_affine*): -score */
int64_t awo_cigar_penalty(const awo_params_t* params, const uint8_t* ops, size_t n) {
    pen_t pn;
    if (pen_from_params_ex(params, &pn, 0)) return INT64_MIN;
    int64_t total = 0;
    size_t i = 0;
    while (i < n) {
        uint8_t op = ops[i];
        size_t j = i;
        while (j < n && ops[j] == op) ++j;
        int64_t L = (int64_t)(j - i);
        if (op == 'M') total += L * params->match_score;
        else if (op == 'X') total += L * pn.x;
        else if (op == 'I' || op == 'D') {
            int64_t c1 = pn.o1 + L * pn.e1;
            if (pn.two_piece) {
                int64_t c2 = pn.o2 + L * pn.e2;
                if (c2 < c1) c1 = c2;
            }
            total += c1;
        }
        i = j;
    }
    return total;
}

static int align_impl(const awo_params_t* params, const uint8_t* pattern, int plen, const uint8_t* text, int tlen,
                      awo_alignment_t* out, int bidirectional) {
    memset(out, 0, sizeof(*out));
    pen_t pn;
    if (pen_from_params(params, &pn) || plen < 0 || tlen < 0) {
        out->status = -1;
        out->score = INT_MAX;
        return -1;
    }
    bialigner_t b;
    b.work = &out->work;
    aligner_init(&b.fwd, &pn, 1, &out->work);
    aligner_init(&b.rev, &pn, 1, &out->work);
    aligner_init(&b.sub, &pn, 0, &out->work);
    cigar_t cigar = {NULL, 0, 0};
    int st;
    if (bidirectional) {
        /* wavefront_bialign: short sequences go straight to the base case */
        const int min_length = MAXI(plen, tlen) <= AW_BIALIGN_FALLBACK_MIN_LENGTH;
        st = bialign_alignment(&b, pattern, plen, text, tlen, AW_COMP_M, AW_COMP_M, min_length ? 0 : INT_MAX, &cigar, 0);
    } else {
        if (tlen == 0) { cigar_push_n(&cigar, 'D', (size_t)plen); st = ST_OK; }
        else if (plen == 0) { cigar_push_n(&cigar, 'I', (size_t)tlen); st = ST_OK; }
        else st = unialign_and_backtrace(&b.sub, pattern, plen, text, tlen, AW_COMP_M, AW_COMP_M, &cigar);
    }
    aligner_destroy(&b.fwd);
    aligner_destroy(&b.rev);
    aligner_destroy(&b.sub);
    if (st != ST_OK) {
        free(cigar.ops);
        out->status = st ? st : -1;
        out->score = INT_MAX;
        return -1;
    }
    out->status = 0;
    out->cigar = cigar.ops ? cigar.ops : (uint8_t*)malloc(1);
    out->cigar_len = cigar.n;
    /* cigar->score = -(penalty) with match == 0 (wavefront_compute_classic_score) */
    out->score = (int32_t)(-awo_cigar_penalty(params, out->cigar, out->cigar_len));
    return 0;
}

int awo_wfa_align(const awo_params_t* params, const uint8_t* pattern, int plen, const uint8_t* text, int tlen,
                  awo_alignment_t* out) {
    return align_impl(params, pattern, plen, text, tlen, out, 1);
}
int awo_wfa_align_unidirectional(const awo_params_t* params, const uint8_t* pattern, int plen, const uint8_t* text,
                                 int tlen, awo_alignment_t* out) {
    return align_impl(params, pattern, plen, text, tlen, out, 0);
}
void awo_alignment_free(awo_alignment_t* a) {
    if (a) {
        free(a->cigar);
        a->cigar = NULL;
        a->cigar_len = 0;
    }
}
