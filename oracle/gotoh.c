/*
 * gotoh.c -- independent O(nm) Gotoh dynamic program for gap-affine / two-piece gap-affine
 * penalties.  It shares no code with wfa2_biwfa.c and is used ONLY to certify that the oracle's
 * (and the CUDA path's) score is optimal on small inputs.  TEST INFRASTRUCTURE ONLY.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "aw_oracle.h"

#define INF64 ((int64_t)1 << 60)
static inline int64_t min2(int64_t a, int64_t b) { return a < b ? a : b; }

int64_t awo_gotoh_penalty(const awo_params_t* p, const uint8_t* pat, int plen, const uint8_t* txt, int tlen) {
    int mode = awo_mode_from_params(p);
    int64_t x = p->mismatch_penalty, o1, e1, o2 = 0, e2 = 0;
    int two = (mode == AWO_MODE_AFFINE2P);
    if (mode == AWO_MODE_EDIT) { o1 = x; e1 = x; }
    else { o1 = p->gap_open; e1 = p->gap_extend; }
    if (two) { o2 = p->gap2_open; e2 = p->gap2_extend; }
    size_t W = (size_t)tlen + 1;
    int64_t* H = (int64_t*)malloc(sizeof(int64_t) * W * 2);
    int64_t* D1 = (int64_t*)malloc(sizeof(int64_t) * W); /* vertical gap ending at (i,j), piece 1 */
    int64_t* D2 = (int64_t*)malloc(sizeof(int64_t) * W);
    int64_t* prev = H, *cur = H + W;
    prev[0] = 0;
    for (int j = 1; j <= tlen; ++j) {
        int64_t g = o1 + (int64_t)j * e1;
        if (two) g = min2(g, o2 + (int64_t)j * e2);
        prev[j] = g;
    }
    for (int j = 0; j <= tlen; ++j) D1[j] = D2[j] = INF64;
    for (int i = 1; i <= plen; ++i) {
        int64_t g = o1 + (int64_t)i * e1;
        if (two) g = min2(g, o2 + (int64_t)i * e2);
        cur[0] = g;
        int64_t I1 = INF64, I2 = INF64; /* horizontal gaps along this row */
        for (int j = 1; j <= tlen; ++j) {
            D1[j] = min2(prev[j] + o1 + e1, D1[j] + e1);
            I1 = min2(cur[j - 1] + o1 + e1, I1 + e1);
            int64_t best = prev[j - 1] + (pat[i - 1] == txt[j - 1] ? 0 : x);
            best = min2(best, min2(D1[j], I1));
            if (two) {
                D2[j] = min2(prev[j] + o2 + e2, D2[j] + e2);
                I2 = min2(cur[j - 1] + o2 + e2, I2 + e2);
                best = min2(best, min2(D2[j], I2));
            }
            cur[j] = best;
        }
        int64_t* tmp = prev; prev = cur; cur = tmp;
    }
    int64_t r = prev[tlen];
    free(H); free(D1); free(D2);
    return r;
}
