// aw_wfa.cuh -- the alignment hot path as one persistent sm_100a kernel per pair class:
//   K4 wavefront compute (gap-affine / gap-affine-2p M,I1,D1,I2,D2 recurrences)
//   K5 match extend (XOR + ffs/clz longest common prefix on 2-bit packed or byte words)
//   K6 biWFA breakpoint search (forward/reverse wavefronts, overlap test, DFS recursion stack)
//   K7 base case: full-history unidirectional WFA + WFA2-ordered backtrace
//   K8 CIGAR run-length encoding, op statistics, score, PAF text
// One CTA (NT threads; NT=32 is the one-warp-per-pair variant) owns one pair at a time and
// pulls pairs from a global work counter.  Replaces lib_wfa2's AffineWavefronts::align as
// driven by /root/reference/src/alignment.rs:201-261 and the CIGAR/PAF passes of
// src/alignment.rs:292-376 + src/lib.rs:71-112.  Tie-break constants: include/aw_wfa2_compat.h.
#pragma once
#include <limits.h>

#include "aw_common.cuh"

namespace awk {

#define AW_KFLAG_COUNT_ONLY 0x100u  // internal: statistics only (orientation passes), no text output
#ifndef AW_CTAS_PER_SM_256
#define AW_CTAS_PER_SM_256 4  // resident 256-thread CTAs per SM the register budget is planned for (measured best: 4 x 64 regs)
#endif
constexpr int NRED = 16;       // reduction slots, see RED_* below
constexpr int MAX_STACK = 96;  // DFS depth bound of the biWFA recursion
constexpr int HIST_META_INTS = 16;
constexpr int EDGE_ZONE = 64;  // I/D in-bounds tracking is only done this close to a wavefront end
constexpr int SEQ_SMEM_WORDS = 2048;  // 8 KB of shared memory for the pair's packed sequences (guards included)
// ---- shared-memory diagonal-band engine (phase 1 of the breakpoint search) ----
constexpr int BAND_T = 32;            // scores advanced per band without leaving shared memory
constexpr int BAND_WP = 1024;         // diagonals held per tile (output range + BAND_T halo either side)
constexpr int BAND_WT = BAND_WP - 2 * BAND_T;
constexpr int BAND_MROWS = 27;        // M history rows (max_score_scope <= 27)
constexpr int BAND_ROWS = BAND_MROWS + 3 + 3 + 2 + 2 + 1;  // + I1, D1, I2, D2 windows + one all-NULL row
constexpr int BAND_NULL_ROW = BAND_ROWS - 1;
constexpr int BAND_TAB = 13;          // per-step table: RED_* layout
constexpr int BAND_MIN_SCORE = 600;   // sub-problems expected to score less use the generic path
constexpr short NULL16 = -16384;
struct Meta16 {
    int lo[5], hi[5], akM, akAll;
};
__host__ __device__ constexpr size_t band_smem_bytes(int scope) {
    return sizeof(short) * BAND_ROWS * BAND_WP + sizeof(int) * 2 * (BAND_T + 1) * BAND_TAB + sizeof(Meta16) * 2 * (scope + BAND_T + 2) + 16;
}

enum { IN_MX = 0, IN_MO1, IN_I1E, IN_D1E, IN_MO2, IN_I2E, IN_D2E };
enum { ST_OK = 0, ST_END_REACHED = 1, ST_FAIL_WORKSPACE = 2 };
// red[] layout: per component c: RED_HI+c = max in-bounds k, RED_LO+c = max(-k); then the two
// antidiagonal bounds and the value at the end cell
enum { RED_HI = 0, RED_LO = 5, RED_AKM = 10, RED_AKALL = 11, RED_END = 12 };

struct SlotMeta {
    int lo[5], hi[5];   // trimmed range per component; empty iff lo > hi
    int akM;            // max antidiagonal 2*off-k over valid extended M cells (INT_MIN if none)
    int akAll;          // upper bound of 2*off-k over every component's non-null cells
    int clo, width;     // computed (allocated) range of this wavefront (history mode)
    int off;            // workspace offset of the wavefront's component block (history mode)
};

struct In {
    int off;     // element k lives at ws[off + k]
    int lo, hi;  // trimmed range; empty iff lo > hi
};

struct SubProblem {
    int pb, pe, tb, te;  // half-open ranges in the pair's pattern / text
    int cb, ce;          // begin / end component
    int rem;             // score_remaining
};

struct Breakpoint {
    int score, score_f, score_r, k_f, off_f, comp;
};

struct KParams {
    // sequence store
    const AwSlot* slots;
    const uint32_t* packed;
    const uint8_t* ascii;
    const char* ids;
    const uint32_t* id_off;  // n+1 offsets into ids
    // work list
    const aw_pair* pairs;
    const uint8_t* is_reverse;  // per pair, from the orientation stage
    const uint32_t* order;      // processing order (heaviest first) or NULL
    uint32_t npairs;
    unsigned int* next_pair;
    AwPen pen;
    uint32_t flags;
    // per-CTA workspace: [ring: 2 x (scope+1) x ncomp x W ints][history arena: hist_ints]
    int* ws;
    unsigned long long ws_ints_per_cta;
    int W;                   // allocated diagonals per ring wavefront
    int hist_ints;           // history arena size (ints)
    long long ring16_int_off; // int offset inside the CTA workspace of the int16 band ring (2 x (scope+BAND_T+2) x ncomp x W halfwords), or -1
    int* ws_hist_meta;       // [cta][hist_max_scores][HIST_META_INTS]
    int hist_max_scores;
    uint32_t* ws_runs;       // [cta][2][runs_cap]: pair runs, then leaf scratch
    unsigned long long runs_cap;
    // outputs
    AwPairOut* out;          // [npairs]
    char* text;
    unsigned long long* text_cursor;
    unsigned long long text_cap;
    uint8_t* bytes;
    unsigned long long* bytes_cursor;
    unsigned long long bytes_cap;
};

// ------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ void cta_sync() {
    if (NT == 32) __syncwarp();
    else __syncthreads();
}

template <int NT>
__device__ __forceinline__ void red_max(int* red, int idx, int v) {
    v = __reduce_max_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v != INT_MIN) {
        if (NT == 32) red[idx] = v;
        else atomicMax(&red[idx], v);
    }
}

// ---- K5: longest common prefix on packed words -------------------------------------------
// BITS = 2 (2-bit packed, 16 symbols / word) or 8 (ASCII, 4 symbols / word).  Guard words
// either side of every sequence make the over-reads legal; the result is clamped to maxlen.
template <int BITS>
__device__ __forceinline__ uint32_t load_fwd(const uint32_t* __restrict__ w, int pos) {
    constexpr int SPW = 32 / BITS;
    const unsigned idx = (unsigned)pos / SPW;  // pos >= 0
    const int sh = ((unsigned)pos % SPW) * BITS;
    return __funnelshift_r(w[idx], w[idx + 1], sh);
}
// symbols pos, pos-1, ... with `pos` in the most significant bits
template <int BITS>
__device__ __forceinline__ uint32_t load_rev(const uint32_t* __restrict__ w, int pos) {
    constexpr int SPW = 32 / BITS;
    const unsigned idx = (unsigned)pos / SPW;
    const int sh = (((unsigned)pos % SPW) + 1) * BITS;  // in [BITS, 32]
    return __funnelshift_rc(w[(int)idx - 1], w[idx], sh);
}
template <int BITS>
__device__ __forceinline__ int lcp_fwd(const uint32_t* __restrict__ pw, int pv, const uint32_t* __restrict__ tw, int th, int maxlen) {
    constexpr int SPW = 32 / BITS;
    int n = 0;
    while (n < maxlen) {
        const uint32_t x = load_fwd<BITS>(pw, pv + n) ^ load_fwd<BITS>(tw, th + n);
        if (x) {
            n += (__ffs(x) - 1) / BITS;
            break;
        }
        n += SPW;
    }
    return min(n, maxlen);
}
template <int BITS>
__device__ __forceinline__ int lcp_rev(const uint32_t* __restrict__ pw, int pv, const uint32_t* __restrict__ tw, int th, int maxlen) {
    constexpr int SPW = 32 / BITS;
    int n = 0;
    while (n < maxlen) {
        const uint32_t x = load_rev<BITS>(pw, pv - n) ^ load_rev<BITS>(tw, th - n);
        if (x) {
            n += __clz(x) / BITS;
            break;
        }
        n += SPW;
    }
    return min(n, maxlen);
}

// one aligner's view of the sub-problem sequences (wavefront_sequences_set_bounds): the reverse
// aligner reads both sub-ranges back to front
struct SeqView {
    const uint32_t* pw;
    const uint32_t* tw;
    int p0, t0;  // absolute symbol of v=0 / h=0 (forward: range begin; reverse: range end - 1)
    int plen, tlen;
    bool rev;
};
template <int BITS>
__device__ __forceinline__ int extend_cell(const SeqView& s, int k, int off) {
    const int v = off - k, h = off;
    const int maxlen = min(s.plen - v, s.tlen - h);
    if (maxlen <= 0) return off;
    const int n = s.rev ? lcp_rev<BITS>(s.pw, s.p0 - v, s.tw, s.t0 - h, maxlen) : lcp_fwd<BITS>(s.pw, s.p0 + v, s.tw, s.t0 + h, maxlen);
    return off + n;
}

// first (branch-free) round of the match extension: up to one word of symbols.  `more` is set
// when the whole word matched and symbols remain (the caller continues with lcp_fwd/lcp_rev).
template <int BITS>
__device__ __forceinline__ void extend_first(const SeqView& s, int k, int off, int& n, bool& more) {
    constexpr int SPW = 32 / BITS;
    int v = off - k, h = off;
    const int maxlen = min(s.plen - v, s.tlen - h);
    if (maxlen <= 0) v = h = 0;  // nothing to compare: keep the (unused) loads inside the sequences
    uint32_t x;
    int cnt;
    if (s.rev) {
        x = load_rev<BITS>(s.pw, s.p0 - v) ^ load_rev<BITS>(s.tw, s.t0 - h);
        cnt = __clz(x) / BITS;  // clz(0) = 32 -> SPW
    } else {
        x = load_fwd<BITS>(s.pw, s.p0 + v) ^ load_fwd<BITS>(s.tw, s.t0 + h);
        cnt = x ? (__ffs(x) - 1) / BITS : SPW;
    }
    n = max(0, min(cnt, maxlen));
    more = (x == 0) && (maxlen > SPW);
}
template <int BITS>
__device__ __forceinline__ int extend_rest(const SeqView& s, int k, int off) {  // off already advanced by one word
    constexpr int SPW = 32 / BITS;
    const int v = off - k, h = off;
    const int maxlen = min(s.plen - v, s.tlen - h);
    if (maxlen <= 0) return off;
    const int n = s.rev ? lcp_rev<BITS>(s.pw, s.p0 - v, s.tw, s.t0 - h, maxlen) : lcp_fwd<BITS>(s.pw, s.p0 + v, s.tw, s.t0 + h, maxlen);
    (void)SPW;
    return off + n;
}

// workspace element type: int (any length) or short (offsets < 32000; halves the L2/HBM footprint).
// Every negative offset means "null", so the int16 form stores one canonical negative value.
template <class WS>
__device__ __forceinline__ WS to_ws(int v);
template <>
__device__ __forceinline__ int to_ws<int>(int v) { return v; }
template <>
__device__ __forceinline__ short to_ws<short>(int v) { return (short)(v < 0 ? (int)NULL16 : v); }

struct StepOut {
    int lo[5], hi[5];
    int akM, akAll;
    int endval;
    bool ambiguous;  // an I/D trim end fell outside the tracked edge zones (needs an exact rescan)
};

// ---- K4+K5: compute wavefront s from its inputs, extend M, and accumulate the block-wide
// reductions (trim ends, antidiagonal bounds, end-cell value) into `red`.  No barrier here.
// Restates wavefront_compute_affine2p_idm + wavefront_extend_matches_packed_end2end(_max)
// (SURVEY A.2, A.3).  `red` must hold INT_MIN on entry.
template <int NT, int BITS, bool TWO, class WS>
__device__ __forceinline__ void wf_cells(WS* __restrict__ ws, const In (&in)[7], const int (&out)[5], int lo, int hi, const SeqView& sv, int k_end,
                                         int comp_end, int* red) {
    const int tid = threadIdx.x;
    const unsigned tlen = (unsigned)sv.tlen, plen = (unsigned)sv.plen;
    // interior cells whose every input (k-1, k, k+1) is inside every input's range need no checks
    int fast_lo = INT_MIN, fast_hi = INT_MAX;
    {
        bool all_present = true;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            if (!TWO && i >= IN_MO2) continue;
            all_present = all_present && (in[i].lo <= in[i].hi);
            fast_lo = max(fast_lo, in[i].lo + 1);
            fast_hi = min(fast_hi, in[i].hi - 1);
        }
        if (!all_present) {
            fast_lo = 1;
            fast_hi = 0;
        }
    }
    const bool narrow = (hi - lo) < 2 * EDGE_ZONE;
    int akM = INT_MIN, akAll = INT_MIN;
    // in-bounds ends of one component inside an edge row: one ballot, the leader lane publishes both ends
    auto edge_track = [&](int c, bool inb, int krow) {
        const unsigned amask = __activemask();
        const unsigned mask = __ballot_sync(amask, inb);
        if (mask != 0 && (int)(tid & 31) == __ffs(amask) - 1) {
            atomicMax(&red[RED_HI + c], krow + 31 - __clz(mask));
            atomicMax(&red[RED_LO + c], -(krow + __ffs(mask) - 1));
        }
    };
    // one running base pointer (ws + k) plus a 32-bit element offset per input / output array
    const int k0 = lo + tid;
    WS* pk = ws + k0;
    const int o_mx = in[IN_MX].off, o_mo1 = in[IN_MO1].off, o_i1e = in[IN_I1E].off, o_d1e = in[IN_D1E].off;
    const int o_mo2 = in[IN_MO2].off, o_i2e = in[IN_I2E].off, o_d2e = in[IN_D2E].off;
    const int w_m = out[AW_COMP_M], w_i1 = out[AW_COMP_I1], w_d1 = out[AW_COMP_D1], w_i2 = out[AW_COMP_I2], w_d2 = out[AW_COMP_D2];
    // one cell: recurrences, bounds, extend, stores, trim / antidiagonal tracking
    auto cell = [&](int k, int j, int mo1l, int mo1r, int i1l, int d1r, int mo2l, int mo2r, int i2l, int d2r, int mx) {
        const int i1 = max(mo1l, i1l) + 1;
        const int d1 = max(mo1r, d1r);
        int i2 = AW_NULLV, d2 = AW_NULLV, ins = i1, del = d1;
        if (TWO) {
            i2 = max(mo2l, i2l) + 1;
            d2 = max(mo2r, d2r);
            ins = max(i1, i2);
            del = max(d1, d2);
        }
        int m = max(del, max(mx + 1, ins));
        if (m >= 0) akAll = max(akAll, 2 * m - k);  // m (pre-null) dominates every component at k
        if ((unsigned)m > tlen || (unsigned)(m - k) > plen) m = AW_NULLV;
        if (m >= 0) {
            m = extend_cell<BITS>(sv, k, m);
            akM = max(akM, 2 * m - k);
        }
        (void)j;
        pk[w_m] = to_ws<WS>(m);
        pk[w_i1] = to_ws<WS>(i1);
        pk[w_d1] = to_ws<WS>(d1);
        if (TWO) {
            pk[w_i2] = to_ws<WS>(i2);
            pk[w_d2] = to_ws<WS>(d2);
        }
        // wavefront_compute_trim_ends keeps [first, last] in-bounds cell of every component: only rows
        // (32 consecutive diagonals of one warp) that touch an edge zone can hold those ends
        const int krow = k - (int)(tid & 31);
        if (narrow || krow - lo < EDGE_ZONE || hi - (krow + 31) < EDGE_ZONE) {
            edge_track(AW_COMP_M, m >= 0, krow);
            edge_track(AW_COMP_I1, (unsigned)i1 <= tlen && (unsigned)(i1 - k) <= plen, krow);
            edge_track(AW_COMP_D1, (unsigned)d1 <= tlen && (unsigned)(d1 - k) <= plen, krow);
            if (TWO) {
                edge_track(AW_COMP_I2, (unsigned)i2 <= tlen && (unsigned)(i2 - k) <= plen, krow);
                edge_track(AW_COMP_D2, (unsigned)d2 <= tlen && (unsigned)(d2 - k) <= plen, krow);
            }
        }
        if (k == k_end) red[RED_END] = (comp_end == AW_COMP_M) ? m : (comp_end == AW_COMP_I1) ? i1 : (comp_end == AW_COMP_D1) ? d1 : (comp_end == AW_COMP_I2) ? i2 : d2;
    };
    int k = k0;
    for (; k <= hi; k += NT) {
        int mo1l, mo1r, i1l, d1r, mo2l = AW_NULLV, mo2r = AW_NULLV, i2l = AW_NULLV, d2r = AW_NULLV, mx;
        if (k >= fast_lo && k <= fast_hi) {  // interior: immediate-offset loads, no checks
            mo1l = pk[o_mo1 - 1];
            mo1r = pk[o_mo1 + 1];
            i1l = pk[o_i1e - 1];
            d1r = pk[o_d1e + 1];
            if (TWO) {
                mo2l = pk[o_mo2 - 1];
                mo2r = pk[o_mo2 + 1];
                i2l = pk[o_i2e - 1];
                d2r = pk[o_d2e + 1];
            }
            mx = pk[o_mx];
        } else {
            auto ck = [&](const In& w, int o, int d) -> int { return (k + d >= w.lo && k + d <= w.hi) ? (int)pk[o + d] : AW_NULLV; };
            mo1l = ck(in[IN_MO1], o_mo1, -1);
            mo1r = ck(in[IN_MO1], o_mo1, 1);
            i1l = ck(in[IN_I1E], o_i1e, -1);
            d1r = ck(in[IN_D1E], o_d1e, 1);
            if (TWO) {
                mo2l = ck(in[IN_MO2], o_mo2, -1);
                mo2r = ck(in[IN_MO2], o_mo2, 1);
                i2l = ck(in[IN_I2E], o_i2e, -1);
                d2r = ck(in[IN_D2E], o_d2e, 1);
            }
            mx = ck(in[IN_MX], o_mx, 0);
        }
        cell(k, 0, mo1l, mo1r, i1l, d1r, mo2l, mo2r, i2l, d2r, mx);
        pk += NT;
    }
    red_max<NT>(red, RED_AKM, akM);
    red_max<NT>(red, RED_AKALL, akAll);
}

// after the barrier: trimmed ranges (wavefront_compute_trim_ends) from the reductions
template <bool TWO>
__device__ __forceinline__ void wf_finish(const int* red, int lo, int hi, StepOut& so) {
    const bool narrow = (hi - lo) < 2 * EDGE_ZONE;
    so.ambiguous = false;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        if (!TWO && (c == AW_COMP_I2 || c == AW_COMP_D2)) {
            so.lo[c] = 1;
            so.hi[c] = 0;
            continue;
        }
        const int h = red[RED_HI + c], l = red[RED_LO + c];
        if (h == INT_MIN) {  // nothing in bounds (among the tracked cells)
            so.lo[c] = lo;
            so.hi[c] = lo - 1;
            if (!narrow) so.ambiguous = true;
        } else {
            so.lo[c] = -l;
            so.hi[c] = h;
            // exact only if each end was found inside its tracked zone
            if (!narrow && ((-l) - lo >= EDGE_ZONE || hi - h >= EDGE_ZONE)) so.ambiguous = true;
        }
    }
    so.akM = red[RED_AKM];
    so.akAll = red[RED_AKALL];
    so.endval = red[RED_END];
}

// exact trim of every component by re-reading the stored wavefront (rare slow path)
template <int NT, bool TWO, class WS>
__device__ __noinline__ void wf_rescan(const WS* __restrict__ ws, const int (&out)[5], int lo, int hi, int plen_, int tlen_, int* red, StepOut& so) {
    const unsigned tlen = (unsigned)tlen_, plen = (unsigned)plen_;
    cta_sync<NT>();
    if (threadIdx.x < NRED) red[threadIdx.x] = INT_MIN;
    cta_sync<NT>();
    int vhi[5], vlo[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) vhi[c] = vlo[c] = INT_MIN;
    for (int k = lo + (int)threadIdx.x; k <= hi; k += NT) {
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            if (!TWO && (c == AW_COMP_I2 || c == AW_COMP_D2)) continue;
            const int v = ws[out[c] + k];
            if ((unsigned)v <= tlen && (unsigned)(v - k) <= plen) {
                if (vlo[c] == INT_MIN) vlo[c] = -k;
                vhi[c] = k;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        if (!TWO && (c == AW_COMP_I2 || c == AW_COMP_D2)) continue;
        red_max<NT>(red, RED_HI + c, vhi[c]);
        red_max<NT>(red, RED_LO + c, vlo[c]);
    }
    cta_sync<NT>();
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        if (!TWO && (c == AW_COMP_I2 || c == AW_COMP_D2)) continue;
        const int h = red[RED_HI + c], l = red[RED_LO + c];
        if (h == INT_MIN) {
            so.lo[c] = lo;
            so.hi[c] = lo - 1;
        } else {
            so.lo[c] = -l;
            so.hi[c] = h;
        }
    }
    so.ambiguous = false;
    cta_sync<NT>();
    if (threadIdx.x < NRED) red[threadIdx.x] = INT_MIN;
    cta_sync<NT>();
}

__device__ __forceinline__ bool end_reached(const StepOut& so, int comp_end, int k_end, int tlen) {
    const int l = (comp_end == AW_COMP_M) ? so.lo[0] : (comp_end == AW_COMP_I1) ? so.lo[1] : (comp_end == AW_COMP_I2) ? so.lo[2] : (comp_end == AW_COMP_D1) ? so.lo[3] : so.lo[4];
    const int h = (comp_end == AW_COMP_M) ? so.hi[0] : (comp_end == AW_COMP_I1) ? so.hi[1] : (comp_end == AW_COMP_I2) ? so.hi[2] : (comp_end == AW_COMP_D1) ? so.hi[3] : so.hi[4];
    return l <= k_end && k_end <= h && so.endval >= tlen;
}

// number of decimal digits of v
__device__ __forceinline__ int ndigits(unsigned long long v) {
    int n = 1;
    while (v >= 10) {
        v /= 10;
        ++n;
    }
    return n;
}
__device__ __forceinline__ char* put_u64(char* p, unsigned long long v) {
    const int n = ndigits(v);
    for (int i = n - 1; i >= 0; --i) {
        p[i] = (char)('0' + (int)(v % 10));
        v /= 10;
    }
    return p + n;
}

// {:.6} of matches/alignment_length (src/lib.rs:83-87,108): IEEE division, then exact
// round-half-even of x * 10^6 (what Rust's flt2dec and C's printf both produce)
__device__ __forceinline__ unsigned identity_micro(unsigned long long matches, unsigned long long alnlen) {
    if (alnlen == 0 || matches == 0) return 0;
    const double x = (double)matches / (double)alnlen;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(x);
    const int e = (int)((bits >> 52) & 0x7ff);
    const unsigned long long mant = (bits & ((1ull << 52) - 1)) | (1ull << 52);
    const int shift = 1075 - e;  // x = mant * 2^-shift, shift >= 52 because x <= 1
    const unsigned __int128 P = (unsigned __int128)mant * 1000000ull;
    if (shift >= 120) return 0;
    const unsigned __int128 one = (unsigned __int128)1 << shift;
    unsigned long long q = (unsigned long long)(P >> shift);
    const unsigned __int128 rem = P & (one - 1), half = one >> 1;
    if (rem > half || (rem == half && (q & 1))) ++q;
    return (unsigned)q;
}

template <int NT>
__device__ __forceinline__ unsigned long long block_excl_scan(unsigned long long v, unsigned long long* sbuf, unsigned long long* total) {
    // sbuf: NT entries of shared memory
    const int tid = threadIdx.x;
    cta_sync<NT>();
    sbuf[tid] = v;
    cta_sync<NT>();
    for (int d = 1; d < NT; d <<= 1) {
        unsigned long long t = (tid >= d) ? sbuf[tid - d] : 0;
        cta_sync<NT>();
        sbuf[tid] += t;
        cta_sync<NT>();
    }
    *total = sbuf[NT - 1];
    return sbuf[tid] - v;
}

// ------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------
template <int NT, int BITS, bool TWO, class WS>
__global__ void __launch_bounds__(NT, (NT >= 256) ? AW_CTAS_PER_SM_256 : (NT == 128 ? 4 : 1)) aw_align_kernel(const KParams P) {
    constexpr int NCOMP = TWO ? 5 : 3;
    extern __shared__ unsigned long long smem_raw[];
    const int scope = P.pen.scope;
    const int ring_n = scope + 1;  // one spare slot: the reverse step is computed speculatively
    SlotMeta* ring_meta = reinterpret_cast<SlotMeta*>(smem_raw);                                     // [2][ring_n]
    int* cand = reinterpret_cast<int*>(ring_meta + 2 * ring_n);                                     // [scope*5] candidate tests
    int* hitk = cand + scope * 5;                                                                    // [scope*5] first hit per candidate
    unsigned long long* scanbuf = reinterpret_cast<unsigned long long*>(hitk + scope * 5);  // [NT]; 8-byte aligned: 60*2*ring_n + 40*scope
#ifndef AW_ENABLE_BAND
#define AW_ENABLE_BAND 0  // the shared-memory diagonal-band engine is experimental: measured slower than the generic loop
#endif
    constexpr bool BAND = (NT == 256) && (AW_ENABLE_BAND != 0);
    const int rn16 = scope + BAND_T + 2;  // slots of the int16 band ring
    short* band_rows = reinterpret_cast<short*>(scanbuf + NT);                    // [BAND_ROWS][BAND_WP]
    int* band_tab = reinterpret_cast<int*>(band_rows + BAND_ROWS * BAND_WP);      // [2][BAND_T+1][BAND_TAB]
    Meta16* meta16 = reinterpret_cast<Meta16*>(band_tab + 2 * (BAND_T + 1) * BAND_TAB);  // [2][rn16]
    __shared__ int s_band_flag;
    __shared__ int red[2][3][NRED];
    __shared__ SubProblem stack[MAX_STACK];
    __shared__ unsigned s_next;
    __shared__ unsigned s_nruns;
    __shared__ int s_ncand;
    __shared__ unsigned long long s_acc[8];
    __shared__ unsigned long long s_text_off, s_bytes_off;
    __shared__ uint32_t s_seq[SEQ_SMEM_WORDS];

    const int tid = threadIdx.x;
    const AwPen pen = P.pen;
    int* const ws_i = P.ws + (size_t)blockIdx.x * P.ws_ints_per_cta;
    WS* const ws = reinterpret_cast<WS*>(ws_i);  // all wavefront offsets below are in WS elements
    const int W = P.W;
    const int hist_base = 2 * ring_n * NCOMP * W;  // history arena starts after the rings
    int* const hist_meta = P.ws_hist_meta + (size_t)blockIdx.x * (size_t)P.hist_max_scores * HIST_META_INTS;
    uint32_t* const pair_runs = P.ws_runs + (size_t)blockIdx.x * 2 * P.runs_cap;
    uint32_t* const leaf_runs = pair_runs + P.runs_cap;

    if (tid < 2 * 3 * NRED) (&red[0][0][0])[tid] = INT_MIN;
    int red_i = 0;  // rotating reduction buffer index (uniform)
    cta_sync<NT>();

    auto comp_idx = [](int c) -> int { return TWO ? c : (c == AW_COMP_D1 ? 2 : c); };
    auto rotate_red = [&]() {
        // recycle the buffers used two steps ago (everybody finished reading them before the last barrier)
        const int nxt = (red_i + 2) % 3;
        if (tid < NRED) {
            red[0][nxt][tid] = INT_MIN;
            red[1][nxt][tid] = INT_MIN;
        }
        red_i = (red_i + 1) % 3;
    };

    for (;;) {
        if (tid == 0) s_next = atomicAdd(P.next_pair, 1u);
        cta_sync<NT>();
        const unsigned work_i = s_next;
        cta_sync<NT>();
        if (work_i >= P.npairs) break;
        const unsigned pair_i = P.order ? P.order[work_i] : work_i;
        const aw_pair pr = P.pairs[pair_i];
        const unsigned is_rev = P.is_reverse ? P.is_reverse[pair_i] : 0u;
        const AwSlot qs = P.slots[2 * pr.query_idx + is_rev];
        const AwSlot ts = P.slots[2 * pr.target_idx];
        const uint32_t* pw = (BITS == 2) ? P.packed + qs.packed_off : reinterpret_cast<const uint32_t*>(P.ascii + qs.ascii_off);
        const uint32_t* tw = (BITS == 2) ? P.packed + ts.packed_off : reinterpret_cast<const uint32_t*>(P.ascii + ts.ascii_off);
        const int PLEN = (int)qs.len, TLEN = (int)ts.len;
        {
            // stage both sequences (with 2 guard words either side) in shared memory when they fit
            constexpr int SPW = 32 / BITS;
            const int pwords = PLEN / SPW + 1, twords = TLEN / SPW + 1;
            if (pwords + twords + 8 <= SEQ_SMEM_WORDS) {
                for (int i = tid; i < pwords + 4; i += NT) s_seq[i] = pw[i - 2];
                for (int i = tid; i < twords + 4; i += NT) s_seq[pwords + 4 + i] = tw[i - 2];
                pw = s_seq + 2;
                tw = s_seq + pwords + 4 + 2;
            }
        }
        const int koff = min(PLEN + 1, W / 2);  // diagonal k lives at index k + koff
        const int kmin_alloc = -koff, kmax_alloc = W - 1 - koff;

        int status = ST_OK;
        unsigned long long cyc[6] = {0, 0, 0, 0, 0, 0};
        long long tmark = clock64();
        auto lap = [&](int i) {
            const long long t = clock64();
            cyc[i] += (unsigned long long)(t - tmark);
            tmark = t;
        };
        unsigned long long w_cells = 0;
        unsigned w_steps = 0, w_bps = 0, w_base = 0, w_maxbase = 0;
        if (tid == 0) s_nruns = 0;
        int sp_n = 0;
        {
            // wavefront_bialign: short sequences go straight to the base case
            SubProblem top = {0, PLEN, 0, TLEN, AW_COMP_M, AW_COMP_M, (max(PLEN, TLEN) <= AW_BIALIGN_FALLBACK_MIN_LENGTH) ? 0 : INT_MAX};
            stack[sp_n++] = top;
        }
        cta_sync<NT>();

        // appends a run to the pair's CIGAR (thread 0 only), merging with the previous run
        auto emit_run = [&](unsigned op, unsigned len) {
            if (len == 0) return;
            unsigned n = s_nruns;
            if (n > 0 && (pair_runs[n - 1] & 3u) == op) {
                pair_runs[n - 1] += len << 2;
            } else if (n < P.runs_cap) {
                pair_runs[n] = (len << 2) | op;
                s_nruns = n + 1;
            }
        };
        auto set_empty = [&](SlotMeta& m) {
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                m.lo[c] = 1;
                m.hi[c] = 0;
            }
            m.akM = m.akAll = INT_MIN;
            m.clo = 0;
            m.width = 0;
            m.off = 0;
        };
        auto store_meta = [&](SlotMeta& m, const StepOut& so) {
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                m.lo[c] = so.lo[c];
                m.hi[c] = so.hi[c];
            }
            m.akM = so.akM;
            m.akAll = so.akAll;
        };
        // wavefront_compute_limits_input over the non-empty inputs
        auto limits = [&](const In (&in)[7], int& lo, int& hi) {
            lo = INT_MAX;
            hi = INT_MIN;
            auto lim = [&](const In& w, int dl, int dh) {
                if (w.lo <= w.hi) {
                    lo = min(lo, w.lo + dl);
                    hi = max(hi, w.hi + dh);
                }
            };
            lim(in[IN_MX], 0, 0);
            lim(in[IN_MO1], -1, 1);
            lim(in[IN_I1E], 1, 1);
            lim(in[IN_D1E], -1, -1);
            if (TWO) {
                lim(in[IN_MO2], -1, 1);
                lim(in[IN_I2E], 1, 1);
                lim(in[IN_D2E], -1, -1);
            }
        };

        while (sp_n > 0 && status == ST_OK) {
            const SubProblem sp = stack[--sp_n];
            const int plen = sp.pe - sp.pb, tlen = sp.te - sp.tb;
            // ---- wavefront_bialign_alignment: trivial cases ----
            if (tlen == 0 || plen == 0) {
                if (tid == 0) {
                    if (tlen == 0) emit_run(AW_OP_D, (unsigned)plen);
                    else emit_run(AW_OP_I, (unsigned)tlen);
                }
                cta_sync<NT>();
                continue;
            }
            const int k_end = tlen - plen;
            bool do_base = sp.rem <= AW_BIALIGN_FALLBACK_MIN_SCORE;
            Breakpoint bp;
            bp.score = INT_MAX;
            bp.score_f = bp.score_r = bp.k_f = bp.off_f = 0;
            bp.comp = AW_COMP_M;

            if (!do_base) {
                // =========== K6: wavefront_bialign_find_breakpoint ===========
                SeqView svd[2];
                svd[0] = SeqView{pw, tw, sp.pb, sp.tb, plen, tlen, false};
                svd[1] = SeqView{pw, tw, sp.pe - 1, sp.te - 1, plen, tlen, true};
                const int cbeg[2] = {sp.cb, sp.ce}, cend[2] = {sp.ce, sp.cb};
                int cur_slot[2] = {0, 0};  // ring slot of the newest committed/computed score per direction
                auto ring_off = [&](int d, int slot, int c) -> int { return ((d * ring_n + slot) * NCOMP + comp_idx(c)) * W + koff; };
                auto slot_back = [&](int slot, int back) -> int {  // slot of (score - back), back <= scope
                    const int s = slot - back;
                    return s < 0 ? s + ring_n : s;
                };
                // score-0 wavefront of direction d (wavefront_unialign_init_end2end) + extend
                auto init_dir = [&](int d, StepOut& so) {
                    int* r = red[d][red_i];
                    if (tid == 0) {
                        int m = 0;
                        if (cbeg[d] == AW_COMP_M) m = extend_cell<BITS>(svd[d], 0, 0);
                        ws[ring_off(d, 0, cbeg[d])] = to_ws<WS>(m);
                        r[RED_AKM] = (cbeg[d] == AW_COMP_M) ? 2 * m : INT_MIN;
                        r[RED_AKALL] = 2 * m;
                        if (cbeg[d] == AW_COMP_M && cend[d] == AW_COMP_M && k_end == 0) r[RED_END] = m;
                    }
                    cta_sync<NT>();
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        so.lo[c] = (c == cbeg[d]) ? 0 : 1;
                        so.hi[c] = 0;
                    }
                    so.akM = r[RED_AKM];
                    so.akAll = r[RED_AKALL];
                    so.endval = r[RED_END];
                    so.ambiguous = false;
                    cur_slot[d] = 0;
                    store_meta(ring_meta[d * ring_n + 0], so);
                    rotate_red();
                };
                // what survives of a step until its barrier: the computed range (null step iff lo > hi)
                struct Range {
                    int lo, hi;
                };
                auto out_offsets = [&](int d, int slot, int (&out)[5]) {
#pragma unroll
                    for (int c = 0; c < 5; ++c) out[c] = ring_off(d, slot, (TWO || c == AW_COMP_M || c == AW_COMP_I1 || c == AW_COMP_D1) ? c : AW_COMP_M);
                };
                // descriptors + limits + cell loop of wavefront `s` (ring slot `slot`) of direction d; no barrier
                auto launch_dir = [&](int d, int s, int slot) -> Range {
                    In in[7];
                    auto fetch = [&](int c, int back) -> In {
                        In w;
                        if (s - back < 0) {
                            w.off = 0;
                            w.lo = 1;
                            w.hi = 0;
                            return w;
                        }
                        const int sl = slot_back(slot, back);
                        const SlotMeta& m = ring_meta[d * ring_n + sl];
                        w.off = ring_off(d, sl, c);
                        w.lo = m.lo[c];
                        w.hi = m.hi[c];
                        return w;
                    };
                    in[IN_MX] = fetch(AW_COMP_M, pen.x);
                    in[IN_MO1] = fetch(AW_COMP_M, pen.o1 + pen.e1);
                    in[IN_I1E] = fetch(AW_COMP_I1, pen.e1);
                    in[IN_D1E] = fetch(AW_COMP_D1, pen.e1);
                    if (TWO) {
                        in[IN_MO2] = fetch(AW_COMP_M, pen.o2 + pen.e2);
                        in[IN_I2E] = fetch(AW_COMP_I2, pen.e2);
                        in[IN_D2E] = fetch(AW_COMP_D2, pen.e2);
                    } else {
                        in[IN_MO2].off = in[IN_I2E].off = in[IN_D2E].off = 0;
                        in[IN_MO2].lo = in[IN_I2E].lo = in[IN_D2E].lo = 1;
                        in[IN_MO2].hi = in[IN_I2E].hi = in[IN_D2E].hi = 0;
                    }
                    Range r;
                    limits(in, r.lo, r.hi);
                    ++w_steps;
                    if (r.lo > r.hi) return r;
                    if (r.lo < kmin_alloc || r.hi > kmax_alloc) {
                        status = ST_FAIL_WORKSPACE;
                        return r;
                    }
                    if (status != ST_OK) return r;
                    int out[5];
                    out_offsets(d, slot, out);
                    wf_cells<NT, BITS, TWO, WS>(ws, in, out, r.lo, r.hi, d == 0 ? svd[0] : svd[1], k_end, d == 0 ? cend[0] : cend[1], red[d][red_i]);
                    w_cells += (unsigned long long)(r.hi - r.lo + 1) * NCOMP;
                    return r;
                };
                // after the barrier: trimmed ranges -> ring meta; returns END_REACHED of this wavefront
                auto finish_dir = [&](int d, int slot, const Range& r) -> bool {
                    SlotMeta& mt = ring_meta[d * ring_n + slot];
                    if (r.lo > r.hi || status != ST_OK) {
                        set_empty(mt);
                        return false;
                    }
                    StepOut so;
                    wf_finish<TWO>(red[d][red_i], r.lo, r.hi, so);
                    if (so.ambiguous) {
                        int out[5];
                        out_offsets(d, slot, out);
                        wf_rescan<NT, TWO, WS>(ws, out, r.lo, r.hi, plen, tlen, red[d][red_i], so);
                    }
                    store_meta(mt, so);
                    return end_reached(so, d == 0 ? cend[0] : cend[1], k_end, tlen);
                };
                auto next_slot = [&](int slot) -> int { return slot + 1 == ring_n ? 0 : slot + 1; };

                // wavefront_bialign_overlap: A0 = direction d0 at score s0, A1 = direction d1 at scores s1..s1-scope+1
                auto overlap = [&](int d0, int d1, int s0, int s1) {
                    const int slot0 = cur_slot[d0];
                    const SlotMeta& m0 = ring_meta[d0 * ring_n + slot0];
                    const int kinv = tlen - plen;
                    const int bp_entry = bp.score;
                    // test t = i*5 + oi: reverse score s1-i, component order D2,I2,D1,I1,M
                    auto test_comp = [](int oi) -> int { return oi == 0 ? AW_COMP_D2 : oi == 1 ? AW_COMP_I2 : oi == 2 ? AW_COMP_D1 : oi == 3 ? AW_COMP_I1 : AW_COMP_M; };
                    auto credit_of = [&](int c) -> int { return (c == AW_COMP_M) ? 0 : ((c == AW_COMP_D1 || c == AW_COMP_I1) ? pen.o1 : pen.o2); };
                    // candidate tests (gate + range intersection + antidiagonal bound), kept in order, by warp 0
                    if (tid < 32) {
                        int ncand = 0;
                        const int ntests = min(scope, s1 + 1) * 5;
                        for (int base = 0; base < ntests; base += 32) {
                            const int t = base + tid;
                            bool ok = false;
                            if (t < ntests) {
                                const int i = t / 5, oi = t - 5 * i, c = test_comp(oi);
                                const int si = s1 - i;
                                if (TWO || (c != AW_COMP_D2 && c != AW_COMP_I2)) {
                                    const SlotMeta& m1 = ring_meta[d1 * ring_n + slot_back(cur_slot[d1], i)];
                                    const int lo0 = m0.lo[c], hi0 = m0.hi[c], lo1 = kinv - m1.hi[c], hi1 = kinv - m1.lo[c];
                                    ok = (s0 + si - credit_of(c) < bp_entry) && lo0 <= hi0 && m1.lo[c] <= m1.hi[c] && !(hi1 < lo0 || hi0 < lo1);
                                    if (ok) {
                                        const long long a0 = (c == AW_COMP_M) ? m0.akM : m0.akAll, a1 = (c == AW_COMP_M) ? m1.akM : m1.akAll;
                                        ok = a0 + a1 >= (long long)plen + tlen;  // necessary for off0 + off1 >= tlen
                                    }
                                }
                            }
                            const unsigned mask = __ballot_sync(0xffffffffu, ok);
                            if (ok) {
                                const int pos = ncand + __popc(mask & ((1u << tid) - 1u));
                                cand[pos] = t;
                                hitk[pos] = INT_MAX;
                            }
                            ncand += __popc(mask);
                        }
                        if (tid == 0) s_ncand = ncand;
                    }
                    cta_sync<NT>();
                    const int ncand = s_ncand;
                    for (int j = 0; j < ncand; ++j) {
                        const int t = cand[j], i = t / 5, c = test_comp(t - 5 * i);
                        const int sl1 = slot_back(cur_slot[d1], i);
                        const SlotMeta& m1 = ring_meta[d1 * ring_n + sl1];
                        const int lo_1 = kinv - m1.hi[c], hi_1 = kinv - m1.lo[c];
                        const int max_lo = max(m0.lo[c], lo_1), min_hi = min(m0.hi[c], hi_1);
                        const WS* p0 = ws + ring_off(d0, slot0, c);
                        const WS* p1 = ws + ring_off(d1, sl1, c);
                        int best = INT_MAX;
                        // 4 diagonals per thread per round: the 8 loads are independent and issue together
                        for (int kb0 = max_lo + tid; kb0 <= min_hi && best == INT_MAX; kb0 += 4 * NT) {
                            int h0[4], h1[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int k0 = min(kb0 + j * NT, min_hi);  // clamped duplicates are re-tested harmlessly
                                h0[j] = p0[k0];
                                h1[j] = p1[kinv - k0];
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int k0 = kb0 + j * NT;
                                if (k0 > min_hi || best != INT_MAX) continue;
                                if (h0[j] + h1[j] >= tlen) {
                                    if (c != AW_COMP_M) {  // indel2indel: the forward cell must be in bounds
                                        const int kk = (d0 == 0) ? k0 : kinv - k0, hh = (d0 == 0) ? h0[j] : h1[j];
                                        if (hh - kk > plen || hh > tlen) continue;
                                    }
                                    best = k0;
                                }
                            }
                        }
                        best = __reduce_min_sync(0xffffffffu, best);
                        if ((tid & 31) == 0 && best != INT_MAX) atomicMin(&hitk[j], best);
                    }
                    cta_sync<NT>();
                    // replay the candidate tests in WFA2's order with the live breakpoint score
                    for (int j = 0; j < ncand; ++j) {
                        const int k0 = hitk[j];
                        if (k0 == INT_MAX) continue;
                        const int t = cand[j], i = t / 5, c = test_comp(t - 5 * i);
                        const int si = s1 - i, credit = credit_of(c);
                        if (s0 + si - credit >= bp.score) continue;
                        const int sl1 = slot_back(cur_slot[d1], i);
                        const int k1 = kinv - k0;
                        const int h0 = ws[ring_off(d0, slot0, c) + k0], h1 = ws[ring_off(d1, sl1, c) + k1];
                        if (d0 == 0) {
                            bp.score_f = s0;
                            bp.score_r = si;
                            bp.k_f = k0;
                            bp.off_f = h0;
                        } else {
                            bp.score_f = si;
                            bp.score_r = s0;
                            bp.k_f = k1;
                            bp.off_f = h1;
                        }
                        bp.score = s0 + si - credit;
                        bp.comp = c;
                    }
                    cta_sync<NT>();  // cand/hitk are rewritten by the next call
                };

                StepOut so;
                int score_f = 0, score_r = 0, f_ak = 0, r_ak = 0;
                bool fb_end = false;  // END_REACHED -> fall back to the base case
                init_dir(0, so);
                if (end_reached(so, cend[0], k_end, tlen)) fb_end = true;
                f_ak = max(0, so.akM);
                if (!fb_end) {
                    init_dir(1, so);
                    if (end_reached(so, cend[1], k_end, tlen)) fb_end = true;
                    r_ak = max(0, so.akM);
                }
                bool last_forward = false;
                bool rev_pending = false;  // reverse wavefront score_r+1 already sits in slot next_slot(cur_slot[1])
                const int max_antidiagonal = plen + tlen - 1;
                bool rev_pending_done = false;  // END_REACHED flag of that speculative wavefront
                lap(4);
                // =========== shared-memory diagonal-band engine for phase 1 ===========
                // Advances both directions BAND_T scores at a time inside shared memory (int16 offsets,
                // NULL-filled halos, no range checks, one barrier per step), streams every produced row to
                // an int16 ring in global memory, then replays WFA2's alternation over the per-step table
                // to find the exact break.  Anything it cannot resolve locally (an out-of-bounds positive
                // I/D offset, whose trimming depends on the whole row) aborts the band: the generic loop
                // below then redoes phase 1 from score 0.  Hand-over converts the last `scope` rows to the
                // generic int32 ring so that phase 2 (overlap) runs on the proven path.
                bool band_done = false;
                if constexpr (BAND)
                if (!fb_end && P.ring16_int_off >= 0 && scope <= BAND_MROWS && pen.e1 <= 2 && (!TWO || pen.e2 <= 1) && sp.rem > BAND_MIN_SCORE &&
                    2 * tlen + plen < 32000 && 2 * plen + tlen < 32000) {
                    short* const ring16 = reinterpret_cast<short*>(ws_i + P.ring16_int_off);
                    auto r16_off = [&](int d, int slot, int c) -> long long { return ((long long)((d * rn16 + slot) * NCOMP + comp_idx(c))) * W + koff; };
                    // row 0 of both directions (written by init_dir) -> int16 ring + meta
                    if (tid < 2) {
                        const int d = tid;
                        ring16[r16_off(d, 0, cbeg[d])] = (short)ws[ring_off(d, 0, cbeg[d])];
                    }
                    for (int d = 0; d < 2; ++d) {
                        const SlotMeta& m = ring_meta[d * ring_n + 0];
                        Meta16& q = meta16[d * rn16 + 0];
#pragma unroll
                        for (int c = 0; c < 5; ++c) {
                            q.lo[c] = m.lo[c];
                            q.hi[c] = m.hi[c];
                        }
                        q.akM = m.akM;
                        q.akAll = m.akAll;
                    }
                    if (tid == 0) s_band_flag = 0;
                    cta_sync<NT>();
                    int S = 0;        // both directions are complete up to score S
                    int slotS = 0;    // int16 ring slot of score S
                    bool band_fail = false, band_break = false;
                    // smem row of a score (negative score -> the all-NULL row)
                    auto row_of = [&](int base, int mod, int cur, int back, int s) -> int {
                        if (s - back < 0) return BAND_NULL_ROW;
                        int r = cur - back;
                        if (r < 0) r += mod;
                        return base + r;
                    };
                    // all-NULL row
                    for (int i = tid; i < BAND_WP; i += NT) band_rows[BAND_NULL_ROW * BAND_WP + i] = NULL16;
                    while (!band_fail && !band_break && !fb_end) {
                        if (f_ak + r_ak >= max_antidiagonal) {
                            band_break = true;
                            break;
                        }
                        // ---- per-step table ----
                        for (int i = tid; i < 2 * (BAND_T + 1) * BAND_TAB; i += NT) band_tab[i] = INT_MIN;
                        cta_sync<NT>();
                        const int m27 = S % BAND_MROWS, m3 = S % 3, m2 = S % 2;
                        for (int d = 0; d < 2; ++d) {
                            const SeqView& sv = (d == 0) ? svd[0] : svd[1];
                            const int ce = (d == 0) ? cend[0] : cend[1];
                            // span of the history rows, widened by the band length
                            int mn = INT_MAX, mxk = INT_MIN;
                            for (int j = 0; j < scope && S - j >= 0; ++j) {
                                int sl = slotS - j;
                                if (sl < 0) sl += rn16;
                                const Meta16& q = meta16[d * rn16 + sl];
#pragma unroll
                                for (int c = 0; c < 5; ++c)
                                    if (q.lo[c] <= q.hi[c]) {
                                        mn = min(mn, q.lo[c]);
                                        mxk = max(mxk, q.hi[c]);
                                    }
                            }
                            if (mn > mxk) {  // every history row is empty: cannot happen in exact mode
                                band_fail = true;
                                break;
                            }
                            const int span_lo = mn - BAND_T, span_hi = mxk + BAND_T;
                            for (int ka = span_lo; ka <= span_hi; ka += BAND_WT) {
                                const int kb = min(ka + BAND_WT - 1, span_hi);
                                const int base_k = ka - BAND_T;  // diagonal of shared-memory column 0
                                // ---- load the history window (NULL outside each row's trimmed range) ----
                                auto load_row = [&](int row, int c, int score) {
                                    short* dst = band_rows + row * BAND_WP;
                                    int lo_r = 1, hi_r = 0;
                                    const short* src = ring16;
                                    if (score >= 0) {
                                        int sl = slotS - (S - score);
                                        if (sl < 0) sl += rn16;
                                        const Meta16& q = meta16[d * rn16 + sl];
                                        lo_r = q.lo[c];
                                        hi_r = q.hi[c];
                                        src = ring16 + r16_off(d, sl, c);
                                    }
                                    for (int i = tid; i < BAND_WP; i += NT) {
                                        const int k = base_k + i;
                                        dst[i] = (k >= lo_r && k <= hi_r) ? src[k] : NULL16;
                                    }
                                };
                                for (int j = 0; j < BAND_MROWS - 1; ++j) {
                                    int r = m27 - j;
                                    if (r < 0) r += BAND_MROWS;
                                    load_row(r, AW_COMP_M, S - j);
                                }
                                for (int j = 0; j < 2; ++j) {
                                    int r = m3 - j;
                                    if (r < 0) r += 3;
                                    load_row(BAND_MROWS + r, AW_COMP_I1, S - j);
                                    load_row(BAND_MROWS + 3 + r, AW_COMP_D1, S - j);
                                }
                                if (TWO) {
                                    load_row(BAND_MROWS + 6 + m2, AW_COMP_I2, S);
                                    load_row(BAND_MROWS + 8 + m2, AW_COMP_D2, S);
                                }
                                cta_sync<NT>();
                                // ---- BAND_T steps inside shared memory ----
                                int c27 = m27, c3 = m3, c2 = m2, sl16 = slotS;
                                for (int t = 1; t <= BAND_T; ++t) {
                                    const int s = S + t;
                                    c27 = (c27 + 1 == BAND_MROWS) ? 0 : c27 + 1;
                                    c3 = (c3 + 1 == 3) ? 0 : c3 + 1;
                                    c2 ^= 1;
                                    sl16 = (sl16 + 1 == rn16) ? 0 : sl16 + 1;
                                    const short* r_mx = band_rows + row_of(0, BAND_MROWS, c27, pen.x, s) * BAND_WP;
                                    const short* r_mo1 = band_rows + row_of(0, BAND_MROWS, c27, pen.o1 + pen.e1, s) * BAND_WP;
                                    const short* r_i1 = band_rows + row_of(BAND_MROWS, 3, c3, pen.e1, s) * BAND_WP;
                                    const short* r_d1 = band_rows + row_of(BAND_MROWS + 3, 3, c3, pen.e1, s) * BAND_WP;
                                    const short* r_mo2 = band_rows + (TWO ? row_of(0, BAND_MROWS, c27, pen.o2 + pen.e2, s) : BAND_NULL_ROW) * BAND_WP;
                                    const short* r_i2 = band_rows + (TWO ? row_of(BAND_MROWS + 6, 2, c2, pen.e2, s) : BAND_NULL_ROW) * BAND_WP;
                                    const short* r_d2 = band_rows + (TWO ? row_of(BAND_MROWS + 8, 2, c2, pen.e2, s) : BAND_NULL_ROW) * BAND_WP;
                                    short* w_m = band_rows + c27 * BAND_WP;
                                    short* w_i1 = band_rows + (BAND_MROWS + c3) * BAND_WP;
                                    short* w_d1 = band_rows + (BAND_MROWS + 3 + c3) * BAND_WP;
                                    short* w_i2 = band_rows + (BAND_MROWS + 6 + c2) * BAND_WP;
                                    short* w_d2 = band_rows + (BAND_MROWS + 8 + c2) * BAND_WP;
                                    short* g_m = ring16 + r16_off(d, sl16, AW_COMP_M);
                                    short* g_i1 = ring16 + r16_off(d, sl16, AW_COMP_I1);
                                    short* g_d1 = ring16 + r16_off(d, sl16, AW_COMP_D1);
                                    short* g_i2 = ring16 + r16_off(d, sl16, TWO ? AW_COMP_I2 : AW_COMP_M);
                                    short* g_d2 = ring16 + r16_off(d, sl16, TWO ? AW_COMP_D2 : AW_COMP_M);
                                    int lo_c[5], hi_c[5];
#pragma unroll
                                    for (int c = 0; c < 5; ++c) lo_c[c] = hi_c[c] = INT_MIN;
                                    int akM = INT_MIN, akAll = INT_MIN, endval = INT_MIN;
                                    bool oob = false;
                                    const unsigned utl = (unsigned)tlen, upl = (unsigned)plen;
                                    // only diagonals the row can occupy at this step: [mn - t, mxk + t], inside the trapezoid
                                    const int i_lo = max(t, mn - t - base_k), i_hi = min(BAND_WP - 1 - t, mxk + t - base_k);
                                    // cells just outside the clipped range still hold an older row: later steps of this
                                    // band read up to (BAND_T - t) + 1 columns beyond it, so NULL that margin
                                    {
                                        const int mg = BAND_T - t + 1;
                                        if (tid < 2 * mg) {
                                            const int i = (tid < mg) ? (i_lo - 1 - tid) : (i_hi + 1 + (tid - mg));
                                            if (i >= 0 && i < BAND_WP) {
                                                w_m[i] = NULL16;
                                                w_i1[i] = NULL16;
                                                w_d1[i] = NULL16;
                                                if (TWO) {
                                                    w_i2[i] = NULL16;
                                                    w_d2[i] = NULL16;
                                                }
                                            }
                                        }
                                    }
                                    constexpr int J = 4;  // cells per thread advanced in lockstep (independent chains)
                                    for (int i0 = i_lo + tid; i0 <= i_hi; i0 += J * NT) {
                                        int ii[J], kk[J], vm[J], vpre[J], vi1[J], vd1[J], vi2[J], vd2[J], ext[J];
                                        bool act[J], more[J];
#pragma unroll
                                        for (int j = 0; j < J; ++j) {
                                            act[j] = (i0 + j * NT) <= i_hi;
                                            ii[j] = act[j] ? i0 + j * NT : i_lo;
                                            kk[j] = base_k + ii[j];
                                        }
#pragma unroll
                                        for (int j = 0; j < J; ++j) {
                                            const int i = ii[j];
                                            int i1 = max((int)r_mo1[i - 1], (int)r_i1[i - 1]) + 1;
                                            const int d1 = max((int)r_mo1[i + 1], (int)r_d1[i + 1]);
                                            int i2 = NULL16, d2 = NULL16, ins = i1, del = d1;
                                            if (TWO) {
                                                i2 = max((int)r_mo2[i - 1], (int)r_i2[i - 1]) + 1;
                                                d2 = max((int)r_mo2[i + 1], (int)r_d2[i + 1]);
                                                ins = max(i1, i2);
                                                del = max(d1, d2);
                                            }
                                            int m = max(del, max((int)r_mx[i] + 1, ins));
                                            vpre[j] = m;
                                            if ((unsigned)m > utl || (unsigned)(m - kk[j]) > upl) m = NULL16;
                                            vm[j] = m;
                                            vi1[j] = (i1 < 0) ? (int)NULL16 : i1;  // all negative offsets are equivalent: stop the +1 drift
                                            vd1[j] = d1;
                                            vi2[j] = (i2 < 0) ? (int)NULL16 : i2;
                                            vd2[j] = d2;
                                        }
#pragma unroll
                                        for (int j = 0; j < J; ++j) {  // first extend round, branch-free
                                            const bool valid = vm[j] >= 0;
                                            extend_first<BITS>(sv, valid ? kk[j] : 0, valid ? vm[j] : 0, ext[j], more[j]);
                                            more[j] = more[j] && valid;
                                            if (valid) vm[j] += ext[j];
                                        }
#pragma unroll
                                        for (int j = 0; j < J; ++j)  // long match runs (rare)
                                            if (more[j]) vm[j] = extend_rest<BITS>(sv, kk[j], vm[j]);
#pragma unroll
                                        for (int j = 0; j < J; ++j) {
                                            if (!act[j]) continue;
                                            const int i = ii[j], k = kk[j], m = vm[j], i1 = vi1[j], d1 = vd1[j], i2 = vi2[j], d2 = vd2[j];
                                            w_m[i] = (short)m;
                                            w_i1[i] = (short)i1;
                                            w_d1[i] = (short)d1;
                                            if (TWO) {
                                                w_i2[i] = (short)i2;
                                                w_d2[i] = (short)d2;
                                            }
                                            if (k >= ka && k <= kb) {  // cells this tile owns
                                                if (k >= kmin_alloc && k <= kmax_alloc) {
                                                    g_m[k] = (short)m;
                                                    g_i1[k] = (short)i1;
                                                    g_d1[k] = (short)d1;
                                                    if (TWO) {
                                                        g_i2[k] = (short)i2;
                                                        g_d2[k] = (short)d2;
                                                    }
                                                } else if (m >= 0 || i1 >= 0 || d1 >= 0 || i2 >= 0 || d2 >= 0) {
                                                    oob = true;  // valid data outside the allocated diagonals
                                                }
                                                if (vpre[j] >= 0) akAll = max(akAll, 2 * vpre[j] - k);
                                                if (m >= 0) {
                                                    if (lo_c[0] == INT_MIN) lo_c[0] = -k;
                                                    hi_c[0] = k;
                                                    akM = max(akM, 2 * m - k);
                                                }
                                                auto track = [&](int c, int v) {
                                                    if (v >= 0) {
                                                        if ((unsigned)v > utl || (unsigned)(v - k) > upl) oob = true;
                                                        if (lo_c[c] == INT_MIN) lo_c[c] = -k;
                                                        hi_c[c] = k;
                                                    }
                                                };
                                                track(AW_COMP_I1, i1);
                                                track(AW_COMP_D1, d1);
                                                if (TWO) {
                                                    track(AW_COMP_I2, i2);
                                                    track(AW_COMP_D2, d2);
                                                }
                                                if (k == k_end) endval = (ce == AW_COMP_M) ? m : (ce == AW_COMP_I1) ? i1 : (ce == AW_COMP_D1) ? d1 : (ce == AW_COMP_I2) ? i2 : d2;
                                            }
                                        }
                                    }
                                    int* tab = band_tab + (d * (BAND_T + 1) + t) * BAND_TAB;
#pragma unroll
                                    for (int c = 0; c < 5; ++c) {
                                        if (!TWO && (c == AW_COMP_I2 || c == AW_COMP_D2)) continue;
                                        red_max<256>(tab, RED_HI + c, hi_c[c]);
                                        red_max<256>(tab, RED_LO + c, lo_c[c]);
                                    }
                                    red_max<256>(tab, RED_AKM, akM);
                                    red_max<256>(tab, RED_AKALL, akAll);
                                    red_max<256>(tab, RED_END, endval);
                                    if (__any_sync(0xffffffffu, oob) && (tid & 31) == 0) s_band_flag = 1;
                                    cta_sync<NT>();
                                }
                                w_cells += (unsigned long long)BAND_T * (unsigned long long)(kb - ka + 1) * NCOMP;
                            }
                            w_steps += BAND_T;
                        }
                        cta_sync<NT>();
                        if (band_fail || s_band_flag) {
                            band_fail = true;
                            break;
                        }
                        // ---- replay WFA2's alternation over the band's steps; publish the rows' metadata ----
                        int sl = slotS;
                        for (int t = 1; t <= BAND_T && !band_break && !fb_end; ++t) {
                            sl = (sl + 1 == rn16) ? 0 : sl + 1;
                            bool done_d[2];
                            int ak_d[2];
                            for (int d = 0; d < 2; ++d) {
                                const int* tab = band_tab + (d * (BAND_T + 1) + t) * BAND_TAB;
                                Meta16& q = meta16[d * rn16 + sl];
#pragma unroll
                                for (int c = 0; c < 5; ++c) {
                                    const int h = tab[RED_HI + c], l = tab[RED_LO + c];
                                    q.lo[c] = (h == INT_MIN) ? 1 : -l;
                                    q.hi[c] = (h == INT_MIN) ? 0 : h;
                                }
                                q.akM = tab[RED_AKM];
                                q.akAll = tab[RED_AKALL];
                                done_d[d] = tab[RED_END] >= tlen;
                                ak_d[d] = q.akM;
                            }
                            // forward step S+t
                            score_f = S + t;
                            f_ak = max(f_ak, max(0, done_d[0] ? 0 : ak_d[0]));
                            last_forward = true;
                            if (AW_BIALIGN_PHASE1_END_REACHED_RETURNS && done_d[0]) {
                                fb_end = true;
                                break;
                            }
                            if (f_ak + r_ak >= max_antidiagonal) {
                                band_break = true;  // reverse stays at S+t-1
                                break;
                            }
                            // reverse step S+t
                            score_r = S + t;
                            r_ak = max(r_ak, max(0, done_d[1] ? 0 : ak_d[1]));
                            last_forward = false;
                            if (AW_BIALIGN_PHASE1_END_REACHED_RETURNS && done_d[1]) {
                                fb_end = true;
                                break;
                            }
                            if (f_ak + r_ak >= max_antidiagonal) band_break = true;
                        }
                        if (!band_break && !fb_end) {
                            S += BAND_T;
                            slotS += BAND_T;
                            if (slotS >= rn16) slotS -= rn16;
                        }
                        cta_sync<NT>();
                    }
                    if (band_fail) {
                        cyc[5] += 1;
                        // redo phase 1 on the generic path from score 0 (rows 0 are still in the int32 ring)
                        score_f = score_r = 0;
                        f_ak = max(0, ring_meta[0 * ring_n + 0].akM);
                        r_ak = max(0, ring_meta[1 * ring_n + 0].akM);
                        last_forward = false;
                        cur_slot[0] = cur_slot[1] = 0;
                    } else {
                        band_done = true;
                        if (!fb_end) {
                            // ---- hand-over: last `scope` rows of each direction -> generic int32 ring ----
                            for (int d = 0; d < 2; ++d) {
                                const int sd = (d == 0) ? score_f : score_r;
                                int sl_top = slotS + (sd - S);
                                if (sl_top >= rn16) sl_top -= rn16;
                                for (int j = 0; j < scope && sd - j >= 0; ++j) {
                                    int sl16 = sl_top - j;
                                    if (sl16 < 0) sl16 += rn16;
                                    const int sl32 = (sd - j) % ring_n;
                                    const Meta16& q = meta16[d * rn16 + sl16];
                                    SlotMeta& mt = ring_meta[d * ring_n + sl32];
#pragma unroll
                                    for (int c = 0; c < 5; ++c) {
                                        mt.lo[c] = q.lo[c];
                                        mt.hi[c] = q.hi[c];
                                        if (!TWO && (c == AW_COMP_I2 || c == AW_COMP_D2)) continue;
                                        const short* src = ring16 + r16_off(d, sl16, c);
                                        WS* dst = ws + ring_off(d, sl32, c);
                                        for (int k = q.lo[c] + tid; k <= q.hi[c]; k += NT) {
                                            const int v = src[k];
                                            dst[k] = to_ws<WS>(v < 0 ? AW_NULLV : v);
                                        }
                                    }
                                    mt.akM = q.akM;
                                    mt.akAll = q.akAll;
                                }
                                cur_slot[d] = sd % ring_n;
                            }
                            cta_sync<NT>();
                        }
                    }
                }
                // ---- phase 1: forward step s_f+1 and (speculative) reverse step s_r+1 share one barrier ----
                while (!band_done && !fb_end && status == ST_OK) {
                    if (f_ak + r_ak >= max_antidiagonal) break;
                    const int slot_f = next_slot(cur_slot[0]), slot_r = next_slot(cur_slot[1]);
                    const Range rf = launch_dir(0, score_f + 1, slot_f);
                    const Range rr = launch_dir(1, score_r + 1, slot_r);
                    cta_sync<NT>();
                    bool done = finish_dir(0, slot_f, rf);
                    const int akM_f = ring_meta[0 * ring_n + slot_f].akM;
                    const bool done_r = finish_dir(1, slot_r, rr);
                    const int akM_r = ring_meta[1 * ring_n + slot_r].akM;
                    rotate_red();
                    if (status != ST_OK) break;
                    // commit forward
                    ++score_f;
                    cur_slot[0] = slot_f;
                    f_ak = max(f_ak, max(0, done ? 0 : akM_f));
                    last_forward = true;
                    if (AW_BIALIGN_PHASE1_END_REACHED_RETURNS && done) {
                        fb_end = true;
                        break;
                    }
                    if (f_ak + r_ak >= max_antidiagonal) {
                        rev_pending = true;
                        rev_pending_done = done_r;
                        break;
                    }
                    // commit reverse
                    ++score_r;
                    cur_slot[1] = slot_r;
                    r_ak = max(r_ak, max(0, done_r ? 0 : akM_r));
                    last_forward = false;
                    if (AW_BIALIGN_PHASE1_END_REACHED_RETURNS && done_r) {
                        fb_end = true;
                        break;
                    }
                }
                lap(0);
                // one committed step of direction d (phase 2): cells, barrier, finish
                auto step_dir = [&](int d, int s) -> bool {
                    const int slot = next_slot(cur_slot[d]);
                    const Range r = launch_dir(d, s, slot);
                    cta_sync<NT>();
                    const bool done = finish_dir(d, slot, r);
                    rotate_red();
                    cur_slot[d] = slot;
                    return done;
                };
                // ---- phase 2: advance until no better breakpoint is possible ----
                const int gap_opening = TWO ? pen.o2 : pen.o1;
                while (!fb_end && status == ST_OK) {
                    if (last_forward) {
                        const int min_r = (score_r > scope - 1) ? score_r - (scope - 1) : 0;
                        if (score_f + min_r - gap_opening >= bp.score) break;
                        overlap(0, 1, score_f, score_r);
                        ++score_r;
                        bool done;
                        if (rev_pending) {  // computed speculatively in phase 1
                            rev_pending = false;
                            cur_slot[1] = next_slot(cur_slot[1]);
                            done = rev_pending_done;
                        } else {
                            done = step_dir(1, score_r);
                        }
                        if (AW_BIALIGN_PHASE2_END_REACHED_RETURNS && done) {
                            fb_end = true;
                            break;
                        }
                        if (status != ST_OK) break;
                    }
                    const int min_f = (score_f > scope - 1) ? score_f - (scope - 1) : 0;
                    if (min_f + score_r - gap_opening >= bp.score) break;
                    overlap(1, 0, score_r, score_f);
                    ++score_f;
                    const bool done = step_dir(0, score_f);
                    if (AW_BIALIGN_PHASE2_END_REACHED_RETURNS && done) {
                        fb_end = true;
                        break;
                    }
                    last_forward = true;
                }
                lap(1);
                if (status != ST_OK) break;
                if (fb_end) {
                    do_base = true;
                } else {
                    ++w_bps;
                    const int bh = bp.off_f, bv = bp.off_f - bp.k_f;
                    if (bv < 0 || bv > plen || bh < 0 || bh > tlen || sp_n + 2 > MAX_STACK) {
                        status = ST_FAIL_WORKSPACE;
                        break;
                    }
                    // right half is pushed first so that the left half is aligned (and emitted) first
                    SubProblem right = {sp.pb + bv, sp.pe, sp.tb + bh, sp.te, bp.comp, sp.ce, bp.score_r};
                    SubProblem left = {sp.pb, sp.pb + bv, sp.tb, sp.tb + bh, sp.cb, bp.comp, bp.score_f};
                    stack[sp_n++] = right;
                    stack[sp_n++] = left;
                    cta_sync<NT>();
                    continue;
                }
            }

            // =========== K7: wavefront_bialign_base: full-history WFA + backtrace ===========
            {
                ++w_base;
                lap(5);
                const SeqView sv = SeqView{pw, tw, sp.pb, sp.tb, plen, tlen, false};
                long long hist_used = 0;
                // component block of a history wavefront: element (c,k) at ws[hist_off(off,width,clo,c) + k]
                auto hist_off = [&](int off, int width, int clo, int c) -> int { return hist_base + off + comp_idx(c) * width - clo; };
                auto write_hist_meta = [&](int s, const SlotMeta& m) {
                    if (tid == 0) {
                        int* g = hist_meta + (size_t)s * HIST_META_INTS;
#pragma unroll
                        for (int c = 0; c < 5; ++c) {
                            g[c] = m.lo[c];
                            g[5 + c] = m.hi[c];
                        }
                        g[10] = m.clo;
                        g[11] = m.width;
                        g[12] = m.off;
                    }
                };
                auto slot_back = [&](int slot, int back) -> int {
                    const int s = slot - back;
                    return s < 0 ? s + ring_n : s;
                };
                StepOut so;
                int score = 0, slot = 0;
                bool done;
                {  // score 0
                    int* r = red[0][red_i];
                    SlotMeta& mt = ring_meta[0];
                    if (tid == 0) {
                        int m = 0;
                        if (sp.cb == AW_COMP_M) m = extend_cell<BITS>(sv, 0, 0);
                        ws[hist_off(0, 1, 0, sp.cb)] = to_ws<WS>(m);
                        if (sp.cb == AW_COMP_M && sp.ce == AW_COMP_M && k_end == 0) r[RED_END] = m;
                    }
                    cta_sync<NT>();
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        so.lo[c] = (c == sp.cb) ? 0 : 1;
                        so.hi[c] = 0;
                    }
                    so.akM = so.akAll = INT_MIN;
                    so.endval = r[RED_END];
                    so.ambiguous = false;
                    store_meta(mt, so);
                    mt.clo = 0;
                    mt.width = 1;
                    mt.off = 0;
                    hist_used = NCOMP;
                    write_hist_meta(0, mt);
                    rotate_red();
                    done = end_reached(so, sp.ce, k_end, tlen);
                }
                while (!done) {
                    ++score;
                    slot = (slot + 1 == ring_n) ? 0 : slot + 1;
                    if (score >= P.hist_max_scores) {
                        status = ST_FAIL_WORKSPACE;
                        break;
                    }
                    In in[7];
                    auto fetch = [&](int c, int back) -> In {
                        In w;
                        if (score - back < 0) {
                            w.off = 0;
                            w.lo = 1;
                            w.hi = 0;
                            return w;
                        }
                        const SlotMeta& m = ring_meta[slot_back(slot, back)];
                        w.off = hist_off(m.off, m.width, m.clo, c);
                        w.lo = m.lo[c];
                        w.hi = m.hi[c];
                        return w;
                    };
                    in[IN_MX] = fetch(AW_COMP_M, pen.x);
                    in[IN_MO1] = fetch(AW_COMP_M, pen.o1 + pen.e1);
                    in[IN_I1E] = fetch(AW_COMP_I1, pen.e1);
                    in[IN_D1E] = fetch(AW_COMP_D1, pen.e1);
                    if (TWO) {
                        in[IN_MO2] = fetch(AW_COMP_M, pen.o2 + pen.e2);
                        in[IN_I2E] = fetch(AW_COMP_I2, pen.e2);
                        in[IN_D2E] = fetch(AW_COMP_D2, pen.e2);
                    } else {
                        in[IN_MO2].off = in[IN_I2E].off = in[IN_D2E].off = 0;
                        in[IN_MO2].lo = in[IN_I2E].lo = in[IN_D2E].lo = 1;
                        in[IN_MO2].hi = in[IN_I2E].hi = in[IN_D2E].hi = 0;
                    }
                    int lo, hi;
                    limits(in, lo, hi);
                    SlotMeta& mt = ring_meta[slot];
                    ++w_steps;
                    if (lo > hi) {
                        set_empty(mt);
                        write_hist_meta(score, mt);
                        continue;
                    }
                    const int width = hi - lo + 1;
                    if (hist_used + (long long)NCOMP * width > (long long)P.hist_ints) {
                        status = ST_FAIL_WORKSPACE;
                        break;
                    }
                    const int off = (int)hist_used;
                    hist_used += (long long)NCOMP * width;
                    int out[5];
#pragma unroll
                    for (int c = 0; c < 5; ++c) out[c] = hist_off(off, width, lo, (TWO || c == AW_COMP_M || c == AW_COMP_I1 || c == AW_COMP_D1) ? c : AW_COMP_M);
                    wf_cells<NT, BITS, TWO, WS>(ws, in, out, lo, hi, sv, k_end, sp.ce, red[0][red_i]);
                    cta_sync<NT>();
                    wf_finish<TWO>(red[0][red_i], lo, hi, so);
                    if (so.ambiguous) wf_rescan<NT, TWO, WS>(ws, out, lo, hi, plen, tlen, red[0][red_i], so);
                    w_cells += (unsigned long long)width * NCOMP;
                    store_meta(mt, so);
                    mt.clo = lo;
                    mt.width = width;
                    mt.off = off;
                    write_hist_meta(score, mt);
                    rotate_red();
                    done = end_reached(so, sp.ce, k_end, tlen);
                }
                if (status != ST_OK) break;
                w_maxbase = max(w_maxbase, (unsigned)score);
                cta_sync<NT>();  // history + meta visible to warp 0
                lap(2);

                // ---- wavefront_backtrace_affine by warp 0: lanes evaluate the candidates ----
                unsigned n_leaf = 0;  // runs pushed (reverse order) into leaf_runs; uniform within warp 0
                if (tid < 32) {
                    const int lane = tid;
                    unsigned last_run = 0;  // pending run (merged before being stored)
                    auto push = [&](unsigned op, unsigned len) {
                        if (len == 0) return;
                        if (last_run != 0 && (last_run & 3u) == op) {
                            last_run += len << 2;
                        } else {
                            if (last_run != 0) {
                                if (lane == 0 && n_leaf < P.runs_cap) leaf_runs[n_leaf] = last_run;
                                ++n_leaf;
                            }
                            last_run = (len << 2) | op;
                        }
                    };
                    int type = sp.ce, sc = score, k = k_end, offset = tlen;
                    int v = plen, h = tlen;
                    // lane l evaluates backtrace candidate type l+1 (AW_BT_*)
                    while (v > 0 && h > 0 && sc > 0) {
                        const int bt = lane + 1;
                        int cand_v = INT_MIN;
                        if (lane < 9) {
                            int comp_src, cost, dk, add;
                            bool active;
                            switch (bt) {
                                case AW_BT_M: comp_src = AW_COMP_M; cost = pen.x; dk = 0; add = 1; active = (type == AW_COMP_M); break;
                                case AW_BT_I1_OPEN: comp_src = AW_COMP_M; cost = pen.o1 + pen.e1; dk = -1; add = 1; active = (type == AW_COMP_M || type == AW_COMP_I1); break;
                                case AW_BT_I1_EXT: comp_src = AW_COMP_I1; cost = pen.e1; dk = -1; add = 1; active = (type == AW_COMP_M || type == AW_COMP_I1); break;
                                case AW_BT_I2_OPEN: comp_src = AW_COMP_M; cost = pen.o2 + pen.e2; dk = -1; add = 1; active = TWO && (type == AW_COMP_M || type == AW_COMP_I2); break;
                                case AW_BT_I2_EXT: comp_src = AW_COMP_I2; cost = pen.e2; dk = -1; add = 1; active = TWO && (type == AW_COMP_M || type == AW_COMP_I2); break;
                                case AW_BT_D1_OPEN: comp_src = AW_COMP_M; cost = pen.o1 + pen.e1; dk = 1; add = 0; active = (type == AW_COMP_M || type == AW_COMP_D1); break;
                                case AW_BT_D1_EXT: comp_src = AW_COMP_D1; cost = pen.e1; dk = 1; add = 0; active = (type == AW_COMP_M || type == AW_COMP_D1); break;
                                case AW_BT_D2_OPEN: comp_src = AW_COMP_M; cost = pen.o2 + pen.e2; dk = 1; add = 0; active = TWO && (type == AW_COMP_M || type == AW_COMP_D2); break;
                                default: comp_src = AW_COMP_D2; cost = pen.e2; dk = 1; add = 0; active = TWO && (type == AW_COMP_M || type == AW_COMP_D2); break;
                            }
                            const int ss = sc - cost;
                            if (active && ss >= 0) {
                                const int* g = hist_meta + (size_t)ss * HIST_META_INTS;
                                const int kk = k + dk;
                                if (g[comp_src] <= kk && kk <= g[5 + comp_src]) {
                                    const int val = ws[hist_off(g[12], g[11], g[10], comp_src) + kk];
                                    if (val >= 0) cand_v = ((val + add) << AW_BT_TYPE_BITS) | bt;
                                }
                            }
                        }
                        const int max_all = __reduce_max_sync(0xffffffffu, cand_v);
                        if (max_all == INT_MIN) {  // cannot happen on a valid path
                            status = ST_FAIL_WORKSPACE;
                            break;
                        }
                        if (type == AW_COMP_M) {
                            const int max_offset = max_all >> AW_BT_TYPE_BITS;
                            push(AW_OP_M, (unsigned)max(0, offset - max_offset));
                            offset = max_offset;
                            v = offset - k;
                            h = offset;
                            if (v <= 0 || h <= 0) break;
                        }
                        const int b = max_all & 0xF;
                        switch (b) {
                            case AW_BT_M: sc -= pen.x; type = AW_COMP_M; break;
                            case AW_BT_I1_OPEN: sc -= pen.o1 + pen.e1; type = AW_COMP_M; break;
                            case AW_BT_I1_EXT: sc -= pen.e1; type = AW_COMP_I1; break;
                            case AW_BT_I2_OPEN: sc -= pen.o2 + pen.e2; type = AW_COMP_M; break;
                            case AW_BT_I2_EXT: sc -= pen.e2; type = AW_COMP_I2; break;
                            case AW_BT_D1_OPEN: sc -= pen.o1 + pen.e1; type = AW_COMP_M; break;
                            case AW_BT_D1_EXT: sc -= pen.e1; type = AW_COMP_D1; break;
                            case AW_BT_D2_OPEN: sc -= pen.o2 + pen.e2; type = AW_COMP_M; break;
                            default: sc -= pen.e2; type = AW_COMP_D2; break;
                        }
                        if (b == AW_BT_M) {
                            push(AW_OP_X, 1);
                            --offset;
                        } else if (b <= AW_BT_I2_EXT) {
                            push(AW_OP_I, 1);
                            --k;
                            --offset;
                        } else {
                            push(AW_OP_D, 1);
                            ++k;
                        }
                        v = offset - k;
                        h = offset;
                    }
                    if (v > 0 && h > 0) {
                        const int nm = min(v, h);
                        push(AW_OP_M, (unsigned)nm);
                        v -= nm;
                        h -= nm;
                    }
                    if (v > 0) push(AW_OP_D, (unsigned)v);
                    if (h > 0) push(AW_OP_I, (unsigned)h);
                    if (last_run != 0) {
                        if (lane == 0 && n_leaf < P.runs_cap) leaf_runs[n_leaf] = last_run;
                        ++n_leaf;
                    }
                    if (lane == 0) {
                        s_acc[0] = n_leaf;
                        s_acc[1] = (unsigned long long)status;
                    }
                }
                cta_sync<NT>();
                n_leaf = (unsigned)s_acc[0];
                if (s_acc[1] != ST_OK) status = (int)s_acc[1];
                if (status != ST_OK) break;
                // append the leaf's runs (stored back to front) to the pair's CIGAR
                unsigned base_n = s_nruns, skip = 0;
                if (n_leaf > 0 && base_n > 0 && (pair_runs[base_n - 1] & 3u) == (leaf_runs[n_leaf - 1] & 3u)) skip = 1;
                cta_sync<NT>();
                if (base_n + n_leaf > P.runs_cap) {
                    status = ST_FAIL_WORKSPACE;
                    break;
                }
                if (tid == 0 && skip) pair_runs[base_n - 1] += leaf_runs[n_leaf - 1] & ~3u;
                for (unsigned i = skip + tid; i < n_leaf; i += NT) pair_runs[base_n + i - skip] = leaf_runs[n_leaf - 1 - i];
                if (tid == 0) s_nruns = base_n + n_leaf - skip;
                cta_sync<NT>();
                lap(3);
            }
        }  // DFS over sub-problems

        // =========== K8: statistics, score, PAF text ===========
        cta_sync<NT>();
        const unsigned nruns = (status == ST_OK) ? s_nruns : 0;
        if (tid < 8) s_acc[tid] = 0;
        cta_sync<NT>();
        {
            unsigned long long cnt[4] = {0, 0, 0, 0}, textlen = 0, penalty = 0;
            for (unsigned r = tid; r < nruns; r += NT) {
                const unsigned run = pair_runs[r], op = run & 3u;
                const unsigned long long len = run >> 2;
                cnt[op] += len;
                textlen += ndigits(len) + 1;
                if (op == AW_OP_X) penalty += len * pen.x;
                else if (op != AW_OP_M) {
                    unsigned long long c1 = pen.o1 + len * pen.e1;
                    if (TWO) c1 = min(c1, pen.o2 + len * (unsigned long long)pen.e2);
                    penalty += c1;
                }
            }
            unsigned long long vals[6] = {cnt[0], cnt[1], cnt[2], cnt[3], textlen, penalty};
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                unsigned long long v = vals[i];
                for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                if ((tid & 31) == 0 && v) atomicAdd(&s_acc[i], v);
            }
        }
        cta_sync<NT>();
        const unsigned long long n_m = s_acc[0], n_x = s_acc[1], n_i = s_acc[2], n_d = s_acc[3], cg_len = s_acc[4], penalty = s_acc[5];
        const unsigned long long q_end = n_m + n_x + n_d, t_end = n_m + n_x + n_i;  // parse_cigar_lengths
        const unsigned long long block_len = max(q_end, t_end);
        const unsigned long long aln_len = n_m + n_x;                               // count_cigar_operations
        const uint32_t qid0 = P.id_off[pr.query_idx], qid1 = P.id_off[pr.query_idx + 1];
        const uint32_t tid0 = P.id_off[pr.target_idx], tid1 = P.id_off[pr.target_idx + 1];
        const bool want_paf = !(P.flags & AW_FLAG_NO_PAF);
        const bool count_only = (P.flags & AW_KFLAG_COUNT_ONLY) != 0;
        // header: q qlen qs qe strand t tlen ts te matches block 60 gi:f:x.xxxxxx cg:Z:
        const unsigned hdr_len = (qid1 - qid0) + 1 + ndigits(PLEN) + 1 + 1 + 1 + ndigits(q_end) + 1 + 1 + 1 + (tid1 - tid0) + 1 + ndigits(TLEN) + 1 + 1 + 1 +
                                 ndigits(t_end) + 1 + ndigits(n_m) + 1 + ndigits(block_len) + 1 + 2 + 1 + 5 + 8 + 1 + 5;
        const unsigned long long line_len = count_only ? 0 : (want_paf ? hdr_len + cg_len : cg_len);
        const unsigned long long nbytes = (!count_only && (P.flags & AW_FLAG_CIGAR_BYTES)) ? (n_m + n_x + n_i + n_d) : 0;
        if (tid == 0) {
            s_text_off = atomicAdd(P.text_cursor, line_len);
            s_bytes_off = nbytes ? atomicAdd(P.bytes_cursor, nbytes) : 0ull;
        }
        cta_sync<NT>();
        const unsigned long long text_off = s_text_off, bytes_off = s_bytes_off;
        if (status == ST_OK && (text_off + line_len > P.text_cap || bytes_off + nbytes > P.bytes_cap)) status = ST_FAIL_WORKSPACE;
        if (status == ST_OK && !count_only) {
            char* line = P.text + text_off;
            char* cg = line;
            if (want_paf) {
                cg = line + hdr_len;
                if (tid == 0) {
                    char* p = line;
                    for (uint32_t i = qid0; i < qid1; ++i) *p++ = P.ids[i];
                    *p++ = '\t';
                    p = put_u64(p, (unsigned long long)PLEN);
                    *p++ = '\t';
                    *p++ = '0';
                    *p++ = '\t';
                    p = put_u64(p, q_end);
                    *p++ = '\t';
                    *p++ = is_rev ? '-' : '+';
                    *p++ = '\t';
                    for (uint32_t i = tid0; i < tid1; ++i) *p++ = P.ids[i];
                    *p++ = '\t';
                    p = put_u64(p, (unsigned long long)TLEN);
                    *p++ = '\t';
                    *p++ = '0';
                    *p++ = '\t';
                    p = put_u64(p, t_end);
                    *p++ = '\t';
                    p = put_u64(p, n_m);
                    *p++ = '\t';
                    p = put_u64(p, block_len);
                    *p++ = '\t';
                    *p++ = '6';
                    *p++ = '0';
                    *p++ = '\t';
                    const char gi[5] = {'g', 'i', ':', 'f', ':'};
                    for (int i = 0; i < 5; ++i) *p++ = gi[i];
                    const unsigned micro = identity_micro(n_m, aln_len);
                    *p++ = (char)('0' + micro / 1000000u);
                    *p++ = '.';
                    unsigned fr = micro % 1000000u;
                    for (int i = 5; i >= 0; --i) {
                        p[i] = (char)('0' + fr % 10u);
                        fr /= 10u;
                    }
                    p += 6;
                    *p++ = '\t';
                    const char cz[5] = {'c', 'g', ':', 'Z', ':'};
                    for (int i = 0; i < 5; ++i) *p++ = cz[i];
                }
            }
            // cigar_bytes_to_string: contiguous chunk of runs per thread, offsets by block scan
            const unsigned chunk = (nruns + NT - 1) / NT;
            const unsigned r0 = min(nruns, tid * chunk), r1 = min(nruns, r0 + chunk);
            unsigned long long my_text = 0, my_cols = 0;
            for (unsigned r = r0; r < r1; ++r) {
                const unsigned run = pair_runs[r];
                my_text += ndigits(run >> 2) + 1;
                my_cols += run >> 2;
            }
            unsigned long long total;
            unsigned long long toff = block_excl_scan<NT>(my_text, scanbuf, &total);
            {
                char* p = cg + toff;
                for (unsigned r = r0; r < r1; ++r) {
                    const unsigned run = pair_runs[r], op = run & 3u;
                    p = put_u64(p, run >> 2);
                    *p++ = (op == AW_OP_M) ? '=' : (op == AW_OP_X) ? 'X' : (op == AW_OP_I) ? 'D' : 'I';
                }
            }
            if (nbytes) {
                unsigned long long boff = block_excl_scan<NT>(my_cols, scanbuf, &total);
                uint8_t* b = P.bytes + bytes_off + boff;
                for (unsigned r = r0; r < r1; ++r) {
                    const unsigned run = pair_runs[r], op = run & 3u, len = run >> 2;
                    const uint8_t ch = (op == AW_OP_M) ? 'M' : (op == AW_OP_X) ? 'X' : (op == AW_OP_I) ? 'I' : 'D';
                    for (unsigned i = 0; i < len; ++i) *b++ = ch;
                }
            }
        }
        lap(4);
        if (tid == 0) {
            AwPairOut o;
            o.status = (status == ST_OK) ? AW_OK : AW_EWORKSPACE;
            o.score = (status == ST_OK) ? -(int32_t)penalty : INT_MAX;
            o.is_reverse = is_rev;
            o.nruns = nruns;
            o.n_m = n_m;
            o.n_x = n_x;
            o.n_i = n_i;
            o.n_d = n_d;
            o.paf_off = text_off;
            o.paf_len = (uint32_t)line_len;
            o.cg_off = want_paf ? hdr_len : 0;
            o.bytes_off = bytes_off;
            o.runs_off = 0;
            o.cells = w_cells;
            o.steps = w_steps;
            o.n_breakpoints = w_bps;
            o.n_base = w_base;
            o.max_base_score = w_maxbase;
            for (int i = 0; i < 6; ++i) o.cyc[i] = cyc[i];
            P.out[pair_i] = o;
        }
        cta_sync<NT>();
    }
}

}  // namespace awk
