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

constexpr int NRED = 16;       // reduction slots: [0,5) max in-bounds k, [5,10) max -k, [10,15) max antidiagonal, [15] end value
constexpr int MAX_STACK = 96;  // DFS depth bound of the biWFA recursion
constexpr int HIST_META_INTS = 16;

enum { IN_MX = 0, IN_MO1, IN_I1E, IN_D1E, IN_MO2, IN_I2E, IN_D2E };
enum { ST_OK = 0, ST_END_REACHED = 1, ST_FAIL_WORKSPACE = 2 };

struct SlotMeta {
    int lo[5], hi[5], ak[5];
    int clo, width;     // computed (allocated) range of this wavefront (history mode)
    unsigned off;       // history arena offset (history mode)
};

struct In {
    const int* p;  // p[k] is the offset of diagonal k
    int lo, hi;    // trimmed range; empty iff lo > hi
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
    // per-CTA workspace
    int* ws_ring;            // [cta][2][scope][ncomp][W]
    unsigned long long ring_ints_per_cta;
    int W;                   // allocated diagonals per wavefront
    int* ws_hist;            // [cta][hist_ints]
    unsigned long long hist_ints_per_cta;
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
    const int idx = pos / SPW;  // pos >= 0
    const int sh = (pos % SPW) * BITS;
    return __funnelshift_r(__ldg(w + idx), __ldg(w + idx + 1), sh);
}
// symbols pos, pos-1, ... with `pos` in the most significant bits
template <int BITS>
__device__ __forceinline__ uint32_t load_rev(const uint32_t* __restrict__ w, int pos) {
    constexpr int SPW = 32 / BITS;
    const int idx = pos / SPW;
    const int sh = ((pos % SPW) + 1) * BITS;  // in [BITS, 32]
    return __funnelshift_rc(__ldg(w + idx - 1), __ldg(w + idx), sh);
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

__device__ __forceinline__ int ld_in(const In& w, int k) { return (k >= w.lo && k <= w.hi) ? w.p[k] : AW_NULLV; }

struct StepOut {
    int lo[5], hi[5], ak[5];
    int endval;
};

// ---- K4+K5: compute wavefront s from its inputs, extend M, trim, detect termination --------
// Restates wavefront_compute_affine2p_idm + wavefront_extend_matches_packed_end2end(_max) +
// wavefront_compute_trim_ends + wavefront_termination_end2end.  Exactly one
// CTA barrier; `red` must be pre-initialised to INT_MIN (triple-buffered by the caller).
template <int NT, int BITS, bool TWO>
__device__ __forceinline__ void wf_step(const In (&in)[7], int* const (&out)[5], int lo, int hi, const SeqView& sv, int k_end, int comp_end,
                                        int* red, StepOut& so) {
    const int tid = threadIdx.x;
    const unsigned tlen = (unsigned)sv.tlen, plen = (unsigned)sv.plen;
    int khi[5], klo[5], akm[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) khi[c] = klo[c] = akm[c] = INT_MIN;
    int endval = INT_MIN;
    for (int k = lo + tid; k <= hi; k += NT) {
        int vals[5];
        const int i1 = max(ld_in(in[IN_MO1], k - 1), ld_in(in[IN_I1E], k - 1)) + 1;
        const int d1 = max(ld_in(in[IN_MO1], k + 1), ld_in(in[IN_D1E], k + 1));
        int ins = i1, del = d1;
        vals[AW_COMP_I1] = i1;
        vals[AW_COMP_D1] = d1;
        if (TWO) {
            const int i2 = max(ld_in(in[IN_MO2], k - 1), ld_in(in[IN_I2E], k - 1)) + 1;
            const int d2 = max(ld_in(in[IN_MO2], k + 1), ld_in(in[IN_D2E], k + 1));
            vals[AW_COMP_I2] = i2;
            vals[AW_COMP_D2] = d2;
            ins = max(ins, i2);
            del = max(del, d2);
        } else {
            vals[AW_COMP_I2] = vals[AW_COMP_D2] = AW_NULLV;
        }
        const int mis = ld_in(in[IN_MX], k) + 1;
        int m = max(del, max(mis, ins));
        if ((unsigned)m > tlen || (unsigned)(m - k) > plen) m = AW_NULLV;
        if (m >= 0) m = extend_cell<BITS>(sv, k, m);
        vals[AW_COMP_M] = m;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            if (!TWO && (c == AW_COMP_I2 || c == AW_COMP_D2)) continue;
            const int v = vals[c];
            out[c][k] = v;
            if ((unsigned)v <= tlen && (unsigned)(v - k) <= plen) {  // in bounds: survives the trim
                if (klo[c] == INT_MIN) klo[c] = -k;
                khi[c] = k;
            }
            if (v >= 0) akm[c] = max(akm[c], 2 * v - k);
        }
        if (k == k_end) endval = vals[comp_end];
    }
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        if (!TWO && (c == AW_COMP_I2 || c == AW_COMP_D2)) continue;
        red_max<NT>(red, c, khi[c]);
        red_max<NT>(red, 5 + c, klo[c]);
        red_max<NT>(red, 10 + c, akm[c]);
    }
    red_max<NT>(red, 15, endval);
    cta_sync<NT>();
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const int h = red[c], l = red[5 + c];
        if (h == INT_MIN) {  // nothing in bounds: trimmed to empty (lo kept, hi = lo-1)
            so.lo[c] = lo;
            so.hi[c] = lo - 1;
        } else {
            so.lo[c] = -l;
            so.hi[c] = h;
        }
        so.ak[c] = red[10 + c];
    }
    so.endval = red[15];
}

__device__ __forceinline__ bool end_reached(const StepOut& so, int comp_end, int k_end, int tlen) {
    return so.lo[comp_end] <= k_end && k_end <= so.hi[comp_end] && so.endval >= tlen;
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
template <int NT, int BITS, bool TWO>
__global__ void __launch_bounds__(NT) aw_align_kernel(const KParams P) {
    constexpr int NCOMP = TWO ? 5 : 3;
    extern __shared__ unsigned long long smem_raw[];
    const int scope = P.pen.scope;
    SlotMeta* ring_meta = reinterpret_cast<SlotMeta*>(smem_raw);          // [2][scope]
    int* hitk = reinterpret_cast<int*>(ring_meta + 2 * scope);           // [scope*5]
    unsigned long long* scanbuf = reinterpret_cast<unsigned long long*>(hitk + ((scope * 5 + 1) & ~1));  // [NT]
    __shared__ int red[3][NRED];
    __shared__ SubProblem stack[MAX_STACK];
    __shared__ unsigned s_next;
    __shared__ unsigned s_nruns;
    __shared__ unsigned long long s_acc[8];
    __shared__ unsigned long long s_text_off, s_bytes_off;

    const int tid = threadIdx.x;
    const AwPen pen = P.pen;
    int* const ring_base = P.ws_ring + (size_t)blockIdx.x * P.ring_ints_per_cta;
    int* const hist_base = P.ws_hist + (size_t)blockIdx.x * P.hist_ints_per_cta;
    int* const hist_meta = P.ws_hist_meta + (size_t)blockIdx.x * (size_t)P.hist_max_scores * HIST_META_INTS;
    uint32_t* const pair_runs = P.ws_runs + (size_t)blockIdx.x * 2 * P.runs_cap;
    uint32_t* const leaf_runs = pair_runs + P.runs_cap;
    const int W = P.W;

    if (tid < 3 * NRED) (&red[0][0])[tid] = INT_MIN;
    int red_i = 0;  // rotating reduction buffer (uniform)
    cta_sync<NT>();

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
        const uint32_t* const pw = (BITS == 2) ? P.packed + qs.packed_off : reinterpret_cast<const uint32_t*>(P.ascii + qs.ascii_off);
        const uint32_t* const tw = (BITS == 2) ? P.packed + ts.packed_off : reinterpret_cast<const uint32_t*>(P.ascii + ts.ascii_off);
        const int PLEN = (int)qs.len, TLEN = (int)ts.len;
        const int koff = min(PLEN + 1, W / 2);  // diagonal k lives at index k + koff
        const int kmin_alloc = -koff, kmax_alloc = W - 1 - koff;

        int status = ST_OK;
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
                auto ring_ptr = [&](int d, int slot, int c) -> int* {
                    return ring_base + ((size_t)((d * scope + slot) * NCOMP + (TWO ? c : (c == AW_COMP_D1 ? 2 : c)))) * (size_t)W + koff;
                };
                auto set_empty = [&](SlotMeta& m) {
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        m.lo[c] = 1;
                        m.hi[c] = 0;
                        m.ak[c] = INT_MIN;
                    }
                };
                // score-0 wavefront of direction d (wavefront_unialign_init_end2end) + extend
                auto init_dir = [&](int d, StepOut& so) {
                    int* r = red[red_i];
                    if (tid == 0) {
                        int m = 0;
                        if (cbeg[d] == AW_COMP_M) m = extend_cell<BITS>(svd[d], 0, 0);
                        ring_ptr(d, 0, cbeg[d])[0] = m;
                        r[cbeg[d]] = 0;
                        r[5 + cbeg[d]] = 0;
                        r[10 + cbeg[d]] = 2 * m;
                        if (cbeg[d] == AW_COMP_M && cend[d] == AW_COMP_M && k_end == 0) r[15] = m;
                    }
                    cta_sync<NT>();
                    SlotMeta& mt = ring_meta[d * scope + 0];
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        const bool on = (c == cbeg[d]);
                        so.lo[c] = on ? 0 : 1;
                        so.hi[c] = on ? 0 : 0;
                        so.ak[c] = on ? r[10 + c] : INT_MIN;
                        mt.lo[c] = so.lo[c];
                        mt.hi[c] = on ? 0 : 0;
                        mt.ak[c] = so.ak[c];
                    }
                    so.endval = r[15];
                    // recycle the buffer used two steps ago
                    const int nxt = (red_i + 2) % 3;
                    if (tid < NRED) red[nxt][tid] = INT_MIN;
                    red_i = (red_i + 1) % 3;
                };
                // compute + extend wavefront `s` of direction d; returns true on END_REACHED
                auto step_dir = [&](int d, int s, StepOut& so) -> bool {
                    In in[7];
                    auto fetch = [&](int c, int score) -> In {
                        In w;
                        if (score < 0) {
                            w.p = nullptr;
                            w.lo = 1;
                            w.hi = 0;
                            return w;
                        }
                        const int slot = score % scope;
                        const SlotMeta& m = ring_meta[d * scope + slot];
                        w.p = ring_ptr(d, slot, c);
                        w.lo = m.lo[c];
                        w.hi = m.hi[c];
                        return w;
                    };
                    in[IN_MX] = fetch(AW_COMP_M, s - pen.x);
                    in[IN_MO1] = fetch(AW_COMP_M, s - pen.o1 - pen.e1);
                    in[IN_I1E] = fetch(AW_COMP_I1, s - pen.e1);
                    in[IN_D1E] = fetch(AW_COMP_D1, s - pen.e1);
                    if (TWO) {
                        in[IN_MO2] = fetch(AW_COMP_M, s - pen.o2 - pen.e2);
                        in[IN_I2E] = fetch(AW_COMP_I2, s - pen.e2);
                        in[IN_D2E] = fetch(AW_COMP_D2, s - pen.e2);
                    } else {
                        in[IN_MO2] = in[IN_I2E] = in[IN_D2E] = fetch(AW_COMP_M, -1);
                    }
                    // wavefront_compute_limits_input over the non-empty inputs
                    int lo = INT_MAX, hi = INT_MIN;
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
                    SlotMeta& mt = ring_meta[d * scope + (s % scope)];
                    ++w_steps;
                    if (lo > hi) {  // all inputs null: null step
                        set_empty(mt);
#pragma unroll
                        for (int c = 0; c < 5; ++c) {
                            so.lo[c] = 1;
                            so.hi[c] = 0;
                            so.ak[c] = INT_MIN;
                        }
                        so.endval = INT_MIN;
                        return false;
                    }
                    if (lo < kmin_alloc || hi > kmax_alloc) {
                        status = ST_FAIL_WORKSPACE;
                        return false;
                    }
                    int* out[5];
#pragma unroll
                    for (int c = 0; c < 5; ++c) out[c] = ring_ptr(d, s % scope, (TWO || c == AW_COMP_M || c == AW_COMP_I1 || c == AW_COMP_D1) ? c : AW_COMP_M);
                    wf_step<NT, BITS, TWO>(in, out, lo, hi, svd[d], k_end, cend[d], red[red_i], so);
                    w_cells += (unsigned long long)(hi - lo + 1) * NCOMP;
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        mt.lo[c] = so.lo[c];
                        mt.hi[c] = so.hi[c];
                        mt.ak[c] = so.ak[c];
                    }
                    const int nxt = (red_i + 2) % 3;
                    if (tid < NRED) red[nxt][tid] = INT_MIN;
                    red_i = (red_i + 1) % 3;
                    return end_reached(so, cend[d], k_end, tlen);
                };
                // wavefront_bialign_overlap: A0 = direction d0 at score s0, A1 = direction d1 at scores s1..s1-scope+1
                auto overlap = [&](int d0, int d1, int s0, int s1) {
                    const SlotMeta& m0 = ring_meta[d0 * scope + (s0 % scope)];
                    const int kinv = tlen - plen;
                    const int order[5] = {AW_COMP_D2, AW_COMP_I2, AW_COMP_D1, AW_COMP_I1, AW_COMP_M};
                    for (int t = tid; t < scope * 5; t += NT) hitk[t] = INT_MAX;
                    cta_sync<NT>();
                    const int bp_entry = bp.score;
                    for (int i = 0; i < scope; ++i) {
                        const int si = s1 - i;
                        if (si < 0) break;
                        const SlotMeta& m1 = ring_meta[d1 * scope + (si % scope)];
#pragma unroll
                        for (int oi = 0; oi < 5; ++oi) {
                            const int c = order[oi];
                            if (!TWO && (c == AW_COMP_D2 || c == AW_COMP_I2)) continue;
                            const int credit = (c == AW_COMP_M) ? 0 : ((c == AW_COMP_D1 || c == AW_COMP_I1) ? pen.o1 : pen.o2);
                            if (s0 + si - credit >= bp_entry) continue;
                            if (m0.lo[c] > m0.hi[c] || m1.lo[c] > m1.hi[c]) continue;
                            const int lo_1 = kinv - m1.hi[c], hi_1 = kinv - m1.lo[c];
                            if (hi_1 < m0.lo[c] || m0.hi[c] < lo_1) continue;
                            // necessary condition for any hit: antidiagonals must meet
                            if ((long long)m0.ak[c] + (long long)m1.ak[c] < (long long)plen + tlen) continue;
                            const int max_lo = max(m0.lo[c], lo_1), min_hi = min(m0.hi[c], hi_1);
                            const int* p0 = ring_ptr(d0, s0 % scope, c);
                            const int* p1 = ring_ptr(d1, si % scope, c);
                            int best = INT_MAX;
                            for (int k0 = max_lo + tid; k0 <= min_hi; k0 += NT) {
                                const int k1 = kinv - k0;
                                const int h0 = p0[k0], h1 = p1[k1];
                                if (h0 + h1 >= tlen) {
                                    if (c != AW_COMP_M) {  // indel2indel: the forward cell must be in bounds
                                        const int kk = (d0 == 0) ? k0 : k1, hh = (d0 == 0) ? h0 : h1;
                                        if (hh - kk > plen || hh > tlen) continue;
                                    }
                                    best = k0;
                                    break;
                                }
                            }
                            best = __reduce_min_sync(0xffffffffu, best);
                            if ((tid & 31) == 0 && best != INT_MAX) atomicMin(&hitk[i * 5 + oi], best);
                        }
                    }
                    cta_sync<NT>();
                    // replay the tests in WFA2's order with the live breakpoint score
                    for (int i = 0; i < scope; ++i) {
                        const int si = s1 - i;
                        if (si < 0) break;
#pragma unroll
                        for (int oi = 0; oi < 5; ++oi) {
                            const int c = order[oi];
                            if (!TWO && (c == AW_COMP_D2 || c == AW_COMP_I2)) continue;
                            const int credit = (c == AW_COMP_M) ? 0 : ((c == AW_COMP_D1 || c == AW_COMP_I1) ? pen.o1 : pen.o2);
                            if (s0 + si - credit >= bp.score) continue;
                            const int k0 = hitk[i * 5 + oi];
                            if (k0 == INT_MAX) continue;
                            const int k1 = kinv - k0;
                            const int h0 = ring_ptr(d0, s0 % scope, c)[k0], h1 = ring_ptr(d1, si % scope, c)[k1];
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
                    }
                    cta_sync<NT>();  // hitk is re-initialised by the next call
                };

                StepOut so;
                int score_f = 0, score_r = 0, f_ak = 0, r_ak = 0;
                bool fb_end = false;  // END_REACHED -> fall back to the base case
                init_dir(0, so);
                if (end_reached(so, cend[0], k_end, tlen)) fb_end = true;
                f_ak = max(0, so.ak[AW_COMP_M]);
                if (!fb_end) {
                    init_dir(1, so);
                    if (end_reached(so, cend[1], k_end, tlen)) fb_end = true;
                    r_ak = max(0, so.ak[AW_COMP_M]);
                }
                bool last_forward = false;
                const int max_antidiagonal = plen + tlen - 1;
                while (!fb_end && status == ST_OK) {
                    if (f_ak + r_ak >= max_antidiagonal) break;
                    ++score_f;
                    bool done = step_dir(0, score_f, so);
                    f_ak = max(f_ak, max(0, done ? 0 : so.ak[AW_COMP_M]));
                    last_forward = true;
                    if (AW_BIALIGN_PHASE1_END_REACHED_RETURNS && done) {
                        fb_end = true;
                        break;
                    }
                    if (status != ST_OK) break;
                    if (f_ak + r_ak >= max_antidiagonal) break;
                    ++score_r;
                    done = step_dir(1, score_r, so);
                    r_ak = max(r_ak, max(0, done ? 0 : so.ak[AW_COMP_M]));
                    last_forward = false;
                    if (AW_BIALIGN_PHASE1_END_REACHED_RETURNS && done) {
                        fb_end = true;
                        break;
                    }
                }
                const int gap_opening = TWO ? pen.o2 : pen.o1;
                while (!fb_end && status == ST_OK) {
                    if (last_forward) {
                        const int min_r = (score_r > scope - 1) ? score_r - (scope - 1) : 0;
                        if (score_f + min_r - gap_opening >= bp.score) break;
                        overlap(0, 1, score_f, score_r);
                        ++score_r;
                        const bool done = step_dir(1, score_r, so);
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
                    const bool done = step_dir(0, score_f, so);
                    if (AW_BIALIGN_PHASE2_END_REACHED_RETURNS && done) {
                        fb_end = true;
                        break;
                    }
                    last_forward = true;
                }
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
                const SeqView sv = SeqView{pw, tw, sp.pb, sp.tb, plen, tlen, false};
                unsigned long long hist_used = 0;
                auto hist_ptr = [&](unsigned off, int width, int clo, int c) -> int* {
                    const int ci = TWO ? c : (c == AW_COMP_D1 ? 2 : c);
                    return hist_base + off + (size_t)ci * width - clo;
                };
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
                        g[12] = (int)m.off;
                    }
                };
                StepOut so;
                int score = 0;
                bool done;
                {  // score 0
                    int* r = red[red_i];
                    SlotMeta& mt = ring_meta[0];
                    if (tid == 0) {
                        int m = 0;
                        if (sp.cb == AW_COMP_M) m = extend_cell<BITS>(sv, 0, 0);
                        hist_ptr(0, 1, 0, sp.cb)[0] = m;
                        r[10 + sp.cb] = 2 * m;
                        if (sp.cb == AW_COMP_M && sp.ce == AW_COMP_M && k_end == 0) r[15] = m;
                    }
                    cta_sync<NT>();
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        const bool on = (c == sp.cb);
                        mt.lo[c] = so.lo[c] = on ? 0 : 1;
                        mt.hi[c] = so.hi[c] = 0;
                        mt.ak[c] = so.ak[c] = on ? r[10 + c] : INT_MIN;
                    }
                    mt.clo = 0;
                    mt.width = 1;
                    mt.off = 0;
                    so.endval = r[15];
                    hist_used = NCOMP;
                    write_hist_meta(0, mt);
                    const int nxt = (red_i + 2) % 3;
                    if (tid < NRED) red[nxt][tid] = INT_MIN;
                    red_i = (red_i + 1) % 3;
                    done = end_reached(so, sp.ce, k_end, tlen);
                }
                while (!done) {
                    ++score;
                    if (score >= P.hist_max_scores) {
                        status = ST_FAIL_WORKSPACE;
                        break;
                    }
                    In in[7];
                    auto fetch = [&](int c, int sc) -> In {
                        In w;
                        if (sc < 0) {
                            w.p = nullptr;
                            w.lo = 1;
                            w.hi = 0;
                            return w;
                        }
                        const SlotMeta& m = ring_meta[sc % scope];
                        w.p = hist_ptr(m.off, m.width, m.clo, c);
                        w.lo = m.lo[c];
                        w.hi = m.hi[c];
                        return w;
                    };
                    in[IN_MX] = fetch(AW_COMP_M, score - pen.x);
                    in[IN_MO1] = fetch(AW_COMP_M, score - pen.o1 - pen.e1);
                    in[IN_I1E] = fetch(AW_COMP_I1, score - pen.e1);
                    in[IN_D1E] = fetch(AW_COMP_D1, score - pen.e1);
                    if (TWO) {
                        in[IN_MO2] = fetch(AW_COMP_M, score - pen.o2 - pen.e2);
                        in[IN_I2E] = fetch(AW_COMP_I2, score - pen.e2);
                        in[IN_D2E] = fetch(AW_COMP_D2, score - pen.e2);
                    } else {
                        in[IN_MO2] = in[IN_I2E] = in[IN_D2E] = fetch(AW_COMP_M, -1);
                    }
                    int lo = INT_MAX, hi = INT_MIN;
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
                    SlotMeta& mt = ring_meta[score % scope];
                    ++w_steps;
                    if (lo > hi) {
#pragma unroll
                        for (int c = 0; c < 5; ++c) {
                            mt.lo[c] = 1;
                            mt.hi[c] = 0;
                            mt.ak[c] = INT_MIN;
                        }
                        mt.clo = 0;
                        mt.width = 0;
                        mt.off = 0;
                        write_hist_meta(score, mt);
                        continue;
                    }
                    const int width = hi - lo + 1;
                    if (hist_used + (unsigned long long)NCOMP * width > P.hist_ints_per_cta) {
                        status = ST_FAIL_WORKSPACE;
                        break;
                    }
                    const unsigned off = (unsigned)hist_used;
                    hist_used += (unsigned long long)NCOMP * width;
                    int* out[5];
#pragma unroll
                    for (int c = 0; c < 5; ++c) out[c] = hist_ptr(off, width, lo, (TWO || c == AW_COMP_M || c == AW_COMP_I1 || c == AW_COMP_D1) ? c : AW_COMP_M);
                    wf_step<NT, BITS, TWO>(in, out, lo, hi, sv, k_end, sp.ce, red[red_i], so);
                    w_cells += (unsigned long long)width * NCOMP;
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        mt.lo[c] = so.lo[c];
                        mt.hi[c] = so.hi[c];
                        mt.ak[c] = so.ak[c];
                    }
                    mt.clo = lo;
                    mt.width = width;
                    mt.off = off;
                    write_hist_meta(score, mt);
                    const int nxt = (red_i + 2) % 3;
                    if (tid < NRED) red[nxt][tid] = INT_MIN;
                    red_i = (red_i + 1) % 3;
                    done = end_reached(so, sp.ce, k_end, tlen);
                }
                if (status != ST_OK) break;
                w_maxbase = max(w_maxbase, (unsigned)score);
                cta_sync<NT>();  // history + meta visible to warp 0

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
                    // candidate table: lane -> (bt type); cost / component / dk / add derive from it
                    while (v > 0 && h > 0 && sc > 0) {
                        const int bt = lane + 1;  // 1..9 for lanes 0..8
                        int cand = INT_MIN;
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
                                    const int val = hist_ptr((unsigned)g[12], g[11], g[10], comp_src)[kk];
                                    if (val >= 0) cand = ((val + add) << AW_BT_TYPE_BITS) | bt;
                                }
                            }
                        }
                        const int max_all = __reduce_max_sync(0xffffffffu, cand);
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
        // header: q qlen qs qe strand t tlen ts te matches block 60 gi:f:x.xxxxxx cg:Z:
        const unsigned hdr_len = (qid1 - qid0) + 1 + ndigits(PLEN) + 1 + 1 + 1 + ndigits(q_end) + 1 + 1 + 1 + (tid1 - tid0) + 1 + ndigits(TLEN) + 1 + 1 + 1 +
                                 ndigits(t_end) + 1 + ndigits(n_m) + 1 + ndigits(block_len) + 1 + 2 + 1 + 5 + 8 + 1 + 5;
        const unsigned long long line_len = want_paf ? hdr_len + cg_len : cg_len;
        const unsigned long long nbytes = (P.flags & AW_FLAG_CIGAR_BYTES) ? (n_m + n_x + n_i + n_d) : 0;
        if (tid == 0) {
            s_text_off = atomicAdd(P.text_cursor, line_len);
            s_bytes_off = nbytes ? atomicAdd(P.bytes_cursor, nbytes) : 0ull;
        }
        cta_sync<NT>();
        const unsigned long long text_off = s_text_off, bytes_off = s_bytes_off;
        if (status == ST_OK && (text_off + line_len > P.text_cap || bytes_off + nbytes > P.bytes_cap)) status = ST_FAIL_WORKSPACE;
        if (status == ST_OK) {
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
            P.out[pair_i] = o;
        }
        cta_sync<NT>();
    }
}

}  // namespace awk
