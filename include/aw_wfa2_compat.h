/*
 * aw_wfa2_compat.h -- every constant / ordering of WFA2-lib's biWFA that decides WHICH
 * co-optimal alignment is reported.  Shared by the CPU oracle (oracle/) and the CUDA
 * kernels (allwave_b200/csrc/) so a later correction is a one-line change for both.
 *
 * PARITY UNPINNED: WFA2-lib (reached by the reference through the un-vendored git
 * dependency lib_wfa2@2f9d9a48, /root/reference/Cargo.toml:27, Cargo.lock:599-601) is not
 * on disk and cannot be fetched; the values below are a restatement of its published
 * algorithm and upstream source layout (wavefront_bialign.c, wavefront_backtrace.c,
 * wavefront_compute*.c, wavefront_extend.c), see SURVEY.md Appendix A.
 */
#ifndef AW_WFA2_COMPAT_H
#define AW_WFA2_COMPAT_H

#include <stdint.h>

/* wavefront offsets (wavefront_offset.h): NULL = INT32_MIN/2 so that NULL+small stays "null-ish" */
#define AW_OFFSET_NULL (INT32_MIN / 2)

/* wavefront_bialign.c: fall back to plain (full-history) WFA below these bounds */
#define AW_BIALIGN_FALLBACK_MIN_SCORE 250  /* score_remaining <= this  -> base case        */
#define AW_BIALIGN_FALLBACK_MIN_LENGTH 100 /* max(plen,tlen) <= this   -> score_remaining=0 */

/* wavefront_bialign_find_breakpoint: a direction that reaches the far corner during the
 * alternating phase (before the wavefronts are close to collision) aborts the search with
 * END_REACHED and the caller falls back to the base case (1); during the overlap phase the
 * extend result is ignored (0). */
#define AW_BIALIGN_PHASE1_END_REACHED_RETURNS 1
#define AW_BIALIGN_PHASE2_END_REACHED_RETURNS 0

/* wavefront_bialign_find_breakpoint, overlap phase: the search stops once
 * score_0 + min_score_1 - gap_opening >= breakpoint.score, where gap_opening is the largest gap-opening
 * credit an indel-to-indel meeting can earn (SURVEY.md A.5 `gap_open_max`): o1 for gap-affine,
 * max(o1, o2) for gap-affine-2p. */
#define AW_BIALIGN_GAP_OPENING(two_piece, o1, o2) ((two_piece) ? ((o1) > (o2) ? (o1) : (o2)) : (o1))

/* wavefront_penalties_set_affine / _affine2p with a match score M < 0 (a bonus): WFA2 aligns with shifted penalties
 *   x' = 2x - 2M,  o' = 2o,  e' = 2e - M        (every column of the alignment gives up M/2 per consumed character)
 * which ranks global alignments exactly like the original scoring, and reports the score of the final CIGAR under the
 * ORIGINAL penalties: score = -(x*#X + sum of gap costs) - M*#M  (= -(s' + M*(plen+tlen)) / 2 for the shifted penalty s').
 * M > 0 is rejected.  [recalled from wavefront_penalties.c / cigar_score_gap_affine*; see PARITY UNPINNED above] */
#define AW_SHIFT_MISMATCH(x, m) (2 * (x) - 2 * (m))
#define AW_SHIFT_GAP_OPEN(o, m) (2 * (o))
#define AW_SHIFT_GAP_EXTEND(e, m) (2 * (e) - (m))

/* wavefront_compute_process_ends: trim the [lo,hi] of every output component (1) or of M only (0) */
#define AW_TRIM_ALL_COMPONENTS 1

/* alignment components (affine2p_matrix_type) */
enum { AW_COMP_M = 0, AW_COMP_I1 = 1, AW_COMP_I2 = 2, AW_COMP_D1 = 3, AW_COMP_D2 = 4 };

/* wavefront_backtrace.c: candidates are (offset << 4) | type and the MAX wins, so equal
 * offsets are broken by the larger type value. */
#define AW_BT_TYPE_BITS 4
enum {
    AW_BT_I1_OPEN = 1,
    AW_BT_I1_EXT = 2,
    AW_BT_I2_OPEN = 3,
    AW_BT_I2_EXT = 4,
    AW_BT_D1_OPEN = 5,
    AW_BT_D1_EXT = 6,
    AW_BT_D2_OPEN = 7,
    AW_BT_D2_EXT = 8,
    AW_BT_M = 9
};

/* wavefront_bialign_overlap: per reverse score i (descending from s1), components are
 * tested in the order D2, I2, D1, I1, M; inside one component test diagonals k0 ascend
 * in the coordinates of aligner 0 and the first hit returns. */

/* orientation sketch (src/alignment.rs:70,75) and mash (src/mash.rs:12,15) */
#define AW_ORIENT_K 15
#define AW_SKETCH_SIZE 1000

#endif
