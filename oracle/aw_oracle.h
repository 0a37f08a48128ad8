/*
 * aw_oracle.h -- CPU restatement (plain C) of the allwave alignment hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under allwave_b200/ may include, link or execute this;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY UNPINNED: the arithmetic of the reference lives in WFA2-lib (C), reached through the
 * un-vendored git dependency lib_wfa2@2f9d9a48addee5185d8ff6ed0594182558d60818
 * (/root/reference/Cargo.toml:27, Cargo.lock:599-601).  It is not on disk, there is no Rust
 * toolchain, and the reference's tests hold no exact CIGAR / score / hash golden vector
 * (SURVEY.md section 4, 8c).  This oracle restates WFA2-lib's published biWFA algorithm and the
 * in-tree Rust glue; it is pinned against (i) the SipHash reference vectors, (ii) the WFA2-lib
 * README known-answer example, (iii) the hand-derived cases of tests/debug/ *.rs, (iv) every
 * property the reference's own tests assert, and (v) an independent Gotoh two-piece DP for
 * score optimality.
 */
#ifndef AW_ORACLE_H
#define AW_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mirrors AlignmentParams, /root/reference/src/types.rs:37-45 */
typedef struct {
    int32_t match_score;
    int32_t mismatch_penalty;
    int32_t gap_open;
    int32_t gap_extend;
    int32_t gap2_open;   /* valid iff has_gap2_open   */
    int32_t gap2_extend; /* valid iff has_gap2_extend */
    uint8_t has_gap2_open;
    uint8_t has_gap2_extend;
} awo_params_t;

/* AlignmentMode::from_params, src/types.rs:105-117 */
enum { AWO_MODE_EDIT = 0, AWO_MODE_AFFINE = 1, AWO_MODE_AFFINE2P = 2 };
int awo_mode_from_params(const awo_params_t* p);

/* work counters of one alignment (used for the roofline's "algorithmic work") */
typedef struct {
    uint64_t cells;       /* (score,diagonal) cells computed x components                 */
    uint64_t steps;       /* wavefront compute steps (all directions / levels)            */
    uint64_t ext_chars;   /* base comparisons done by extend                              */
    uint32_t n_breakpoints;
    uint32_t n_base;      /* base-case (unidirectional) sub-alignments                    */
    uint32_t max_base_score;
    uint32_t max_level;
    uint32_t n_end_reached_fallbacks;
    uint64_t max_width;
} awo_work_t;

/* result of AffineWavefronts::align + score() + cigar() (src/alignment.rs:231-236) */
typedef struct {
    int32_t status; /* 0 = Completed */
    int32_t score;  /* WFA2 convention: -(penalty) with match==0 */
    uint8_t* cigar; /* WFA2 op letters M X I D, malloc'ed */
    size_t cigar_len;
    awo_work_t work;
} awo_alignment_t;

/* ---- L0/L1: WFA2-lib biWFA (MemoryMode::Ultralow, End2End, no heuristic) ---- */
int awo_wfa_align(const awo_params_t* params, const uint8_t* pattern, int plen, const uint8_t* text,
                  int tlen, awo_alignment_t* out);
/* MemoryMode::High variant (plain unidirectional WFA + backtrace); used only by tests */
int awo_wfa_align_unidirectional(const awo_params_t* params, const uint8_t* pattern, int plen,
                                 const uint8_t* text, int tlen, awo_alignment_t* out);
void awo_alignment_free(awo_alignment_t* a);
/* if non-NULL, every breakpoint / base case is logged here (debug aid for the CUDA path) */
void awo_set_trace_file(void* stdio_file);
/* opt-in fast mode of the biWFA restatement for the CPU *baseline* legs (free lists, unchecked interior loop, 8-byte extend);
 * results are identical to the plain path (tests/test_oracle.py::test_fast_mode_agrees) */
void awo_set_fast(int on);
int awo_get_fast(void);
void awo_pool_release(void);

/* penalty of a WFA2 op string under params (gap run L costs min(o1+L*e1, o2+L*e2)) */
int64_t awo_cigar_penalty(const awo_params_t* params, const uint8_t* ops, size_t n);

/* independent O(nm) Gotoh DP (1- or 2-piece affine): optimal penalty, for certification only */
int64_t awo_gotoh_penalty(const awo_params_t* params, const uint8_t* pattern, int plen,
                          const uint8_t* text, int tlen);

/* ---- Rust DefaultHasher (SipHash-1-3, zero keys) ---- */
uint64_t awo_siphash(int c_rounds, int d_rounds, uint64_t k0, uint64_t k1, const uint8_t* msg, size_t len);
uint64_t awo_hash_bytes(const uint8_t* bytes, size_t len); /* <[u8] as Hash>: len(u64 LE) || bytes */
uint64_t awo_hash_str(const uint8_t* bytes, size_t len);   /* <str as Hash>: bytes || 0xFF         */

/* ---- sketches / orientation (src/alignment.rs:69-154,178-190; src/mash.rs:40-114) ---- */
/* returns count (<= sketch_size); out must hold sketch_size entries */
size_t awo_sketch_stranded(const uint8_t* seq, size_t len, int k, size_t sketch_size, uint64_t* out);
size_t awo_sketch_canonical(const uint8_t* seq, size_t len, int k, size_t sketch_size, uint64_t* out);
/* intersection / union sizes over the deduplicated sets */
void awo_jaccard_counts(const uint64_t* a, size_t na, const uint64_t* b, size_t nb, uint64_t* inter,
                        uint64_t* uni);
double awo_jaccard(const uint64_t* a, size_t na, const uint64_t* b, size_t nb);
double awo_mash_distance(const uint64_t* a, size_t na, const uint64_t* b, size_t nb, int k);
void awo_reverse_complement(const uint8_t* seq, size_t len, uint8_t* out);
/* determine_orientation_mash: returns 1 if the query must be reverse-complemented */
int awo_orientation_mash(const uint8_t* query, size_t qlen, const uint8_t* target, size_t tlen);

/* ---- CIGAR passes + PAF (src/alignment.rs:292-376, src/lib.rs:71-112) ---- */
typedef struct {
    uint64_t query_idx, target_idx;
    uint64_t query_start, query_end, target_start, target_end;
    uint8_t is_reverse;
    int32_t score;
    uint64_t num_matches, alignment_length;
    uint8_t* cigar_bytes;
    size_t cigar_len;
    awo_work_t work;
} awo_result_t;

/* "<n><op>..." with M->=, X->X, I->D, D->I; returns malloc'ed NUL-terminated string */
char* awo_cigar_bytes_to_string(const uint8_t* ops, size_t n);
/* align_pair (src/alignment.rs:25-66); use_mash: 1 = stranded min-hash orientation, 0 = 2x WFA */
void awo_align_pair(const uint8_t* query, size_t qlen, const uint8_t* target, size_t tlen,
                    uint64_t query_idx, uint64_t target_idx, const awo_params_t* params,
                    const awo_params_t* orientation_params, int use_mash, awo_result_t* out);
void awo_result_free(awo_result_t* r);
/* alignment_to_paf; returns malloc'ed line without trailing newline */
char* awo_alignment_to_paf(const awo_result_t* r, const char* qname, size_t qlen, const char* tname,
                           size_t tlen);
/* parse_scores (src/lib.rs:116-153): 0 ok, -1 parse error, -2 wrong count */
int awo_parse_scores(const char* s, awo_params_t* out);

/* ---- pair scheduling (src/iterator.rs:30-92,256-334; src/knn_graph.rs:12-52,112-174) ---- */
enum { AWO_SPARS_NONE = 0, AWO_SPARS_RANDOM = 1, AWO_SPARS_AUTO = 2, AWO_SPARS_GIANT = 3, AWO_SPARS_TREE = 4 };
typedef struct {
    int kind;
    double fraction; /* RANDOM: keep fraction; GIANT: connectivity probability; TREE: random fraction */
    uint64_t k_nearest, k_farthest;
    int kmer_size; /* 0 -> default 15 */
} awo_sparsification_t;
double awo_connectivity_probability(uint64_t n, double connectivity_prob);
/* returns malloc'ed array of 2*npairs u64 (i,j); seqs may be NULL unless kind==TREE */
uint64_t* awo_pair_list(uint64_t n, const char* const* ids, const uint8_t* const* seqs, const size_t* lens,
                        int exclude_self, const awo_sparsification_t* sp, uint64_t* npairs);
/* mash distance matrix n x n (row-major doubles) with canonical sketches (src/mash.rs:141-165) */
double* awo_distance_matrix(uint64_t n, const uint8_t* const* seqs, const size_t* lens, int k,
                            size_t sketch_size);
/* build_knn_graph on a literal matrix (src/knn_graph.rs:112-143); returns 2*npairs u64 */
uint64_t* awo_build_knn_graph(const double* matrix, uint64_t n, uint64_t k_neighbors, int farthest,
                              uint64_t* npairs);

/* ---- whole-job driver: aligns pairs with T threads, returns PAF lines (CPU baseline leg) ---- */
typedef struct {
    uint64_t npairs;
    char** paf;       /* npairs lines, pair order */
    int32_t* scores;  /* npairs */
    uint64_t sum_block_len;
    awo_work_t work;  /* summed */
    double seconds;   /* wall time of the alignment loop */
} awo_job_t;
int awo_run_pairs(uint64_t n, const char* const* ids, const uint8_t* const* seqs, const size_t* lens,
                  const uint64_t* pairs, uint64_t npairs, const awo_params_t* params, int use_mash,
                  int threads, awo_job_t* out);
void awo_job_free(awo_job_t* j);

#ifdef __cplusplus
}
#endif
#endif
