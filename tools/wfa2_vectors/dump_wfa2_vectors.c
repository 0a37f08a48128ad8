/* Pins the parity oracle against a REAL WFA2-lib build (not available in this repo's build container: "PARITY UNPINNED").
 *
 * A maintainer with WFA2-lib at the commit allwave pins (lib_wfa2 -> WFA2-lib 2f9d9a48addee5185d8ff6ed0594182558d60818,
 * reference Cargo.lock:599-601) runs, from the WFA2-lib checkout after `make`:
 *
 *     gcc -O2 -I. /path/to/tools/wfa2_vectors/dump_wfa2_vectors.c lib/libwfa.a -lm -fopenmp -o dump_wfa2_vectors
 *     gzip -dc /path/to/tests/golden/wfa2_inputs.tsv.gz | ./dump_wfa2_vectors | gzip > /path/to/tests/golden/wfa2_lib_vectors.tsv.gz
 *
 * and commits the output.  tests/test_oracle.py::test_real_wfa2_vectors then checks the oracle (score and every CIGAR
 * operation) against it, and tests/test_gpu_round2.py::test_real_wfa2_vectors_gpu the CUDA path; both skip, saying so, while
 * the file is absent.
 *
 * Input lines:  id <TAB> match <TAB> mismatch <TAB> o1 <TAB> e1 <TAB> o2 <TAB> e2 <TAB> pattern <TAB> text      (o2 = e2 = -1: one-piece affine)
 * Output lines: id <TAB> score <TAB> operations (M X I D, one character per column, WFA2's orientation: I consumes text)
 * The aligner is configured exactly as allwave does (reference src/alignment.rs:222-290, src/wfa.rs:185-224): gap-affine or
 * gap-affine-2p penalties, MemoryMode::Ultralow (biWFA), end-to-end span, full alignment scope, no heuristic. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "wavefront/wavefront_align.h"

int main(void) {
    size_t cap = 1u << 26;
    char* line = (char*)malloc(cap);
    if (!line) return 1;
    while (fgets(line, (int)cap, stdin)) {
        char* f[9];
        int n = 0;
        for (char* p = strtok(line, "\t\n"); p && n < 9; p = strtok(NULL, "\t\n")) f[n++] = p;
        if (n == 8) f[n++] = (char*)"";  /* empty text */
        if (n < 9) continue;
        const int match = atoi(f[1]), x = atoi(f[2]), o1 = atoi(f[3]), e1 = atoi(f[4]), o2 = atoi(f[5]), e2 = atoi(f[6]);
        wavefront_aligner_attr_t attr = wavefront_aligner_attr_default;
        if (o2 < 0) {
            attr.distance_metric = gap_affine;
            attr.affine_penalties.match = match;
            attr.affine_penalties.mismatch = x;
            attr.affine_penalties.gap_opening = o1;
            attr.affine_penalties.gap_extension = e1;
        } else {
            attr.distance_metric = gap_affine_2p;
            attr.affine2p_penalties.match = match;
            attr.affine2p_penalties.mismatch = x;
            attr.affine2p_penalties.gap_opening1 = o1;
            attr.affine2p_penalties.gap_extension1 = e1;
            attr.affine2p_penalties.gap_opening2 = o2;
            attr.affine2p_penalties.gap_extension2 = e2;
        }
        attr.alignment_scope = compute_alignment;
        attr.alignment_form.span = alignment_end2end;
        attr.memory_mode = wavefront_memory_ultralow;
        attr.heuristic.strategy = wf_heuristic_none;
        wavefront_aligner_t* wf = wavefront_aligner_new(&attr);
        const int status = wavefront_align(wf, f[7], (int)strlen(f[7]), f[8], (int)strlen(f[8]));
        printf("%s\t", f[0]);
        if (status != 0) {
            printf("ERR%d\t\n", status);
        } else {
            printf("%d\t", wf->cigar->score);
            fwrite(wf->cigar->operations + wf->cigar->begin_offset, 1, (size_t)(wf->cigar->end_offset - wf->cigar->begin_offset), stdout);
            printf("\n");
        }
        wavefront_aligner_delete(wf);
    }
    free(line);
    return 0;
}
