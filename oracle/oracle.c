/*
 * oracle/oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU-only restatement of the algorithm of MotifScan's native
 * scoring extension (reference: motifscan/motif/cscore.c).  It exists so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the
 * CUDA path against an independent implementation of the same arithmetic.
 * Nothing under motifscan_b200/ may import, link or call this file.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   - the reference's own known-answer tests (tests/test_motif_score.py:6-32),
 *   - tests/golden/*.json, produced by running the unmodified reference
 *     extension (oracle/_ref, compiled from /root/reference by oracle/Makefile)
 *     on seeded inputs (generator: tests/golden/make_golden.py),
 *   - and, when oracle/_ref is present, live differential runs against it.
 *
 * Arithmetic notes (why results are bit-identical to the reference):
 *   - scores accumulate in double, forward and reverse strand in the SAME
 *     ascending-column loop, skipping non-ACGT bases (cscore.c:344-354);
 *   - there is no multiply on the path, so no FMA contraction can differ;
 *   - the threshold predicate is `score - cutoff >= -1e-10` (cscore.c:358,375).
 *
 * Inputs are flat arrays instead of Python objects:
 *   mats     concatenated PWMs, motif m at mats + mat_off[m], row-major 4 x L_m
 *            (row = base A,C,G,T; the reference's `double *matrix[4]`,
 *            cscore.c:6-11)
 *   seqs     concatenated ASCII sequence bytes, sequence i is
 *            seqs[seq_off[i] .. seq_off[i+1])
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* cscore.c:81-114 convert_seq: A/a 0, C/c 1, G/g 2, T/t 3, anything else -1 */
static inline int8_t orc_code(char ch) {
    switch (ch) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

void orc_encode(const char *seq, int64_t n, int8_t *out) {
    for (int64_t i = 0; i < n; i++) out[i] = orc_code(seq[i]);
}

/* cscore.c:36-48 get_max_raw_score: per column max over the 4 rows, floored
 * at 0 (col_max starts at 0), summed in ascending column order in double. */
double orc_max_raw_score(const double *mat, int32_t L) {
    double total = 0;
    for (int32_t j = 0; j < L; j++) {
        double col_max = 0;
        for (int r = 0; r < 4; r++) {
            double v = mat[(size_t) r * L + j];
            if (v > col_max) col_max = v;
        }
        total += col_max;
    }
    return total;
}

/* One window: cscore.c:344-354 (scan) == cscore.c:195-205 (score). */
static inline void orc_window(const double *mat, int32_t L, const int8_t *code,
                              int strand, double *fwd, double *rev) {
    double f = 0, r = 0;
    for (int32_t col = 0; col < L; col++) {
        int8_t row = code[col];
        if (row != -1) {
            if (strand & 1) f += mat[(size_t) row * L + col];
            if (strand & 2) r += mat[(size_t) (3 - row) * L + (L - 1 - col)];
        }
    }
    *fwd = f;
    *rev = r;
}

/*
 * cscore.c:174-229 motif_score_thread / :231-302 motif_score  (`c_score`).
 * Scores ONLY the window at offset 0 of every sequence, for every motif.
 * strand: 1 fwd, 2 rev, 3 max(fwd, rev) with `fwd > rev ? fwd : rev`
 * (cscore.c:215-221); result divided by max_raw_score (cscore.c:223).
 * out is n_pwms x n_seqs, row-major.
 * The reference does not check len(seq) >= L (undefined behaviour); this
 * restatement returns -1 in that case instead of reading out of bounds.
 */
int orc_score(int32_t n_pwms, const int32_t *lens, const double *mats,
              const int64_t *mat_off, int64_t n_seqs, const char *seqs,
              const int64_t *seq_off, int strand, double *out) {
    int32_t lmax = 0;
    for (int32_t m = 0; m < n_pwms; m++) if (lens[m] > lmax) lmax = lens[m];
    for (int64_t i = 0; i < n_seqs; i++)
        if (seq_off[i + 1] - seq_off[i] < lmax) return -1;
    int8_t *code = (int8_t *) malloc((size_t) (lmax > 0 ? lmax : 1));
    if (!code) return -2;
    for (int32_t m = 0; m < n_pwms; m++) {
        const double *mat = mats + mat_off[m];
        int32_t L = lens[m];
        double max_raw = orc_max_raw_score(mat, L);
        for (int64_t i = 0; i < n_seqs; i++) {
            orc_encode(seqs + seq_off[i], L, code);
            double f, r, s = 0;
            orc_window(mat, L, code, strand, &f, &r);
            switch (strand) {
                case 1: s = f; break;
                case 2: s = r; break;
                case 3: s = (f > r) ? f : r; break;
            }
            out[(size_t) m * n_seqs + i] = s / max_raw;
        }
    }
    free(code);
    return 0;
}

/* Growable site store (replaces the reference's malloc'd linked list,
 * cscore.c:304-315; order of insertion is preserved). */
typedef struct {
    int64_t n, cap;
    int64_t *seq_idx;
    int64_t *start;
    double *score;
    int8_t *strand;
} orc_sites;

static int orc_push(orc_sites *s, int64_t seq_idx, int64_t start, double score,
                    int strand) {
    if (s->n == s->cap) {
        int64_t cap = s->cap ? s->cap * 2 : 1024;
        int64_t *a = (int64_t *) realloc(s->seq_idx, sizeof(int64_t) * cap);
        if (!a) return -1;
        s->seq_idx = a;
        int64_t *b = (int64_t *) realloc(s->start, sizeof(int64_t) * cap);
        if (!b) return -1;
        s->start = b;
        double *c = (double *) realloc(s->score, sizeof(double) * cap);
        if (!c) return -1;
        s->score = c;
        int8_t *d = (int8_t *) realloc(s->strand, sizeof(int8_t) * cap);
        if (!d) return -1;
        s->strand = d;
        s->cap = cap;
    }
    s->seq_idx[s->n] = seq_idx;
    s->start[s->n] = start;
    s->score[s->n] = score;
    s->strand[s->n] = (int8_t) strand;
    s->n++;
    return 0;
}

/*
 * cscore.c:317-397 scan_motif_thread / :399-476 scan_motif (`c_scan_motif`).
 * For motif m, sequence i ascending (skipped when shorter than the motif,
 * cscore.c:337), start j = 0 .. len-L ascending: forward site first
 * (strand 1), then reverse site (strand 2), both with the forward-strand
 * start j (cscore.c:356-389).  Hit iff score/max_raw - cutoff >= -1e-10.
 *
 * Outputs: counts[n_pwms]; site arrays are motif-major, in the reference's
 * list order.  The arrays are malloc'd here; release with orc_free().
 * Motifs are independent, so the per-motif loop may run under OpenMP
 * (n_threads > 1) without changing any result or its order -- the same
 * property the reference's pthread pool relies on (cscore.c:323-328).
 */
int orc_scan(int32_t n_pwms, const int32_t *lens, const double *mats,
             const int64_t *mat_off, const double *cutoffs, int64_t n_seqs,
             const char *seqs, const int64_t *seq_off, int strand,
             int n_threads, int64_t *counts, int64_t **o_seq_idx,
             int64_t **o_start, double **o_score, int8_t **o_strand) {
    int64_t total_bp = seq_off[n_seqs];
    int8_t *code = (int8_t *) malloc((size_t) (total_bp > 0 ? total_bp : 1));
    if (!code) return -2;
    orc_encode(seqs, total_bp, code);
    orc_sites *per = (orc_sites *) calloc((size_t) (n_pwms > 0 ? n_pwms : 1),
                                          sizeof(orc_sites));
    if (!per) { free(code); return -2; }
    int failed = 0;
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (int32_t m = 0; m < n_pwms; m++) {
        const double *mat = mats + mat_off[m];
        int32_t L = lens[m];
        double max_raw = orc_max_raw_score(mat, L);
        double cutoff = cutoffs ? cutoffs[m] : 1.0; /* cscore.c:71-75 */
        for (int64_t i = 0; i < n_seqs; i++) {
            int64_t len = seq_off[i + 1] - seq_off[i];
            if (len < L) continue;
            const int8_t *c = code + seq_off[i];
            for (int64_t j = 0; j < len - L + 1; j++) {
                double f, r;
                orc_window(mat, L, c + j, strand, &f, &r);
                if (strand & 1) {
                    f = f / max_raw;
                    if (f - cutoff >= -1e-10)
                        if (orc_push(&per[m], i, j, f, 1)) failed = 1;
                }
                if (strand & 2) {
                    r = r / max_raw;
                    if (r - cutoff >= -1e-10)
                        if (orc_push(&per[m], i, j, r, 2)) failed = 1;
                }
            }
        }
    }
    free(code);
    int64_t total = 0;
    for (int32_t m = 0; m < n_pwms; m++) { counts[m] = per[m].n; total += per[m].n; }
    int64_t alloc = total > 0 ? total : 1;
    int64_t *seq_idx = (int64_t *) malloc(sizeof(int64_t) * alloc);
    int64_t *start = (int64_t *) malloc(sizeof(int64_t) * alloc);
    double *score = (double *) malloc(sizeof(double) * alloc);
    int8_t *strd = (int8_t *) malloc(sizeof(int8_t) * alloc);
    if (!seq_idx || !start || !score || !strd) failed = 1;
    int64_t at = 0;
    for (int32_t m = 0; m < n_pwms; m++) {
        if (!failed && per[m].n) {
            memcpy(seq_idx + at, per[m].seq_idx, sizeof(int64_t) * per[m].n);
            memcpy(start + at, per[m].start, sizeof(int64_t) * per[m].n);
            memcpy(score + at, per[m].score, sizeof(double) * per[m].n);
            memcpy(strd + at, per[m].strand, sizeof(int8_t) * per[m].n);
            at += per[m].n;
        }
        free(per[m].seq_idx); free(per[m].start); free(per[m].score); free(per[m].strand);
    }
    free(per);
    if (failed) { free(seq_idx); free(start); free(score); free(strd); return -2; }
    *o_seq_idx = seq_idx; *o_start = start; *o_score = score; *o_strand = strd;
    return 0;
}

void orc_free(void *p) { free(p); }
