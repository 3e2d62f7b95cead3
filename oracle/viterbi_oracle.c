/*
 * oracle/viterbi_oracle.c -- CPU restatement of PAGAN2's pairwise graph Viterbi (fill + end corner +
 * traceback) on the flat job layout of include/pagan2_b200.h.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this; the product library (pagan2_msa_b200/csrc) never links or calls it.
 *
 * Parity pin: the reference ships no tests or golden vectors, so this file is
 * pinned against the reference ITSELF: oracle/_ref (the unmodified reference sources compiled here,
 * oracle/Makefile) dumps every job it runs; tests/test_oracle.py demands bit-identical scores
 * and paths on those dumps, and tests/golden/ holds committed dumps for boxes without /root/reference.
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference/src).
 * Evaluation order and FP64 association are kept exactly: ties are decided by first-wins strict '>'
 * (main/basic_alignment.h:449-462).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "../include/pagan2_b200.h"

/* Matrix_pointer (main/basic_alignment.h:33-50) without the probability fields that are unused when
 * compute_full_score is false (basic_alignment.h:563). */
typedef struct {
    double score;
    int32_t x_ind, y_ind, x_edge_ind, y_edge_ind, matrix;
} cell_t;

static const cell_t EMPTY_CELL = {-HUGE_VAL, -1, -1, -1, -1, -1};

typedef struct {
    const pg2_graph *L, *R;
    const pg2_model_desc *m;
    int lx, ly;           /* matrix dims: left.n_sites-1, right.n_sites-1 (viterbi_alignment.cpp:243) */
    int banded;
    const int32_t *upper, *lower;
    int *blo, *bhi;       /* clipped per-row band [max(0,upper), min(lower, ly-1)] (utils/tunnel_matrix.h:194) */
    int64_t *row_off;
    cell_t *X, *Y, *M;
    cell_t outside;       /* shared out-of-band entry, score -inf (viterbi_alignment.cpp:238-241) */
    int no_terminal_edges, reduced;
} dp_t;

/* Tunnel_slice::at (utils/tunnel_matrix.h:85-98): reads outside the band see the empty entry. */
static const cell_t *rd(const dp_t *d, const cell_t *mat, int i, int j) {
    if (j < d->blo[i] || j > d->bhi[i]) return &d->outside;
    return &mat[d->row_off[i] + (j - d->blo[i])];
}
static cell_t *wr(dp_t *d, cell_t *mat, int i, int j) { return &mat[d->row_off[i] + (j - d->blo[i])]; }

/* first_is_bigger (main/basic_alignment.h:449-462) */
static int first_is_bigger(double a, double b) {
    if (a == -HUGE_VAL && b == -HUGE_VAL) return 0;
    return a > b;
}

/* get_log_gap_open_penalty (main/basic_alignment.h:490-513); pair_end_reads is always false (:565) */
static float log_gap_open_penalty(const dp_t *d, int prev_site) {
    if (d->reduced && prev_site == 0) return 0;
    return d->m->log_gap_open;
}

/* iterate_bwd_edges_for_gap + score_gap_ext / score_gap_double / score_gap_open
 * (main/viterbi_alignment.cpp:1328-1349, 2116-2219).  is_x: gap in X (walk left edges, fixed column)
 * else gap in Y (walk right edges, fixed row). */
static void gap_cell(dp_t *d, int i, int j, int is_x, int end_gap, cell_t *max) {
    const pg2_graph *g = is_x ? d->L : d->R;
    int site = is_x ? i : j;
    const cell_t *zm = is_x ? d->X : d->Y; /* same-type matrix */
    const cell_t *wm = is_x ? d->Y : d->X; /* other gap matrix */
    for (int k = g->bwd_off[site]; k < g->bwd_off[site + 1]; k++) {
        int p = g->edge_start[k];
        int pi = is_x ? p : i, pj = is_x ? j : p;
        /* score_gap_ext :2116-2149 (no edge weight on gap moves, :2118,2121) */
        double s = rd(d, zm, pi, pj)->score + (end_gap ? d->m->log_gap_end_ext : d->m->log_gap_ext);
        if (first_is_bigger(s, max->score)) {
            max->score = s;
            max->matrix = is_x ? PG2_X_MAT : PG2_Y_MAT;
            if (is_x) { max->x_ind = p; max->x_edge_ind = g->edge_index[k]; }
            else { max->y_ind = p; max->y_edge_ind = g->edge_index[k]; }
        }
        /* score_gap_double :2158-2180: + log_gap_close() (== 0.0f, utils/evol_model.h:77) + log_gap_open() */
        s = rd(d, wm, pi, pj)->score + 0.0f + d->m->log_gap_open;
        if (first_is_bigger(s, max->score)) {
            max->score = s;
            max->matrix = is_x ? PG2_Y_MAT : PG2_X_MAT;
            if (is_x) { max->x_ind = p; max->x_edge_ind = g->edge_index[k]; }
            else { max->y_ind = p; max->y_edge_ind = g->edge_index[k]; }
        }
        /* score_gap_open :2190-2211 */
        s = rd(d, d->M, pi, pj)->score + d->m->log_non_gap + log_gap_open_penalty(d, p);
        if (first_is_bigger(s, max->score)) {
            max->score = s;
            max->matrix = PG2_M_MAT;
            if (is_x) { max->x_ind = p; max->x_edge_ind = g->edge_index[k]; }
            else { max->y_ind = p; max->y_edge_ind = g->edge_index[k]; }
        }
    }
}

/* score_m_match / score_x_match / score_y_match (main/viterbi_alignment.cpp:2029-2112) */
static void match_pair(dp_t *d, int kl, int kr, const cell_t *src, int which, double log_match, cell_t *max) {
    int pl = d->L->edge_start[kl], pr = d->R->edge_start[kr];
    double wl = d->L->edge_logw[kl], wr_ = d->R->edge_logw[kr];
    double s = rd(d, src, pl, pr)->score + log_match + wl + wr_;
    if (first_is_bigger(s, max->score)) {
        max->score = s;
        max->x_ind = pl;
        max->y_ind = pr;
        max->x_edge_ind = d->L->edge_index[kl];
        max->y_edge_ind = d->R->edge_index[kr];
        max->matrix = which;
    }
}

/* iterate_bwd_edges_for_match (main/viterbi_alignment.cpp:1353-1436): pairs (l0,r0),(l0,r1..),(l1,r0),... */
static void match_cell(dp_t *d, int i, int j, cell_t *max) {
    const pg2_graph *L = d->L, *R = d->R;
    int l0 = L->bwd_off[i], l1 = L->bwd_off[i + 1], r0 = R->bwd_off[j], r1 = R->bwd_off[j + 1];
    if (l0 == l1 || r0 == r1) return;
    float lng = d->m->log_non_gap;
    double ls = d->m->log_score[(size_t)L->state[i] + (size_t)R->state[j] * (size_t)d->m->fas]; /* :1363 */
    double m_log = 2 * lng + ls;          /* :1364  int*float stays float, then widened */
    double x_log = 0.0f + lng + ls;       /* :1366  close penalty is 0 (basic_alignment.h:515-542) */
    double y_log = 0.0f + lng + ls;       /* :1367 */
    for (int kl = l0; kl < l1; kl++)
        for (int kr = r0; kr < r1; kr++) {
            match_pair(d, kl, kr, d->M, PG2_M_MAT, m_log, max);
            match_pair(d, kl, kr, d->X, PG2_X_MAT, x_log, max);
            match_pair(d, kl, kr, d->Y, PG2_Y_MAT, y_log, max);
        }
}

/* compute_fwd_scores (main/viterbi_alignment.cpp:856-971) */
static void compute_fwd_scores(dp_t *d, int i, int j) {
    if (i == 0 && j == 0) return;
    int j_end = (j == 0 || j == d->ly - 1) && !d->no_terminal_edges; /* :864-868 */
    int i_end = (i == 0 || i == d->lx - 1) && !d->no_terminal_edges; /* :875-879 */
    cell_t *mx = wr(d, d->X, i, j), *my = wr(d, d->Y, i, j), *mm = wr(d, d->M, i, j);
    if (i > 0) {
        gap_cell(d, i, j, 1, j_end, mx);
        mx->y_ind = j; /* :913 */
    } else {
        mx->x_ind = mx->y_ind = mx->matrix = -1;
        mm->x_ind = mm->y_ind = mm->matrix = -1;
    }
    if (j > 0) {
        gap_cell(d, i, j, 0, i_end, my);
        my->x_ind = i; /* :942 */
    } else {
        my->x_ind = my->y_ind = my->matrix = -1;
        mm->x_ind = mm->y_ind = mm->matrix = -1;
    }
    if (i > 0 && j > 0) match_cell(d, i, j, mm);
    else mm->x_ind = mm->y_ind = mm->matrix = -1;
}

/* score_gap_close (main/viterbi_alignment.cpp:2221-2255); close penalty == 0.0f */
static void gap_close(dp_t *d, const pg2_graph *g, int k, int is_x, cell_t *max) {
    int prev = g->edge_start[k];
    double s = (is_x ? rd(d, d->X, prev, d->ly - 1) : rd(d, d->Y, d->lx - 1, prev))->score + 0.0f;
    if (first_is_bigger(s, max->score)) {
        max->score = s;
        if (is_x) {
            max->matrix = PG2_X_MAT; max->x_ind = prev; max->x_edge_ind = g->edge_index[k]; max->y_edge_ind = -1;
        } else {
            max->matrix = PG2_Y_MAT; max->y_ind = prev; max->y_edge_ind = g->edge_index[k]; max->x_edge_ind = -1;
        }
    }
}

/* iterate_bwd_edges_for_end_corner (main/viterbi_alignment.cpp:1440-1552) */
static void end_corner(dp_t *d, cell_t *max) {
    const pg2_graph *L = d->L, *R = d->R;
    int ls = d->lx, rs = d->ly; /* stop sites */
    int l0 = L->bwd_off[ls], l1 = L->bwd_off[ls + 1], r0 = R->bwd_off[rs], r1 = R->bwd_off[rs + 1];
    *max = EMPTY_CELL;
    if (l0 == l1 || r0 == r1) return;
    double m_log = d->m->log_non_gap; /* :1451 */
    double best;
    match_pair(d, l0, r0, d->M, PG2_M_MAT, m_log, max);
    best = max->score;
    gap_close(d, L, l0, 1, max);
    if (first_is_bigger(max->score, best)) { best = max->score; max->y_ind = d->ly - 1; }
    gap_close(d, R, r0, 0, max);
    if (first_is_bigger(max->score, best)) { best = max->score; max->x_ind = d->lx - 1; }
    for (int kr = r0 + 1; kr < r1; kr++) { /* :1479-1500 */
        match_pair(d, l0, kr, d->M, PG2_M_MAT, m_log, max);
        if (first_is_bigger(max->score, best)) best = max->score;
        gap_close(d, R, kr, 0, max);
        if (first_is_bigger(max->score, best)) { best = max->score; max->x_ind = d->lx - 1; }
    }
    for (int kl = l0 + 1; kl < l1; kl++) { /* :1504-1550 */
        match_pair(d, kl, r0, d->M, PG2_M_MAT, m_log, max);
        if (first_is_bigger(max->score, best)) best = max->score;
        gap_close(d, L, kl, 1, max);
        if (first_is_bigger(max->score, best)) { best = max->score; max->y_ind = d->ly - 1; }
        for (int kr = r0 + 1; kr < r1; kr++) {
            match_pair(d, kl, kr, d->M, PG2_M_MAT, m_log, max);
            if (first_is_bigger(max->score, best)) best = max->score;
            gap_close(d, R, kr, 0, max);
            if (first_is_bigger(max->score, best)) { best = max->score; max->x_ind = d->lx - 1; }
        }
    }
}

typedef struct {
    pg2_step *v;
    int n, cap;
} stack_t;

static int push(stack_t *s, const cell_t *c, int real) {
    if (s->n >= s->cap) return -1;
    pg2_step *e = &s->v[s->n++];
    e->score = c->score;
    e->matrix = c->matrix;
    e->x_ind = c->x_ind;
    e->y_ind = c->y_ind;
    e->x_edge_ind = c->x_edge_ind;
    e->y_edge_ind = c->y_edge_ind;
    e->real_site = real;
    return 0;
}

/* insert_gap_path_pointer + insert_preexisting_gap (main/viterbi_alignment.h:127-193) */
static int preexisting_gap(stack_t *s, int *i, int *j, int x_ind, int y_ind) {
    while (x_ind < *i) {
        cell_t g = {-1, *i - 1, *j, -1, -1, PG2_X_MAT};
        if (push(s, &g, 0)) return -1;
        --*i;
    }
    while (y_ind < *j) {
        cell_t g = {-1, *i, *j - 1, -1, -1, PG2_Y_MAT};
        if (push(s, &g, 0)) return -1;
        --*j;
    }
    return 0;
}

/* is_used(true) marks in the order backtrack_new_path sets them (:1054-1057, 1079-1101, 1128, 1155) */
typedef struct {
    int32_t *l, *r;
    int nl, nr, cap;
} marks_t;
static void mark(int32_t *v, int *n, int cap, int e) {
    if (v && e >= 0 && *n < cap) v[(*n)++] = e;
}
/* Sequence::get_fwd_edge_index_at_site(from, Edge(from, stop)): the edge from `from` into the stop site, or -1 */
static int terminal_fwd_edge(const pg2_graph *g, int from, int stop_site) {
    for (int k = g->bwd_off[stop_site]; k < g->bwd_off[stop_site + 1]; k++)
        if (g->edge_start[k] == from) return g->edge_index[k];
    return -1;
}

/* backtrack_new_path (main/viterbi_alignment.cpp:1038-1189); returns PG2_JOB_* status. */
static int backtrack(dp_t *d, const cell_t *fp, stack_t *st, marks_t *mk) {
    int vit = fp->matrix, x_ind = fp->x_ind, y_ind = fp->y_ind;
    int j = d->ly - 1, i = d->lx - 1;
    int first_x = 1, first_y = 1;
    mark(mk->l, &mk->nl, mk->cap, fp->x_edge_ind); /* :1054-1057 */
    mark(mk->r, &mk->nr, mk->cap, fp->y_edge_ind);
    if (preexisting_gap(st, &i, &j, x_ind, y_ind)) return -1;
    if (i > 0 || j > 0)
        if (push(st, fp, 1)) return -1;
    while (j >= 0) {
        while (i >= 0) {
            const cell_t *c;
            if (vit == PG2_M_MAT) { c = rd(d, d->M, i, j); }
            else if (vit == PG2_X_MAT) { c = rd(d, d->X, i, j); }
            else if (vit == PG2_Y_MAT) { c = rd(d, d->Y, i, j); }
            else return PG2_JOB_BROKEN_PATH;
            int was = vit;
            if ((was == PG2_M_MAT || was == PG2_X_MAT) && first_x) { /* :1079-1086, :1118-1125 */
                mark(mk->l, &mk->nl, mk->cap, terminal_fwd_edge(d->L, x_ind, d->lx));
                first_x = 0;
            }
            if ((was == PG2_M_MAT || was == PG2_Y_MAT) && first_y) { /* :1087-1094, :1145-1152 */
                mark(mk->r, &mk->nr, mk->cap, terminal_fwd_edge(d->R, y_ind, d->ly));
                first_y = 0;
            }
            if (was == PG2_M_MAT || was == PG2_X_MAT) mark(mk->l, &mk->nl, mk->cap, c->x_edge_ind); /* :1100, :1128 */
            if (was == PG2_M_MAT || was == PG2_Y_MAT) mark(mk->r, &mk->nr, mk->cap, c->y_edge_ind); /* :1101, :1155 */
            vit = c->matrix; x_ind = c->x_ind; y_ind = c->y_ind;
            if (was == PG2_M_MAT) { i--; j--; }
            else if (was == PG2_X_MAT) i--;
            else j--;
            if (preexisting_gap(st, &i, &j, x_ind, y_ind)) return -1;
            if (i > 0 || j > 0)
                if (push(st, c, 1)) return -1;
            if (i < 1 && j < 1) break;
        }
        if (i < 1 && j < 1) break;
    }
    return PG2_JOB_OK;
}

/* Band validity as required by Tunnel_matrix (utils/tunnel_matrix.h:163-168: "monotonically increasing") */
static int band_ok(const int32_t *up, const int32_t *lo, int lx) {
    if (up[0] > 0 || lo[0] < 0) return 0; /* (0,0) must be in band: initialise_array_corner writes it (:729) */
    for (int i = 1; i < lx; i++)
        if (up[i] < up[i - 1] || lo[i] < lo[i - 1]) return 0;
    return 1;
}

/*
 * Align one job.  out_steps (capacity step_cap) receives the reference's forward path
 * (vector<Path_pointer>, viterbi_alignment.cpp:1183-1187).  Returns PG2_JOB_* or -1 when step_cap is
 * too small / allocation failed.  *score = max_end.score.
 * order: 0 = reference's unbanded loop order (j outer, :275-281), 1 = row-major (banded order, :262-271);
 * results do not depend on it (every dependency precedes in both).
 */
int pg2o_align_marks(const pg2_job *job, const pg2_model_desc *model, double *score, pg2_step *out_steps,
                     int32_t step_cap, int32_t *n_steps, int64_t *cells_out, int32_t *used_left, int32_t *n_used_left,
                     int32_t *used_right, int32_t *n_used_right);
int pg2o_align(const pg2_job *job, const pg2_model_desc *model, double *score, pg2_step *out_steps,
               int32_t step_cap, int32_t *n_steps, int64_t *cells_out) {
    return pg2o_align_marks(job, model, score, out_steps, step_cap, n_steps, cells_out, 0, 0, 0, 0);
}

/* The same, also returning the edge indices marked is_used(true), in marking order (capacity step_cap each). */
int pg2o_align_marks(const pg2_job *job, const pg2_model_desc *model, double *score, pg2_step *out_steps,
                     int32_t step_cap, int32_t *n_steps, int64_t *cells_out, int32_t *used_left, int32_t *n_used_left,
                     int32_t *used_right, int32_t *n_used_right) {
    dp_t d;
    marks_t mk = {used_left, used_right, 0, 0, step_cap};
    if (n_used_left) *n_used_left = 0;
    if (n_used_right) *n_used_right = 0;
    memset(&d, 0, sizeof d);
    d.L = &job->left;
    d.R = &job->right;
    d.m = model;
    d.lx = job->left.n_sites - 1;
    d.ly = job->right.n_sites - 1;
    d.banded = job->upper && job->lower;
    d.upper = job->upper;
    d.lower = job->lower;
    d.outside = EMPTY_CELL;
    d.no_terminal_edges = (job->flags & PG2_FLAG_NO_TERMINAL_EDGES) != 0;
    d.reduced = (job->flags & PG2_FLAG_REDUCED_TERMINAL_GAP_PENALTIES) != 0;
    *n_steps = 0;
    *score = -HUGE_VAL;
    if (d.lx < 1 || d.ly < 1) return PG2_JOB_BAD_GRAPH;
    if (d.banded && !band_ok(job->upper, job->lower, d.lx)) return PG2_JOB_BAD_BAND;

    d.blo = (int *)malloc(sizeof(int) * d.lx);
    d.bhi = (int *)malloc(sizeof(int) * d.lx);
    d.row_off = (int64_t *)malloc(sizeof(int64_t) * (d.lx + 1));
    if (!d.blo || !d.bhi || !d.row_off) return -1;
    int64_t total = 0;
    for (int i = 0; i < d.lx; i++) {
        int lo = 0, hi = d.ly - 1;
        if (d.banded) {
            lo = job->upper[i] > 0 ? job->upper[i] : 0;
            hi = job->lower[i] < d.ly - 1 ? job->lower[i] : d.ly - 1;
        }
        d.blo[i] = lo;
        d.bhi[i] = hi;
        d.row_off[i] = total;
        if (hi >= lo) total += hi - lo + 1;
    }
    d.row_off[d.lx] = total;
    if (cells_out) *cells_out = total;
    d.X = (cell_t *)malloc(sizeof(cell_t) * (total ? total : 1));
    d.Y = (cell_t *)malloc(sizeof(cell_t) * (total ? total : 1));
    d.M = (cell_t *)malloc(sizeof(cell_t) * (total ? total : 1));
    int rc = -1;
    if (d.X && d.Y && d.M) {
        for (int64_t k = 0; k < total; k++) d.X[k] = d.Y[k] = d.M[k] = EMPTY_CELL;
        /* initialise_array_corner (:725-733); (0,0) is in band for every band the reference builds */
        if (d.blo[0] == 0 && d.bhi[0] >= 0) wr(&d, d.M, 0, 0)->score = 0.0;
        for (int i = 0; i < d.lx; i++)
            for (int j = d.blo[i]; j <= d.bhi[i]; j++) compute_fwd_scores(&d, i, j);
        cell_t max_end;
        end_corner(&d, &max_end);
        *score = max_end.score;
        if (max_end.score == -HUGE_VAL) rc = PG2_JOB_NO_PATH;
        else {
            stack_t st = {out_steps, 0, step_cap};
            rc = backtrack(&d, &max_end, &st, &mk);
            if (rc == PG2_JOB_OK) {
                if (n_used_left) *n_used_left = mk.nl;
                if (n_used_right) *n_used_right = mk.nr;
                /* stack -> forward order (:1183-1187) */
                for (int a = 0, b = st.n - 1; a < b; a++, b--) {
                    pg2_step t = out_steps[a];
                    out_steps[a] = out_steps[b];
                    out_steps[b] = t;
                }
                *n_steps = st.n;
            }
        }
    }
    free(d.X); free(d.Y); free(d.M); free(d.blo); free(d.bhi); free(d.row_off);
    return rc;
}

/* Dump of one full matrix of scores for debugging kernels: out[(i*ly + j)*3 + {0:X,1:Y,2:M}] */
int pg2o_abi_version(void) { return PG2_ABI_VERSION; }
