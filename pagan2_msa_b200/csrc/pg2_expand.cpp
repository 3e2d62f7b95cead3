// pg2_expand.cpp -- host-side unpacker of one job's packed traceback (no CUDA, no DP).
//
// Rebuilds what backtrack_new_path leaves in `vector<Path_pointer> path` (reference
// src/main/viterbi_alignment.cpp:1038-1189, insert_preexisting_gap / insert_gap_path_pointer
// src/main/viterbi_alignment.h:127-200) from the pointers the traceback kernel emitted, and replays the
// score of every element with the reference's operation order (score_* :2029-2255).  The replayed end
// score must reproduce the device's score bit for bit, which makes this a consistency check of the
// fill kernel's arithmetic as well.
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../include/pagan2_b200.h"

namespace {

struct Visited {
    int mat, i, j;     // the cell that was read
    int src;           // matrix its pointer names (PG2_PTR_NONE at the start corner)
    int kl, kr;        // CSR positions of the edges used (-1 when not applicable)
};

inline int csr_pos(const pg2_graph &g, int site, int ord) {
    int k = g.bwd_off[site] + ord;
    return (k >= g.bwd_off[site] && k < g.bwd_off[site + 1]) ? k : -1;
}

struct Replay {
    const pg2_job *job;
    const pg2_model_desc *m;
    int lx, ly;
    bool term, reduced;
    double open_pen(int prev) const { return (reduced && prev == 0) ? 0.0 : (double)m->log_gap_open; }
    // score of cell `c` given the score of the cell its pointer names
    double step(const Visited &c, double sc) const {
        const pg2_graph &L = job->left, &R = job->right;
        if (c.mat == PG2_X_MAT) {
            bool end_gap = term && (c.j == 0 || c.j == ly - 1);
            int p = L.edge_start[c.kl];
            if (c.src == PG2_X_MAT) return sc + (double)(end_gap ? m->log_gap_end_ext : m->log_gap_ext);
            if (c.src == PG2_Y_MAT) return (sc + 0.0) + (double)m->log_gap_open;
            return (sc + (double)m->log_non_gap) + open_pen(p);
        }
        if (c.mat == PG2_Y_MAT) {
            bool end_gap = term && (c.i == 0 || c.i == lx - 1);
            int q = R.edge_start[c.kr];
            if (c.src == PG2_Y_MAT) return sc + (double)(end_gap ? m->log_gap_end_ext : m->log_gap_ext);
            if (c.src == PG2_X_MAT) return (sc + 0.0) + (double)m->log_gap_open;
            return (sc + (double)m->log_non_gap) + open_pen(q);
        }
        float lng = m->log_non_gap;
        double ls = (double)m->log_score[(size_t)L.state[c.i] + (size_t)R.state[c.j] * (size_t)m->fas];
        double base = (c.src == PG2_M_MAT) ? (double)(2 * lng) + ls : (double)(0.0f + lng) + ls;
        return ((sc + base) + (double)L.edge_logw[c.kl]) + (double)R.edge_logw[c.kr];
    }
};

}  // namespace

extern "C" int pg2_expand_path(const pg2_job *job, const pg2_model_desc *model, const pg2_result *result, const uint16_t *steps,
                               pg2_step *out_steps, int32_t *n_out, int32_t *used_left, int32_t *n_used_left,
                               int32_t *used_right, int32_t *n_used_right) {
    if (!job || !model || !result || !steps || !out_steps || !n_out) return PG2_ERR_INVALID;
    *n_out = 0;
    if (n_used_left) *n_used_left = 0;
    if (n_used_right) *n_used_right = 0;
    if (result->status != PG2_JOB_OK) return PG2_ERR_INVALID;
    // compact plain chains (pagan2_b200.h): write their CSR out once, the walk below reads explicit arrays
    pg2_job explicit_job = *job;
    std::vector<int32_t> chain_store[2][3];
    std::vector<float> chain_w[2];
    for (int side = 0; side < 2; ++side) {
        pg2_graph &g = side == 0 ? explicit_job.left : explicit_job.right;
        if (g.bwd_off) continue;
        if (g.edge_start || g.edge_logw || g.n_edges != g.n_sites - 1 || g.n_sites < 2) return PG2_ERR_INVALID;
        std::vector<int32_t> &off = chain_store[side][0], &es = chain_store[side][1], &ei = chain_store[side][2];
        off.resize((size_t)g.n_sites + 1);
        es.resize((size_t)g.n_edges);
        chain_w[side].assign((size_t)g.n_edges, 0.0f);
        for (int s2 = 0; s2 <= g.n_sites; ++s2) off[(size_t)s2] = s2 > 0 ? s2 - 1 : 0;
        for (int k = 0; k < g.n_edges; ++k) es[(size_t)k] = k;
        g.bwd_off = off.data();
        g.edge_start = es.data();
        g.edge_logw = chain_w[side].data();
        if (!g.edge_index) {
            ei.resize((size_t)g.n_edges);
            for (int k = 0; k < g.n_edges; ++k) ei[(size_t)k] = k + 1;
            g.edge_index = ei.data();
        }
    }
    job = &explicit_job;
    const pg2_graph &L = job->left, &R = job->right;
    const int lx = L.n_sites - 1, ly = R.n_sites - 1;
    const int cap = L.n_sites + R.n_sites;
    // run-length decoding (pagan2_b200.h): a word with bit 15 set repeats the previous pointer (word & 0x7fff) times
    std::vector<uint16_t> raw;
    {
        const uint16_t *enc = steps + result->step_off;
        raw.reserve((size_t)cap);
        for (int k = 0; k < result->n_steps; ++k) {
            const uint16_t w = enc[k];
            if (!(w & 0x8000u)) { raw.push_back(w); continue; }
            if (raw.empty() || raw.size() + (size_t)(w & 0x7fffu) > (size_t)cap + 1) return PG2_ERR_INVALID;
            raw.insert(raw.end(), (size_t)(w & 0x7fffu), raw.back());
        }
    }
    const uint16_t *rec = raw.data();
    const int n_rec = (int)raw.size();
    if (n_rec < 1) return PG2_ERR_INVALID;

    int nl = 0, nr = 0;
    auto mark_l = [&](int e) { if (used_left && e >= 0 && nl < cap) used_left[nl++] = e; };
    auto mark_r = [&](int e) { if (used_right && e >= 0 && nr < cap) used_right[nr++] = e; };

    // ---- end pointer (max_end, :293-296) ----
    uint32_t ep = rec[0];
    int vit = PG2_PTR_MATRIX(ep);
    int x_ind, y_ind, x_edge = -1, y_edge = -1, end_kl = -1, end_kr = -1;
    if (vit == PG2_M_MAT) {
        end_kl = csr_pos(L, lx, PG2_PTR_LEFT(ep));
        end_kr = csr_pos(R, ly, PG2_PTR_RIGHT(ep));
        if (end_kl < 0 || end_kr < 0) return PG2_ERR_INVALID;
        x_ind = L.edge_start[end_kl]; y_ind = R.edge_start[end_kr];
        x_edge = L.edge_index[end_kl]; y_edge = R.edge_index[end_kr];
    } else if (vit == PG2_X_MAT) {
        end_kl = csr_pos(L, lx, PG2_PTR_LEFT(ep));
        if (end_kl < 0) return PG2_ERR_INVALID;
        x_ind = L.edge_start[end_kl]; y_ind = ly - 1;
        x_edge = L.edge_index[end_kl];
    } else if (vit == PG2_Y_MAT) {
        end_kr = csr_pos(R, ly, PG2_PTR_RIGHT(ep));
        if (end_kr < 0) return PG2_ERR_INVALID;
        x_ind = lx - 1; y_ind = R.edge_start[end_kr];
        y_edge = R.edge_index[end_kr];
    } else {
        return PG2_ERR_INVALID;
    }
    const int end_vit = vit;

    // ---- backward walk over the emitted records (backtrack_new_path :1038-1181) ----
    // elements are collected in push order (= reverse path order); scores are filled in afterwards
    struct Elem { pg2_step s; int score_from; };  // score_from: index into `visited` (+1), 0 = final score, -1 = literal
    std::vector<Elem> stack;
    std::vector<Visited> visited;
    stack.reserve((size_t)n_rec + 16);
    visited.reserve((size_t)n_rec);

    auto push_gap = [&](int gi, int gj, int mat) {  // insert_gap_path_pointer: Matrix_pointer mp(-1,i,j,matrix)
        Elem e;
        e.s.score = -1.0; e.s.matrix = mat; e.s.x_ind = gi; e.s.y_ind = gj; e.s.x_edge_ind = -1; e.s.y_edge_ind = -1; e.s.real_site = 0;
        e.score_from = -1;
        stack.push_back(e);
    };
    auto preexisting_gap = [&](int &i, int &j, int xi, int yi) {
        while (xi < i) { push_gap(i - 1, j, PG2_X_MAT); --i; }
        while (yi < j) { push_gap(i, j - 1, PG2_Y_MAT); --j; }
    };
    auto push_real = [&](int i, int j, int mat, int xi, int yi, int xe, int ye, int score_from) {
        if (i > 0 || j > 0) {  // insert_new_path_pointer
            Elem e;
            e.s.score = 0; e.s.matrix = mat; e.s.x_ind = xi; e.s.y_ind = yi; e.s.x_edge_ind = xe; e.s.y_edge_ind = ye; e.s.real_site = 1;
            e.score_from = score_from;
            stack.push_back(e);
        }
    };
    auto terminal_fwd_edge = [&](const pg2_graph &g, int from, int stop_site) {  // get_fwd_edge_index_at_site(x_ind, Edge(x_ind,max_i))
        for (int k = g.bwd_off[stop_site]; k < g.bwd_off[stop_site + 1]; ++k)
            if (g.edge_start[k] == from) return g.edge_index[k];
        return -1;
    };

    mark_l(x_edge);
    mark_r(y_edge);
    int i = lx - 1, j = ly - 1;
    bool first_x = true, first_y = true;
    preexisting_gap(i, j, x_ind, y_ind);
    push_real(i, j, vit, x_ind, y_ind, x_edge, y_edge, 0);

    int k = 1;
    bool done = false;
    while (j >= 0 && !done) {
        while (i >= 0) {
            if (vit != PG2_M_MAT && vit != PG2_X_MAT && vit != PG2_Y_MAT) return PG2_ERR_INVALID;
            if (k >= n_rec) return PG2_ERR_INVALID;
            if ((size_t)stack.size() > (size_t)cap + 8) return PG2_ERR_INVALID;
            uint32_t q = rec[k++];
            Visited v;
            v.mat = vit; v.i = i; v.j = j; v.src = PG2_PTR_MATRIX(q); v.kl = -1; v.kr = -1;
            int nx = -1, ny = -1, ex = -1, ey = -1;
            if (vit == PG2_M_MAT || vit == PG2_X_MAT) {
                if (first_x) { mark_l(terminal_fwd_edge(L, x_ind, lx)); first_x = false; }
            }
            if (vit == PG2_M_MAT || vit == PG2_Y_MAT) {
                if (first_y) { mark_r(terminal_fwd_edge(R, y_ind, ly)); first_y = false; }
            }
            if (v.src != PG2_PTR_NONE) {
                if (vit == PG2_M_MAT || vit == PG2_X_MAT) {
                    v.kl = csr_pos(L, i, PG2_PTR_LEFT(q));
                    if (v.kl < 0) return PG2_ERR_INVALID;
                    nx = L.edge_start[v.kl]; ex = L.edge_index[v.kl];
                }
                if (vit == PG2_M_MAT || vit == PG2_Y_MAT) {
                    v.kr = csr_pos(R, j, PG2_PTR_RIGHT(q));
                    if (v.kr < 0) return PG2_ERR_INVALID;
                    ny = R.edge_start[v.kr]; ey = R.edge_index[v.kr];
                }
            }
            visited.push_back(v);
            int from = (int)visited.size();  // element copied from this cell carries this cell's score
            if (vit == PG2_M_MAT) {
                x_ind = nx; y_ind = ny;
                mark_l(ex); mark_r(ey);
                vit = v.src;
                --i; --j;
                preexisting_gap(i, j, x_ind, y_ind);
                push_real(i, j, v.src == PG2_PTR_NONE ? -1 : v.src, x_ind, y_ind, ex, ey, from);
            } else if (vit == PG2_X_MAT) {
                x_ind = nx; y_ind = j;  // max_x->y_ind = j (:913)
                mark_l(ex);
                vit = v.src;
                --i;
                preexisting_gap(i, j, x_ind, y_ind);
                push_real(i, j, v.src == PG2_PTR_NONE ? -1 : v.src, x_ind, y_ind, ex, -1, from);
            } else {
                x_ind = i; y_ind = ny;  // max_y->x_ind = i (:942)
                mark_r(ey);
                vit = v.src;
                --j;
                preexisting_gap(i, j, x_ind, y_ind);
                push_real(i, j, v.src == PG2_PTR_NONE ? -1 : v.src, x_ind, y_ind, -1, ey, from);
            }
            if (i < 1 && j < 1) { done = true; break; }
        }
        if (i < 1 && j < 1) break;
    }

    // ---- forward replay of the scores along the visited cells ----
    Replay rp;
    rp.job = job; rp.m = model; rp.lx = lx; rp.ly = ly;
    rp.term = !(job->flags & PG2_FLAG_NO_TERMINAL_EDGES);
    rp.reduced = (job->flags & PG2_FLAG_REDUCED_TERMINAL_GAP_PENALTIES) != 0;
    std::vector<double> cell_score(visited.size() + 1);
    double sc = 0.0;  // M(0,0) (:729)
    for (int t = (int)visited.size() - 1; t >= 0; --t) {
        const Visited &v = visited[t];
        if (v.src == PG2_PTR_NONE) {
            // only the start corner M(0,0) carries no pointer; it is read in the degenerate empty case
            sc = (v.mat == PG2_M_MAT && v.i == 0 && v.j == 0) ? 0.0 : -HUGE_VAL;
        } else {
            sc = rp.step(v, sc);
        }
        cell_score[t + 1] = sc;
    }
    double final_score;
    if (end_vit == PG2_M_MAT)
        final_score = ((sc + (double)model->log_non_gap) + (double)L.edge_logw[end_kl]) + (double)R.edge_logw[end_kr];
    else
        final_score = sc + 0.0;
    cell_score[0] = final_score;
    if (memcmp(&final_score, &result->score, sizeof(double)) != 0) return PG2_ERR_INVALID;  // replay must be bit-exact

    if ((int)stack.size() > cap) return PG2_ERR_CAPACITY;
    int n = (int)stack.size();
    for (int a = 0; a < n; ++a) {  // stack -> forward order (:1183-1187)
        const Elem &e = stack[n - 1 - a];
        out_steps[a] = e.s;
        if (e.score_from >= 0) out_steps[a].score = cell_score[e.score_from];
    }
    *n_out = n;
    if (n_used_left) *n_used_left = nl;
    if (n_used_right) *n_used_right = nr;
    return PG2_OK;
}
