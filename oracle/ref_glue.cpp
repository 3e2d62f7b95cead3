// oracle/ref_glue.cpp -- TEST INFRASTRUCTURE, never linked into the product.
//
// Glue that turns the UNMODIFIED reference sources under /root/reference/src into
//   oracle/_ref/pagan2_ref        a non-NCBI `pagan2` command-line binary, and
//   oracle/_ref/libpagan2ref.so   the same objects (minus main) as a library
// (recipe: oracle/Makefile).  Nothing from the reference is copied; its .cpp files are compiled
// where they lie.  This file supplies only what the reference tree does not carry:
//   * stubs for the two translation units that cannot build here (Exonerate wrapper: needs
//     boost::regex + an external binary; version check: needs libcurl);
//   * a link-time interposer (GNU ld --wrap) around Viterbi_alignment::align
//     (viterbi_alignment.cpp:187) that serialises every (left graph, right graph, model, band)
//     -> (score, path) job the reference executes into a job-stream file, so real progressive /
//     placement / pileup runs become fixtures for the oracle restatement and the CUDA path;
//   * pagan2_ref_align_flat(): runs the reference Viterbi_alignment::align on a job given in the
//     flat CSR layout of include/pagan2_b200.h (graphs rebuilt as reference Sequence objects).
//
// Job-stream container (little endian): file = records; record = u32 'PJOB', u32 n_fields;
// field = u32 name_len, name, u32 dtype (0=i32,1=f32,2=f64), u64 count, payload.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <ctime>
#include <string>
#include <vector>
#include <map>
#include <set>
#include <stack>
#include <sstream>
#include <iostream>
#include <fstream>
#include <algorithm>
#include <mutex>
#include <thread>
#include <functional>
#include <stdexcept>
#include <iomanip>
#include <sys/stat.h>
#include <unistd.h>
#include <stdint.h>

// built with -fno-access-control (oracle/Makefile): the glue reads private members, it never edits the classes
#include "utils/settings.h"
#include "utils/settings_handle.h"
#include "utils/db_matrix.h"
#include "utils/int_matrix.h"
#include "utils/evol_model.h"
#include "utils/model_factory.h"
#include "main/sequence.h"
#include "main/basic_alignment.h"
#include "main/viterbi_alignment.h"
#include "utils/exonerate_queries.h"
#include "utils/check_version.h"

using namespace std;
using namespace ppa;

// --------------------------------------------------------------------------------------------
// Stubs for the TUs that cannot be built in this image.
// --------------------------------------------------------------------------------------------
Exonerate_queries::Exonerate_queries() {}
bool Exonerate_queries::test_executable() { return false; }
void Exonerate_queries::local_alignment(map<string, string> *, Fasta_entry *, map<string, hit> *, bool, bool) {}
void Exonerate_queries::local_alignment(Node *, Fasta_entry *, std::multimap<std::string, std::string> *,
                                        std::map<std::string, hit> *, bool, bool, bool) {}
void Exonerate_queries::preselect_targets(map<string, string> *, vector<Fasta_entry> *, map<string, string> *,
                                          map<string, multimap<string, hit> > *, bool) {}
void Exonerate_queries::local_pairwise_alignment(string *, string *, vector<Substring_hit> *, int *) {}

Check_version::Check_version(float) {}

// --------------------------------------------------------------------------------------------
// Job-stream writer
// --------------------------------------------------------------------------------------------
namespace {

struct Field {
    string name;
    uint32_t dtype;
    uint64_t count;
    const void *data;
};

FILE *g_dump = 0;
bool g_dump_checked = false;
std::mutex g_dump_mutex;
vector<uint64_t> g_table_hashes;
bool g_flat_mode = false;          // set by pagan2_ref_align_flat: skip build_ancestral_sequence
double g_build_seconds = 0;        // time spent inside build_ancestral_sequence of the current align
long long g_dump_max_cells = -1;   // PAGAN2_ORACLE_DUMP_MAX_CELLS: skip path-less huge jobs? (-1 = dump all)

double now_seconds() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

void open_dump_once() {
    if (g_dump_checked) return;
    g_dump_checked = true;
    const char *p = getenv("PAGAN2_ORACLE_DUMP");
    if (p && *p) {
        g_dump = fopen(p, "wb");
        if (!g_dump) { fprintf(stderr, "ref_glue: cannot open dump file %s\n", p); exit(2); }
    }
    const char *m = getenv("PAGAN2_ORACLE_DUMP_MAX_CELLS");
    if (m && *m) g_dump_max_cells = atoll(m);
}

void write_record(const vector<Field> &fields) {
    uint32_t magic = 0x424f4a50u; // 'PJOB'
    uint32_t n = (uint32_t)fields.size();
    fwrite(&magic, 4, 1, g_dump);
    fwrite(&n, 4, 1, g_dump);
    for (size_t i = 0; i < fields.size(); i++) {
        const Field &f = fields[i];
        uint32_t nl = (uint32_t)f.name.size();
        fwrite(&nl, 4, 1, g_dump);
        fwrite(f.name.data(), 1, nl, g_dump);
        fwrite(&f.dtype, 4, 1, g_dump);
        fwrite(&f.count, 8, 1, g_dump);
        size_t es = f.dtype == 2 ? 8 : 4;
        if (f.count) fwrite(f.data, es, f.count, g_dump);
    }
    fflush(g_dump);
}

struct Flat_graph {
    vector<int> state, off, start, eidx;
    vector<float> logw;
};

// Sequence graph -> CSR of backward edges in the reference's list order
// (Site::get_first_bwd_edge / get_next_bwd_edge, sequence.h:395-417).
void flatten(Sequence *s, Flat_graph *g) {
    int n = s->sites_length();
    g->state.resize(n);
    g->off.assign(n + 1, 0);
    for (int i = 0; i < n; i++) {
        Site *site = s->get_site_at(i);
        g->state[i] = site->get_state();
        g->off[i] = (int)g->start.size();
        if (site->has_bwd_edge()) {
            Edge *e = site->get_first_bwd_edge();
            for (;;) {
                g->start.push_back(e->get_start_site_index());
                g->logw.push_back((float)e->get_log_posterior_weight());
                g->eidx.push_back(e->get_index());
                if (!site->has_next_bwd_edge()) break;
                e = site->get_next_bwd_edge();
            }
        }
    }
    g->off[n] = (int)g->start.size();
}

// indices of the edges whose is_used() flag is set (the marks backtrack_new_path leaves, viterbi_alignment.cpp:1054-1155)
void used_edges(Sequence *s, vector<int> *out) {
    out->clear();
    vector<Edge> *edges = s->get_edges();
    for (size_t e = 0; e < edges->size(); e++)
        if (edges->at(e).is_used()) out->push_back((int)e);
}
vector<int> g_flat_used[2];  // used edges of the last pagan2_ref_align_flat call (left, right)

uint64_t fnv1a(const void *p, size_t n) {
    const unsigned char *b = (const unsigned char *)p;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

void model_table(Evol_model *m, int *fas, vector<float> *t) {
    int n = m->logCharPr->x;
    *fas = n;
    t->resize((size_t)n * n);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) (*t)[(size_t)i + (size_t)j * n] = m->log_score(i, j);
}

void path_arrays(vector<Path_pointer> &path, vector<int> *p, vector<double> *ps, double *score) {
    p->resize(path.size() * 6);
    ps->resize(path.size());
    *score = -HUGE_VAL;
    for (size_t k = 0; k < path.size(); k++) {
        const Matrix_pointer &mp = path[k].mp;
        (*p)[6 * k + 0] = mp.matrix;
        (*p)[6 * k + 1] = mp.x_ind;
        (*p)[6 * k + 2] = mp.y_ind;
        (*p)[6 * k + 3] = mp.x_edge_ind;
        (*p)[6 * k + 4] = mp.y_edge_ind;
        (*p)[6 * k + 5] = path[k].real_site ? 1 : 0;
        (*ps)[k] = mp.score;
        if (path[k].real_site) *score = mp.score; // last real element = end corner (viterbi_alignment.cpp:1069)
    }
}

} // namespace

// --------------------------------------------------------------------------------------------
// Link-time interposers (ld --wrap=<mangled>): the reference objects call these instead of the
// real members; __real_<mangled> is the untouched reference code.
// --------------------------------------------------------------------------------------------
#define ALIGN_SYM _ZN3ppa17Viterbi_alignment5alignEPNS_8SequenceES2_PNS_10Evol_modelEffb
#define BUILD_SYM _ZN3ppa15Basic_alignment24build_ancestral_sequenceEPNS_8SequenceEPSt6vectorINS_12Path_pointerESaIS4_EEb
#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)

extern "C" void CAT(__real_, ALIGN_SYM)(Viterbi_alignment *, Sequence *, Sequence *, Evol_model *, float, float, bool);
extern "C" void CAT(__real_, BUILD_SYM)(Basic_alignment *, Sequence *, vector<Path_pointer> *, bool);

extern "C" void CAT(__wrap_, BUILD_SYM)(Basic_alignment *self, Sequence *seq, vector<Path_pointer> *path, bool is_reads) {
    if (g_flat_mode) return;
    double t0 = now_seconds();
    CAT(__real_, BUILD_SYM)(self, seq, path, is_reads);
    g_build_seconds += now_seconds() - t0;
}

static double g_total_align_seconds = 0, g_total_build_seconds = 0;
static long long g_total_cells = 0, g_total_jobs = 0;

extern "C" void CAT(__wrap_, ALIGN_SYM)(Viterbi_alignment *self, Sequence *left, Sequence *right, Evol_model *model,
                                       float lbl, float rbl, bool is_reads) {
    open_dump_once();

    // inputs must be captured before the call: align() marks child edges and the caller may
    // later edit the graphs.
    Flat_graph L, R;
    vector<int> l_used_before, r_used_before;
    bool dumping = g_dump != 0;
    if (dumping) { flatten(left, &L); flatten(right, &R); used_edges(left, &l_used_before); used_edges(right, &r_used_before); }

    g_build_seconds = 0;
    double t0 = now_seconds();
    CAT(__real_, ALIGN_SYM)(self, left, right, model, lbl, rbl, is_reads);
    double dt = now_seconds() - t0;

    {
        std::lock_guard<std::mutex> lk(g_dump_mutex);
        g_total_align_seconds += dt;
        g_total_build_seconds += g_build_seconds;
        g_total_jobs++;
        long long lx = left->sites_length() - 1, ly = right->sites_length() - 1, cells = 0;
        if (self->tunnel_defined && (long long)self->upper_bound.size() == lx) {
            for (long long i = 0; i < lx; i++) {
                long long lo = max(0, self->upper_bound[i]), hi = min((long long)self->lower_bound[i], ly - 1);
                if (hi >= lo) cells += hi - lo + 1;
            }
        } else cells = lx * ly;
        g_total_cells += cells;

        if (!dumping) return;

        int fas;
        vector<float> table;
        model_table(model, &fas, &table);
        uint64_t h = fnv1a(table.data(), table.size() * 4);
        int table_id = -1;
        for (size_t i = 0; i < g_table_hashes.size(); i++)
            if (g_table_hashes[i] == h) table_id = (int)i;
        bool new_table = table_id < 0;
        if (new_table) { table_id = (int)g_table_hashes.size(); g_table_hashes.push_back(h); }

        int flags = 0;
        if (Settings_handle::st.is("no-terminal-edges")) flags |= 1;
        if (self->reduced_terminal_gap_penalties) flags |= 2;
        bool banded = self->tunnel_defined && self->upper_bound.size() > 0;

        int meta[10] = {fas, flags, model->get_data_type(), is_reads ? 1 : 0, banded ? 1 : 0,
                        left->sites_length(), right->sites_length(), table_id, (int)(cells & 0x7fffffff), (int)(cells >> 31)};
        float scal[5] = {model->log_gap_open(), model->log_gap_ext(), model->log_gap_end_ext(), model->log_gap_break_ext(),
                         model->log_non_gap()};
        float dist[3] = {model->distance, lbl, rbl};
        vector<int> p;
        vector<double> ps;
        double score;
        path_arrays(self->path, &p, &ps, &score);
        double times[2] = {dt, g_build_seconds};

        vector<Field> f;
        f.push_back({"meta", 0, 10, meta});
        f.push_back({"model", 1, 5, scal});
        f.push_back({"dist", 1, 3, dist});
        if (new_table) f.push_back({"table", 1, (uint64_t)table.size(), table.data()});
        f.push_back({"l_state", 0, (uint64_t)L.state.size(), L.state.data()});
        f.push_back({"l_off", 0, (uint64_t)L.off.size(), L.off.data()});
        f.push_back({"l_start", 0, (uint64_t)L.start.size(), L.start.data()});
        f.push_back({"l_logw", 1, (uint64_t)L.logw.size(), L.logw.data()});
        f.push_back({"l_eidx", 0, (uint64_t)L.eidx.size(), L.eidx.data()});
        f.push_back({"r_state", 0, (uint64_t)R.state.size(), R.state.data()});
        f.push_back({"r_off", 0, (uint64_t)R.off.size(), R.off.data()});
        f.push_back({"r_start", 0, (uint64_t)R.start.size(), R.start.data()});
        f.push_back({"r_logw", 1, (uint64_t)R.logw.size(), R.logw.data()});
        f.push_back({"r_eidx", 0, (uint64_t)R.eidx.size(), R.eidx.data()});
        if (banded) {
            f.push_back({"upper", 0, (uint64_t)self->upper_bound.size(), self->upper_bound.data()});
            f.push_back({"lower", 0, (uint64_t)self->lower_bound.size(), self->lower_bound.data()});
        }
        f.push_back({"score", 2, 1, &score});
        f.push_back({"path", 0, (uint64_t)p.size(), p.data()});
        f.push_back({"path_score", 2, (uint64_t)ps.size(), ps.data()});
        // edge marks: what was set before the call and what is set after it (the call only ever sets marks)
        vector<int> l_used_after, r_used_after;
        used_edges(left, &l_used_after);
        used_edges(right, &r_used_after);
        f.push_back({"l_used_before", 0, (uint64_t)l_used_before.size(), l_used_before.data()});
        f.push_back({"l_used_after", 0, (uint64_t)l_used_after.size(), l_used_after.data()});
        f.push_back({"r_used_before", 0, (uint64_t)r_used_before.size(), r_used_before.data()});
        f.push_back({"r_used_after", 0, (uint64_t)r_used_after.size(), r_used_after.data()});
        f.push_back({"time", 2, 2, times});
        write_record(f);
    }
}

// Totals over every align() the process ran: seconds in align (incl. graph build), seconds of that
// spent in build_ancestral_sequence, DP cells, jobs.  Printed by pagan2_ref at exit when
// PAGAN2_ORACLE_STATS is set (used by bench.py --impl reference).
extern "C" void pagan2_ref_totals(double *align_s, double *build_s, long long *cells, long long *jobs) {
    *align_s = g_total_align_seconds;
    *build_s = g_total_build_seconds;
    *cells = g_total_cells;
    *jobs = g_total_jobs;
}

namespace {
struct Stats_at_exit {
    ~Stats_at_exit() {
        const char *p = getenv("PAGAN2_ORACLE_STATS");
        if (p && *p) {
            FILE *f = fopen(p, "w");
            if (f) {
                fprintf(f, "{\"align_seconds\": %.9g, \"build_seconds\": %.9g, \"cells\": %lld, \"jobs\": %lld}\n",
                        g_total_align_seconds, g_total_build_seconds, g_total_cells, g_total_jobs);
                fclose(f);
            }
        }
        if (g_dump) { fclose(g_dump); g_dump = 0; }
    }
} g_stats_at_exit;
}

// --------------------------------------------------------------------------------------------
// Reference alignment of one flat job.
// --------------------------------------------------------------------------------------------
namespace {

Model_factory *g_flat_mf = 0;

Sequence *unflatten(int n, const int *state, const int *off, const int *start, const float *logw, const int *eidx) {
    Sequence *s = new Sequence(n, Model_factory::dna);
    int n_csr = off[n];
    int max_e = 0;
    for (int k = 0; k < n_csr; k++) max_e = max(max_e, eidx[k]);
    for (int i = 0; i < n; i++) {
        int type = i == 0 ? Site::start_site : (i == n - 1 ? Site::stop_site : Site::real_site);
        int pstate = (i == 0 || i == n - 1) ? Site::ends_site : Site::terminal;
        Site site(s->get_edges(), type, pstate);
        site.set_state(state[i]);
        site.set_empty_children();
        s->push_back_site(site);
    }
    for (int e = 0; e <= max_e; e++) {
        Edge dummy(-1, -1);
        s->push_back_edge(dummy);
    }
    for (int i = 0; i < n; i++)
        for (int k = off[i]; k < off[i + 1]; k++) {
            Edge &e = s->get_edges()->at(eidx[k]);
            e.start_site_index = start[k];
            e.end_site_index = i;
            e.log_posterior_weight = logw[k];
            e.posterior_weight = expf(logw[k]);
            s->get_site_at(start[k])->add_new_fwd_edge_index(eidx[k]);
            s->get_site_at(i)->add_new_bwd_edge_index(eidx[k]);
        }
    return s;
}

// the reference's option table with its defaults (library callers have no command line)
void flat_settings() {
    static bool inited = false;
    if (!inited) {
        const char *argv[] = {"pagan2_ref", "--silent"};
        Settings_handle::st.read_command_line_arguments(2, (char **)argv);
        g_flat_mf = new Model_factory(Model_factory::dna);
        inited = true;
    }
}

} // namespace

extern "C" int pagan2_ref_align_flat(int fas, const float *table, const float *scalars,
                                     int ln, const int *lstate, const int *loff, const int *lstart, const float *llogw, const int *leidx,
                                     int rn, const int *rstate, const int *roff, const int *rstart, const float *rlogw, const int *reidx,
                                     const int *upper, const int *lower, int flags,
                                     double *score_out, int *path_out, double *path_score_out, int path_cap, int *path_len_out) {
    flat_settings();
    if (flags & 1) Settings_handle::st.vm.set("no-terminal-edges", "", false);
    else Settings_handle::st.vm.erase("no-terminal-edges");
    if (flags & 2) Settings_handle::st.vm.erase("no-reduced-terminal-penalties");
    else Settings_handle::st.vm.set("no-reduced-terminal-penalties", "", false);

    Evol_model model(Model_factory::dna, 0.1f);
    delete model.logCharPr;
    model.logCharPr = new Db_matrix(fas, fas, "logP_char");
    for (int j = 0; j < fas; j++)
        for (int i = 0; i < fas; i++) model.logCharPr->s(table[(size_t)i + (size_t)j * fas], i, j);
    model.log_id_prob = scalars[0];
    model.log_ext_prob = scalars[1];
    model.log_end_ext_prob = scalars[2];
    model.log_break_ext_prob = scalars[3];
    model.log_match_prob = scalars[4];

    Sequence *L = unflatten(ln, lstate, loff, lstart, llogw, leidx);
    Sequence *R = unflatten(rn, rstate, roff, rstart, rlogw, reidx);

    Viterbi_alignment va;
    if (upper && lower) {
        va.upper_bound.assign(upper, upper + (ln - 1));
        va.lower_bound.assign(lower, lower + (ln - 1));
        va.tunnel_defined = true;
    }
    g_flat_mode = true;
    va.align(L, R, &model, 0.1f, 0.1f, false);
    g_flat_mode = false;

    vector<int> p;
    vector<double> ps;
    double score;
    path_arrays(va.path, &p, &ps, &score);
    *score_out = score;
    *path_len_out = (int)va.path.size();
    int rc = 0;
    if ((int)va.path.size() > path_cap) rc = 1;
    else {
        if (!p.empty()) memcpy(path_out, p.data(), p.size() * sizeof(int));
        if (!ps.empty()) memcpy(path_score_out, ps.data(), ps.size() * sizeof(double));
    }
    used_edges(L, &g_flat_used[0]);
    used_edges(R, &g_flat_used[1]);
    delete va.ancestral_sequence;
    delete L;
    delete R;
    return rc;
}

// Edge indices marked is_used(true) by the last pagan2_ref_align_flat call (side 0 = left, 1 = right), ascending.
// Returns the count (the first `cap` are written).
extern "C" int pagan2_ref_last_used(int side, int *out, int cap) {
    const vector<int> &v = g_flat_used[side ? 1 : 0];
    for (int k = 0; k < (int)v.size() && k < cap; k++) out[k] = v[k];
    return (int)v.size();
}

// --------------------------------------------------------------------------------------------
// pagan2_ref_prefix_anchors(): the reference's own Find_anchors::find_long_substrings
// (find_anchors.cpp:35-127) on two plain character strings; hits come back as (start_1, start_2, length) triples
// in the reference's order.  Returns the number of hits (which may exceed cap: only cap are written).
// --------------------------------------------------------------------------------------------
#include "utils/find_anchors.h"
extern "C" int pagan2_ref_prefix_anchors(const char *seq1, int len1, const char *seq2, int len2, int min_length, int *out, int cap) {
    std::string s1(seq1, (size_t)len1), s2(seq2, (size_t)len2);
    std::vector<Substring_hit> hits;
    Find_anchors fa;
    fa.find_long_substrings(&s1, &s2, &hits, min_length);
    for (int k = 0; k < (int)hits.size() && k < cap; k++) {
        out[3 * k] = hits[k].start_site_1;
        out[3 * k + 1] = hits[k].start_site_2;
        out[3 * k + 2] = hits[k].length;
    }
    return (int)hits.size();
}

// --------------------------------------------------------------------------------------------
// pagan2_ref_anchor_band(): the reference's own Find_anchors::define_tunnel (find_anchors.cpp:320-489) -- hits -> the
// per-row band bounds -- with --anchors-offset = width.  upper / lower receive len1 + 1 values each.
// --------------------------------------------------------------------------------------------
extern "C" int pagan2_ref_anchor_band(const int *hits, int n_hits, const char *str1, int len1, const char *str2, int len2, int width,
                                      int *upper, int *lower) {
    flat_settings();
    Settings_handle::st.vm.erase("anchors-offset");
    Settings_handle::st.vm.set("anchors-offset", std::to_string(width), false);
    std::vector<Substring_hit> h((size_t)n_hits);
    for (int k = 0; k < n_hits; k++) {
        h[k].start_site_1 = hits[3 * k];
        h[k].start_site_2 = hits[3 * k + 1];
        h[k].length = hits[3 * k + 2];
        h[k].score = hits[3 * k + 2];
    }
    std::string s1(str1, (size_t)len1), s2(str2, (size_t)len2);
    std::vector<int> ub, lb;
    Find_anchors fa;
    fa.define_tunnel(&h, &ub, &lb, &s1, &s2);
    if ((int)ub.size() != len1 + 1 || (int)lb.size() != len1 + 1) return -1;
    for (int i = 0; i <= len1; i++) { upper[i] = ub[i]; lower[i] = lb[i]; }
    return 0;
}
