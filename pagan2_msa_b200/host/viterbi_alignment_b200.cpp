// viterbi_alignment_b200.cpp -- see viterbi_alignment_b200.h.  Host C++ only: graph packing, the C-ABI calls and the
// reference's own post-processing.  No DP arithmetic happens here.
//
// Built with -fno-access-control against the UNMODIFIED reference headers: it sets and reads the members
// Viterbi_alignment::align itself sets and reads (left, right, model, path, ancestral_sequence, the band vectors);
// it never edits the classes.
#include "viterbi_alignment_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/pagan2_b200.h"
#include "utils/log_output.h"
#include "utils/settings_handle.h"

using namespace ppa;

namespace {

// One Sequence graph as the CSR of backward edges the C-ABI takes, in the reference's list order
// (Site::get_first_bwd_edge / get_next_bwd_edge, sequence.h:395-417 -- the order decides ties).
struct Packed_sequence {
    std::vector<int32_t> state, off, start, eidx;
    std::vector<float> logw;
    pg2_graph view() const {
        pg2_graph g;
        g.n_sites = (int32_t)state.size();
        g.n_edges = (int32_t)start.size();
        g.state = state.data();
        g.bwd_off = off.data();
        g.edge_start = start.data();
        g.edge_logw = logw.data();
        g.edge_index = eidx.data();
        return g;
    }
};

void pack_sequence(Sequence *s, Packed_sequence *g) {
    const int n = s->sites_length();
    g->state.resize(n);
    g->off.assign(n + 1, 0);
    for (int i = 0; i < n; i++) {
        Site *site = s->get_site_at(i);
        g->state[i] = site->get_state();
        g->off[i] = (int32_t)g->start.size();
        if (site->has_bwd_edge()) {
            Edge *e = site->get_first_bwd_edge();
            for (;;) {
                g->start.push_back(e->get_start_site_index());
                g->logw.push_back((float)e->get_log_posterior_weight());  // Edge::log_posterior_weight is a float (sequence.h:43)
                g->eidx.push_back(e->get_index());
                if (!site->has_next_bwd_edge()) break;
                e = site->get_next_bwd_edge();
            }
        }
    }
    g->off[n] = (int32_t)g->start.size();
}

// The engine: one pg2_ctx per process, model handles cached by what determines the table (alphabet, distance and
// the five scalars; Model_factory::alignment_model is a pure function of the distance, model_factory.cpp:1871).
struct Engine {
    std::mutex mutex;
    pg2_ctx *ctx = nullptr;
    int device = 0;
    typedef std::tuple<int, int, float, float, float, float, float, float> Model_key;
    struct Model_entry {
        int32_t handle;
        std::vector<float> table;
        pg2_model_desc desc;
    };
    std::map<Model_key, Model_entry> models;
    ppa_b200::Totals totals = {0, 0, 0, 0.0, 0.0};

    void fatal(const char *what, int rc) {
        Log_output::write_out(std::string("pagan2_b200: ") + what + " failed (" + std::to_string(rc) + "): " + pg2_last_error() +
                                  "\nThis build aligns on a B200 only; there is no CPU path.\n",
                              0);
        exit(1);
    }
    void ensure() {
        if (ctx) return;
        const char *d = getenv("PAGAN2_B200_DEVICE");
        if (d && *d) device = atoi(d);
        int rc = pg2_ctx_create(device, &ctx);
        if (rc != PG2_OK) fatal("pg2_ctx_create", rc);
    }
    const Model_entry &model_for(Evol_model *m) {
        const int fas = m->logCharPr->x;
        Model_key key(fas, m->get_data_type(), m->distance, m->log_gap_open(), m->log_gap_ext(), m->log_gap_end_ext(),
                      m->log_gap_break_ext(), m->log_non_gap());
        auto it = models.find(key);
        if (it != models.end()) return it->second;
        Model_entry e;
        e.table.resize((size_t)fas * fas);
        for (int j = 0; j < fas; j++)
            for (int i = 0; i < fas; i++) e.table[(size_t)i + (size_t)j * fas] = m->log_score(i, j);  // float, as the DP reads it
        it = models.emplace(key, std::move(e)).first;
        Model_entry &me = it->second;
        me.desc.fas = fas;
        me.desc.log_score = me.table.data();
        me.desc.log_gap_open = m->log_gap_open();
        me.desc.log_gap_ext = m->log_gap_ext();
        me.desc.log_gap_end_ext = m->log_gap_end_ext();
        me.desc.log_gap_break_ext = m->log_gap_break_ext();
        me.desc.log_non_gap = m->log_non_gap();
        int rc = pg2_model_upload(ctx, &me.desc, &me.handle);
        if (rc != PG2_OK) fatal("pg2_model_upload", rc);
        return me;
    }
};

Engine &engine() {
    static Engine e;
    return e;
}

// viterbi_alignment.cpp:191-231: what align() does before it allocates the matrices
void prepare(const ppa_b200::Alignment_job &j) {
    Viterbi_alignment *va = j.va;
    va->left = j.left;
    va->right = j.right;
    va->model = j.model;
    va->left_branch_length = j.left_branch_length;
    va->right_branch_length = j.right_branch_length;
    va->debug_print_input_sequences(3);
    va->set_basic_settings();
    if (j.is_reads_sequence || Settings_handle::st.is("keep-all-edges")) va->set_reads_alignment_settings();
    va->set_additional_settings();
    if (va->reduced_terminal_gap_penalties) va->mark_no_gap_penalty_sites(va->left, va->right);
    va->log_edge_weight = &ppa::Viterbi_alignment::edge_log_posterior_weight;
    va->edge_weight = &ppa::Viterbi_alignment::edge_posterior_weight;
    va->transform_edge_weight = &ppa::Viterbi_alignment::square_root_edge_weight;
    if (Settings_handle::st.is("no-weight-transform")) va->transform_edge_weight = &ppa::Viterbi_alignment::plain_edge_weight;
    if (Settings_handle::st.is("cuberoot-weight-transform")) va->transform_edge_weight = &ppa::Viterbi_alignment::cube_root_edge_weight;
    if (va->compute_full_score || Settings_handle::st.is("sample-path") || Settings_handle::st.is("mpost-posterior-plot-file")) {
        Log_output::write_out(
            "pagan2_b200: --full-probability / --sample-path / --sample-additional-paths / posterior plots are not part of the "
            "device alignment path; run the CPU build of pagan2 for these modes.\n",
            0);
        exit(1);
    }
    Log_output::write_out("Viterbi_alignment: lengths: " + Log_output::itos(j.left->sites_length()) + " " +
                              Log_output::itos(j.right->sites_length()),
                          3);
}

// viterbi_alignment.cpp:379-392: path -> vector<Path_pointer>, used-edge marks, ancestral sequence
void finish(const ppa_b200::Alignment_job &j, const pg2_job &job, const pg2_model_desc &desc, const pg2_result &res, const uint16_t *steps) {
    Viterbi_alignment *va = j.va;
    const int cap = job.left.n_sites + job.right.n_sites;
    std::vector<pg2_step> out(cap);
    std::vector<int32_t> used_l(cap), used_r(cap);
    int32_t n = 0, nl = 0, nr = 0;
    int rc = pg2_expand_path(&job, &desc, &res, steps, out.data(), &n, used_l.data(), &nl, used_r.data(), &nr);
    if (rc != PG2_OK) {
        Log_output::write_out("Viterbi_alignment: incorrect backward pointer (device path could not be expanded)\n", 0);
        exit(1);  // the reference exits on a broken traceback (viterbi_alignment.cpp:1167-1171)
    }
    std::vector<Edge> *left_edges = va->left->get_edges(), *right_edges = va->right->get_edges();
    for (int k = 0; k < nl; k++) left_edges->at(used_l[k]).is_used(true);
    for (int k = 0; k < nr; k++) right_edges->at(used_r[k]).is_used(true);

    va->path.clear();
    va->path.reserve(n);
    int last_real = -1;
    for (int k = 0; k < n; k++) {
        const pg2_step &s = out[k];
        Matrix_pointer mp(s.score, s.x_ind, s.y_ind, s.matrix);
        mp.x_edge_ind = s.x_edge_ind;
        mp.y_edge_ind = s.y_edge_ind;
        if (s.real_site) {
            va->path.push_back(Path_pointer(mp, true));
            last_real = k;
        } else {
            // insert_gap_path_pointer (viterbi_alignment.h:127-144): one skipped site, one branch further
            va->path.push_back(Path_pointer(mp, false, s.matrix == Viterbi_alignment::x_mat ? va->left_branch_length : va->right_branch_length, 1));
        }
    }
    if (last_real >= 0) {  // max_end (:293-296)
        va->path[last_real].mp.bwd_score = 1.0;
        va->path[last_real].mp.full_score = 1.0;
    }
    Log_output::write_out("Viterbi_alignment: path found", 3);
    va->ancestral_sequence = new Sequence(va->path.size(), va->model->get_data_type());
    va->build_ancestral_sequence(va->ancestral_sequence, &va->path, j.is_reads_sequence);
    Log_output::write_out("Viterbi_alignment: sequence built", 3);
}

}  // namespace

namespace ppa_b200 {

void set_device(int device) { engine().device = device; }

Totals totals() { return engine().totals; }

void Alignment_batch::add(Viterbi_alignment *va, Sequence *left, Sequence *right, Evol_model *model, float l_branch_length,
                          float r_branch_length, bool is_reads_sequence) {
    Alignment_job j = {va, left, right, model, l_branch_length, r_branch_length, is_reads_sequence};
    jobs.push_back(j);
}

void Alignment_batch::run() {
    if (jobs.empty()) return;
    Engine &E = engine();
    std::lock_guard<std::mutex> lock(E.mutex);  // a pg2_ctx is not thread-safe; the reference's worker threads share one
    E.ensure();

    // graphs that several jobs share (a placement target under many reads) are packed once: the engine
    // recognises them by the identity of their arrays and uploads them once
    std::map<Sequence *, Packed_sequence> packed;
    std::vector<pg2_job> pj(jobs.size());
    std::vector<const Engine::Model_entry *> me(jobs.size());
    int64_t step_cap = 0;
    for (size_t k = 0; k < jobs.size(); k++) {
        const Alignment_job &j = jobs[k];
        prepare(j);
        for (Sequence *s : {j.left, j.right})
            if (!packed.count(s)) pack_sequence(s, &packed[s]);
    }
    for (size_t k = 0; k < jobs.size(); k++) {
        const Alignment_job &j = jobs[k];
        Viterbi_alignment *va = j.va;
        me[k] = &E.model_for(j.model);
        pg2_job &p = pj[k];
        p.left = packed[j.left].view();
        p.right = packed[j.right].view();
        p.model = me[k]->handle;
        p.flags = (Settings_handle::st.is("no-terminal-edges") ? PG2_FLAG_NO_TERMINAL_EDGES : 0u) |
                  (va->reduced_terminal_gap_penalties ? PG2_FLAG_REDUCED_TERMINAL_GAP_PENALTIES : 0u);
        const bool banded = va->tunnel_defined && (int)va->upper_bound.size() >= p.left.n_sites - 1 &&
                            (int)va->lower_bound.size() >= p.left.n_sites - 1;
        p.upper = banded ? va->upper_bound.data() : nullptr;
        p.lower = banded ? va->lower_bound.data() : nullptr;
        step_cap += p.left.n_sites + p.right.n_sites;
    }

    std::vector<pg2_result> res(jobs.size());
    std::vector<uint16_t> steps((size_t)step_cap);
    int rc = pg2_align_batch(E.ctx, (int32_t)jobs.size(), pj.data(), res.data(), steps.data(), step_cap);
    if (rc != PG2_OK) E.fatal("pg2_align_batch", rc);
    pg2_stats st;
    if (pg2_get_stats(E.ctx, &st) == PG2_OK) {
        E.totals.fill_ms += st.fill_ms;
        E.totals.traceback_ms += st.traceback_ms;
        E.totals.cells += st.cells;
    }
    E.totals.jobs += (long long)jobs.size();
    E.totals.batches++;

    // the band admits no path: the reference refills the whole matrix (viterbi_alignment.cpp:298-317)
    std::vector<size_t> retry;
    for (size_t k = 0; k < jobs.size(); k++)
        if (res[k].status == PG2_JOB_NO_PATH && pj[k].upper) retry.push_back(k);
    if (!retry.empty()) {
        std::vector<pg2_job> rj;
        int64_t cap2 = 0;
        for (size_t k : retry) {
            Log_output::write_msg("anchored alignment failed: trying again", 1);
            pg2_job p = pj[k];
            p.upper = p.lower = nullptr;
            rj.push_back(p);
            cap2 += p.left.n_sites + p.right.n_sites;
        }
        std::vector<pg2_result> rres(rj.size());
        std::vector<uint16_t> rsteps((size_t)cap2);
        rc = pg2_align_batch(E.ctx, (int32_t)rj.size(), rj.data(), rres.data(), rsteps.data(), cap2);
        if (rc != PG2_OK) E.fatal("pg2_align_batch (unbanded refill)", rc);
        for (size_t t = 0; t < retry.size(); t++) {
            const size_t k = retry[t];
            if (rres[t].status != PG2_JOB_OK) {
                Log_output::write_out("\nViterbi_alignment: max_end.score==-HUGE_VAL\n", 1);
                exit(1);
            }
            pj[k].upper = pj[k].lower = nullptr;
            finish(jobs[k], pj[k], me[k]->desc, rres[t], rsteps.data());
            res[k].status = -1;  // done
        }
    }
    for (size_t k = 0; k < jobs.size(); k++) {
        if (res[k].status == -1) continue;
        if (res[k].status != PG2_JOB_OK) {
            if (res[k].status == PG2_JOB_NO_PATH) Log_output::write_out("\nViterbi_alignment: max_end.score==-HUGE_VAL\n", 1);
            else Log_output::write_out("Viterbi_alignment: the device rejected an alignment job (status " + Log_output::itos(res[k].status) + ")\n", 0);
            exit(1);
        }
        finish(jobs[k], pj[k], me[k]->desc, res[k], steps.data());
    }
    jobs.clear();
}

void align_on_device(Viterbi_alignment *va, Sequence *left, Sequence *right, Evol_model *model, float l_branch_length,
                     float r_branch_length, bool is_reads_sequence) {
    Alignment_batch b;
    b.add(va, left, right, model, l_branch_length, r_branch_length, is_reads_sequence);
    b.run();
}

}  // namespace ppa_b200
