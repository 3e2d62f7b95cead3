// viterbi_alignment_b200.cpp -- see viterbi_alignment_b200.h.  Host C++ only: graph packing, the batch schedulers, the C-ABI
// calls and the reference's own post-processing.  No DP arithmetic happens here.
//
// Built with -fno-access-control against the UNMODIFIED reference headers: it sets and reads the members
// Viterbi_alignment::align itself sets and reads (left, right, model, path, ancestral_sequence, the band vectors) and
// calls the members Node / Reads_aligner call among themselves; it never edits the classes.
#include "viterbi_alignment_b200.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../../include/pagan2_b200.h"
#include "main/node.h"
#include "main/reads_aligner.h"
#include "utils/fasta_reader.h"
#include "utils/find_anchors.h"
#include "utils/log_output.h"
#include "utils/model_factory.h"
#include "utils/settings_handle.h"

using namespace ppa;

namespace {

uint64_t fnv1a(const void *p, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

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
    // what the DP depends on (the edge indices only name the edges for the used marks)
    uint64_t content_hash() const {
        uint64_t h = fnv1a(state.data(), state.size() * 4);
        h = fnv1a(off.data(), off.size() * 4, h);
        h = fnv1a(start.data(), start.size() * 4, h);
        h = fnv1a(logw.data(), logw.size() * 4, h);
        return fnv1a(eidx.data(), eidx.size() * 4, h);
    }
};

void pack_sequence(Sequence *s, Packed_sequence *g) {
    const int n = s->sites_length();
    g->state.resize(n);
    g->off.assign(n + 1, 0);
    g->start.clear(); g->logw.clear(); g->eidx.clear();
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

typedef std::tuple<int, int, float, float, float, float, float, float> Model_key;
Model_key model_key(Evol_model *m) {
    return Model_key(m->logCharPr->x, m->get_data_type(), m->distance, m->log_gap_open(), m->log_gap_ext(), m->log_gap_end_ext(),
                     m->log_gap_break_ext(), m->log_non_gap());
}

// One alignment between the reference's settings block (viterbi_alignment.cpp:191-231) and its post-processing (:379-392)
struct Pending {
    ppa_b200::Alignment_job job;
    Packed_sequence left_store, right_store;
    const Packed_sequence *left = nullptr, *right = nullptr;
    Model_key mkey;
    uint32_t flags = 0;
    const int32_t *upper = nullptr, *lower = nullptr;
    pg2_result res;
    std::vector<uint16_t> steps;  // the job's own path words (res.step_off == 0)
    bool done = false;
};

// The engine: one pg2_ctx per device named in PAGAN2_B200_DEVICES ("0-7", "0,2,3"; default PAGAN2_B200_DEVICE or 0), model
// handles cached by what determines the table (alphabet, distance and the five scalars; Model_factory::alignment_model is
// a pure function of the distance, model_factory.cpp:1871).
// host wall time of the mirror's own steps, summed over threads (nanoseconds; PAGAN2_B200_STATS)
std::atomic<long long> g_ns_stage(0), g_ns_engine(0), g_ns_expand(0), g_ns_build(0), g_ns_model(0), g_ns_anchor(0), g_anchor_calls(0);
struct Scoped_ns {
    std::atomic<long long> &acc;
    std::chrono::steady_clock::time_point t0;
    explicit Scoped_ns(std::atomic<long long> &a) : acc(a), t0(std::chrono::steady_clock::now()) {}
    ~Scoped_ns() { acc += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); }
};

struct Engine {
    std::mutex mutex;  // one launch batch at a time: a pg2_ctx is not thread-safe, the reference's worker threads share the devices
    std::vector<int> devices;
    std::vector<pg2_ctx *> ctxs;
    int default_device = 0;
    struct Model_entry {
        std::vector<int32_t> handle;  // per ctx, -1 until uploaded there
        std::vector<float> table;
        pg2_model_desc desc;
    };
    std::map<Model_key, Model_entry> models;
    ppa_b200::Totals totals = {0, 0, 0, 0.0, 0.0, 0, 0, 0, 0, 0};

    void fatal(const char *what, int rc) {
        Log_output::write_out(std::string("pagan2_b200: ") + what + " failed (" + std::to_string(rc) + "): " + pg2_last_error() +
                                  "\nThis build aligns on a B200 only; there is no CPU path.\n",
                              0);
        exit(1);
    }
    void ensure() {
        if (!ctxs.empty()) return;
        const char *list = getenv("PAGAN2_B200_DEVICES");
        if (list && *list) {
            for (const char *q = list; *q;) {
                char *e;
                long a = strtol(q, &e, 10);
                if (e == q) break;
                long b = a;
                if (*e == '-') { b = strtol(e + 1, &e, 10); }
                for (long d = a; d <= b; d++) devices.push_back((int)d);
                q = *e ? e + 1 : e;
            }
        }
        if (devices.empty()) {
            const char *d = getenv("PAGAN2_B200_DEVICE");
            devices.push_back((d && *d) ? atoi(d) : default_device);
        }
        for (int d : devices) {
            pg2_ctx *c = nullptr;
            int rc = pg2_ctx_create(d, &c);
            if (rc != PG2_OK) fatal("pg2_ctx_create", rc);
            ctxs.push_back(c);
        }
    }
    Model_entry &model_for(Evol_model *m) {
        const Model_key key = model_key(m);
        auto it = models.find(key);
        if (it != models.end()) return it->second;
        const int fas = m->logCharPr->x;
        Model_entry e;
        e.table.resize((size_t)fas * fas);
        for (int j = 0; j < fas; j++)
            for (int i = 0; i < fas; i++) e.table[(size_t)i + (size_t)j * fas] = m->log_score(i, j);  // float, as the DP reads it
        it = models.emplace(key, std::move(e)).first;
        Model_entry &me = it->second;
        me.handle.assign(ctxs.size(), -1);
        me.desc.fas = fas;
        me.desc.log_score = me.table.data();
        me.desc.log_gap_open = m->log_gap_open();
        me.desc.log_gap_ext = m->log_gap_ext();
        me.desc.log_gap_end_ext = m->log_gap_end_ext();
        me.desc.log_gap_break_ext = m->log_gap_break_ext();
        me.desc.log_non_gap = m->log_non_gap();
        return me;
    }
    int32_t handle_on(Model_entry &me, size_t ci) {
        if (me.handle[ci] < 0) {
            int rc = pg2_model_upload(ctxs[ci], &me.desc, &me.handle[ci]);
            if (rc != PG2_OK) fatal("pg2_model_upload", rc);
        }
        return me.handle[ci];
    }

    // One launch batch: fill + traceback of every job on the device(s); the jobs come back with their own path words.
    // With several devices the batch is cut by contiguous index range, balanced by cell count; every device runs its
    // share on its own context and host thread, and the results meet in host memory.
    void device_align(const std::vector<Pending *> &batch, const std::vector<Model_entry *> &me) {
        Scoped_ns timed(g_ns_engine);
        ensure();
        const size_t n = batch.size();
        if (!n) return;
        std::vector<size_t> cut(1, 0);
        {
            const size_t nd = std::min(ctxs.size(), std::max<size_t>(1, n / 2));
            std::vector<double> cells(n + 1, 0.0);
            for (size_t k = 0; k < n; k++)
                cells[k + 1] = cells[k] + (double)batch[k]->left->state.size() * (double)batch[k]->right->state.size();
            for (size_t d = 1; d < nd; d++) {
                const double want = cells[n] * (double)d / (double)nd;
                size_t pos = (size_t)(std::lower_bound(cells.begin(), cells.end(), want) - cells.begin());
                pos = std::min(std::max(pos, cut.back() + 1), n - (nd - d));
                cut.push_back(pos);
            }
            cut.push_back(n);
        }
        const size_t shards = cut.size() - 1;
        // model handles are uploaded here, on the calling thread (one upload per table and device)
        std::vector<std::vector<pg2_job> > pj(shards);
        for (size_t s = 0; s < shards; s++) {
            pj[s].resize(cut[s + 1] - cut[s]);
            for (size_t k = cut[s]; k < cut[s + 1]; k++) {
                pg2_job &p = pj[s][k - cut[s]];
                p.left = batch[k]->left->view();
                p.right = batch[k]->right->view();
                p.model = handle_on(*me[k], s);
                p.flags = batch[k]->flags;
                p.upper = batch[k]->upper;
                p.lower = batch[k]->lower;
            }
        }
        std::vector<int> rcs(shards, PG2_OK);
        std::vector<std::string> errs(shards);
        std::vector<pg2_stats> stats(shards);
        auto run_shard = [&](size_t s) {
            const size_t lo = cut[s], cnt = cut[s + 1] - cut[s];
            int64_t cap = 0;
            for (size_t k = 0; k < cnt; k++) cap += pj[s][k].left.n_sites + pj[s][k].right.n_sites;
            std::vector<pg2_result> res(cnt);
            std::vector<uint16_t> steps((size_t)cap);
            int rc = pg2_align_batch(ctxs[s], (int32_t)cnt, pj[s].data(), res.data(), steps.data(), cap);
            if (rc != PG2_OK) { rcs[s] = rc; errs[s] = pg2_last_error(); return; }
            pg2_get_stats(ctxs[s], &stats[s]);
            // the band admits no path: the reference refills the whole matrix (viterbi_alignment.cpp:298-317)
            std::vector<size_t> retry;
            for (size_t k = 0; k < cnt; k++)
                if (res[k].status == PG2_JOB_NO_PATH && pj[s][k].upper) retry.push_back(k);
            std::vector<pg2_result> rres;
            std::vector<uint16_t> rsteps;
            if (!retry.empty()) {
                std::vector<pg2_job> rj;
                int64_t cap2 = 0;
                for (size_t k : retry) {
                    Log_output::write_msg("anchored alignment failed: trying again", 1);
                    pg2_job p = pj[s][k];
                    p.upper = p.lower = nullptr;
                    rj.push_back(p);
                    cap2 += p.left.n_sites + p.right.n_sites;
                }
                rres.resize(rj.size());
                rsteps.resize((size_t)cap2);
                rc = pg2_align_batch(ctxs[s], (int32_t)rj.size(), rj.data(), rres.data(), rsteps.data(), cap2);
                if (rc != PG2_OK) { rcs[s] = rc; errs[s] = pg2_last_error(); return; }
            }
            size_t rpos = 0;
            for (size_t k = 0; k < cnt; k++) {
                Pending *p = batch[lo + k];
                const pg2_result *r = &res[k];
                const uint16_t *words = steps.data();
                if (rpos < retry.size() && retry[rpos] == k) {
                    r = &rres[rpos++];
                    words = rsteps.data();
                    p->upper = p->lower = nullptr;  // the path was found in the full matrix
                }
                p->res = *r;
                p->steps.assign(words + r->step_off, words + r->step_off + (r->status == PG2_JOB_OK ? r->n_steps : 0));
                p->res.step_off = 0;
            }
        };
        if (shards == 1) run_shard(0);
        else {
            std::vector<std::thread> pool;
            for (size_t s = 0; s < shards; s++) pool.emplace_back(run_shard, s);
            for (auto &t : pool) t.join();
        }
        for (size_t s = 0; s < shards; s++) {
            if (rcs[s] != PG2_OK) {
                Log_output::write_out("pagan2_b200: pg2_align_batch failed (" + std::to_string(rcs[s]) + "): " + errs[s] +
                                          "\nThis build aligns on a B200 only; there is no CPU path.\n", 0);
                exit(1);
            }
            totals.fill_ms += stats[s].fill_ms;
            totals.traceback_ms += stats[s].traceback_ms;
            totals.cells += stats[s].cells;
        }
        totals.jobs += (long long)n;
        totals.batches++;
        if (shards > 1) totals.sharded_batches++;
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

uint32_t job_flags(Viterbi_alignment *va) {
    return (Settings_handle::st.is("no-terminal-edges") ? PG2_FLAG_NO_TERMINAL_EDGES : 0u) |
           (va->reduced_terminal_gap_penalties ? PG2_FLAG_REDUCED_TERMINAL_GAP_PENALTIES : 0u);
}

// settings, packing, band: everything of one alignment that needs no device
void stage(Pending &p) {
    Scoped_ns timed(g_ns_stage);
    prepare(p.job);
    Viterbi_alignment *va = p.job.va;
    if (!p.left) { pack_sequence(p.job.left, &p.left_store); p.left = &p.left_store; }
    if (!p.right) { pack_sequence(p.job.right, &p.right_store); p.right = &p.right_store; }
    p.mkey = model_key(p.job.model);
    p.flags = job_flags(va);
    const int lx1 = (int)p.left->state.size() - 1;
    const bool banded = va->tunnel_defined && (int)va->upper_bound.size() >= lx1 && (int)va->lower_bound.size() >= lx1;
    p.upper = banded ? va->upper_bound.data() : nullptr;
    p.lower = banded ? va->lower_bound.data() : nullptr;
}

// viterbi_alignment.cpp:379-392: path -> vector<Path_pointer>, used-edge marks, ancestral sequence.  Runs on the caller's
// thread, in the order the reference makes its align() calls: the marks accumulate on the child graphs and
// build_ancestral_sequence reads them (basic_alignment.cpp:587-646).
void finish(Pending &p, const pg2_model_desc &desc) {
    const ppa_b200::Alignment_job &j = p.job;
    Viterbi_alignment *va = j.va;
    if (p.res.status != PG2_JOB_OK) {
        if (p.res.status == PG2_JOB_NO_PATH) Log_output::write_out("\nViterbi_alignment: max_end.score==-HUGE_VAL\n", 1);
        else Log_output::write_out("Viterbi_alignment: the device rejected an alignment job (status " + Log_output::itos(p.res.status) + ")\n", 0);
        exit(1);
    }
    pg2_job job;
    job.left = p.left->view();
    job.right = p.right->view();
    job.model = 0;
    job.flags = p.flags;
    job.upper = p.upper;
    job.lower = p.lower;
    const int cap = job.left.n_sites + job.right.n_sites;
    std::vector<pg2_step> out(cap);
    std::vector<int32_t> used_l(cap), used_r(cap);
    int32_t n = 0, nl = 0, nr = 0;
    std::unique_ptr<Scoped_ns> timed(new Scoped_ns(g_ns_expand));
    int rc = pg2_expand_path(&job, &desc, &p.res, p.steps.data(), out.data(), &n, used_l.data(), &nl, used_r.data(), &nr);
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
    timed.reset(new Scoped_ns(g_ns_build));
    va->ancestral_sequence = new Sequence(va->path.size(), va->model->get_data_type());
    va->build_ancestral_sequence(va->ancestral_sequence, &va->path, j.is_reads_sequence);
    Log_output::write_out("Viterbi_alignment: sequence built", 3);
}

// ------------------------------------------------------------------------------------------------------------------
// Wave combiner: the worker threads of a guide-tree wave (node.cpp:240-264) each reach Viterbi_alignment::align with
// one job; they are collected here and the last one to arrive submits the whole wave as ONE launch batch.
// ------------------------------------------------------------------------------------------------------------------
struct Combiner {
    std::mutex m;
    std::condition_variable cv;
    int expected = 0;  // > 0: a wave of this many alignments is in flight
    std::vector<Pending *> queue;

    void flush_locked(std::unique_lock<std::mutex> &lk) {
        std::vector<Pending *> batch;
        batch.swap(queue);
        expected = 0;
        lk.unlock();
        Engine &E = engine();
        {
            std::lock_guard<std::mutex> el(E.mutex);
            E.ensure();
            std::vector<Engine::Model_entry *> me(batch.size());
            for (size_t k = 0; k < batch.size(); k++) me[k] = &E.model_for(batch[k]->job.model);
            E.device_align(batch, me);
            if (batch.size() > 1) E.totals.wave_batches++;
        }
        lk.lock();
        for (Pending *p : batch) p->done = true;
        cv.notify_all();
    }
    // true: the job was aligned as part of the wave
    bool submit(Pending *p) {
        std::unique_lock<std::mutex> lk(m);
        if (expected <= 0) return false;
        queue.push_back(p);
        if ((int)queue.size() >= expected) flush_locked(lk);
        else cv.wait(lk, [&] { return p->done; });
        return true;
    }
    // a worker of the wave ended without an alignment (nothing in the reference's wave loops does, but a wave must not hang)
    void leave() {
        std::unique_lock<std::mutex> lk(m);
        if (expected <= 0) return;
        expected--;
        if (expected > 0 && (int)queue.size() >= expected) flush_locked(lk);
    }
};
Combiner &combiner() {
    static Combiner c;
    return c;
}
thread_local bool t_in_wave = false, t_submitted = false;

// ------------------------------------------------------------------------------------------------------------------
// Placement prefetch: the trial alignments of query placement (Reads_aligner::read_match_score, reads_aligner.cpp:3467:
// tree node LEFT at distance 0.001, read RIGHT) are independent of each other, and the reference makes them one by one in
// nested loops over reads and candidate nodes (:2613-2893, :1996-2277).  When such a call misses the result cache, the
// alignments the same loops are about to ask for -- the next reads of the query file against every candidate node --
// are computed as ONE launch batch and cached by CONTENT (hash of both packed graphs, model, flags); each later call only
// expands its path, replays its edge marks and builds its ancestral sequence, in the reference's own order.
// A prediction that is off costs device time, never correctness: a call the cache does not hold is aligned on its own.
// ------------------------------------------------------------------------------------------------------------------
struct Cache_key {
    uint64_t l, r;
    bool operator==(const Cache_key &o) const { return l == o.l && r == o.r; }
};
struct Cache_hash {
    size_t operator()(const Cache_key &k) const { return (size_t)(k.l * 0x9E3779B97F4A7C15ull ^ k.r); }
};
struct Cache_entry {
    Model_key mkey;
    uint32_t flags;
    int32_t ln, rn;
    pg2_result res;
    std::vector<uint16_t> steps;
};
struct Read_item {
    Packed_sequence g;
    uint64_t h;
    size_t read;
};
struct Placement {
    bool active = false;
    Reads_aligner *ra = nullptr;
    Node *root = nullptr;
    std::vector<Fasta_entry> reads;
    std::vector<uint64_t> read_hash;   // content hash of the packed read graph, forward strand (0: not built yet)
    size_t cursor = 0;
    size_t chunk = 2048;               // reads per prefetch window (PAGAN2_B200_PREFETCH_READS)
    // the window being served: packed reads [win_lo, win_hi) (both strands with --both-strands) and the target graphs
    // (by content) they have been aligned against
    size_t win_lo = 0, win_hi = 0;
    std::vector<Read_item> items;
    std::vector<uint64_t> done_targets;
    std::unordered_map<Cache_key, Cache_entry, Cache_hash> cache;
    std::mutex m;
};
Placement &placement() {
    static Placement p;
    return p;
}

bool cache_take(Pending &p) {
    Placement &P = placement();
    if (!P.active || p.upper) return false;
    std::lock_guard<std::mutex> lk(P.m);
    if (P.cache.empty()) return false;
    auto it = P.cache.find(Cache_key{p.left->content_hash(), p.right->content_hash()});
    if (it == P.cache.end()) return false;
    Cache_entry &e = it->second;
    if (e.mkey != p.mkey || e.flags != p.flags || e.ln != (int32_t)p.left->state.size() || e.rn != (int32_t)p.right->state.size()) return false;
    p.res = e.res;
    p.steps.swap(e.steps);
    P.cache.erase(it);
    engine().totals.cache_hits++;
    return true;
}

Sequence *read_sequence(Fasta_entry &read) {
    // as Node::add_sequence(*read, read->data_type, false, true) in read_match_score (reads_aligner.cpp:3476-3480)
    return new Sequence(read, read.data_type, false, true, false);
}

// the batch around a missed trial alignment; returns true when p itself was aligned by it
bool placement_prefetch(Pending &p) {
    Placement &P = placement();
    if (!P.active || p.upper || p.job.left_branch_length != 0.001f || !p.job.is_reads_sequence) return false;
    std::unique_lock<std::mutex> lk(P.m);
    const uint64_t rhash = p.right->content_hash(), lhash = p.left->content_hash();
    const size_t n = P.reads.size();
    const bool both = Settings_handle::st.is("both-strands") && p.job.model->get_data_type() == Model_factory::dna;
    // which read of the query file is this?  (the loops take them in file order)
    auto hash_of = [&](size_t r) {
        if (!P.read_hash[r]) {
            Sequence *s = read_sequence(P.reads[r]);
            Packed_sequence g;
            pack_sequence(s, &g);
            delete s;
            P.read_hash[r] = g.content_hash() | 1ull;
        }
        return P.read_hash[r];
    };
    size_t ri = n;
    for (const Read_item &it : P.items)
        if (it.h == rhash && it.read >= P.cursor) { ri = it.read; break; }
    for (size_t d = 0; d < std::min<size_t>(n, 64) && ri == n; d++) {
        const size_t r = (P.cursor + d) % n;
        if (hash_of(r) == (rhash | 1ull)) ri = r;
    }
    if (ri == n) return false;
    P.cursor = ri;
    if (ri < P.win_lo || ri >= P.win_hi) {
        // a new window of reads: pack them once; what the cache still holds belongs to reads the loops have passed
        P.cache.clear();
        P.done_targets.clear();
        P.items.clear();
        P.win_lo = ri;
        P.win_hi = std::min(n, ri + P.chunk);
        for (size_t r = P.win_lo; r < P.win_hi; r++)
            for (int strand = 0; strand < (both ? 2 : 1); strand++) {
                Fasta_entry e = P.reads[r];
                if (strand) e.sequence = P.ra->reverse_complement(e.sequence);
                Sequence *s = read_sequence(e);
                Read_item it;
                pack_sequence(s, &it.g);
                delete s;
                it.h = it.g.content_hash();
                it.read = r;
                P.items.push_back(std::move(it));
            }
    }
    // candidate nodes as the reference's loops enumerate them, minus the graphs this window already went through
    Node *root = P.ra->global_root ? P.ra->global_root : P.root;
    std::multimap<std::string, std::string> tid_nodes;
    bool ignore_tid_tags = true;
    P.ra->get_target_node_names(root, &tid_nodes, &ignore_tid_tags);
    std::map<std::string, Node *> nodes;
    root->get_all_nodes(&nodes);
    struct Target { Packed_sequence g; uint64_t h; std::string tid; };
    std::vector<Target> targets;
    std::vector<Sequence *> seen;
    for (auto &kv : tid_nodes) {
        auto nit = nodes.find(kv.second);
        if (nit == nodes.end() || !nit->second->node_has_sequence_object) continue;
        Sequence *seq = nit->second->get_sequence();
        if (std::find(seen.begin(), seen.end(), seq) != seen.end()) continue;
        seen.push_back(seq);
        Target t;
        pack_sequence(seq, &t.g);
        t.h = t.g.content_hash();
        t.tid = kv.first;
        if (std::find(P.done_targets.begin(), P.done_targets.end(), t.h) != P.done_targets.end()) continue;
        targets.push_back(std::move(t));
    }
    if (targets.empty()) return false;
    std::vector<std::pair<size_t, size_t> > pairs;  // (target, item)
    for (size_t t = 0; t < targets.size(); t++) {
        P.done_targets.push_back(targets[t].h);
        for (size_t k = 0; k < P.items.size(); k++) {
            if (P.items[k].read < ri) continue;
            if (!ignore_tid_tags && targets[t].tid != P.reads[P.items[k].read].tid) continue;
            pairs.push_back(std::make_pair(t, k));
        }
    }
    if (pairs.empty()) return false;
    std::vector<Pending> jobs(pairs.size());
    std::vector<Pending *> batch(pairs.size());
    for (size_t k = 0; k < pairs.size(); k++) {
        Pending &q = jobs[k];
        q.job = p.job;
        q.left = &targets[pairs[k].first].g;
        q.right = &P.items[pairs[k].second].g;
        q.mkey = p.mkey;
        q.flags = p.flags;
        batch[k] = &q;
    }
    {
        Engine &E = engine();
        std::lock_guard<std::mutex> el(E.mutex);
        E.ensure();
        std::vector<Engine::Model_entry *> me(batch.size(), &E.model_for(p.job.model));
        E.device_align(batch, me);
        E.totals.prefetch_batches++;
    }
    bool served = false;
    for (size_t k = 0; k < jobs.size(); k++) {
        Pending &q = jobs[k];
        const Cache_key key{targets[pairs[k].first].h, P.items[pairs[k].second].h};
        if (!served && key.l == lhash && key.r == rhash && q.left->state.size() == p.left->state.size()) {
            p.res = q.res;
            p.steps.swap(q.steps);
            served = true;
            continue;
        }
        Cache_entry e;
        e.mkey = q.mkey;
        e.flags = q.flags;
        e.ln = (int32_t)q.left->state.size();
        e.rn = (int32_t)q.right->state.size();
        e.res = q.res;
        e.steps.swap(q.steps);
        P.cache[key] = std::move(e);
    }
    return served;
}

}  // namespace

namespace ppa_b200 {

void set_device(int device) { engine().default_device = device; }

Totals totals() {
    Totals t = engine().totals;
    t.host_stage_ms = g_ns_stage.load() * 1e-6;
    t.host_engine_ms = g_ns_engine.load() * 1e-6;
    t.host_expand_ms = g_ns_expand.load() * 1e-6;
    t.host_build_ms = g_ns_build.load() * 1e-6;
    t.host_model_ms = g_ns_model.load() * 1e-6;
    t.host_anchor_ms = g_ns_anchor.load() * 1e-6;
    t.anchor_calls = g_anchor_calls.load();
    return t;
}

void Alignment_batch::add(Viterbi_alignment *va, Sequence *left, Sequence *right, Evol_model *model, float l_branch_length,
                          float r_branch_length, bool is_reads_sequence) {
    Alignment_job j = {va, left, right, model, l_branch_length, r_branch_length, is_reads_sequence};
    jobs.push_back(j);
}

void Alignment_batch::run() {
    if (jobs.empty()) return;
    Engine &E = engine();
    // graphs that several jobs share (a placement target under many reads) are packed once: the engine recognises them
    // by the identity of their arrays and uploads them once
    std::map<Sequence *, Packed_sequence> packed;
    std::vector<Pending> pend(jobs.size());
    std::vector<Pending *> batch(jobs.size());
    for (size_t k = 0; k < jobs.size(); k++) {
        Pending &p = pend[k];
        p.job = jobs[k];
        for (Sequence *s : {p.job.left, p.job.right})
            if (!packed.count(s)) pack_sequence(s, &packed[s]);
        p.left = &packed[p.job.left];
        p.right = &packed[p.job.right];
        stage(p);
        batch[k] = &p;
    }
    std::vector<Engine::Model_entry *> me(jobs.size());
    {
        std::lock_guard<std::mutex> lock(E.mutex);
        E.ensure();
        for (size_t k = 0; k < jobs.size(); k++) me[k] = &E.model_for(jobs[k].model);
        E.device_align(batch, me);
    }
    for (size_t k = 0; k < jobs.size(); k++) finish(pend[k], me[k]->desc);
    jobs.clear();
}

void align_on_device(Viterbi_alignment *va, Sequence *left, Sequence *right, Evol_model *model, float l_branch_length,
                     float r_branch_length, bool is_reads_sequence) {
    Pending p;
    p.job = Alignment_job{va, left, right, model, l_branch_length, r_branch_length, is_reads_sequence};
    stage(p);
    Engine &E = engine();
    bool done = cache_take(p);
    if (!done && t_in_wave) {
        done = combiner().submit(&p);
        t_submitted = t_submitted || done;
    }
    if (!done) done = placement_prefetch(p);
    Engine::Model_entry *me;
    {
        std::lock_guard<std::mutex> lock(E.mutex);
        E.ensure();
        me = &E.model_for(model);
        if (!done) {
            std::vector<Pending *> one(1, &p);
            std::vector<Engine::Model_entry *> mes(1, me);
            E.device_align(one, mes);
        }
    }
    finish(p, me->desc);
}

// Guide-tree alignment in waves (the reference's own wave structure, node.cpp:227-285: every node whose two children
// have their sequences is ready): the ready nodes run Node::align_sequences_this_node_openmp -- the reference's code,
// model construction, anchoring and memory checks included -- on one host thread each; their Viterbi_alignment::align
// calls meet in the combiner and become ONE launch batch; path expansion and build_ancestral_sequence then run on the
// nodes' own threads, side by side.
void align_tree_in_waves(Node *root, Model_factory *mf, int n_threads, bool boost_variant) {
    (void)n_threads;  // the wave decides the width: one thread per ready node
    // The OpenMP per-node function copies its model through Evol_model::operator=, which sizes the tables for the DNA
    // alphabet whatever the data (evol_model.cpp:114-119): with codon data it reads out of bounds in the reference itself
    // (SURVEY section 5).  Codon trees therefore take the --boost per-node function, whose model is copy-constructed.
    const bool threaded_fn = boost_variant || mf->get_sequence_data_type() == Model_factory::codon || Settings_handle::st.is("codons");
    Node::number_of_nodes = root->get_number_of_leaves() - 1;
    Node::alignment_number = 1;
    std::vector<Node *> wait_nodes, run_nodes;
    root->build_queues(wait_nodes, run_nodes);
    const size_t max_wave = 256;
    for (;;) {
        for (size_t lo = 0; lo < run_nodes.size(); lo += max_wave) {
            const size_t hi = std::min(run_nodes.size(), lo + max_wave);
            {
                std::lock_guard<std::mutex> lk(combiner().m);
                combiner().expected = (int)(hi - lo);
            }
            std::vector<std::thread> pool;
            for (size_t k = lo; k < hi; k++) {
                Node *node = run_nodes[k];
                pool.emplace_back([node, mf, threaded_fn]() {
                    t_in_wave = true;
                    t_submitted = false;
                    if (threaded_fn) node->align_sequences_this_node_threaded(mf);
                    else node->align_sequences_this_node_openmp(mf);
                    if (!t_submitted) combiner().leave();
                    t_in_wave = false;
                });
            }
            for (auto &t : pool) t.join();
        }
        if (wait_nodes.empty()) break;
        run_nodes.clear();
        for (auto it = wait_nodes.begin(); it != wait_nodes.end();) {
            if ((*it)->left_child->node_has_sequence_object && (*it)->right_child->node_has_sequence_object) {
                run_nodes.push_back(*it);
                it = wait_nodes.erase(it);
            } else ++it;
        }
        if (run_nodes.empty()) break;  // a malformed tree: the reference would spin here
    }
    if (Settings_handle::st.is("output-ancestors") || Settings_handle::st.is("ancestors")) root->reconstruct_parsimony_ancestor(mf);
}

// Query placement with the trial alignments prefetched in batches (see Placement above): loads the query file the way
// Reads_aligner::align does (reads_aligner.cpp:51-75) so that the batches can look ahead, then runs the reference's
// own placement code.
void placement_begin(Reads_aligner *ra, Node *root) {
    Placement &P = placement();
    P.active = false;
    if (Settings_handle::st.is("pileup-alignment") || Settings_handle::st.is("align-reads-at-root") || Settings_handle::st.is("find-orfs") ||
        !Settings_handle::st.is("no-anchors") || !Settings_handle::st.is("queryfile"))
        return;
    if (const char *off = getenv("PAGAN2_B200_NO_PREFETCH")) if (atoi(off)) return;
    const std::string file = Settings_handle::st.get("queryfile").as<std::string>();
    Fasta_reader fr;
    std::vector<Fasta_entry> reads;
    try {
        fr.read(file, reads, true);
        fr.remove_gaps(&reads);
    } catch (ppa::IOException &e) {
        return;  // the reference reports it
    }
    const int data_type = fr.check_sequence_data_type(&reads);
    fr.check_alphabet(&reads, data_type);
    if (reads.empty()) return;
    P.ra = ra;
    P.root = root;
    P.reads.swap(reads);
    P.read_hash.assign(P.reads.size(), 0);
    P.cursor = 0;
    P.win_lo = P.win_hi = 0;
    P.items.clear();
    P.done_targets.clear();
    P.cache.clear();
    if (const char *ch = getenv("PAGAN2_B200_PREFETCH_READS")) if (atoi(ch) > 0) P.chunk = (size_t)atoi(ch);
    P.active = true;
}

void placement_end() {
    Placement &P = placement();
    P.active = false;
    P.cache.clear();
    P.items.clear();
    P.reads.clear();
}

// Model_factory::alignment_model (model_factory.cpp:1871-2230) is a pure function of the distance: the reference
// rebuilds the substitution tables for every node (node.cpp:70-71) and twice for every trial alignment of a read
// (reads_aligner.cpp:3487, 3493) -- 0.23 s per call for the 1892-state codon tables.  Here the first model built for a
// distance is kept and later calls get a deep copy (same doubles, bit for bit: they were computed by the reference's code).
namespace {
struct Model_cache {
    std::mutex m;
    std::map<std::pair<Model_factory *, uint64_t>, std::shared_ptr<Evol_model>> models;
    size_t bytes = 0;
} g_model_cache;

size_t model_bytes(const Evol_model &m) {
    const size_t f = (size_t)m.logCharPr->x;
    return f * f * (2 * sizeof(double) + sizeof(int)) + f * 2 * sizeof(double) + (size_t)m.char_as * m.char_as * sizeof(int);
}

void copy_model(const Evol_model &src, Evol_model &dst) {
    auto copy_db = [](Db_matrix *d, const Db_matrix *s) { memcpy(d->data, s->data, sizeof(double) * (size_t)s->x * (size_t)s->y); };  // (y == 1 for vectors)
    copy_db(dst.charPi, src.charPi);
    copy_db(dst.charPr, src.charPr);
    copy_db(dst.logCharPi, src.logCharPi);
    copy_db(dst.logCharPr, src.logCharPr);
    memcpy(dst.parsimony_table->data, src.parsimony_table->data, sizeof(int) * (size_t)src.parsimony_table->x * (size_t)src.parsimony_table->y);
    memcpy(dst.mostcommon_table->data, src.mostcommon_table->data, sizeof(int) * (size_t)src.mostcommon_table->x * (size_t)src.mostcommon_table->y);
    dst.data_type = src.data_type; dst.char_as = src.char_as; dst.distance = src.distance;
    dst.id_prob = src.id_prob; dst.ext_prob = src.ext_prob; dst.end_ext_prob = src.end_ext_prob; dst.break_ext_prob = src.break_ext_prob;
    dst.match_prob = src.match_prob;
    dst.log_id_prob = src.log_id_prob; dst.log_ext_prob = src.log_ext_prob; dst.log_end_ext_prob = src.log_end_ext_prob;
    dst.log_break_ext_prob = src.log_break_ext_prob; dst.log_match_prob = src.log_match_prob;
    dst.ins_rate = src.ins_rate; dst.del_rate = src.del_rate; dst.ins_prob = src.ins_prob; dst.del_prob = src.del_prob;
    dst.full_char_alphabet = src.full_char_alphabet;
    dst.ambiguity_type = src.ambiguity_type;
}
}  // namespace

// (Evol_model has no copy constructor of its own -- the implicit one shares the tables -- so, like the reference's code,
// this relies on the returned local being constructed in place: ONE return statement, naming the local.)
static Evol_model clone_cached_model(Model_factory *mf, double distance, Evol_model (*build)(Model_factory *, double));

Evol_model cached_alignment_model(Model_factory *mf, double distance, Evol_model (*build)(Model_factory *, double)) {
    Scoped_ns timed(g_ns_model);
    static const bool off = getenv("PAGAN2_B200_NO_MODEL_CACHE") && atoi(getenv("PAGAN2_B200_NO_MODEL_CACHE"));
    if (off) return build(mf, distance);
    return clone_cached_model(mf, distance, build);
}

static Evol_model clone_cached_model(Model_factory *mf, double distance, Evol_model (*build)(Model_factory *, double)) {
    uint64_t bits;
    memcpy(&bits, &distance, 8);
    const std::pair<Model_factory *, uint64_t> key(mf, bits);
    // the lookup (and a first build) under the lock; the copy -- 71 MB of tables for codon models -- outside it, so that the
    // host threads of a wave copy side by side (the master stays alive through the shared pointer if the cache is emptied)
    std::shared_ptr<Evol_model> master;
    {
        std::lock_guard<std::mutex> lk(g_model_cache.m);
        auto it = g_model_cache.models.find(key);
        if (it == g_model_cache.models.end()) {
            Evol_model built = build(mf, distance);  // (as the reference's callers take it)
            master = std::make_shared<Evol_model>(built.data_type, built.distance);
            copy_model(built, *master);
            const size_t b = model_bytes(*master);
            if (g_model_cache.bytes + b > ((size_t)2 << 30)) {  // bounded: a tree of distinct branch lengths gains nothing from the cache
                g_model_cache.models.clear();
                g_model_cache.bytes = 0;
            }
            g_model_cache.bytes += b;
            g_model_cache.models.emplace(key, master);
        } else {
            master = it->second;
            engine().totals.model_cache_hits++;
        }
    }
    Evol_model out(master->data_type, master->distance);
    copy_model(*master, out);
    return out;
}


// ------------------------------------------------------------------------------------------------------------------
// Prefix anchors (utils/find_anchors.cpp:35-127)
// ------------------------------------------------------------------------------------------------------------------
void find_long_substrings(Find_anchors *fa, std::string *seq1, std::string *seq2, std::vector<Substring_hit> *hits, int min_length,
                          void (*reference_fn)(Find_anchors *, std::string *, std::string *, std::vector<Substring_hit> *, int)) {
    Scoped_ns timed(g_ns_anchor);
    g_anchor_calls++;
    static const bool use_reference = getenv("PAGAN2_B200_REF_ANCHORS") && atoi(getenv("PAGAN2_B200_REF_ANCHORS"));
    if (use_reference || !hits->empty()) { reference_fn(fa, seq1, seq2, hits, min_length); return; }
    fa->len1 = (int)seq1->length();
    fa->len2 = (int)seq2->length();
    std::vector<pg2_anchor_hit> found((size_t)std::max(1024, std::min(fa->len1, fa->len2) / 16));
    int32_t n = 0;
    int rc = pg2_find_prefix_anchors(seq1->data(), fa->len1, seq2->data(), fa->len2, min_length, found.data(), (int32_t)found.size(), &n);
    if (rc == PG2_ERR_CAPACITY) {
        found.resize((size_t)n);
        rc = pg2_find_prefix_anchors(seq1->data(), fa->len1, seq2->data(), fa->len2, min_length, found.data(), (int32_t)found.size(), &n);
    }
    if (rc != PG2_OK) {
        Log_output::write_out("pagan2_b200: pg2_find_prefix_anchors failed (" + std::to_string(rc) + ")\n", 0);
        exit(1);
    }
    hits->reserve((size_t)n);
    for (int32_t k = 0; k < n; k++) {
        Substring_hit s;  // (both strands plus by construction, substring_hit.h:40)
        s.start_site_1 = found[(size_t)k].start_1;
        s.start_site_2 = found[(size_t)k].start_2;
        s.length = found[(size_t)k].length;
        s.score = found[(size_t)k].length;
        hits->push_back(s);
    }
}


void define_tunnel(Find_anchors *fa, std::vector<Substring_hit> *hits, std::vector<int> *upper_bound, std::vector<int> *lower_bound,
                   std::string *str1, std::string *str2,
                   void (*reference_fn)(Find_anchors *, std::vector<Substring_hit> *, std::vector<int> *, std::vector<int> *, std::string *,
                                        std::string *)) {
    Scoped_ns timed(g_ns_anchor);
    static const bool use_reference = getenv("PAGAN2_B200_REF_ANCHORS") && atoi(getenv("PAGAN2_B200_REF_ANCHORS"));
    if (use_reference || Settings_handle::st.is("plot-anchors-for-R")) { reference_fn(fa, hits, upper_bound, lower_bound, str1, str2); return; }
    const int width = Settings_handle::st.get("anchors-offset").as<int>();
    std::vector<pg2_anchor_hit> h(hits->size());
    for (size_t k = 0; k < hits->size(); k++) {
        h[k].start_1 = hits->at(k).start_site_1;
        h[k].start_2 = hits->at(k).start_site_2;
        h[k].length = hits->at(k).length;
    }
    const int len1 = (int)str1->length();
    std::vector<int32_t> up((size_t)len1 + 1), lo((size_t)len1 + 1);
    int rc = pg2_anchor_band(h.data(), (int32_t)h.size(), str1->data(), len1, str2->data(), (int)str2->length(), width, up.data(), lo.data());
    if (rc != PG2_OK) { reference_fn(fa, hits, upper_bound, lower_bound, str1, str2); return; }  // (out-of-range hits: the reference throws)
    // the reference appends the upper bounds and puts every lower bound in FRONT of what the vector holds (:391, :419)
    upper_bound->insert(upper_bound->end(), up.begin(), up.end());
    lower_bound->insert(lower_bound->begin(), lo.begin(), lo.end());
}

}  // namespace ppa_b200
