// dropin_interposer.cpp -- turns the UNMODIFIED reference objects into `pagan2_b200`: GNU ld --wrap redirects every call
// of ppa::Viterbi_alignment::align (src/main/viterbi_alignment.cpp:187) to ppa_b200::align_on_device.  This is the
// binding a maintainer would otherwise make by editing align() itself (INTEGRATION.md shows that edit); the link-time
// form needs no change to the reference tree and is what tests/test_dropin_e2e.py runs against `pagan2_ref`.
//
// Also here: the two translation units of the reference that cannot be built in this image (Exonerate wrapper:
// boost::regex + an external binary; version check: libcurl) as empty stubs.  Neither touches the alignment path.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "viterbi_alignment_b200.h"
#include "main/node.h"
#include "main/reads_aligner.h"
#include "utils/check_version.h"
#include "utils/exonerate_queries.h"
#include "utils/find_anchors.h"

using namespace std;
using namespace ppa;

Exonerate_queries::Exonerate_queries() {}
bool Exonerate_queries::test_executable() { return false; }
void Exonerate_queries::local_alignment(map<string, string> *, Fasta_entry *, map<string, hit> *, bool, bool) {}
void Exonerate_queries::local_alignment(Node *, Fasta_entry *, std::multimap<std::string, std::string> *, std::map<std::string, hit> *, bool,
                                        bool, bool) {}
void Exonerate_queries::preselect_targets(map<string, string> *, vector<Fasta_entry> *, map<string, string> *,
                                          map<string, multimap<string, hit> > *, bool) {}
void Exonerate_queries::local_pairwise_alignment(string *, string *, vector<Substring_hit> *, int *) {}
Check_version::Check_version(float) {}

#define ALIGN_SYM _ZN3ppa17Viterbi_alignment5alignEPNS_8SequenceES2_PNS_10Evol_modelEffb
#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)

extern "C" void CAT(__wrap_, ALIGN_SYM)(Viterbi_alignment *self, Sequence *left, Sequence *right, Evol_model *model, float lbl, float rbl,
                                       bool is_reads) {
    ppa_b200::align_on_device(self, left, right, model, lbl, rbl, is_reads);
}

// The schedulers.  `--threads N` (N > 1) sends the guide-tree traversal through Node::start_openmp_alignment or, with
// --boost, Node::start_threaded_alignment (main.cpp:185-191): both become the wave scheduler.  Reads_aligner::align
// (main.cpp) is bracketed by the placement prefetch.  Both reach the reference's own code for everything but the DP.
#define OMP_SYM _ZN3ppa4Node22start_openmp_alignmentEPNS_13Model_factoryEi
#define THR_SYM _ZN3ppa4Node24start_threaded_alignmentEPNS_13Model_factoryEi
#define RAL_SYM _ZN3ppa13Reads_aligner5alignEPNS_4NodeEPNS_13Model_factoryEi
extern "C" void CAT(__wrap_, OMP_SYM)(Node *root, Model_factory *mf, int n_threads) { ppa_b200::align_tree_in_waves(root, mf, n_threads); }
extern "C" void CAT(__wrap_, THR_SYM)(Node *root, Model_factory *mf, int n_threads) { ppa_b200::align_tree_in_waves(root, mf, n_threads, true); }
extern "C" void CAT(__real_, RAL_SYM)(Reads_aligner *self, Node *root, Model_factory *mf, int count);
extern "C" void CAT(__wrap_, RAL_SYM)(Reads_aligner *self, Node *root, Model_factory *mf, int count) {
    ppa_b200::placement_begin(self, root);
    CAT(__real_, RAL_SYM)(self, root, mf, count);
    ppa_b200::placement_end();
}

// Model_factory::alignment_model: models kept by distance (the reference rebuilds them per node and per trial alignment)
#define AMD_SYM _ZN3ppa13Model_factory15alignment_modelEd
extern "C" Evol_model CAT(__real_, AMD_SYM)(Model_factory *self, double distance);
extern "C" Evol_model CAT(__wrap_, AMD_SYM)(Model_factory *self, double distance) {
    return ppa_b200::cached_alignment_model(self, distance, &CAT(__real_, AMD_SYM));
}

// Find_anchors::find_long_substrings (--use-prefix-anchors): same hits, linear hit thinning
#define FLS_SYM _ZN3ppa12Find_anchors20find_long_substringsEPNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEES7_PSt6vectorINS_13Substring_hitESaIS9_EEi
extern "C" void CAT(__real_, FLS_SYM)(Find_anchors *self, std::string *s1, std::string *s2, std::vector<Substring_hit> *hits, int min_length);
extern "C" void CAT(__wrap_, FLS_SYM)(Find_anchors *self, std::string *s1, std::string *s2, std::vector<Substring_hit> *hits, int min_length) {
    ppa_b200::find_long_substrings(self, s1, s2, hits, min_length, &CAT(__real_, FLS_SYM));
}

// Find_anchors::define_tunnel(hits, upper, lower, str1, str2): the band from the hits, linear time
#define FDT_SYM _ZN3ppa12Find_anchors13define_tunnelEPSt6vectorINS_13Substring_hitESaIS2_EEPS1_IiSaIiEES8_PNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEESF_
extern "C" void CAT(__real_, FDT_SYM)(Find_anchors *self, std::vector<Substring_hit> *hits, std::vector<int> *ub, std::vector<int> *lb, std::string *s1,
                                     std::string *s2);
extern "C" void CAT(__wrap_, FDT_SYM)(Find_anchors *self, std::vector<Substring_hit> *hits, std::vector<int> *ub, std::vector<int> *lb, std::string *s1,
                                     std::string *s2) {
    ppa_b200::define_tunnel(self, hits, ub, lb, s1, s2, &CAT(__real_, FDT_SYM));
}

namespace {
struct Stats_at_exit {
    ~Stats_at_exit() {
        const char *p = getenv("PAGAN2_B200_STATS");
        if (p && *p) {
            ppa_b200::Totals t = ppa_b200::totals();
            FILE *f = fopen(p, "w");
            if (f) {
                fprintf(f, "{\"jobs\": %lld, \"cells\": %lld, \"batches\": %lld, \"fill_ms\": %.6f, \"traceback_ms\": %.6f, "
                           "\"wave_batches\": %lld, \"prefetch_batches\": %lld, \"cache_hits\": %lld, \"sharded_batches\": %lld, "
                           "\"model_cache_hits\": %lld, \"host_ms\": {\"stage\": %.3f, \"engine_calls\": %.3f, \"expand_path\": %.3f, "
                           "\"build_ancestral_sequence\": %.3f, \"alignment_model\": %.3f, \"prefix_anchors\": %.3f}, \"anchor_calls\": %lld}\n",
                        t.jobs, t.cells, t.batches, t.fill_ms, t.traceback_ms, t.wave_batches, t.prefetch_batches, t.cache_hits, t.sharded_batches,
                        t.model_cache_hits, t.host_stage_ms, t.host_engine_ms, t.host_expand_ms, t.host_build_ms, t.host_model_ms, t.host_anchor_ms, t.anchor_calls);
                fclose(f);
            }
        }
    }
} g_stats_at_exit;
}  // namespace
