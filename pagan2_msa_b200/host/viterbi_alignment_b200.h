// viterbi_alignment_b200.h -- host-side mirror of the reference's alignment seam over the C-ABI.
//
// The reference has no plugin API; the seam `Node` uses is the class surface of ppa::Viterbi_alignment
// (reference src/main/node.cpp:77-159):
//
//     Viterbi_alignment va;  va.define_tunnel(...);  va.align(left, right, &model, lbl, rbl, is_reads);
//     Sequence *anc = va.get_simple_sequence();
//
// This header gives that seam a device back end with the same names, argument meaning and error behaviour:
//
//   ppa_b200::align_on_device(va, left, right, model, lbl, rbl, is_reads)
//       the body of Viterbi_alignment::align (src/main/viterbi_alignment.cpp:187-465) with the matrix allocation,
//       fill loops, end-corner scan and backtrack (:238-383) replaced by pg2_align_batch + pg2_expand_path.
//       Settings (:191-231) and build_ancestral_sequence (:389-392) are the reference's own member functions,
//       called on the caller's object, so `va` ends in the same state: path, ancestral_sequence, used-edge marks
//       on `left` / `right`.
//   ppa_b200::Alignment_batch
//       the launch-batch form for the schedulers (a guide-tree wave, node.cpp:240-264; the trial / final
//       alignments of many reads, reads_aligner.cpp:983-1216): add() jobs, run() them in ONE pg2_align_batch,
//       then every `va` is finished exactly as above, in the order the jobs were added.
//   ppa_b200::align_tree_in_waves(root, mf, n_threads)
//       the batch scheduler of the guide-tree traversal: stands where Node::start_openmp_alignment /
//       Node::start_threaded_alignment stand (node.cpp:196-285).  Every ready node of a wave runs the reference's own
//       per-node code on a host thread; their align() calls are combined into one launch batch per wave.
//   ppa_b200::placement_begin(ra, root) / placement_end()
//       bracket Reads_aligner::align (reads_aligner.cpp:35): while active, a trial alignment of query placement that
//       is not cached yet triggers ONE launch batch over the next reads of the query file x every candidate node; the
//       reference's own loops then find their alignments in the cache and only post-process them, in their own order.
//   With PAGAN2_B200_DEVICES=0-7 a launch batch is cut by index range over the devices (one pg2_ctx and one host
//   thread each) and the results meet in host memory.
//
// Modes that are not part of the device contract (--full-probability, --sample-path,
// --sample-additional-paths, posterior plots) are refused with the reference's own style of fatal message; there
// is no silent CPU path.
#ifndef PAGAN2_VITERBI_ALIGNMENT_B200_H
#define PAGAN2_VITERBI_ALIGNMENT_B200_H

// the reference headers rely on their includers for the standard headers
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <stack>
#include <string>
#include <vector>

#include "main/sequence.h"
#include "main/viterbi_alignment.h"
#include "utils/evol_model.h"

namespace ppa {
class Node;
class Model_factory;
class Reads_aligner;
}

namespace ppa_b200 {

// One pending alignment: the arguments of Viterbi_alignment::align plus the object they were made on.
struct Alignment_job {
    ppa::Viterbi_alignment *va;
    ppa::Sequence *left, *right;
    ppa::Evol_model *model;
    float left_branch_length, right_branch_length;
    bool is_reads_sequence;
};

class Alignment_batch {
   public:
    void add(ppa::Viterbi_alignment *va, ppa::Sequence *left, ppa::Sequence *right, ppa::Evol_model *model, float l_branch_length = 0,
             float r_branch_length = 0, bool is_reads_sequence = false);
    // Aligns every added job on the device and finishes every Viterbi_alignment (path, ancestral sequence, edge
    // marks).  Exits like the reference on a broken traceback (viterbi_alignment.cpp:1167-1171).
    void run();
    size_t size() const { return jobs.size(); }

   private:
    std::vector<Alignment_job> jobs;
};

// Drop-in body of Viterbi_alignment::align for one alignment.
void align_on_device(ppa::Viterbi_alignment *va, ppa::Sequence *left, ppa::Sequence *right, ppa::Evol_model *model,
                     float l_branch_length = 0, float r_branch_length = 0, bool is_reads_sequence = false);

// CUDA device the engine binds to (default 0; also PAGAN2_B200_DEVICE).  Call before the first alignment.
void set_device(int device);

// Guide-tree alignment in waves: one launch batch per wave of ready nodes (node.cpp:227-285).
void align_tree_in_waves(ppa::Node *root, ppa::Model_factory *mf, int n_threads, bool boost_variant = false);

// Query placement: batched prefetch of the trial alignments while Reads_aligner::align runs.
void placement_begin(ppa::Reads_aligner *ra, ppa::Node *root);
void placement_end();

// Model_factory::alignment_model with the models kept by distance (model_factory.cpp:1871): `build`
// is the reference's own function; a later call for the same distance gets a deep copy of what it returned.
ppa::Evol_model cached_alignment_model(ppa::Model_factory *mf, double distance, ppa::Evol_model (*build)(ppa::Model_factory *, double));

// Find_anchors::find_long_substrings (utils/find_anchors.cpp:35-127) over pg2_find_prefix_anchors: the same
// hits in the same order, without the reference's quadratic erase loop.  `reference_fn` is the reference's own function; it is
// called instead when PAGAN2_B200_REF_ANCHORS is set or the hit vector is not empty on entry.
void find_long_substrings(ppa::Find_anchors *fa, std::string *seq1, std::string *seq2, std::vector<ppa::Substring_hit> *hits, int min_length,
                          void (*reference_fn)(ppa::Find_anchors *, std::string *, std::string *, std::vector<ppa::Substring_hit> *, int));

// Find_anchors::define_tunnel (utils/find_anchors.cpp:320-435) over pg2_anchor_band: the same band in linear time.
void define_tunnel(ppa::Find_anchors *fa, std::vector<ppa::Substring_hit> *hits, std::vector<int> *upper_bound, std::vector<int> *lower_bound,
                   std::string *str1, std::string *str2,
                   void (*reference_fn)(ppa::Find_anchors *, std::vector<ppa::Substring_hit> *, std::vector<int> *, std::vector<int> *, std::string *,
                                        std::string *));

// Device-side totals since process start (for the drop-in binary's stats file, PAGAN2_B200_STATS).
struct Totals {
    long long jobs, cells, batches;
    double fill_ms, traceback_ms;
    long long wave_batches;      // launch batches that combined the alignments of a guide-tree wave
    long long prefetch_batches;  // launch batches of prefetched placement trial alignments
    long long cache_hits;        // align() calls served from a prefetched batch
    long long sharded_batches;   // launch batches cut over more than one device
    long long model_cache_hits;  // Model_factory::alignment_model calls answered with a copy of an earlier model
    // host wall time spent inside the host mirror, summed over the calling threads (ms): settings + graph packing, the engine calls
    // (packing of the launch batch, upload, kernels, copy back -- the caller waits), path expansion + edge marks, the reference's own
    // build_ancestral_sequence, and Model_factory::alignment_model (builds and copies)
    double host_stage_ms, host_engine_ms, host_expand_ms, host_build_ms, host_model_ms;
    double host_anchor_ms;       // find_long_substrings
    long long anchor_calls;
};
Totals totals();

}  // namespace ppa_b200

#endif
