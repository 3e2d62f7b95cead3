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

// Device-side totals since process start (for the drop-in binary's --b200-stats line).
struct Totals {
    long long jobs, cells, batches;
    double fill_ms, traceback_ms;
};
Totals totals();

}  // namespace ppa_b200

#endif
