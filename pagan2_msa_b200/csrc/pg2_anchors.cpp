// pg2_anchors.cpp -- prefix anchors (--use-prefix-anchors): the long exact substrings two sequences share, from which the
// reference builds the anchor band of an alignment (Find_anchors::find_long_substrings, src/utils/find_anchors.cpp:35-127,
// called from Viterbi_alignment::define_tunnel, src/main/viterbi_alignment.cpp:70-74).
//
// Why this is here: with the DP on the device, anchoring became 85 % of the wall time of an anchored
// 200 kb alignment (3.5 s of host time per pair against 0.25 s of fill + traceback).  The reference's function spends it in ONE
// place: every overlapping hit is removed from the middle of the hit vector with vector::erase (:106-121) -- 155 000 hits,
// 1 300 survivors, 10^10 element moves.  The search below makes the same decisions in the same order -- the same suffix order
// (the C library's qsort over strcmp of the same strings, so that even ties between identical suffixes fall the same way), the
// same adjacent-pair scan, the same std::sort by length, the same greedy walk -- and keeps the accepted hits with a write
// index instead: 0.17 s per 200 kb pair, identical hits (tests/test_anchors.py compares with the reference itself).
//
// Host code: the hits feed the reference's own host functions (check_hits_order_conflict, define_tunnel, :320-489), and
// what the device needs from them is the band (pg2_job.upper / lower).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/pagan2_b200.h"

namespace {

int suffix_cmp(const void *p, const void *q) { return strcmp(*(const char *const *)p, *(const char *const *)q); }

}  // namespace

extern "C" int pg2_find_prefix_anchors(const char *seq1, int32_t len1, const char *seq2, int32_t len2, int32_t min_length,
                                       pg2_anchor_hit *hits, int32_t cap, int32_t *n_hits) {
    if (!seq1 || !seq2 || len1 < 0 || len2 < 0 || !n_hits || (cap > 0 && !hits)) return PG2_ERR_INVALID;
    *n_hits = 0;
    const size_t m = (size_t)len1 + (size_t)len2;
    if (m == 0) return PG2_OK;
    // both sequences as C strings; every suffix of either is a pointer into them (find_anchors.cpp:43-61)
    std::vector<char> text(m + 2);
    char *c1 = text.data(), *c2 = c1 + len1 + 1;
    memcpy(c1, seq1, (size_t)len1);
    c1[len1] = 0;
    memcpy(c2, seq2, (size_t)len2);
    c2[len2] = 0;
    // The reference keeps both texts in variable-length stack arrays (`char c1[len1]; char c2[len2];`, :43-44) and writes each
    // terminator one element past the end (:52, :61).  As g++ lays them out on x86-64 (each array rounded up to 16 bytes, c2
    // directly below c1) the terminator lands in the array's own padding -- except when len2 is a multiple of 16: then c2[len2]
    // IS c1[0], the first character of sequence 1 becomes the terminator, and the suffix that starts there is empty (a hit at
    // (0, 0) shrinks to (1, 1) or vanishes).  The hits must be the ones the reference program finds, so that is reproduced
    // (tests/test_anchors.py::test_matches_live_reference_all_length_residues pins it against the reference built here).
    if (len2 % 16 == 0 && len1 > 0) c1[0] = 0;
    std::vector<char *> a(m);
    for (int32_t n = 0; n < len1; n++) a[(size_t)n] = c1 + n;
    for (int32_t n = 0; n < len2; n++) a[(size_t)len1 + (size_t)n] = c2 + n;
    qsort(a.data(), m, sizeof(char *), suffix_cmp);  // :66
    // neighbours in suffix order that come from different sequences and share a prefix of min_length or more (:68-85)
    std::vector<pg2_anchor_hit> found;
    auto in1 = [&](const char *p) { return p >= c1 && p < c1 + len1; };
    auto in2 = [&](const char *p) { return p >= c2 && p < c2 + len2; };
    for (size_t i = 0; i + 1 < m; i++) {
        const char *p = a[i], *q = a[i + 1];
        if (!((in1(p) && in2(q)) || (in1(q) && in2(p)))) continue;
        int32_t length = 0;
        for (const char *x = p, *y = q; *x && *x == *y; ++x, ++y) length++;  // identical_prefix_length (find_anchors.h:99-105)
        if (length >= min_length) {
            pg2_anchor_hit h;
            h.start_1 = (int32_t)((in1(p) ? p : q) - c1);
            h.start_2 = (int32_t)((in2(p) ? p : q) - c2);
            h.length = length;
            found.push_back(h);
        }
    }
    // longest first (:87; std::sort as the reference calls it: the order of equally long hits follows from the comparisons alone)
    std::sort(found.begin(), found.end(), [](pg2_anchor_hit p, pg2_anchor_hit q) { return p.length > q.length; });
    // greedy choice in that order: a hit that touches a site an accepted hit covers is dropped (:89-125)
    std::vector<char> hit_site1((size_t)len1, 0), hit_site2((size_t)len2, 0);
    size_t kept = 0;
    for (size_t k = 0; k < found.size(); k++) {
        const pg2_anchor_hit h = found[k];
        bool overlap = false;
        for (int32_t i = h.start_1, j = h.start_2; i < h.start_1 + h.length && j < h.start_2 + h.length; i++, j++)
            if (hit_site1[(size_t)i] || hit_site2[(size_t)j]) { overlap = true; break; }
        if (overlap) continue;
        for (int32_t i = h.start_1, j = h.start_2; i < h.start_1 + h.length && j < h.start_2 + h.length; i++, j++) {
            hit_site1[(size_t)i] = 1;
            hit_site2[(size_t)j] = 1;
        }
        found[kept++] = h;
    }
    *n_hits = (int32_t)kept;
    if ((int64_t)kept > (int64_t)cap) return PG2_ERR_CAPACITY;
    for (size_t k = 0; k < kept; k++) hits[k] = found[k];
    return PG2_OK;
}

// Hits -> anchor band (Find_anchors::define_tunnel, src/utils/find_anchors.cpp:320-435): per DP row i of sequence 1 (0 ..
// len1, gaps of the sequence strings counted) the inclusive column range [upper[i], lower[i]] the alignment may use -- `width`
// columns around the anchored diagonals, widening to the full rectangle between anchors.  The same two sweeps as the reference;
// the lower bounds are written in place where the reference inserts every value at the FRONT of its vector (:419, quadratic:
// 2 s per 200 kb alignment).
extern "C" int pg2_anchor_band(const pg2_anchor_hit *hits, int32_t n_hits, const char *str1, int32_t len1, const char *str2, int32_t len2,
                               int32_t width, int32_t *upper, int32_t *lower) {
    if (n_hits < 0 || len1 < 0 || len2 < 0 || width < 0 || (n_hits > 0 && !hits) || !str1 || !str2 || !upper || !lower) return PG2_ERR_INVALID;
    // positions of the characters in the gapped strings (the hits count characters, the band counts sites; :329-338)
    std::vector<int32_t> index1, index2;
    index1.reserve((size_t)len1);
    index2.reserve((size_t)len2);
    for (int32_t i = 0; i < len1; i++) if (str1[i] != '-') index1.push_back(i + 1);
    for (int32_t i = 0; i < len2; i++) if (str2[i] != '-') index2.push_back(i + 1);
    std::vector<int32_t> diagonals((size_t)len1 + 1, -1);
    for (int32_t h = 0; h < n_hits; h++) {
        const pg2_anchor_hit &hit = hits[h];
        if (hit.start_1 < 0 || hit.start_2 < 0 || hit.length < 0 || (size_t)hit.start_1 + (size_t)hit.length > index1.size() ||
            (size_t)hit.start_2 + (size_t)hit.length > index2.size())
            return PG2_ERR_INVALID;  // (the reference's vector::at would throw)
        for (int32_t i = 0; i < hit.length; i++) diagonals[(size_t)index1[(size_t)(hit.start_1 + i)]] = index2[(size_t)(hit.start_2 + i)];
        const size_t after = (size_t)hit.start_1 + (size_t)hit.length;
        if (after < index1.size() && index1[after] < (int32_t)diagonals.size()) diagonals[(size_t)index1[after]] = -2;  // :354-355
    }
    // upper bounds, top down (:360-394)
    {
        int32_t y1 = 0, y2 = 0, prev_y = 0, m_count = 0;
        for (int32_t i = 0; i <= len1; i++) {
            if (i >= width && diagonals[(size_t)(i - width)] >= 0) y1 = diagonals[(size_t)(i - width)];
            if (diagonals[(size_t)i] >= 0) y2 = diagonals[(size_t)i] - width;
            const bool run = diagonals[(size_t)i] >= 0 && i > 0 && diagonals[(size_t)(i - 1)] + 1 == diagonals[(size_t)i];
            if (run) m_count++;
            else if (diagonals[(size_t)i] == -2) m_count = 0;
            int32_t y = std::max(std::min(y1, y2), 0);
            if (run && m_count >= width) prev_y = y;
            y = std::max(std::min(y, prev_y), 0);
            upper[i] = y;
        }
    }
    // lower bounds, bottom up (:396-420)
    {
        int32_t y1 = len2, y2 = len2, prev_y = len2, m_count = 0;
        for (int32_t i = len1; i >= 0; i--) {
            if (i <= len1 - width && diagonals[(size_t)(i + width)] >= 0) y1 = diagonals[(size_t)(i + width)];
            if (diagonals[(size_t)i] >= 0) y2 = diagonals[(size_t)i] + width;
            const bool run = diagonals[(size_t)i] >= 0 && i < len1 && diagonals[(size_t)(i + 1)] - 1 == diagonals[(size_t)i];
            if (run) m_count++;
            else if (diagonals[(size_t)i] == -2) m_count = 0;
            int32_t y = std::min(std::max(y1, y2), len2);
            if (run && m_count >= width) prev_y = y;
            y = std::min(std::max(y, prev_y), len2);
            lower[i] = y;
        }
    }
    return PG2_OK;
}
