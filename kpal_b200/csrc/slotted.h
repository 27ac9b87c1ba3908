// The SLOTTED packed stream of a FASTA text (pack.cpp), shared with the hybrid upload of
// kpal_count_fasta (cabi.cu).
//
// The text is cut at line starts into segments of about `seg` bytes.  Segment j owns the
// bases [slot_j, slot_{j+1}) of the stream, slot_j = align64(cut_j) + 64 j: a byte emits at
// most one base, so slot_j lies behind everything the text before cut_j can emit, and any
// segment can be packed without knowing how much the others emit -- by another thread, or by
// the device.  Unused slot ends are invalid.  A cut may fall inside a record; the windows
// that span it are restored by a JUNCTION record of its own (64 bases each, behind the last
// slot): the last k - 1 positions emitted before the cut followed by the first k - 1 after it
// hold exactly the k - 1 windows that cross the cut, each once.
#pragma once
#include <stdint.h>

#include <vector>

namespace kpal {

struct SlottedPlan {
    std::vector<uint64_t> cut, slot;      // m + 1 entries: byte offsets (line starts) and first bases
    uint64_t m = 0;                       // segments
    uint64_t first_header = 0;            // offset of the first header line ('>' at a line start)
    // bases of the whole stream when the segments from `first` on are packed as slots (+ their junctions)
    uint64_t stream_bases(uint64_t first) const { return slot[m] + 64 * (m - first); }
    uint64_t junction_slot(uint64_t first, uint64_t j) const { return slot[m] + 64 * (j - first); }
};

// false: the text is not one for slots (no header within the first MB, lines longer than 64 KB)
bool slotted_plan(const char *fasta, uint64_t n_bytes, uint64_t seg, SlottedPlan &plan);

// Packs segment j into its slot and (j > 0) the junction record of its cut; codes / valid
// point at base `base` of the stream (a multiple of 64).  Returns the bases emitted.
uint64_t slotted_pack_segment(const SlottedPlan &plan, const unsigned char *text, uint64_t n_bytes, uint64_t j,
                              uint64_t first, int k, uint32_t *codes, uint32_t *valid, uint64_t base);

bool fasta_segment_fast();              // the 32-bytes-per-step form is available on this host

}  // namespace kpal
