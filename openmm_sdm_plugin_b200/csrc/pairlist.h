// pairlist.h -- what the pair kernel reads of a context's device-resident pair list, and the
// launchers of the kernels that consume it (internal).
#pragma once

#include "nblist_core.h"
#include "sdm_kernels.h"

namespace sdm {

// Work unit of the row kernel: an i-group and a chunk (<= 32 warp steps) of its row of j-atoms.
struct RowUnit {
    int c0n;      // first cluster of the i-group | clusters in it (1..2) << 28
    int begin;    // first row entry (global index into jent)
    int end;      // one past the last
    int seg;      // (mend - begin) | (lend - begin) << 16: entries [begin, mend) carry an allow word
                  // (exclusions / triangle), entries [lend, end) are j-atoms WITHOUT a Lennard-Jones
                  // term (epsilon == 0: water hydrogens) -- their steps skip a third of the arithmetic
};

// Everything the pair kernel reads.
struct PairListView {
    nbl::Grid G;
    const float4* posq;          // [nslot] sorted, wrapped (+ image) positions, .w = q*sqrt(K)
    const float2* par;           // [nslot] (sigma/2, 2*sqrt(eps)); (0,0) for dummies
    const float4* jrec;          // [nslot][2] the same two records side by side (32 bytes = one sector per atom):
                                 // what the pair kernel gathers, one address and one sector per j-atom
    const int* atom;             // [nslot] replica*n + atom, or -1 for a dummy slot
    int nslot_cap;               // accumulator plane stride
    // per-atom j rows (nblist_core.h stage 5)
    const uint32_t* jent;        // row entries: j slot | shift << 26
    const uint16_t* jallow;      // allow word per row entry (read for the masked prefix of a row only)
    const RowUnit* runits;       // [nrunits]
    const int* runit_order;      // [nrunits] the order in which the warps draw the units (longest first), or nullptr
    const int* nrunits;          // device word: units of the current list (0 when its build overflowed)
    int nrunits_ub;              // host-side upper bound (grid sizing)
    int row_group;               // clusters per i-group (1 or 2)
    int dummy_slot;              // a slot that holds a far-away dummy atom (padding lanes)
};

// The pair kernel.  emit != nullptr selects the debug build of the SAME kernel, which also records
// the System indices (i < j) of every pair it accepts for replica emit->replica -- the bit-exact
// pair-set tests read the hot kernel's own decisions, not a re-implementation.
struct PairEmit {
    int* counter;    // device: pairs accepted so far
    int* pairs;      // device: [2*cap]
    int cap;
    int replica;
};
void launch_pair_rows(const Topology& T, const PairListView& V, const double* pos_all,
                      long long* f1acc, double* epart, long long* cpart, int exact,
                      int* unit_counter, int num_sms, const PairEmit* emit, int reserve, cudaStream_t s);

// Per-eval refresh of the sorted positions from the current double positions (same periodic
// image as at build time) + staleness check against the build-time positions; max_disp2 (may be
// null) receives the largest squared displacement since the build, as float bits.
void launch_refresh(const Topology& T, const nbl::Grid& G, const int* d_nslot, int nslot_ub, const double* pos_all,
                    const int* atom, const int* img, const float4* posq_build, float4* posq, float4* jrec,
                    float half_skin2, int* flags, int* list_age, unsigned int* max_disp2, long long* acc_to_clear,
                    cudaStream_t s);

}  // namespace sdm
