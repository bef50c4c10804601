// pairlist.h -- device-resident cluster-pair list of a context and the launchers that build
// and consume it (internal).
#pragma once

#include "nblist_core.h"
#include "sdm_kernels.h"

namespace sdm {

struct Unit {
    int sci;      // supercluster
    int begin;    // first entry (global index into entries)
    int end;      // one past the last entry
    int pad;
};

// Everything the pair kernel reads.
struct PairListView {
    nbl::Grid G;
    const float4* posq;          // [nslot] sorted, wrapped (+ image) positions, .w = q*sqrt(K)
    const float2* par;           // [nslot] (sigma/2, 2*sqrt(eps)); (0,0) for dummies
    const int* atom;             // [nslot] replica*n + atom, or -1 for a dummy slot
    const nbl::SciDesc* sci;     // [nsci]
    const uint2* entries;        // {cj | shift<<26, imask | mask_index<<8}
    const uint32_t* masks;       // [(nmasks+1)*16]; set 0 = all ones
    const Unit* units;           // [nunits]
    int nunits;
    int nslot_cap;               // accumulator plane stride
};

// Tile list: the entries expanded into one record per (i-cluster, j-cluster) tile, grouped by
// i-cluster, for pair_tile_kernel.  Within an i-cluster the tiles of masked entries come first;
// both groups are padded to an even count with dummy records (shift code 63, cluster 0, mask 0).
struct TileUnit {
    int islot;    // first slot of the i-cluster
    int begin;    // first record (index into recs)
    int nrec;     // records of this unit (even, <= chunk)
    int nmask;    // how many of them (from the start) carry exclusion masks (even)
};

struct TileListView {
    const uint2* recs;           // {cj | shift<<26, index of the tile's two mask words (0 = all ones)}
    const TileUnit* units;
    int nunits;
};

void launch_pair_tiles(const Topology& T, const PairListView& V, const TileListView& TL,
                       const double* pos_all, long long* f1acc, double* epart, long long* cpart,
                       int exact, int* unit_counter, int num_sms, cudaStream_t s);

void launch_pair_cluster(const Topology& T, const PairListView& V, const double* pos_all,
                         long long* f1acc, double* epart, long long* cpart, int exact,
                         int* unit_counter, int num_sms, cudaStream_t s);
// Debug / parity: dump the in-cutoff pairs of one replica exactly as the pair kernel decides them.
void launch_pair_emit(const Topology& T, const PairListView& V, const double* pos_all, int exact,
                      int* emit_counter, int* emit_pairs, int emit_cap, int emit_replica,
                      cudaStream_t s);

// Per-eval refresh of the sorted positions from the current double positions (same periodic
// image as at build time) + staleness check against the build-time positions.
void launch_refresh(const Topology& T, const nbl::Grid& G, int nslot, const double* pos_all,
                    const int* atom, const int* img, const float4* posq_build, float4* posq,
                    float half_skin2, int* flags, int* list_age, cudaStream_t s);

}  // namespace sdm
