// nblist_hostcheck.cpp -- CPU checker for the cluster-pair list logic (TEST INFRASTRUCTURE).
//
// Compiles openmm_sdm_plugin_b200/csrc/nblist_core.h -- the very per-item bodies the device
// kernels call -- with g++, drives them serially in the same order pairlist.cu does on the
// device, then walks the list exactly like the pair kernel's (sci, entry, ci, lane) traversal
// and reports every atom pair that passes the masks and the (double-precision) cutoff test.
// tests/test_nblist_host.py compares that with the oracle's pair set: every in-cutoff
// non-excluded pair must be covered exactly once.  Nothing here is part of the product path.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../openmm_sdm_plugin_b200/csrc/nblist_core.h"

using namespace sdm::nbl;

namespace {

bool setup_grid(Grid& G, int n, int R, bool periodic, double rlist, const double lo[3],
                const double ext[3], long long cell_cap) {
    // the very function pairlist.cu sizes its grid with; SDMB200_LAYOUT=cells selects the
    // geometric 3-D cells like there
    const char* e = std::getenv("SDMB200_LAYOUT");
    return size_grid(G, n, R, periodic, rlist, lo, ext, cell_cap, !(e && std::string(e) == "cells"));
}

}  // namespace

extern "C" {


// pos: [R][n][3] doubles.  excl: unique pairs a<b.  Returns the number of covered pairs of
// replica `replica` (sorted (i<j) System indices written to out_pairs up to max_pairs), or <0.
// stats[0..9]: nslot, nsci, nentries, -, row units, evaluated lane-pairs (all replicas), ncell, span,
// row entries, row entries that carry an allow word.
long long hostcheck_pairs(int n, int R, const double* pos, int periodic, const double* box,
                          double rc, double skin, int n_excl, const int* excl, int chunk,
                          int replica, int* out_pairs, long long max_pairs, double* stats) {
    const double rlist = rc + skin;
    Grid G;
    std::memset(&G, 0, sizeof(G));
    double lo[3] = {0, 0, 0}, ext[3];
    const int total = n * R;
    if (periodic) {
        for (int d = 0; d < 3; d++) ext[d] = box[d];
    } else {
        double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
        for (int k = 0; k < total; k++)
            for (int d = 0; d < 3; d++) {
                mn[d] = std::min(mn[d], pos[3 * (size_t)k + d]);
                mx[d] = std::max(mx[d], pos[3 * (size_t)k + d]);
            }
        for (int d = 0; d < 3; d++) { lo[d] = mn[d] - 1e-3; ext[d] = std::max(mx[d] - mn[d] + 2e-3, 1e-2); }
    }
    if (!setup_grid(G, n, R, periodic != 0, rlist, lo, ext, std::max(64, n / 8 + 64))) return -1;
    const int ncells = R * G.ncell;
    const int noff = G.span * G.span * (G.columns ? G.kz : G.span);

    // keys + three stable sorts: (cell, z), then kd refinement by y and by x (key_kernel,
    // refine_key_kernel and the radix sorts of pairlist.cu)
    std::vector<uint64_t> keys(total), ks(total);
    std::vector<int> vals(total), vs(total);
    auto wrapped = [&](int ga, uint32_t* fr) {
        int im[3];
        float xw[3];
        return atom_cell(G, ga / n, pos[3 * (size_t)ga], pos[3 * (size_t)ga + 1], pos[3 * (size_t)ga + 2], xw, im, fr);
    };
    for (int t = 0; t < total; t++) {
        uint32_t fr[3];
        const uint32_t g = wrapped(t, fr);
        keys[t] = make_key(g, 0u, fr[2]);
        vals[t] = t;
    }
    auto sort_pairs = [&]() {
        std::vector<int> order(total);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return keys[a] < keys[b]; });
        for (int k = 0; k < total; k++) { ks[k] = keys[order[k]]; vs[k] = vals[order[k]]; }
    };
    sort_pairs();
    if (G.columns) {
        // column layout: that was the (column, z) order; cut the columns into chunk cells
        // (chunk_key_kernel) and sort by (cell, z rank)
        std::vector<int> col_first(ncells + 1, 0);
        for (int k = 0; k < total; k++) {
            uint32_t c = (uint32_t)(ks[k] >> kSubBits);
            if (k == 0 || (uint32_t)(ks[k - 1] >> kSubBits) != c) col_first[c] = k;
        }
        for (int p = 0; p < total; p++) {
            const uint32_t pc = (uint32_t)(ks[p] >> kSubBits);
            uint32_t rin;
            const uint32_t g = chunk_cell(G, pc, p - col_first[pc], &rin);
            keys[p] = make_key(g, 0u, rin);
            vals[p] = vs[p];
        }
        sort_pairs();
    }
    // cells
    std::vector<int> cell_first(ncells + 1, 0), cell_count(ncells + 1, 0), cell_slot(ncells + 2, 0), cell_sci(ncells + 2, 0);
    for (int k = 0; k < total; k++) {
        uint32_t c = (uint32_t)(ks[k] >> kSubBits);
        if (k == 0 || (uint32_t)(ks[k - 1] >> kSubBits) != c) cell_first[c] = k;
        cell_count[c]++;
    }
    for (int level = 1; level <= 2; level++) {
        for (int p = 0; p < total; p++) {
            const uint32_t c = (uint32_t)(ks[p] >> kSubBits);
            const uint32_t b = kd_bucket(level, p, cell_first[c], cell_count[c], key_bucket(ks[p]));
            uint32_t fr[3];
            wrapped(vs[p], fr);
            keys[p] = make_key(c, b, fr[level == 1 ? 1 : 0]);
            vals[p] = vs[p];
        }
        sort_pairs();
    }
    int nslot = 0, nsci = 0;
    for (int c = 0; c < ncells; c++) {
        cell_slot[c] = nslot;
        cell_sci[c] = nsci;
        int pc = (cell_count[c] + kClusterSize - 1) / kClusterSize * kClusterSize;
        nslot += pc;
        nsci += (pc / kClusterSize + kMaxCi - 1) / kMaxCi;
    }
    cell_slot[ncells] = nslot;
    cell_sci[ncells] = nsci;
    const int ncl = nslot / kClusterSize;

    // slots
    std::vector<float> posq(4 * (size_t)nslot, 0.f);
    std::vector<int> atom(nslot, -1), slot_of(total, -1);
    std::vector<double> posw(3 * (size_t)nslot, 0.0);  // wrapped double coordinates per slot
    for (int s = 0; s < nslot; s++) { posq[4 * (size_t)s] = posq[4 * (size_t)s + 1] = posq[4 * (size_t)s + 2] = kFar; }
    for (int k = 0; k < total; k++) {
        uint32_t c = (uint32_t)(ks[k] >> kSubBits);
        int ga = vs[k];
        int slot = cell_slot[c] + (k - cell_first[c]);
        float xw[3];
        int im[3];
        atom_cell(G, ga / n, pos[3 * (size_t)ga], pos[3 * (size_t)ga + 1], pos[3 * (size_t)ga + 2], xw, im);
        for (int d = 0; d < 3; d++) {
            posq[4 * (size_t)slot + d] = xw[d];
            posw[3 * (size_t)slot + d] = pos[3 * (size_t)ga + d] + (periodic ? im[d] * box[d] : 0.0);
        }
        atom[slot] = ga;
        slot_of[ga] = slot;
    }
    // boxes
    std::vector<BBox> cl_box(ncl), sci_box(nsci);
    for (int c = 0; c < ncl; c++) cl_box[c] = group_bbox(posq.data(), c * kClusterSize, kClusterSize);
    std::vector<SciDesc> sci(nsci);
    std::vector<int> cl_sci(ncl, -1);
    for (int c = 0; c < ncells; c++) {
        int cl0 = cell_slot[c] / kClusterSize, cl1 = cell_slot[c + 1] / kClusterSize, s = cell_sci[c];
        for (int k = cl0; k < cl1; k += kMaxCi, s++) {
            SciDesc d;
            d.c0 = k;
            d.nci = std::min(kMaxCi, cl1 - k);
            d.replica = c / G.ncell;
            d.pad = 0;
            BBox b = cl_box[k];
            cl_sci[k] = s;
            for (int j = 1; j < d.nci; j++) { b = box_union(b, cl_box[k + j]); cl_sci[k + j] = s; }
            sci[s] = d;
            sci_box[s] = b;
        }
    }
    // search: count, scan, fill
    SearchView V;
    V.G = G;
    V.sci = sci.data();
    V.sci_box = sci_box.data();
    V.cl_box = cl_box.data();
    V.cell_slot = cell_slot.data();
    std::vector<BBox> cell_box(ncells);   // cell_box_kernel
    for (int c = 0; c < ncells; c++) {
        BBox b;
        for (int d = 0; d < 3; d++) { b.lo[d] = kBoxEmptyLo; b.hi[d] = -kBoxEmptyLo; }
        for (int k = cell_slot[c] / kClusterSize; k < cell_slot[c + 1] / kClusterSize; k++)
            if (!box_empty(cl_box[k])) b = box_union(b, cl_box[k]);
        cell_box[c] = b;
    }
    V.cell_box = cell_box.data();
    V.posq4 = posq.data();
    const long long nitems = (long long)nsci * noff;
    std::vector<int> item_off(nitems + 1, 0);
    for (long long t = 0; t < nitems; t++)
        item_off[t + 1] = item_off[t] + search_any(V, (int)(t / noff), (int)(t % noff), [](int, uint32_t, uint32_t, bool) {});
    const int nraw = item_off[nitems];
    std::vector<uint32_t> rx(nraw), ry(nraw);
    std::vector<int> rsci(nraw);
    for (long long t = 0; t < nitems; t++) {
        int base = item_off[t];
        search_any(V, (int)(t / noff), (int)(t % noff), [&](int k, uint32_t w0, uint32_t imask, bool diag) {
            rx[base + k] = w0;
            ry[base + k] = imask;
            rsci[base + k] = (int)(t / noff);
        });
    }
    // exact prune + order-preserving compaction (prune_kernel / compact_kernel / sci_off_kernel)
    std::vector<uint32_t> ex, ey;
    std::vector<int> esci, sci_off(nsci + 1, 0);
    {
        int next_sci = 0;
        for (int e = 0; e < nraw; e++) {
            while (next_sci <= rsci[e]) sci_off[next_sci++] = (int)ex.size();
            const SciDesc sd = sci[rsci[e]];
            const uint32_t m = prune_imask(V, sd, rx[e], ry[e] & 0xffu);
            if (!m) continue;
            const int B = (int)(rx[e] & 0x3ffffffu);
            ex.push_back(rx[e]);
            ey.push_back(m);
            esci.push_back(rsci[e]);
        }
        while (next_sci <= nsci) sci_off[next_sci++] = (int)ex.size();
    }
    const int nentries = (int)ex.size();
    // the System's exclusions as the CSR over atoms (both directions) the device path walks
    // (Topology::excl_start / excl_idx)
    std::vector<int> excl_start(n + 1, 0), excl_idx;
    {
        std::vector<std::vector<int>> nb(n);
        for (int k = 0; k < n_excl; k++) { nb[excl[2 * k]].push_back(excl[2 * k + 1]); nb[excl[2 * k + 1]].push_back(excl[2 * k]); }
        for (int a = 0; a < n; a++) {
            std::sort(nb[a].begin(), nb[a].end());
            excl_start[a + 1] = excl_start[a] + (int)nb[a].size();
            excl_idx.insert(excl_idx.end(), nb[a].begin(), nb[a].end());
        }
    }
    int nunits = 0;

    // traversal exactly like the row kernel (kernels_cluster.cu, pair_row_kernel): every i-group
    // (G consecutive clusters of a supercluster) walks its row of individual j-atoms -- the atoms
    // of its entries' j-clusters that row_hits() keeps and whose allow word is not empty -- and a
    // lane evaluates its j-atom against the allowed i-atoms of the group.
    int Grow = 1;
    if (const char* e = getenv("SDMB200_ROW_GROUP")) Grow = atoi(e) == 2 ? 2 : 1;
    const bool use_walk = getenv("SDM_HOSTCHECK_WALK") != nullptr;
    std::vector<float> par2(2 * (size_t)nslot, 1.0f);   // the checker carries no LJ parameters: every atom has a term
    const int ng = kMaxCi / Grow;
    std::vector<std::pair<int, int>> found;
    const double rc2 = rc * rc;
    double lane_pairs = 0, row_entries = 0, row_masked = 0;
    std::vector<long> row_len((size_t)nsci * ng, 0);
    for (int s = 0; s < nsci; s++) {
        const SciDesc sd = sci[s];
        for (int e = sci_off[s]; e < sci_off[s + 1]; e++) {
            const int cj = (int)(ex[e] & 0x3ffffffu);
            const uint32_t imask = ey[e] & 0xffu;
            const uint32_t code = ex[e] >> 26;
            const int sh[3] = {shift_x(code), shift_y(code), shift_z(code)};
            uint32_t jh_lo, jh_hi;
            entry_hits(V, sd, ex[e], imask, &jh_lo, &jh_hi);
            // the j-cluster's exclusion tiles (cluster_info_kernel), or the CSR walk when SDM_HOSTCHECK_WALK is set
            ClusterTiles tiles;
            const ClusterInfo info = cluster_info(cj, atom.data(), par2.data(), excl_start.data(), excl_idx.data(),
                                                  slot_of.data(), n, &tiles);
            for (int g = 0; g < ng; g++) {
                const uint32_t hits = row_hits(jh_lo, jh_hi, imask, g, Grow);
                for (int tj = 0; tj < kJGroup; tj++) {
                    if (!((hits >> tj) & 1u)) continue;
                    const uint32_t allow = (use_walk || info.npart < 0)
                                               ? row_allow(sd, imask, cj, code, g, Grow, tj, excl_start.data(),
                                                           excl_idx.data(), slot_of.data(), atom.data(), n)
                                               : row_allow(sd, imask, cj, code, g, Grow, tj, info, &tiles);
                    if (!allow) continue;
                    row_entries += 1;
                    row_len[(size_t)s * ng + g] += 1;
                    if (allow != (Grow == 2 ? 0xffffu : 0xffu)) row_masked += 1;
                    lane_pairs += 8 * Grow;
                    if (sd.replica != replica) continue;
                    const int jslot = cj * kJGroup + tj;
                    const int aj = atom[jslot];
                    if (aj < 0) continue;
                    for (int a = 0; a < 8 * Grow; a++) {
                        if (!((allow >> a) & 1u)) continue;
                        const int ci = g * Grow + a / kClusterSize;
                        if (ci >= sd.nci) continue;   // the kernel stages these lanes as dummies
                        const int islot = (sd.c0 + ci) * kClusterSize + a % kClusterSize;
                        const int ai = atom[islot];
                        if (ai < 0) continue;
                        // the image this entry addresses (what the kernel evaluates), in double
                        double d[3];
                        for (int k = 0; k < 3; k++)
                            d[k] = posw[3 * (size_t)islot + k] - (posw[3 * (size_t)jslot + k] + (periodic ? sh[k] * box[k] : 0.0));
                        if (d[0] * d[0] + d[1] * d[1] + d[2] * d[2] > rc2) continue;
                        const int p = ai % n, q = aj % n;
                        found.emplace_back(std::min(p, q), std::max(p, q));
                    }
                }
            }
        }
    }
    for (long len : row_len) nunits += (int)((len + 32L * chunk - 1) / (32L * chunk));   // rows_units_count_kernel
    std::sort(found.begin(), found.end());
    long long m = std::min<long long>((long long)found.size(), max_pairs);
    for (long long k = 0; k < m; k++) { out_pairs[2 * k] = found[k].first; out_pairs[2 * k + 1] = found[k].second; }
    if (stats) {
        stats[0] = nslot; stats[1] = nsci; stats[2] = nentries; stats[3] = 0;
        stats[4] = nunits; stats[5] = lane_pairs; stats[6] = G.ncell; stats[7] = G.span;
        stats[8] = row_entries; stats[9] = row_masked;
    }
    return (long long)found.size();
}

}  // extern "C"
