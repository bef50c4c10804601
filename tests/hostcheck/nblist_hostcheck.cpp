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
// stats[0..7]: nslot, nsci, nentries, nmasks, nunits, evaluated lane-pairs (all replicas),
// ncell, span.
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
    std::vector<int> rflag(nraw, 0), rsci(nraw);
    for (long long t = 0; t < nitems; t++) {
        int base = item_off[t];
        search_any(V, (int)(t / noff), (int)(t % noff), [&](int k, uint32_t w0, uint32_t imask, bool diag) {
            rx[base + k] = w0;
            ry[base + k] = imask;
            rflag[base + k] = diag ? 1 : 0;
            rsci[base + k] = (int)(t / noff);
        });
    }
    // exact prune + order-preserving compaction (prune_kernel / compact_kernel / sci_off_kernel)
    std::vector<uint32_t> ex, ey;
    std::vector<int> flag, esci, sci_off(nsci + 1, 0);
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
            flag.push_back(rflag[e] && ((m >> (B - sd.c0)) & 1u) ? 1 : 0);
            esci.push_back(rsci[e]);
        }
        while (next_sci <= nsci) sci_off[next_sci++] = (int)ex.size();
    }
    const int nentries = (int)ex.size();
    flag.push_back(0);
    // exclusions pass 0
    auto for_excl = [&](int pass, std::vector<uint32_t>* masks) {
        for (int r = 0; r < R; r++)
            for (int k = 0; k < n_excl; k++) {
                int sa = slot_of[r * n + excl[2 * k]], sb = slot_of[r * n + excl[2 * k + 1]];
                int si, sj;
                exclusion_roles(sa, sb, &si, &sj);
                int isci = cl_sci[si / kClusterSize];
                int ci = si / kClusterSize - sci[isci].c0;
                uint32_t cj = (uint32_t)(sj / kJGroup);
                uint32_t bit = mask_bit(si, sj);
                for (int e = sci_off[isci]; e < sci_off[isci + 1]; e++) {
                    if ((ex[e] & 0x3ffffffu) != cj) continue;
                    if (pass == 0) flag[e] = 1;
                    else (*masks)[(size_t)(ey[e] >> 8) * kMaskWords + mask_word(ci, si)] &= ~bit;
                }
            }
    };
    for_excl(0, nullptr);
    int nmasks = 0;
    std::vector<int> midx(nentries, 0);
    for (int e = 0; e < nentries; e++) if (flag[e]) midx[e] = nmasks++;
    std::vector<uint32_t> masks((size_t)(nmasks + 1) * kMaskWords, 0xffffffffu);
    for (int e = 0; e < nentries; e++) {
        if (!flag[e]) continue;
        uint32_t m = (uint32_t)midx[e] + 1u;
        ey[e] = (ey[e] & 0xffu) | (m << 8);
        int cj = (int)(ex[e] & 0x3ffffffu);
        uint32_t code = ex[e] >> 26;
        const SciDesc sd = sci[esci[e]];
        for (int w = 0; w < kMaskWords; w++) {
            uint32_t v = 0xffffffffu;
            if (code == kShiftZero && cj == sd.c0 + (w >> 1)) v = triangle_mask(w & 1);
            masks[(size_t)m * kMaskWords + w] = v;
        }
    }
    for_excl(1, &masks);
    int nunits = 0;
    for (int s = 0; s < nsci; s++) {
        int len = sci_off[s + 1] - sci_off[s];
        nunits += (len + chunk - 1) / chunk;
    }

    if (getenv("SDM_HOSTCHECK_STATS")) {
        // pairing statistics: consecutive unmasked entries of a unit taken two at a time
        long both = 0, one = 0, single_tiles = 0, masked_tiles = 0, npairs = 0, nsingle = 0, nmasked = 0;
        for (int s = 0; s < nsci; s++)
            for (int b = sci_off[s]; b < sci_off[s + 1]; b += chunk) {
                const int e1 = std::min(b + chunk, sci_off[s + 1]);
                std::vector<uint32_t> um;
                for (int e = b; e < e1; e++) {
                    if (ey[e] >> 8) { nmasked++; masked_tiles += __builtin_popcount(ey[e] & 0xffu); }
                    else um.push_back(ey[e] & 0xffu);
                }
                size_t k = 0;
                for (; k + 1 < um.size(); k += 2) {
                    npairs++;
                    both += __builtin_popcount(um[k] & um[k + 1]);
                    one += __builtin_popcount(um[k] ^ um[k + 1]);
                }
                if (k < um.size()) { nsingle++; single_tiles += __builtin_popcount(um[k]); }
            }
        {
            // all entries of a unit (masked too), ordered by imask value, then paired; and a greedy
            // minimum-Hamming-distance pairing for comparison
            long tiles = 0, un_sorted = 0, un_plain = 0, un_greedy = 0, steps = 0;
            for (int s = 0; s < nsci; s++)
                for (int b = sci_off[s]; b < sci_off[s + 1]; b += chunk) {
                    const int e1 = std::min(b + chunk, sci_off[s + 1]);
                    std::vector<uint32_t> m;
                    for (int e = b; e < e1; e++) { m.push_back(ey[e] & 0xffu); tiles += __builtin_popcount(ey[e] & 0xffu); }
                    steps += ((long)m.size() + 1) / 2;
                    for (size_t k = 0; k < m.size(); k += 2) un_plain += __builtin_popcount(m[k] | (k + 1 < m.size() ? m[k + 1] : 0u));
                    std::vector<uint32_t> q = m;
                    std::sort(q.begin(), q.end());
                    for (size_t k = 0; k < q.size(); k += 2) un_sorted += __builtin_popcount(q[k] | (k + 1 < q.size() ? q[k + 1] : 0u));
                    std::vector<char> used(m.size(), 0);
                    for (size_t a = 0; a < m.size(); a++) {
                        if (used[a]) continue;
                        used[a] = 1;
                        int best = -1, bd = 99;
                        for (size_t c = a + 1; c < m.size(); c++)
                            if (!used[c] && __builtin_popcount(m[a] ^ m[c]) < bd) { bd = __builtin_popcount(m[a] ^ m[c]); best = (int)c; }
                        if (best >= 0) { used[best] = 1; un_greedy += __builtin_popcount(m[a] | m[best]); }
                        else un_greedy += __builtin_popcount(m[a]);
                    }
                }
            fprintf(stderr, "union pairing: tiles %ld steps %ld | 2*union plain %ld sorted %ld greedy %ld\n", tiles, steps,
                    2 * un_plain, 2 * un_sorted, 2 * un_greedy);
        }
        fprintf(stderr, "pairing: pairs %ld both %ld one %ld | singles %ld tiles %ld | masked %ld tiles %ld | units %d\n",
                npairs, both, one, nsingle, single_tiles, nmasked, masked_tiles, nunits);
    }

    // traversal exactly like the row kernel (kernels_cluster.cu, pair_row_kernel): every i-group
    // (G consecutive clusters of a supercluster) walks its row of individual j-atoms -- the atoms
    // of its entries' j-clusters that row_hits() keeps and whose allow word is not empty -- and a
    // lane evaluates its j-atom against the allowed i-atoms of the group.
    int Grow = 1;
    if (const char* e = getenv("SDMB200_ROW_GROUP")) Grow = atoi(e) == 2 ? 2 : 1;
    const int ng = kMaxCi / Grow;
    std::vector<std::pair<int, int>> found;
    const double rc2 = rc * rc;
    double lane_pairs = 0, row_entries = 0, row_masked = 0;
    for (int s = 0; s < nsci; s++) {
        const SciDesc sd = sci[s];
        for (int e = sci_off[s]; e < sci_off[s + 1]; e++) {
            const int cj = (int)(ex[e] & 0x3ffffffu);
            const uint32_t imask = ey[e] & 0xffu, m = ey[e] >> 8;
            const uint32_t code = ex[e] >> 26;
            const int sh[3] = {shift_x(code), shift_y(code), shift_z(code)};
            uint32_t jh_lo, jh_hi;
            entry_hits(V, sd, ex[e], imask, &jh_lo, &jh_hi);
            const uint32_t* maskset = m ? &masks[(size_t)m * kMaskWords] : nullptr;
            const bool same_sci = cj >= sd.c0 && cj < sd.c0 + sd.nci;
            for (int g = 0; g < ng; g++) {
                const uint32_t hits = row_hits(jh_lo, jh_hi, imask, g, Grow);
                for (int tj = 0; tj < kJGroup; tj++) {
                    if (!((hits >> tj) & 1u)) continue;
                    const uint32_t allow = row_allow(maskset, imask, same_sci, g, Grow, tj);
                    if (!allow) continue;
                    row_entries += 1;
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
    std::sort(found.begin(), found.end());
    long long m = std::min<long long>((long long)found.size(), max_pairs);
    for (long long k = 0; k < m; k++) { out_pairs[2 * k] = found[k].first; out_pairs[2 * k + 1] = found[k].second; }
    if (stats) {
        stats[0] = nslot; stats[1] = nsci; stats[2] = nentries; stats[3] = nmasks;
        stats[4] = nunits; stats[5] = lane_pairs; stats[6] = G.ncell; stats[7] = G.span;
        stats[8] = row_entries; stats[9] = row_masked;
    }
    return (long long)found.size();
}

}  // extern "C"
