// nblist_core.h -- pair list construction, written once as per-item bodies that compile both
// for the device (thin __global__ wrappers in pairlist.cu) and for the host (the CPU checker
// tests/hostcheck builds from this same header with g++ to validate pair coverage against the
// oracle without a GPU).  No reference code corresponds to this: the reference delegates
// neighbour lists to OpenMM; the layout below is designed for the
// sm_100a pair kernel in kernels_rows.cu.
//
// Layout
//   * replicas live in disjoint cell ranges of ONE global index space: global cell
//     g = r*ncell + cell.
//   * atoms are sorted by global cell and, inside a cell, by a balanced three-level kd split
//     (z, then y, then x; every split point is a multiple of 8 atoms), so the 8-atom clusters
//     are compact boxes (28.8 % -> 31.8 % useful pairs per tile on the 20 k-atom fixture compared
//     with a Morton order); every cell is padded to a multiple of 8 slots with dummy atoms, so a
//     cluster (8 slots) never straddles a cell.
//   * column layout (Grid::columns, the default): the cells are xy columns of the box cut along z
//     into chunks of exactly 64 atoms (the last chunk of a column takes the remainder), found
//     with one extra sort by (column, z).  Every chunk is one full supercluster of 8 full
//     clusters, so padding exists only at the top of a column (0.8 % dummy slots instead of 6.3 %
//     on the 20 k-atom fixture) and the imasks are denser: 9 % fewer tiles, 15 % fewer entries,
//     35.5 % instead of 31.8 % useful pairs per tile than with geometric 3-D cells of ~52 atoms.
//   * an i-supercluster (sci) is a run of <= 8 clusters of one cell; the search gives it j-cluster
//     entries {cj | shift<<26, imask}; imask bit ci says cluster ci of the sci has an atom within
//     rlist of an atom of j-cluster cj.  Every unordered cluster pair is owned by exactly one side
//     (the lower cluster index) and a cluster against itself keeps the triangle j > i, so each atom
//     pair is evaluated once.
//   * the entries are an intermediate: the pair kernel walks per-atom j ROWS derived from them
//     (stage 5 below), with exclusions and the triangle folded into per-entry allow words.
#pragma once

#include <stdint.h>

#include <cmath>
#if defined(__CUDACC__)
#define SDM_HD __host__ __device__ __forceinline__
#else
#define SDM_HD inline
#include <cmath>
#endif

namespace sdm {
namespace nbl {

constexpr int kClusterSize = 8;      // atoms per i-cluster
constexpr int kJGroup = 8;           // atoms per j-group (a whole cluster)
constexpr int kMaxCi = 8;            // clusters per supercluster
constexpr int kCoordBits = 16;       // in-cell coordinate resolution of the sort key
constexpr int kSubBits = kCoordBits + 2;  // sort key = cell << 18 | kd bucket (2 bits) << 16 | coordinate
constexpr int kMaxSpan = 6;          // search stencil is at most kMaxSpan cells per dimension
constexpr int kChunkAtoms = kClusterSize * kMaxCi;  // atoms per chunk cell of the column layout
constexpr float kFar = 1.0e6f;       // coordinate of dummy (padding) atoms
constexpr float kBoxEmptyLo = 3.0e38f;

struct Grid {
    int periodic;
    int nc[3];
    int ncell;          // cells per replica
    int n;              // atoms per replica
    int R;              // replicas
    int span;           // stencil cells per dimension (<= kMaxSpan)
    double lo[3];       // grid origin (0 when periodic)
    double box[3];      // box edges (periodic) or grid extent (non-periodic)
    double cs[3];       // cell side per dimension
    double inv_cs[3];
    float rlist, rlist2;
    float boxf[3];
    // column layout (see "Layout" above): cells are xy columns cut along z into runs of
    // kChunkAtoms atoms; nc[2] == kz is the number of such chunk cells a column can hold and
    // the z index of a cell is the chunk number, not a coordinate.  0 = geometric 3-D cells.
    int columns;
    int kz;
};

struct BBox {
    float lo[3], hi[3];
};

struct SciDesc {
    int c0;        // first cluster (global cluster index)
    int nci;       // clusters in this sci (1..8)
    int replica;
    int pad;
};

// ---- small helpers ----------------------------------------------------------------------------
SDM_HD uint32_t spread3(uint32_t v) {  // 3 bits -> every third bit
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4);
}

SDM_HD bool box_empty(const BBox& b) { return b.lo[0] > b.hi[0]; }

SDM_HD float box_dist2(const BBox& a, const BBox& b, float sx, float sy, float sz) {
    // squared distance between AABB a and AABB b shifted by (sx,sy,sz)
    float d2 = 0.f;
    const float s[3] = {sx, sy, sz};
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int d = 0; d < 3; d++) {
        float g1 = (b.lo[d] + s[d]) - a.hi[d];
        float g2 = a.lo[d] - (b.hi[d] + s[d]);
        float g = g1 > g2 ? g1 : g2;
        if (g > 0.f) d2 += g * g;
    }
    return d2;
}

// Which side of an unordered cluster pair (A != B) carries it in its list: the cluster with the
// lower index plays the i role.  (A parity checkerboard would balance list lengths, but it makes
// neighbouring i-clusters of a supercluster alternate, halving the density of the imasks and with
// it the reuse of every j load; list lengths are balanced by the unit chunking instead.)
SDM_HD bool owner_is_i(int A, int B) { return A < B; }

SDM_HD uint32_t shift_code(int sx, int sy, int sz) {  // each in {-1,0,1}
    return (uint32_t)((sx + 1) | ((sy + 1) << 2) | ((sz + 1) << 4));
}
SDM_HD int shift_x(uint32_t code) { return (int)(code & 3u) - 1; }
SDM_HD int shift_y(uint32_t code) { return (int)((code >> 2) & 3u) - 1; }
SDM_HD int shift_z(uint32_t code) { return (int)((code >> 4) & 3u) - 1; }
constexpr uint32_t kShiftZero = 1u | (1u << 2) | (1u << 4);

// ---- stage 0 (host): grid sizing -----------------------------------------------------------------
// Chooses the cell grid for a box of extent `ext` at origin `lo` holding n atoms per replica.
// Fills periodic / n / R / rlist / nc / cs / inv_cs / lo / box / boxf / ncell / span / columns / kz.
// Returns false when no grid with span <= kMaxSpan and ncell <= cell_cap exists.
inline bool size_grid(Grid& G, int n, int R, bool periodic, double rlist, const double lo[3],
                      const double ext[3], long long cell_cap, bool columns) {
    G.periodic = periodic ? 1 : 0;
    G.n = n;
    G.R = R;
    // + 1e-4 nm: the list is pruned with FP32 distances, the cutoff test may be decided in FP64
    G.rlist = (float)(rlist + 1e-4);
    G.rlist2 = (float)((rlist + 1e-4) * (rlist + 1e-4));
    const double vol = ext[0] * ext[1] * ext[2];
    const double density = vol > 0 ? n / vol : 100.0;
    G.columns = columns ? 1 : 0;
    G.kz = 1;
    for (int d = 0; d < 3; d++) {
        G.lo[d] = lo[d];
        G.box[d] = ext[d];
        G.boxf[d] = (float)ext[d];
    }
    // cells of ~40 atoms on average (one supercluster) or, column layout, xy columns with the
    // cross-section of one 64-atom chunk's cube
    double side = std::cbrt((columns ? (double)kChunkAtoms : 40.0) / (density > 1e-6 ? density : 1e-6));
    if (side < 0.5 * rlist + 1e-3) side = 0.5 * rlist + 1e-3;
    for (int iter = 0; iter < 64; iter++) {
        long long ncell = 1;
        int span = 1;
        const int ngeo = columns ? 2 : 3;
        for (int d = 0; d < ngeo; d++) {
            int nc = (int)std::floor(ext[d] / side + (columns ? 0.5 : 0.0));
            if (nc < 1) nc = 1;
            G.nc[d] = nc;
            G.cs[d] = ext[d] / nc;
            ncell *= nc;
            int sp = (int)std::floor((G.cs[d] + 2 * rlist + 3e-4) / G.cs[d]) + 2;
            if (!G.periodic && sp > nc) sp = nc;
            if (sp > span) span = sp;
        }
        if (columns) {
            // along z the cells are the chunks themselves: kz chunk cells per column, 25 % + 1
            // more than the average column needs (a fuller column keeps the rest in its last
            // cell); one geometric cell along z, so atom_cell returns the column
            const int kz = (int)std::ceil(1.25 * n / (double)(ncell * kChunkAtoms)) + 1;
            G.nc[2] = kz;
            G.cs[2] = ext[2];
            G.kz = kz;
            ncell *= kz;
        }
        for (int d = 0; d < 3; d++) G.inv_cs[d] = 1.0 / G.cs[d];
        if (span <= kMaxSpan && ncell <= cell_cap) {
            G.ncell = (int)ncell;
            G.span = span;
            return true;
        }
        side *= 1.1;
    }
    return false;
}

// ---- stage 1: sort keys -------------------------------------------------------------------------
// Wraps the position into the box (periodic), returns the global cell, the wrapped coordinates
// (relative to the grid origin they are >= 0) and the integer image that was applied
// (xw = x + img*L).
SDM_HD uint32_t atom_cell(const Grid& G, int replica, double x, double y, double z, float* xw_out,
                          int* img_out, uint32_t* frac_out = nullptr) {
    double p[3] = {x, y, z};
    int c[3];
    for (int d = 0; d < 3; d++) {
        double w = p[d];
        int img = 0;
        if (G.periodic) {
            double f = floor(w / G.box[d]);
            w -= f * G.box[d];
            img = -(int)f;
            if (w >= G.box[d]) { w -= G.box[d]; img -= 1; }
            if (w < 0) { w = 0; }
        }
        double t = (w - G.lo[d]) * G.inv_cs[d];
        int k = (int)floor(t);
        if (k < 0) k = 0;
        if (k >= G.nc[d]) k = G.nc[d] - 1;
        c[d] = k;
        xw_out[d] = (float)w;
        img_out[d] = img;
        if (frac_out) {
            // position inside the cell in kCoordBits bits (order-preserving; ties keep their order)
            double fr = (t - (double)k) * (double)(1 << kCoordBits);
            int q = (int)fr;
            if (q < 0) q = 0;
            if (q > (1 << kCoordBits) - 1) q = (1 << kCoordBits) - 1;
            frac_out[d] = (uint32_t)q;
        }
    }
    uint32_t cell = (uint32_t)((c[2] * G.nc[1] + c[1]) * G.nc[0] + c[0]);
    return (uint32_t)replica * (uint32_t)G.ncell + cell;
}

SDM_HD uint64_t make_key(uint32_t gcell, uint32_t bucket, uint32_t cbits) {
    return ((uint64_t)gcell << kSubBits) | ((uint64_t)(bucket & 3u) << kCoordBits) | (uint64_t)cbits;
}
SDM_HD uint32_t key_bucket(uint64_t key) { return (uint32_t)(key >> kCoordBits) & 3u; }

// Column layout, between the (column, z) sort and the (cell, z) sort: the atom at rank `rank` of
// its column (pseudo cell = replica*ncell + column, what atom_cell returns when the grid has one
// geometric cell along z) goes to chunk cell rank / 64 -- the last cell of a column takes what
// is left -- and keeps its z order through its rank inside the chunk.
SDM_HD uint32_t chunk_cell(const Grid& G, uint32_t pseudo_cell, int rank, uint32_t* rank_in_chunk) {
    const uint32_t r = pseudo_cell / (uint32_t)G.ncell, col = pseudo_cell % (uint32_t)G.ncell;
    int chunk = rank / kChunkAtoms;
    if (chunk > G.kz - 1) chunk = G.kz - 1;
    const int rin = rank - chunk * kChunkAtoms;
    *rank_in_chunk = (uint32_t)(rin < (1 << kCoordBits) - 1 ? rin : (1 << kCoordBits) - 1);
    return r * (uint32_t)G.ncell + (uint32_t)chunk * (uint32_t)(G.nc[0] * G.nc[1]) + col;
}

// Atoms that go to the first half when a run of `count` atoms is split: half of its clusters
// (rounded up), i.e. a multiple of 8.
SDM_HD int split_first(int count) {
    const int m = (count + kClusterSize - 1) / kClusterSize;
    const int h = kClusterSize * ((m + 1) / 2);
    return h < count ? h : count;
}

// kd refinement.  The sort runs three times: by (cell, z), by (cell, b1, y), by (cell, b1, b2, x),
// where b1 / b2 say whether the atom fell into the second half of the previous split.  Given the
// position p of an atom in the order produced by round `level` (1 or 2), the first position and
// size of its cell and its previous bucket bits, returns the bucket bits for the next round.
SDM_HD uint32_t kd_bucket(int level, int p, int cell_first, int cell_count, uint32_t prev_bucket) {
    const int h1 = split_first(cell_count);
    if (level == 1) return (uint32_t)((p - cell_first) >= h1) << 1;
    const int b1 = (int)(prev_bucket >> 1) & 1;
    const int seg0 = cell_first + (b1 ? h1 : 0);
    const int segn = b1 ? cell_count - h1 : h1;
    return (uint32_t)(b1 << 1) | (uint32_t)((p - seg0) >= split_first(segn));
}

// ---- stage 2: bounding boxes --------------------------------------------------------------------
// posq: sorted slots (float4 as 4 floats); a dummy slot has x >= kFar/2.
SDM_HD BBox group_bbox(const float* posq4, int first_slot, int count) {
    BBox b;
    for (int d = 0; d < 3; d++) { b.lo[d] = kBoxEmptyLo; b.hi[d] = -kBoxEmptyLo; }
    for (int s = first_slot; s < first_slot + count; s++) {
        const float* p = posq4 + 4 * (size_t)s;
        if (p[0] >= 0.5f * kFar) continue;
        for (int d = 0; d < 3; d++) {
            if (p[d] < b.lo[d]) b.lo[d] = p[d];
            if (p[d] > b.hi[d]) b.hi[d] = p[d];
        }
    }
    return b;
}

SDM_HD BBox box_union(const BBox& a, const BBox& b) {
    BBox u;
    for (int d = 0; d < 3; d++) {
        u.lo[d] = a.lo[d] < b.lo[d] ? a.lo[d] : b.lo[d];
        u.hi[d] = a.hi[d] > b.hi[d] ? a.hi[d] : b.hi[d];
    }
    return u;
}

// ---- stage 3: pair search ---------------------------------------------------------------------
// Read-only view of what the search needs.
struct SearchView {
    Grid G;
    const SciDesc* sci;        // [nsci]
    const BBox* sci_box;       // [nsci]
    const BBox* cl_box;        // [ncluster]  8-slot cluster boxes
    const int* cell_slot;      // [R*ncell + 1] first slot of every global cell (padded layout)
    const BBox* cell_box;      // [R*ncell] box of the cell's real atoms (column layout: early exit)
    const float* posq4;        // [nslot][4] sorted positions: exact (atom-level) pruning of imask
};

// Cell-coordinate range (unwrapped for periodic, clamped for non-periodic) the sci must scan.
SDM_HD void search_range(const Grid& G, const BBox& b, int cmin[3], int cmax[3]) {
    const float eps = 1.0e-4f;
    for (int d = 0; d < 3; d++) {
        double a = ((double)b.lo[d] - (double)G.rlist - eps - G.lo[d]) * G.inv_cs[d];
        double c = ((double)b.hi[d] + (double)G.rlist + eps - G.lo[d]) * G.inv_cs[d];
        int ia = (int)floor(a), ic = (int)floor(c);
        if (!G.periodic) {
            if (ia < 0) ia = 0;
            if (ic < 0) ic = 0;
            if (ia >= G.nc[d]) ia = G.nc[d] - 1;
            if (ic >= G.nc[d]) ic = G.nc[d] - 1;
        }
        if (ic - ia + 1 > G.span) ic = ia + G.span - 1;  // guarded by the host-side span check
        cmin[d] = ia;
        cmax[d] = ic;
    }
}

// Is any atom of cluster A within rlist of any atom of cluster B (shifted)?  Dummies never are.
SDM_HD bool any_pair_within(const float* posq4, int A, int B, float sx, float sy, float sz,
                            float rlist2) {
    for (int tj = 0; tj < kJGroup; tj++) {
        const float* pj = posq4 + 4 * (size_t)(B * kJGroup + tj);
        if (pj[0] >= 0.5f * kFar) continue;
        const float xj = pj[0] + sx, yj = pj[1] + sy, zj = pj[2] + sz;
        for (int ti = 0; ti < kClusterSize; ti++) {
            const float* pi = posq4 + 4 * (size_t)(A * kClusterSize + ti);
            if (pi[0] >= 0.5f * kFar) continue;
            const float dx = pi[0] - xj, dy = pi[1] - yj, dz = pi[2] - zj;
            if (dx * dx + dy * dy + dz * dz < rlist2) return true;
        }
    }
    return false;
}

// Clusters of the slot range [s0, s1) (one cell) against the sci, for one periodic image: calls
// emit(k, cj | shift<<26, imask, diag) for every j-cluster that interacts with at least one
// cluster of the sci under the ownership rule; k counts on from `count`.  Returns the new count.
template <class Emit>
SDM_HD int scan_cell_clusters(const SearchView& V, const SciDesc& sd, const BBox& sb, int s0, int s1,
                              float sx, float sy, float sz, uint32_t code, int count, Emit& emit) {
    const Grid& G = V.G;
    for (int B = s0 / kJGroup; B < s1 / kJGroup; B++) {
        const BBox jb = V.cl_box[B];
        if (box_empty(jb)) continue;
        if (box_dist2(sb, jb, sx, sy, sz) >= G.rlist2) continue;
        uint32_t imask = 0;
        bool diag = false;
        for (int ci = 0; ci < sd.nci; ci++) {
            const int A = sd.c0 + ci;
            bool own;
            if (A == B) {
                // a cluster against itself: zero shift -> triangle; a non-zero shift only once
                own = (code == kShiftZero) || (code > kShiftZero);
                if (code == kShiftZero) diag = true;
            } else {
                own = owner_is_i(A, B);
            }
            if (!own) continue;
            if (box_dist2(V.cl_box[A], jb, sx, sy, sz) >= G.rlist2) continue;
            imask |= 1u << ci;   // box level only; the exact atom-pair prune is a separate pass
        }
        if (imask) {
            emit(count, (uint32_t)B | (code << 26), imask, diag && ((imask >> (B - sd.c0)) & 1u));
            count++;
        }
    }
    return count;
}

// One search item = (sci, stencil offset).  Visits the clusters of the addressed cell and calls
// emit(k, cj | shift<<26, imask, diag) for every j-cluster that interacts with at least one
// cluster of the sci under the ownership rule.  Returns the number of entries.
template <class Emit>
SDM_HD int search_item(const SearchView& V, int isci, int off, Emit emit) {
    const Grid& G = V.G;
    const SciDesc sd = V.sci[isci];
    const BBox sb = V.sci_box[isci];
    if (box_empty(sb)) return 0;
    int cmin[3], cmax[3];
    search_range(G, sb, cmin, cmax);
    const int ox = off % G.span, oy = (off / G.span) % G.span, oz = off / (G.span * G.span);
    int u[3] = {cmin[0] + ox, cmin[1] + oy, cmin[2] + oz};
    if (u[0] > cmax[0] || u[1] > cmax[1] || u[2] > cmax[2]) return 0;
    int w[3], sh[3];
    for (int d = 0; d < 3; d++) {
        if (G.periodic) {
            int q = u[d] >= 0 ? u[d] / G.nc[d] : -((-u[d] + G.nc[d] - 1) / G.nc[d]);
            w[d] = u[d] - q * G.nc[d];
            sh[d] = q;
            if (q < -1 || q > 1) return 0;  // excluded by the host-side box-size check
        } else {
            w[d] = u[d];
            sh[d] = 0;
        }
    }
    const float sx = sh[0] * G.boxf[0], sy = sh[1] * G.boxf[1], sz = sh[2] * G.boxf[2];
    const uint32_t code = shift_code(sh[0], sh[1], sh[2]);
    const int gcell = sd.replica * G.ncell + (w[2] * G.nc[1] + w[1]) * G.nc[0] + w[0];
    return scan_cell_clusters(V, sd, sb, V.cell_slot[gcell], V.cell_slot[gcell + 1], sx, sy, sz, code, 0, emit);
}

// Column layout: one search item = (sci, xy stencil offset, chunk cell of that column).  The z
// index of a cell is a chunk number, so every chunk of the neighbour column is a candidate and
// the (up to three) periodic images along z are tried against the cell's box.
template <class Emit>
SDM_HD int search_item_columns(const SearchView& V, int isci, int off, Emit emit) {
    const Grid& G = V.G;
    const SciDesc sd = V.sci[isci];
    const BBox sb = V.sci_box[isci];
    if (box_empty(sb)) return 0;
    int cmin[3], cmax[3];
    search_range(G, sb, cmin, cmax);   // x and y only
    const int ox = off % G.span, oy = (off / G.span) % G.span, chunk = off / (G.span * G.span);
    const int u[2] = {cmin[0] + ox, cmin[1] + oy};
    if (u[0] > cmax[0] || u[1] > cmax[1]) return 0;
    int w[2], sh[2];
    for (int d = 0; d < 2; d++) {
        if (G.periodic) {
            int q = u[d] >= 0 ? u[d] / G.nc[d] : -((-u[d] + G.nc[d] - 1) / G.nc[d]);
            w[d] = u[d] - q * G.nc[d];
            sh[d] = q;
            if (q < -1 || q > 1) return 0;  // excluded by the host-side box-size check
        } else {
            w[d] = u[d];
            sh[d] = 0;
        }
    }
    const int gcell = sd.replica * G.ncell + (chunk * G.nc[1] + w[1]) * G.nc[0] + w[0];
    const int s0 = V.cell_slot[gcell], s1 = V.cell_slot[gcell + 1];
    if (s0 == s1) return 0;
    const BBox cb = V.cell_box[gcell];
    if (box_empty(cb)) return 0;
    const float sx = sh[0] * G.boxf[0], sy = sh[1] * G.boxf[1];
    int count = 0;
    for (int iz = (G.periodic ? -1 : 0); iz <= (G.periodic ? 1 : 0); iz++) {
        const float sz = iz * G.boxf[2];
        if (box_dist2(sb, cb, sx, sy, sz) >= G.rlist2) continue;
        count = scan_cell_clusters(V, sd, sb, s0, s1, sx, sy, sz, shift_code(sh[0], sh[1], iz), count, emit);
    }
    return count;
}

template <class Emit>
SDM_HD int search_any(const SearchView& V, int isci, int off, Emit emit) {
    return V.G.columns ? search_item_columns(V, isci, off, emit) : search_item(V, isci, off, emit);
}

// ---- stage 3b: exact pruning of a raw entry ----------------------------------------------------
// Keeps imask bit ci only if some real atom pair of (cluster c0+ci, j-cluster B shifted by the
// entry's image) is closer than rlist.  Scalar form (host checker); the device pass in
// pairlist.cu evaluates the same predicate with one warp per entry.
SDM_HD uint32_t prune_imask(const SearchView& V, const SciDesc& sd, uint32_t w0, uint32_t imask) {
    const Grid& G = V.G;
    const int B = (int)(w0 & 0x3ffffffu);
    const uint32_t code = w0 >> 26;
    const float sx = shift_x(code) * G.boxf[0], sy = shift_y(code) * G.boxf[1], sz = shift_z(code) * G.boxf[2];
    uint32_t out = 0;
    for (int ci = 0; ci < sd.nci; ci++)
        if (((imask >> ci) & 1u) && any_pair_within(V.posq4, sd.c0 + ci, B, sx, sy, sz, G.rlist2)) out |= 1u << ci;
    return out;
}

// Per-cluster hit bytes of an entry (scalar form of what prune_kernel records next to the pruned
// imask): bit tj of byte ci = j-atom tj of the entry's shifted j-cluster is closer than rlist to
// some real atom of cluster c0+ci.  jh_lo holds clusters 0..3, jh_hi clusters 4..7.
SDM_HD void entry_hits(const SearchView& V, const SciDesc& sd, uint32_t w0, uint32_t imask,
                       uint32_t* jh_lo, uint32_t* jh_hi) {
    const Grid& G = V.G;
    const int B = (int)(w0 & 0x3ffffffu);
    const uint32_t code = w0 >> 26;
    const float sx = shift_x(code) * G.boxf[0], sy = shift_y(code) * G.boxf[1], sz = shift_z(code) * G.boxf[2];
    *jh_lo = *jh_hi = 0u;
    for (int ci = 0; ci < sd.nci; ci++) {
        if (!((imask >> ci) & 1u)) continue;
        uint32_t h8 = 0u;
        for (int tj = 0; tj < kJGroup; tj++) {
            const float* pj = V.posq4 + 4 * (size_t)(B * kJGroup + tj);
            if (pj[0] >= 0.5f * kFar) continue;
            const float xj = pj[0] + sx, yj = pj[1] + sy, zj = pj[2] + sz;
            for (int ti = 0; ti < kClusterSize; ti++) {
                const float* pi = V.posq4 + 4 * (size_t)((sd.c0 + ci) * kClusterSize + ti);
                if (pi[0] >= 0.5f * kFar) continue;
                const float dx = pi[0] - xj, dy = pi[1] - yj, dz = pi[2] - zj;
                if (dx * dx + dy * dy + dz * dz < G.rlist2) { h8 |= 1u << tj; break; }
            }
        }
        if (ci < 4) *jh_lo |= h8 << (8 * ci);
        else *jh_hi |= h8 << (8 * (ci - 4));
    }
}

// ---- stage 5: per-atom j rows -------------------------------------------------------------------
// The pair kernel does not walk the cluster-pair entries themselves: every i-group (G = 1 or 2
// consecutive clusters of a supercluster, 8*G atoms) gets a ROW of individual j-atoms -- the atoms
// of its entries' j-clusters that are within rlist of at least one atom of the group and not
// excluded from all of them.  A j-atom that is out of reach of all 8*G i-atoms costs nothing, which
// raises the share of evaluated lane pairs that are inside the cutoff from 35 % (8 x 8 tiles) to
// 50 % (G = 1) on the 20 k-atom fixture.  Row entries with an exclusion / triangle mask come first
// and carry an allow word (bit a = i-atom a of the group may interact with this j-atom).
constexpr int kRowChunkSteps = 32;   // warp steps (32 j-atoms each) per work unit, at most 32

// What the row construction needs to know about a j-cluster, gathered once per list build: which of
// its atoms carry no Lennard-Jones term, and its EXCLUSION TILES -- the (few) clusters that hold an
// excluded partner of one of its atoms, each with the 8 x 8 bit matrix of the excluded pairs.  With
// this record in registers the allow word of a row entry needs no further memory access (walking
// the exclusion CSR per j-atom -- four dependent loads per partner -- stalled whole warps of the row
// kernels).  npart < 0: more partner clusters than the record holds (0.6 % of the clusters of the
// 20 k-atom fixture: spatially sorted atoms split most molecules over clusters, the mean is five
// partner clusters), the CSR is walked instead.  The matrices live in a separate array
// (ClusterTiles): only the few cells that have a tile on their i-group read one.
constexpr int kMaxPartners = 10;
struct ClusterInfo {
    uint32_t nolj;                  // bit tj: epsilon of atom tj is zero (or the slot is a dummy)
    int npart;
    int part[kMaxPartners];         // partner clusters (global cluster index)
};
struct ClusterTiles {
    uint64_t mask[kMaxPartners];    // bit (8*tj + ia): atom tj of this cluster is excluded with atom ia of part[k]
};

// par2: [nslot][2] floats (sigma/2, 2*sqrt(eps))
SDM_HD ClusterInfo cluster_info(int c, const int* atom, const float* par2, const int* excl_start, const int* excl_idx,
                                const int* slot_of, int n, ClusterTiles* tiles) {
    ClusterInfo ci;
    ci.nolj = 0u;
    ci.npart = 0;
    for (int k = 0; k < kMaxPartners; k++) { ci.part[k] = -1; tiles->mask[k] = 0ull; }
    for (int tj = 0; tj < kJGroup; tj++) {
        const int s = c * kJGroup + tj;
        if (par2[2 * (size_t)s + 1] == 0.f) ci.nolj |= 1u << tj;
        const int ga = atom[s];
        if (ga < 0) continue;
        const int r = ga / n, a = ga - r * n;
        for (int k = excl_start[a]; k < excl_start[a + 1]; k++) {
            const int sp = slot_of[r * n + excl_idx[k]];
            const int A = sp / kClusterSize;
            int w = 0;
            while (w < ci.npart && ci.part[w] != A) w++;
            if (w == ci.npart) {
                if (ci.npart < 0 || ci.npart == kMaxPartners) { ci.npart = -1; continue; }
                ci.part[ci.npart++] = A;
            }
            if (ci.npart >= 0) tiles->mask[w] |= 1ull << (8 * tj + sp % kClusterSize);
        }
    }
    return ci;
}

// bit t of the result = any of bits 4t..4t+3 of a warp ballot over lanes (tj, ti) = (lane>>2, lane&3)
SDM_HD uint32_t compress_nibbles(uint32_t b) {
    uint32_t x = b | (b >> 1);
    x |= x >> 2;
    x &= 0x11111111u;
    x = (x | (x >> 3)) & 0x03030303u;
    x = (x | (x >> 6)) & 0x000f000fu;
    return (x | (x >> 12)) & 0xffu;
}

// j-atoms (bit tj) of an entry that are within rlist of some atom of i-group g: OR of the per-cluster
// hit bytes (jh_lo: clusters 0..3, jh_hi: clusters 4..7) of the group's clusters that are in imask.
SDM_HD uint32_t row_hits(uint32_t jh_lo, uint32_t jh_hi, uint32_t imask, int g, int G) {
    uint32_t h = 0;
    for (int q = 0; q < G; q++) {
        const int ci = g * G + q;
        if (!((imask >> ci) & 1u)) continue;
        h |= ((ci < 4 ? jh_lo >> (8 * ci) : jh_hi >> (8 * (ci - 4))) & 0xffu);
    }
    return h;
}

// Allow word of the j-atom tj of an entry (j-cluster B under image `code`) for i-group g of the
// supercluster sd: bit (8*q + ia) = i-atom ia of the group's q-th cluster may interact with it.
//   * a cluster against itself under the zero shift keeps the triangle j index > i index;
//   * a cluster of the group that is not in imask is switched off when B lies in the same
//     supercluster (the other cluster owns that pair); elsewhere it stays on -- its atoms are further
//     than rlist from the whole j-cluster, and an all-ones word keeps the entry on the unmasked path;
//   * every excluded partner of the j-atom (excl_start / excl_idx: the System's exclusions as a CSR
//     over atoms, both directions) that sits in the group is switched off.
// atom[slot] = replica*n + atom, slot_of = its inverse.
// The part of an allow word that does not depend on exclusions: triangle and ownership.
SDM_HD uint32_t row_allow_base(const SciDesc& sd, uint32_t imask, int B, uint32_t code, int g, int G, int tj) {
    const bool same_sci = B >= sd.c0 && B < sd.c0 + sd.nci;
    uint32_t allow = 0;
    for (int q = 0; q < G; q++) {
        const int ci = g * G + q;
        uint32_t bits = 0xffu;
        if (!((imask >> ci) & 1u)) {
            if (same_sci) bits = 0u;
        } else if (code == kShiftZero && B == sd.c0 + ci) {
            bits = (1u << tj) - 1u;
        }
        allow |= bits << (8 * q);
    }
    return allow;
}

// Exclusions from the j-cluster's exclusion tiles (info.npart >= 0), in two steps so that the tile
// search runs once per (entry, i-group) cell and not once per j-atom: ex[q] = the 8 x 8 exclusion
// matrix (bit 8*tj + ia) between the j-cluster and the group's q-th cluster, zero when there is none.
// Fixed trip counts and constant indices: the record stays in registers.
SDM_HD bool row_excl_tiles(const SciDesc& sd, int g, int G, const ClusterInfo& info, const ClusterTiles* tiles,
                           uint64_t ex[2]) {
    ex[0] = ex[1] = 0ull;
    bool any = false;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int k = 0; k < kMaxPartners; k++) {
        const int q = info.part[k] - (sd.c0 + g * G);
        if (k < info.npart && q >= 0 && q < G) {
            if (q == 0) ex[0] |= tiles->mask[k]; else ex[1] |= tiles->mask[k];
            any = true;
        }
    }
    return any;
}

SDM_HD uint32_t row_allow(const SciDesc& sd, uint32_t imask, int B, uint32_t code, int g, int G, int tj,
                          const uint64_t ex[2]) {
    uint32_t allow = row_allow_base(sd, imask, B, code, g, G, tj);
    allow &= ~(uint32_t)((ex[0] >> (8 * tj)) & 0xffull);
    if (G == 2) allow &= ~((uint32_t)((ex[1] >> (8 * tj)) & 0xffull) << 8);
    return allow;
}

SDM_HD uint32_t row_allow(const SciDesc& sd, uint32_t imask, int B, uint32_t code, int g, int G, int tj,
                          const ClusterInfo& info, const ClusterTiles* tiles) {
    uint64_t ex[2];
    row_excl_tiles(sd, g, G, info, tiles, ex);
    return row_allow(sd, imask, B, code, g, G, tj, ex);
}

// Exclusions by walking the System's exclusion CSR for the j-atom (clusters whose partners do not
// fit the record, and the reference the host checker can compare the tiles with).
SDM_HD uint32_t row_allow(const SciDesc& sd, uint32_t imask, int B, uint32_t code, int g, int G, int tj,
                          const int* excl_start, const int* excl_idx, const int* slot_of, const int* atom, int n) {
    uint32_t allow = row_allow_base(sd, imask, B, code, g, G, tj);
    const int ga = atom[B * kJGroup + tj];
    if (ga < 0) return 0u;
    const int r = ga / n, a = ga - r * n;
    const int first = (sd.c0 + g * G) * kClusterSize;
    for (int k = excl_start[a]; k < excl_start[a + 1]; k++) {
        const int rel = slot_of[r * n + excl_idx[k]] - first;
        if (rel >= 0 && rel < G * kClusterSize) allow &= ~(1u << rel);
    }
    return allow;
}

}  // namespace nbl
}  // namespace sdm
