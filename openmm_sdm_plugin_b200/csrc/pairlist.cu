// pairlist.cu -- builds the pair list on the device (per-item bodies from nblist_core.h, CUB for
// the sorts and the scans): cell-sorted 8-atom clusters, cluster-pair entries from the search and
// the exact prune, and from those the per-atom j rows the pair kernel (kernels_rows.cu) walks.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cub/cub.cuh>

#include "pairlist.h"
#include "sdm_ctx.h"

namespace sdm {

using nbl::BBox;
using nbl::Grid;
using nbl::SciDesc;

struct PairList {
    Grid G{};
    int ncells = 0;          // R * ncell
    int cell_cap = 0;        // allocated cells per replica
    int nslot_cap = 0, ncl_cap = 0, nsci_cap = 0;
    size_t entries_cap = 0, items_cap = 0;
    // sizes of the current list (host copies)
    int nslot = 0, ncl = 0, nsci = 0, nentries = 0, scan_max = 0;

    uint64_t *keys = nullptr, *keys_sorted = nullptr;
    int *vals = nullptr, *vals_sorted = nullptr;
    int *cell_first = nullptr, *cell_count = nullptr, *cell_pcount = nullptr, *cell_nsci = nullptr;
    int *cell_slot = nullptr, *cell_sci = nullptr;
    float4 *posq = nullptr, *posq_build = nullptr;
    float4* jrec = nullptr;     // [nslot][2] (posq, par) side by side: the pair kernel's gather record
    float2* par = nullptr;
    int *atom = nullptr, *img = nullptr, *slot_of = nullptr;
    BBox *cl_box = nullptr, *sci_box = nullptr, *cell_box = nullptr;
    nbl::ClusterInfo* cl_info = nullptr;   // [ncl] partner clusters and LJ-free atoms of every cluster
    nbl::ClusterTiles* cl_tiles = nullptr; // [ncl] its exclusion tiles
    int use_columns = 1;        // column layout (default) or geometric 3-D cells (SDMB200_LAYOUT=cells)
    SciDesc* sci = nullptr;
    int* cl_sci = nullptr;
    int *item_count = nullptr, *item_off = nullptr;
    uint2* entries = nullptr;
    int* entry_sci = nullptr;
    // raw (box-pruned) entries of the search, before the exact prune + compaction
    uint2* raw_entries = nullptr;
    int *raw_sci = nullptr, *raw_c0nci = nullptr, *raw_keep = nullptr, *raw_pos = nullptr;
    size_t raw_cap = 0;
    int nraw = 0;
    // per-atom j rows (nblist_core.h stage 5): the list the pair kernel walks
    int row_group = 1;          // clusters per i-group (SDMB200_ROW_GROUP = 1 | 2)
    int row_chunk = nbl::kRowChunkSteps;   // warp steps per unit (SDMB200_ROW_CHUNK)
    uint2 *raw_jhit = nullptr, *entry_jhit = nullptr;   // per (cluster of the sci, j-atom) hit bits of an entry
    int *seg_total = nullptr, *seg_off = nullptr;       // [nsci * (8 / G) * 3 + 1] row segments
    uint32_t* jent = nullptr;
    uint16_t* jallow = nullptr;
    size_t jent_cap = 0;
    int *row_nunits = nullptr, *row_unit_off = nullptr; // [nsci * (8 / G) + 1]
    RowUnit* runits = nullptr;
    size_t runits_cap = 0;
    int row_lpt = 1;            // draw the longest units first (SDMB200_ROW_LPT=0: list order)
    uint32_t *ru_key = nullptr, *ru_key_sorted = nullptr;
    int *ru_val = nullptr, *ru_order = nullptr;
    int njent = 0, nrunits = 0;
    int dummy_slot = 0;
    int* sci_off = nullptr;     // [nsci+1] first (compacted) entry of every sci
    int* part_off = nullptr;    // [R+1]
    int* unit_counter = nullptr; // work counter of the persistent pair kernel
    unsigned int* max_disp2 = nullptr;   // largest squared displacement since the build (float bits)
    double* epart = nullptr;
    long long* cpart = nullptr;
    double* minmax = nullptr;   // [6] non-periodic extent reduction
    void* cub_tmp = nullptr;
    size_t cub_tmp_bytes = 0;
    int* h_counts = nullptr;    // pinned [16]: counts of the last build (see kCnt*)
    int* d_cnt = nullptr;       // device [16]
    cudaEvent_t ev_counts = nullptr;   // the read-back of d_cnt of the last build has landed in h_counts
    bool counts_pending = false;       // that read-back has not been looked at yet
    bool have_counts = false;          // the host counts describe a complete earlier build
    bool force_sync = false;           // the next build sizes its buffers step by step on the host
    int kd_in_block = 1;               // SDMB200_KD_SORTS=1: the two kd rounds as global radix sorts (development knob)
    int async_builds = 1;              // SDMB200_ASYNC_BUILD=0: every build synchronises (development knob)
    // launch bounds of the device-sized stages (upper bounds of the counts the kernels read from device memory)
    int ub_nsci = 0, ub_nraw = 0, ub_nrunits = 0;
    int64_t n_async = 0, n_sync = 0;
    double* h_minmax = nullptr; // pinned [6]
    std::vector<void*> allocs;
    double density_hint = 0;    // atoms / nm^3 used to size cells
};

namespace {

#define PL_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            char buf_[512];                                                                    \
            snprintf(buf_, sizeof(buf_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                     __FILE__, __LINE__);                                                      \
            return sdm_fail(SDM_ERR_CUDA, buf_);                                               \
        }                                                                                      \
    } while (0)

template <class T>
int pl_alloc(PairList* pl, T** p, size_t count) {
    void* q = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    PL_CUDA(cudaMalloc(&q, bytes));
    PL_CUDA(cudaMemset(q, 0, bytes));
    pl->allocs.push_back(q);
    *p = static_cast<T*>(q);
    return SDM_OK;
}

template <class T>
int pl_realloc(PairList* pl, T** p, size_t count) {
    if (*p) {
        auto it = std::find(pl->allocs.begin(), pl->allocs.end(), (void*)*p);
        if (it != pl->allocs.end()) pl->allocs.erase(it);
        PL_CUDA(cudaFree(*p));
        *p = nullptr;
    }
    return pl_alloc(pl, p, count);
}

// ---- device wrappers around the per-item bodies ------------------------------------------------
// round 0 of the sort: (cell, z)
__global__ void key_kernel(Grid G, const double* __restrict__ pos_all, uint64_t* keys, int* vals) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G.R * G.n) return;
    const int r = t / G.n;
    const double* p = pos_all + 3 * (size_t)t;
    float xw[3];
    int img[3];
    uint32_t fr[3];
    const uint32_t g = nbl::atom_cell(G, r, p[0], p[1], p[2], xw, img, fr);
    keys[t] = nbl::make_key(g, 0u, fr[2]);
    vals[t] = t;
}

// column layout: from the (column, z) order to the key of the chunk cell (see nbl::chunk_cell)
__global__ void chunk_key_kernel(Grid G, int total, const uint64_t* __restrict__ keys_sorted,
                                 const int* __restrict__ vals_sorted, const int* __restrict__ col_first,
                                 uint64_t* keys, int* vals) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const uint32_t pc = (uint32_t)(keys_sorted[p] >> nbl::kSubBits);
    uint32_t rin;
    const uint32_t g = nbl::chunk_cell(G, pc, p - col_first[pc], &rin);
    keys[p] = nbl::make_key(g, 0u, rin);
    vals[p] = vals_sorted[p];
}

__global__ void cell_box_kernel(int ncells, const int* __restrict__ cell_slot, const BBox* __restrict__ cl_box,
                                BBox* cell_box) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    BBox b;
    for (int d = 0; d < 3; d++) { b.lo[d] = nbl::kBoxEmptyLo; b.hi[d] = -nbl::kBoxEmptyLo; }
    for (int k = cell_slot[c] / nbl::kClusterSize; k < cell_slot[c + 1] / nbl::kClusterSize; k++)
        if (!nbl::box_empty(cl_box[k])) b = nbl::box_union(b, cl_box[k]);
    cell_box[c] = b;
}

// rounds 1 and 2: bucket of the previous split + the next coordinate (y, then x)
__global__ void refine_key_kernel(Grid G, int level, int total, const double* __restrict__ pos_all,
                                  const uint64_t* __restrict__ keys_sorted, const int* __restrict__ vals_sorted,
                                  const int* __restrict__ cell_first, const int* __restrict__ cell_count,
                                  uint64_t* keys, int* vals) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const uint64_t k = keys_sorted[p];
    const uint32_t c = (uint32_t)(k >> nbl::kSubBits);
    const int ga = vals_sorted[p];
    const uint32_t b = nbl::kd_bucket(level, p, cell_first[c], cell_count[c], nbl::key_bucket(k));
    const double* q = pos_all + 3 * (size_t)ga;
    float xw[3];
    int img[3];
    uint32_t fr[3];
    nbl::atom_cell(G, ga / G.n, q[0], q[1], q[2], xw, img, fr);
    keys[p] = nbl::make_key(c, b, fr[level == 1 ? 1 : 0]);
    vals[p] = ga;
}

// The two kd refinement rounds of every cell in ONE kernel, one block per cell (replaces two global
// radix sorts): the atoms of the cell arrive in z order; they are ordered by y inside each z half and by
// x inside each y half of those -- the same keys (16-bit in-cell coordinates, ties keep the previous
// order) and the same split points (nbl::kd_bucket) as the sort-based rounds, so the result is the
// order those produce.  Ranks are counted directly (a cell holds about 64 atoms; an overfull one just
// takes longer); `scratch` holds y | x << 16 | position after the first round << 32 per atom.
__global__ void __launch_bounds__(64)
kd_refine_kernel(Grid G, const double* __restrict__ pos_all, const int* __restrict__ vals_sorted,
                 const int* __restrict__ cell_first, const int* __restrict__ cell_count,
                 uint64_t* __restrict__ scratch, int* __restrict__ vals_out) {
    const int c = blockIdx.x;
    const int first = cell_first[c], cnt = cell_count[c];
    if (cnt <= 0) return;
    const int h1 = nbl::split_first(cnt);
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const int ga = vals_sorted[first + i];
        const double* q = pos_all + 3 * (size_t)ga;
        float xw[3];
        int img[3];
        uint32_t fr[3];
        nbl::atom_cell(G, ga / G.n, q[0], q[1], q[2], xw, img, fr);
        scratch[first + i] = (uint64_t)fr[1] | ((uint64_t)fr[0] << 16);
    }
    __syncthreads();
    // round 1: by y inside the z half
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const int b1 = i >= h1;
        const int s0 = b1 ? h1 : 0, s1 = b1 ? cnt : h1;
        const uint32_t yi = (uint32_t)scratch[first + i] & 0xffffu;
        int rank = s0;
        for (int j = s0; j < s1; j++) {
            const uint32_t yj = (uint32_t)scratch[first + j] & 0xffffu;
            rank += (yj < yi || (yj == yi && j < i)) ? 1 : 0;
        }
        scratch[first + i] |= (uint64_t)rank << 32;
    }
    __syncthreads();
    // round 2: by x inside the y half of the z half; then the atom's final place
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const int b1 = i >= h1;
        const int s0 = b1 ? h1 : 0, s1 = b1 ? cnt : h1;
        const int h2 = nbl::split_first(s1 - s0);
        const uint64_t wi = scratch[first + i];
        const int pi = (int)(wi >> 32);
        const int b2 = (pi - s0) >= h2;
        const uint32_t xi = (uint32_t)(wi >> 16) & 0xffffu;
        int rank = s0 + (b2 ? h2 : 0);
        for (int j = s0; j < s1; j++) {
            const uint64_t wj = scratch[first + j];
            const int pj = (int)(wj >> 32);
            if (((pj - s0) >= h2) != b2) continue;
            const uint32_t xj = (uint32_t)(wj >> 16) & 0xffffu;
            rank += (xj < xi || (xj == xi && pj < pi)) ? 1 : 0;
        }
        vals_out[first + rank] = vals_sorted[first + i];
    }
}

__global__ void cell_bounds_kernel(int total, const uint64_t* __restrict__ keys_sorted,
                                   int* cell_first, int* cell_count) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    const uint32_t c = (uint32_t)(keys_sorted[k] >> nbl::kSubBits);
    if (k == 0 || (uint32_t)(keys_sorted[k - 1] >> nbl::kSubBits) != c) cell_first[c] = k;
    if (k == total - 1 || (uint32_t)(keys_sorted[k + 1] >> nbl::kSubBits) != c) {
        // the matching "first" is written by another thread of this launch; store the end and
        // let the next kernel subtract
        cell_count[c] = k + 1;
    }
}

__global__ void cell_sizes_kernel(int ncells, const int* __restrict__ cell_first,
                                  int* cell_count /* in: end or 0, out: count */, int* cell_pcount,
                                  int* cell_nsci) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int end = cell_count[c];
    const int cnt = end > 0 ? end - cell_first[c] : 0;
    cell_count[c] = cnt;
    const int pc = (cnt + nbl::kClusterSize - 1) / nbl::kClusterSize * nbl::kClusterSize;
    cell_pcount[c] = pc;
    const int ncl = pc / nbl::kClusterSize;
    cell_nsci[c] = (ncl + nbl::kMaxCi - 1) / nbl::kMaxCi;
}

__global__ void fill_slots_kernel(Grid G, Topology T, int total, const double* __restrict__ pos_all,
                                  const uint64_t* __restrict__ keys_sorted,
                                  const int* __restrict__ vals_sorted,
                                  const int* __restrict__ cell_first, const int* __restrict__ cell_slot,
                                  float4* posq, float2* par, int* atom, int* img, int* slot_of) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    const uint32_t c = (uint32_t)(keys_sorted[k] >> nbl::kSubBits);
    const int ga = vals_sorted[k];
    const int r = ga / G.n, a = ga - r * G.n;
    const int slot = cell_slot[c] + (k - cell_first[c]);
    const double* p = pos_all + 3 * (size_t)ga;
    float xw[3];
    int im[3];
    nbl::atom_cell(G, r, p[0], p[1], p[2], xw, im);
    const float4 pf = T.parf[a];
    posq[slot] = make_float4(xw[0], xw[1], xw[2], pf.x);
    par[slot] = make_float2(pf.y, pf.z);
    atom[slot] = ga;
    img[slot] = (im[0] + 512) | ((im[1] + 512) << 10) | ((im[2] + 512) << 20);
    slot_of[ga] = slot;
}

__global__ void fill_dummies_kernel(int ncells, const int* __restrict__ cell_count,
                                    const int* __restrict__ cell_slot, float4* posq, float2* par,
                                    int* atom, int* img) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int s0 = cell_slot[c] + cell_count[c], s1 = cell_slot[c + 1];
    for (int s = s0; s < s1; s++) {
        posq[s] = make_float4(nbl::kFar, nbl::kFar, nbl::kFar, 0.f);
        par[s] = make_float2(0.f, 0.f);
        atom[s] = -1;
        img[s] = 512 | (512 << 10) | (512 << 20);
    }
}

__global__ void bbox_kernel(const int* __restrict__ d_nslot, const float4* __restrict__ posq, BBox* cl_box) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= *d_nslot / nbl::kClusterSize) return;
    cl_box[c] = nbl::group_bbox(reinterpret_cast<const float*>(posq), c * nbl::kClusterSize,
                                nbl::kClusterSize);
}

__global__ void sci_kernel(Grid G, int ncells, const int* __restrict__ cell_slot,
                           const int* __restrict__ cell_sci, const BBox* __restrict__ cl_box,
                           SciDesc* sci, BBox* sci_box, int* cl_sci) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int cl0 = cell_slot[c] / nbl::kClusterSize, cl1 = cell_slot[c + 1] / nbl::kClusterSize;
    int s = cell_sci[c];
    for (int k = cl0; k < cl1; k += nbl::kMaxCi, s++) {
        SciDesc d;
        d.c0 = k;
        d.nci = min(nbl::kMaxCi, cl1 - k);
        d.replica = c / G.ncell;
        d.pad = 0;
        BBox b = cl_box[k];
        cl_sci[k] = s;
        for (int j = 1; j < d.nci; j++) {
            b = nbl::box_union(b, cl_box[k + j]);
            cl_sci[k + j] = s;
        }
        sci[s] = d;
        sci_box[s] = b;
    }
}

struct CountEmit {
    __device__ void operator()(int, uint32_t, uint32_t, bool) const {}
};
struct FillEmit {
    uint2* out;
    int* esci;
    int* c0nci;      // first cluster | cluster count << 27 of the owning supercluster (for the prune pass)
    int base, isci, sd_c0nci;
    __device__ void operator()(int k, uint32_t w0, uint32_t imask, bool) const {
        out[base + k] = make_uint2(w0, imask);
        esci[base + k] = isci;
        c0nci[base + k] = sd_c0nci;
    }
};

// The counts a stage reads (superclusters, raw entries, ...) live in device memory; the grids are sized
// with host-side upper bounds (ub_*), so a build needs no host synchronisation.  Items past the real
// count get a zero so that the scans can run over the bound.
__global__ void search_count_kernel(nbl::SearchView V, const int* __restrict__ d_nsci, int ub_items, int noff,
                                    int* item_count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ub_items) return;
    item_count[t] = t < *d_nsci * noff ? nbl::search_any(V, t / noff, t % noff, CountEmit()) : 0;
}

__global__ void search_fill_kernel(nbl::SearchView V, const int* __restrict__ d_nsci, int noff,
                                   const int* __restrict__ item_off, uint2* entries, int* esci, int* c0nci,
                                   int cap) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *d_nsci * noff) return;
    const int base = item_off[t];
    if (item_off[t + 1] > cap) return;   // the whole item must fit (overflow is reported by finish_kernel)
    const SciDesc sd = V.sci[t / noff];
    nbl::search_any(V, t / noff, t % noff, FillEmit{entries, esci, c0nci, base, t / noff, sd.c0 | (sd.nci << 27)});
}

// Exact pruning, one warp per raw entry: imask bit ci survives only if some real atom pair of
// (cluster c0+ci, the entry's shifted j-cluster) is closer than rlist -- the predicate of
// nbl::prune_imask.  Lane (tj, ti) = (lane>>2, lane&3) tests j-atom tj against i-atoms ti, ti+4.
__global__ void __launch_bounds__(128)
prune_kernel(Grid G, const int* __restrict__ d_nraw, int ub_nraw, int raw_cap, uint2* __restrict__ raw,
             const int* __restrict__ raw_c0nci, const float4* __restrict__ posq, int* __restrict__ keep,
             uint2* __restrict__ jhit) {
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (e >= min(*d_nraw, raw_cap)) {
        if (e <= ub_nraw && lane == 0) keep[e] = 0;   // the scan runs over ub_nraw + 1 flags
        return;
    }
    const uint2 ent = raw[e];                 // two independent loads, then the atoms: no
    const int c0 = raw_c0nci[e] & 0x7ffffff;  // dependent descriptor chain per entry
    const int B = (int)(ent.x & 0x3ffffffu);
    const uint32_t code = ent.x >> 26;
    const int tj = lane >> 2, ti = lane & 3;
    float4 xj = posq[B * nbl::kJGroup + tj];
    const bool jreal = xj.x < 0.5f * nbl::kFar;
    xj.x += (float)nbl::shift_x(code) * G.boxf[0];
    xj.y += (float)nbl::shift_y(code) * G.boxf[1];
    xj.z += (float)nbl::shift_z(code) * G.boxf[2];
    uint32_t todo = ent.y & 0xffu;
    uint32_t mine = 0u;   // bit ci: one of this lane's two atom pairs with cluster ci is inside rlist
    while (todo) {
        const int ci = __ffs(todo) - 1;
        todo &= todo - 1u;
        const float4* pi = posq + (size_t)(c0 + ci) * nbl::kClusterSize + ti;
        const float4 a = pi[0], b = pi[4];
        bool hit = false;
        if (jreal) {
            float dx = a.x - xj.x, dy = a.y - xj.y, dz = a.z - xj.z;
            hit = a.x < 0.5f * nbl::kFar && dx * dx + dy * dy + dz * dz < G.rlist2;
            dx = b.x - xj.x; dy = b.y - xj.y; dz = b.z - xj.z;
            hit = hit || (b.x < 0.5f * nbl::kFar && dx * dx + dy * dy + dz * dz < G.rlist2);
        }
        mine |= (hit ? 1u : 0u) << ci;
    }
    // OR over the four ti lanes of a j-atom: the clusters that reach j-atom tj; then every j-atom
    // drops its bits into the per-cluster hit bytes (bit tj of byte ci) and the warp ORs them together
    mine |= __shfl_xor_sync(0xffffffffu, mine, 1);
    mine |= __shfl_xor_sync(0xffffffffu, mine, 2);
    uint32_t jh_lo = 0u, jh_hi = 0u;   // per cluster of the sci: which j-atoms have an i-atom within rlist
#pragma unroll
    for (int ci = 0; ci < 4; ci++) {
        jh_lo |= ((mine >> ci) & 1u) << (8 * ci + tj);
        jh_hi |= ((mine >> (ci + 4)) & 1u) << (8 * ci + tj);
    }
    jh_lo = __reduce_or_sync(0xffffffffu, jh_lo);
    jh_hi = __reduce_or_sync(0xffffffffu, jh_hi);
    const uint32_t out = __reduce_or_sync(0xffffffffu, mine);
    if (lane == 0) {
        jhit[e] = make_uint2(jh_lo, jh_hi);
        raw[e].y = out;
        keep[e] = out != 0u;
    }
}

__global__ void compact_kernel(const int* __restrict__ d_nraw, int raw_cap, const uint2* __restrict__ raw,
                               const int* __restrict__ raw_sci, const int* __restrict__ keep,
                               const int* __restrict__ pos, uint2* entries, int* entry_sci, int cap,
                               const uint2* __restrict__ raw_jhit, uint2* entry_jhit) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= min(*d_nraw, raw_cap) || !keep[e]) return;
    const int k = pos[e];
    if (k >= cap) return;
    entries[k] = raw[e];
    entry_jhit[k] = raw_jhit[e];
    entry_sci[k] = raw_sci[e];
}

__global__ void sci_off_kernel(const int* __restrict__ d_nsci, int noff, const int* __restrict__ d_nraw, int ub_nraw,
                               const int* __restrict__ item_off, const int* __restrict__ pos, int* sci_off) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int nsci = *d_nsci;
    if (s > nsci) return;
    const int first = s < nsci ? item_off[(size_t)s * noff] : *d_nraw;
    sci_off[s] = pos[min(first, ub_nraw)];   // pos: exclusive scan over ub_nraw + 1 flags, zero past the real count
}

// ---- per-atom j rows (nblist_core.h stage 5) -----------------------------------------------------
// A row (supercluster s, i-group g) is laid out [class 0 | class 1 | class 2] -- 0 = entries that
// carry an allow word, 1 = plain, 2 = plain and the j-atom has no Lennard-Jones term (epsilon == 0)
// -- each class in entry order, rows in cluster order.  One block per supercluster, one warp per
// i-group, one lane per entry of a 32-entry chunk: the first pass (FILL = false) only sums the three
// class counts of every row (seg_total); one small global scan over those nsci * groups * 3 totals
// places the segments (seg_off); the second pass repeats the walk, turns the lanes' counts into
// positions with one packed warp scan per chunk and writes the row entries -- consecutive lanes write
// consecutive pieces of the same three segments.  Nothing per (entry, group) cell ever goes to memory.
struct RowsIn {
    int G, n;
    const uint2* entries;
    const uint2* jhit;
    const SciDesc* sci;
    const int* sci_off;
    const nbl::ClusterInfo* cl_info;
    const nbl::ClusterTiles* cl_tiles;
    const int *excl_start, *excl_idx, *slot_of, *atom;
};

// nbl::cluster_info with one lane per atom (32 clusters per block): every lane walks the exclusions of
// its own atom -- the dependent loads of the eight atoms of a cluster overlap -- and the lanes of a
// cluster then take turns to enter their partners into the cluster's record in shared memory, in
// atom order, so the record is the one the scalar routine builds.
__global__ void __launch_bounds__(256)
cluster_info_kernel(const int* __restrict__ d_nslot, int n, const int* __restrict__ atom, const float2* __restrict__ par,
                    const int* __restrict__ excl_start, const int* __restrict__ excl_idx,
                    const int* __restrict__ slot_of, nbl::ClusterInfo* __restrict__ out,
                    nbl::ClusterTiles* __restrict__ tiles) {
    __shared__ nbl::ClusterInfo s_info[32];
    __shared__ nbl::ClusterTiles s_tiles[32];
    const int lc = threadIdx.x >> 3, tj = threadIdx.x & 7;
    const int c = blockIdx.x * 32 + lc;
    const int ncl = *d_nslot / nbl::kClusterSize;
    if (tj == 0) {
        s_info[lc].nolj = 0u;
        s_info[lc].npart = 0;
        for (int k = 0; k < nbl::kMaxPartners; k++) { s_info[lc].part[k] = -1; s_tiles[lc].mask[k] = 0ull; }
    }
    __syncwarp();
    const bool live = c < ncl;
    const int slot = c * nbl::kJGroup + tj;
    const int ga = live ? atom[slot] : -1;
    if (live && par[slot].y == 0.f) atomicOr(&s_info[lc].nolj, 1u << tj);
    int k0 = 0, k1 = 0, r = 0;
    if (ga >= 0) {
        r = ga / n;
        const int a = ga - r * n;
        k0 = excl_start[a];
        k1 = excl_start[a + 1];
    }
    // the partner slots of this lane's atom, fetched before the turns (all lanes' loads in flight
    // together); an atom with more than kPre exclusions fetches the rest during its turn
    constexpr int kPre = 12;
    int pre[kPre];
#pragma unroll
    for (int k = 0; k < kPre; k++) pre[k] = k0 + k < k1 ? slot_of[r * n + excl_idx[k0 + k]] : -1;
    for (int turn = 0; turn < nbl::kJGroup; turn++) {
        if (turn == tj) {
            nbl::ClusterInfo& ci = s_info[lc];
#pragma unroll
            for (int kk = 0; kk < kPre + 1; kk++) {
              const int kend = kk < kPre ? k0 + kk + 1 : k1;
              for (int k = k0 + kk; k < kend && k < k1; k++) {
                const int sp = kk < kPre ? pre[kk] : slot_of[r * n + excl_idx[k]];
                const int A = sp / nbl::kClusterSize;
                int w = 0;
                while (w < ci.npart && ci.part[w] != A) w++;
                if (w == ci.npart) {
                    if (ci.npart < 0 || ci.npart == nbl::kMaxPartners) { ci.npart = -1; continue; }
                    ci.part[ci.npart++] = A;
                }
                if (ci.npart >= 0) s_tiles[lc].mask[w] |= 1ull << (8 * tj + sp % nbl::kClusterSize);
              }
            }
        }
        __syncwarp();
    }
    if (live && tj == 0) {
        out[c] = s_info[lc];
        tiles[c] = s_tiles[lc];
    }
}

template <bool FILL>
__global__ void __launch_bounds__(256)
rows_kernel(RowsIn in, const int* __restrict__ d_nsci, int* __restrict__ seg_total, const int* __restrict__ seg_off,
            uint32_t* __restrict__ jent, uint16_t* __restrict__ jallow, int cap) {
    const int ng = nbl::kMaxCi / in.G;
    const int s = blockIdx.x, g = threadIdx.x >> 5, lane = threadIdx.x & 31;   // blockDim.x = 32 * ng
    if (s >= *d_nsci) {   // the grid covers the bound: rows past the real count are empty
        if (!FILL && lane < 3) seg_total[3 * ((size_t)s * ng + g) + lane] = 0;
        return;
    }
    const uint32_t full = in.G == 2 ? 0xffffu : 0xffu;
    const SciDesc sd = in.sci[s];
    const int e0 = in.sci_off[s], e1 = in.sci_off[s + 1];
    int base0 = 0, base1 = 0, base2 = 0;   // FILL: next free position of the row's three segments
    if (FILL) {
        const int* so = seg_off + 3 * ((size_t)s * ng + g);
        base0 = so[0]; base1 = so[1]; base2 = so[2];
    }
    int t0 = 0, t1 = 0, t2 = 0;            // !FILL: this lane's share of the segment totals
    for (int c0 = e0; c0 < e1; c0 += 32) {
        const int e = c0 + lane;
        uint32_t hits = 0u, code = 0u;
        int B = 0;
        bool special = false;
        nbl::ClusterInfo info;
        info.nolj = 0u;
        info.npart = 0;
        uint32_t imask = 0u;
        uint64_t ex[2] = {0ull, 0ull};
        if (e < e1) {
            const uint2 ent = in.entries[e];
            const uint2 jh = in.jhit[e];
            imask = ent.y & 0xffu;
            code = ent.x >> 26;
            B = (int)(ent.x & 0x3ffffffu);
            info = in.cl_info[B];
            hits = nbl::row_hits(jh.x, jh.y, imask, g, in.G);
            // special cells: an exclusion tile of the j-cluster falls on a cluster of this i-group, the
            // j-cluster is the group's own cluster (triangle) or -- groups of two -- one of the
            // supercluster's (ownership split); its exclusions did not fit the record (CSR walk)
            const bool own = in.G == 1 ? (B == sd.c0 + g && code == nbl::kShiftZero) : (B >= sd.c0 && B < sd.c0 + sd.nci);
            special = nbl::row_excl_tiles(sd, g, in.G, info, in.cl_tiles + B, ex) || own || info.npart < 0;
        }
        // class counts of this lane's cell; nearly every cell is plain: two popcounts
        int n0 = 0, n1 = 0, n2 = 0;
        uint32_t keep = hits;              // special cells: the hits whose allow word is not empty
        if (!special) {
            n1 = __popc(hits & ~info.nolj);
            n2 = __popc(hits & info.nolj);
        } else {
            uint32_t h = hits;
            while (h) {
                const int tj = __ffs(h) - 1;
                h &= h - 1u;
                const uint32_t allow = info.npart >= 0
                                           ? nbl::row_allow(sd, imask, B, code, g, in.G, tj, ex)
                                           : nbl::row_allow(sd, imask, B, code, g, in.G, tj, in.excl_start, in.excl_idx,
                                                            in.slot_of, in.atom, in.n);
                if (allow == 0u) { keep &= ~(1u << tj); continue; }
                if (allow != full) n0++;
                else if ((info.nolj >> tj) & 1u) n2++;
                else n1++;
            }
        }
        if (!FILL) {
            t0 += n0; t1 += n1; t2 += n2;
            continue;
        }
        // positions: exclusive warp scan of the three counts in one packed word (each sum <= 32 * 8)
        const uint32_t packed = (uint32_t)n0 | ((uint32_t)n1 << 10) | ((uint32_t)n2 << 20);
        uint32_t x = packed;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        const uint32_t tot = __shfl_sync(0xffffffffu, x, 31);
        x -= packed;
        int p0 = base0 + (int)(x & 0x3ffu), p1 = base1 + (int)((x >> 10) & 0x3ffu), p2 = base2 + (int)(x >> 20);
        base0 += (int)(tot & 0x3ffu); base1 += (int)((tot >> 10) & 0x3ffu); base2 += (int)(tot >> 20);
        while (keep) {
            const int tj = __ffs(keep) - 1;
            keep &= keep - 1u;
            const bool nolj = ((info.nolj >> tj) & 1u) != 0u;
            uint32_t allow = full;
            if (special)
                allow = info.npart >= 0 ? nbl::row_allow(sd, imask, B, code, g, in.G, tj, ex)
                                        : nbl::row_allow(sd, imask, B, code, g, in.G, tj, in.excl_start, in.excl_idx,
                                                         in.slot_of, in.atom, in.n);
            const bool masked = allow != full;
            const int p = masked ? p0 : nolj ? p2 : p1;
            p0 += masked ? 1 : 0;
            p2 += (!masked && nolj) ? 1 : 0;
            p1 += (!masked && !nolj) ? 1 : 0;
            if (p < cap) {
                jent[p] = (uint32_t)(B * nbl::kJGroup + tj) | (code << 26);
                if (masked) jallow[p] = (uint16_t)allow;   // plain entries: no allow word is read
            }
        }
    }
    if (!FILL) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            t0 += __shfl_xor_sync(0xffffffffu, t0, o);
            t1 += __shfl_xor_sync(0xffffffffu, t1, o);
            t2 += __shfl_xor_sync(0xffffffffu, t2, o);
        }
        if (lane == 0) {
            int* out = seg_total + 3 * ((size_t)s * ng + g);
            out[0] = t0; out[1] = t1; out[2] = t2;
        }
    }
}

// row (s, g): [begin, mend) masked, [mend, lend) plain, [lend, end) plain without LJ (seg_off: the
// exclusive scan of the segment totals, one more element than segments); units of <= chunk steps
__device__ __forceinline__ void row_bounds(int s, int g, int ng, const int* __restrict__ seg_off, int* begin,
                                           int* mend, int* lend, int* end) {
    const int* so = seg_off + 3 * ((size_t)s * ng + g);
    *begin = so[0];
    *mend = so[1];
    *lend = so[2];
    *end = so[3];
}

__global__ void rows_units_count_kernel(const int* __restrict__ d_nsci, int ub_nsci, int ng, int chunk,
                                        const int* __restrict__ seg_off, int* __restrict__ row_nunits) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ub_nsci * ng) return;
    if (t >= *d_nsci * ng) { row_nunits[t] = 0; return; }
    int b, m, l, e;
    row_bounds(t / ng, t % ng, ng, seg_off, &b, &m, &l, &e);
    row_nunits[t] = (e - b + 32 * chunk - 1) / (32 * chunk);
}

__global__ void rows_units_fill_kernel(const int* __restrict__ d_nsci, int ng, int G, int chunk,
                                       const SciDesc* __restrict__ sci, const int* __restrict__ seg_off,
                                       const int* __restrict__ row_unit_off, RowUnit* __restrict__ units, int cap) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *d_nsci * ng) return;
    const int s = t / ng, g = t % ng;
    int b, m, l, e;
    row_bounds(s, g, ng, seg_off, &b, &m, &l, &e);
    const SciDesc sd = sci[s];
    const int ncl = min(G, sd.nci - g * G);
    int u = row_unit_off[t];
    for (int x = b; x < e; x += 32 * chunk, u++) {
        const int xe = min(x + 32 * chunk, e);
        const int mlen = max(x, min(m, xe)) - x, llen = max(x, min(l, xe)) - x;   // <= 1024 each
        if (u < cap) units[u] = RowUnit{(sd.c0 + g * G) | (ncl << 28), x, xe, mlen | (llen << 16)};
    }
}

// sort key of a unit: longer units first; the radix sort is stable, so units of equal length keep
// the list order (neighbouring i-clusters, which share their j-atoms in L1)
__global__ void rows_unit_key_kernel(const int* __restrict__ d_nunits, int ub_nunits, int cap,
                                     const RowUnit* __restrict__ units, uint32_t* key, int* val) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= ub_nunits) return;
    // the sort runs over the bound: padding keys sort behind every unit
    key[u] = u < min(*d_nunits, cap) ? 63u - (uint32_t)min(63, (units[u].end - units[u].begin + 31) >> 5) : 64u;
    val[u] = u;
}

__global__ void rows_part_off_kernel(Grid G, int ng, const int* __restrict__ cell_sci,
                                     const int* __restrict__ row_unit_off, const int* __restrict__ d_nsci, int* part_off) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > G.R) return;
    const int nsci = *d_nsci;
    const int s = r == G.R ? nsci : min(cell_sci[r * G.ncell], nsci);
    part_off[r] = row_unit_off[(size_t)s * ng];
}

// End of a build: the build-time positions the refresh kernel measures displacements against, and the
// 32-byte gather records of the pair kernel.
__global__ void snapshot_kernel(int nslot_cap, const float4* __restrict__ posq, const float2* __restrict__ par,
                                float4* __restrict__ posq_build, float4* __restrict__ jrec) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslot_cap) return;
    const float4 p = posq[s];
    const float2 q = par[s];
    posq_build[s] = p;
    jrec[2 * (size_t)s] = p;
    jrec[2 * (size_t)s + 1] = make_float4(q.x, q.y, 0.f, 0.f);
}

__global__ void dummy_slot_kernel(int slot, float4* posq, float4* posq_build, float2* par, int* atom, int* img) {
    posq[slot] = posq_build[slot] = make_float4(nbl::kFar, nbl::kFar, nbl::kFar, 0.f);
    par[slot] = make_float2(0.f, 0.f);
    atom[slot] = -1;
    img[slot] = 512 | (512 << 10) | (512 << 20);
}

__global__ void minmax_kernel(int total, const double* __restrict__ pos, double* out) {
    // single block reduction; non-periodic systems only, at list-build time
    __shared__ double smin[3][256], smax[3][256];
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int k = threadIdx.x; k < total; k += blockDim.x)
        for (int d = 0; d < 3; d++) {
            const double v = pos[3 * (size_t)k + d];
            lo[d] = fmin(lo[d], v);
            hi[d] = fmax(hi[d], v);
        }
    for (int d = 0; d < 3; d++) { smin[d][threadIdx.x] = lo[d]; smax[d][threadIdx.x] = hi[d]; }
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s)
            for (int d = 0; d < 3; d++) {
                smin[d][threadIdx.x] = fmin(smin[d][threadIdx.x], smin[d][threadIdx.x + s]);
                smax[d][threadIdx.x] = fmax(smax[d][threadIdx.x], smax[d][threadIdx.x + s]);
            }
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int d = 0; d < 3; d++) { out[d] = smin[d][0]; out[3 + d] = smax[d][0]; }
}

inline int blocks(long long n, int t = 256) { return (int)((n + t - 1) / t); }

// Choose the cell grid.  Cells hold ~40 atoms (5 clusters) at the system's density and are at
// least rlist/2 wide so that the search stencil stays within kMaxSpan cells per dimension.
int setup_grid(sdm_ctx* c, PairList* pl, const double lo[3], const double ext[3]) {
    const bool periodic = c->T.method == SDM_CUTOFF_PERIODIC;
    if (!nbl::size_grid(pl->G, c->n, c->R, periodic, c->T.rc + c->opt.skin, lo, ext, pl->cell_cap,
                        pl->use_columns != 0))
        return sdm_fail(SDM_ERR_INVALID, "could not fit a cell grid (box too anisotropic for the cluster path)");
    return SDM_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// counts of a build: device words, read back once, asynchronously
// ---------------------------------------------------------------------------------------------
enum { kCntSlot = 0, kCntSci, kCntRaw, kCntEntries, kCntJent, kCntUnits, kCntOverflow, kCntUnitsLaunch, kCntN = 16 };

// Last kernel of a build: gathers the counts the stages left at the ends of their scans, compares them
// with the capacities and the launch bounds this build ran with, and -- if one of them was exceeded --
// raises SDM_ERR_CAPACITY for every replica and hides the (incomplete) units from the pair kernel.
// Also starts the list's age and displacement tracking (one evaluation with the new list is under way).
__global__ void finish_kernel(const int* d_nslot, const int* d_nsci, int ub_nsci, const int* d_nraw, const int* d_nent,
                              const int* d_njent, const int* d_nunits, int raw_cap, int ub_nraw, int entries_cap,
                              int jent_cap, int runits_cap, int ub_nunits, int R, int* flags, int* cnt,
                              int* list_age, unsigned int* max_disp2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int nraw = *d_nraw, nent = *d_nent, njent = *d_njent, nunits = *d_nunits;
    const bool over = *d_nsci > ub_nsci || nraw > raw_cap || nraw > ub_nraw || nent > entries_cap || njent > jent_cap ||
                      nunits > runits_cap || nunits > ub_nunits;
    cnt[kCntSlot] = *d_nslot;
    cnt[kCntSci] = *d_nsci;
    cnt[kCntRaw] = nraw;
    cnt[kCntEntries] = nent;
    cnt[kCntJent] = njent;
    cnt[kCntUnits] = nunits;
    cnt[kCntOverflow] = over ? 1 : 0;
    cnt[kCntUnitsLaunch] = over ? 0 : nunits;
    if (over)
        for (int r = 0; r < R; r++) flags[r] = SDM_ERR_CAPACITY;
    *list_age = 1;
    *max_disp2 = 0u;
}

// Launch bound for a count that was `prev` at the last build: a quarter more, and sticky -- a bound that
// still has room is kept, so that consecutive builds (and the evaluation graph) launch identical grids.
static int sticky_bound(int bound, int prev, int slack, long long cap) {
    if (bound > 0 && prev <= bound - bound / 12 && prev >= bound / 2) return bound;
    return (int)std::min<long long>(cap, (long long)prev + prev / 4 + slack);
}

// ---------------------------------------------------------------------------------------------
// per-atom j rows from the pruned entries and the System's exclusions (nblist_core.h stage 5)
// ---------------------------------------------------------------------------------------------
static int build_rows(sdm_ctx* c, bool sync, const int* d_nslot, const int* d_nsci) {
    PairList* pl = c->pl;
    cudaStream_t s = c->stream;
    const int G = pl->row_group, ng = nbl::kMaxCi / G;
    const int ub_nsci = pl->ub_nsci;
    const int nrows = ub_nsci * ng, nseg = 3 * nrows;
    cluster_info_kernel<<<blocks(pl->ncl_cap, 32), 256, 0, s>>>(d_nslot, c->n, pl->atom, pl->par, c->T.excl_start,
                                                               c->T.excl_idx, pl->slot_of, pl->cl_info, pl->cl_tiles);
    const RowsIn in{G, c->n, pl->entries, pl->entry_jhit, pl->sci, pl->sci_off, pl->cl_info, pl->cl_tiles,
                    c->T.excl_start, c->T.excl_idx, pl->slot_of, pl->atom};
    if (ub_nsci > 0) rows_kernel<false><<<ub_nsci, 32 * ng, 0, s>>>(in, d_nsci, pl->seg_total, nullptr, nullptr, nullptr, 0);
    PL_CUDA(cudaMemsetAsync(pl->seg_total + nseg, 0, sizeof(int), s));
    PL_CUDA(cub::DeviceScan::ExclusiveSum(pl->cub_tmp, pl->cub_tmp_bytes, pl->seg_total, pl->seg_off, nseg + 1, s));
    rows_units_count_kernel<<<blocks(nrows), 256, 0, s>>>(d_nsci, ub_nsci, ng, pl->row_chunk, pl->seg_off, pl->row_nunits);
    PL_CUDA(cudaMemsetAsync(pl->row_nunits + nrows, 0, sizeof(int), s));
    PL_CUDA(cub::DeviceScan::ExclusiveSum(pl->cub_tmp, pl->cub_tmp_bytes, pl->row_nunits, pl->row_unit_off, nrows + 1, s));
    const int* d_njent = pl->seg_off + nseg;
    const int* d_nunits = pl->row_unit_off + nrows;
    if (sync) {
        PL_CUDA(cudaMemcpyAsync(&pl->h_counts[kCntJent], d_njent, sizeof(int), cudaMemcpyDeviceToHost, s));
        PL_CUDA(cudaMemcpyAsync(&pl->h_counts[kCntUnits], d_nunits, sizeof(int), cudaMemcpyDeviceToHost, s));
        PL_CUDA(cudaStreamSynchronize(s));
        pl->njent = pl->h_counts[kCntJent];
        pl->nrunits = pl->h_counts[kCntUnits];
        if ((size_t)pl->njent > pl->jent_cap) {
            pl->jent_cap = (size_t)(pl->njent * 1.25) + 4096;
            if (int rc = pl_realloc(pl, &pl->jent, pl->jent_cap)) return rc;
            if (int rc = pl_realloc(pl, &pl->jallow, pl->jent_cap)) return rc;
        }
        pl->ub_nrunits = sticky_bound(pl->ub_nrunits, pl->nrunits, 256, 1ll << 30);
        if ((size_t)pl->ub_nrunits > pl->runits_cap) {
            pl->runits_cap = (size_t)pl->ub_nrunits;
            if (int rc = pl_realloc(pl, &pl->runits, pl->runits_cap)) return rc;
            if (int rc = pl_realloc(pl, &pl->epart, pl->runits_cap)) return rc;
            if (int rc = pl_realloc(pl, &pl->cpart, pl->runits_cap)) return rc;
            if (int rc = pl_realloc(pl, &pl->ru_key, pl->runits_cap)) return rc;
            if (int rc = pl_realloc(pl, &pl->ru_key_sorted, pl->runits_cap)) return rc;
            if (int rc = pl_realloc(pl, &pl->ru_val, pl->runits_cap)) return rc;
            if (int rc = pl_realloc(pl, &pl->ru_order, pl->runits_cap)) return rc;
        }
    }
    if (!sync) pl->ub_nrunits = sticky_bound(pl->ub_nrunits, pl->nrunits, 256, (long long)pl->runits_cap);   // previous build's count
    const int ub_nunits = (int)std::min<size_t>(pl->ub_nrunits, pl->runits_cap);
    if (ub_nsci > 0)
        rows_kernel<true><<<ub_nsci, 32 * ng, 0, s>>>(in, d_nsci, nullptr, pl->seg_off, pl->jent, pl->jallow, (int)pl->jent_cap);
    rows_units_fill_kernel<<<blocks(nrows), 256, 0, s>>>(d_nsci, ng, G, pl->row_chunk, pl->sci, pl->seg_off,
                                                        pl->row_unit_off, pl->runits, (int)pl->runits_cap);
    rows_part_off_kernel<<<blocks(c->R + 1), 256, 0, s>>>(pl->G, ng, pl->cell_sci, pl->row_unit_off, d_nsci, pl->part_off);
    if (pl->row_lpt && ub_nunits > 0) {
        rows_unit_key_kernel<<<blocks(ub_nunits), 256, 0, s>>>(d_nunits, ub_nunits, (int)pl->runits_cap, pl->runits,
                                                              pl->ru_key, pl->ru_val);
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, pl->ru_key, pl->ru_key_sorted, pl->ru_val, pl->ru_order, ub_nunits, 0, 7, s);
        if (need > pl->cub_tmp_bytes) {
            if (!sync) return sdm_fail(SDM_ERR_CUDA, "internal: sort scratch too small for an asynchronous build");
            if (int rc = pl_realloc(pl, (char**)&pl->cub_tmp, need)) return rc;
            pl->cub_tmp_bytes = need;
        }
        PL_CUDA(cub::DeviceRadixSort::SortPairs(pl->cub_tmp, pl->cub_tmp_bytes, pl->ru_key, pl->ru_key_sorted, pl->ru_val,
                                                pl->ru_order, ub_nunits, 0, 7, s));
        c->launches += 2;
    }
    c->launches += 7;
    return SDM_OK;
}

// ---------------------------------------------------------------------------------------------
// build
// ---------------------------------------------------------------------------------------------
// Two ways through the same kernels.  sync: the host reads every count as soon as it exists and grows
// the buffers that are too small (first build, non-periodic systems, after an overflow).  Otherwise the
// whole build is enqueued without a single host synchronisation: the kernels read their counts from
// device memory, the grids are sized with bounds derived from the previous build, finish_kernel checks
// the bounds and capacities on the device (SDM_ERR_CAPACITY for every replica -> the caller repeats the
// evaluation, which then builds synchronously), and the counts come back with one asynchronous copy
// that the NEXT build looks at.
static int read_back_counts(sdm_ctx* c) {
    PairList* pl = c->pl;
    if (!pl->counts_pending) return SDM_OK;
    PL_CUDA(cudaEventSynchronize(pl->ev_counts));
    pl->counts_pending = false;
    const int* h = pl->h_counts;
    pl->nslot = h[kCntSlot];
    pl->ncl = pl->nslot / nbl::kClusterSize;
    pl->nsci = h[kCntSci];
    pl->nraw = h[kCntRaw];
    pl->nentries = h[kCntEntries];
    pl->njent = h[kCntJent];
    pl->nrunits = h[kCntUnits];
    pl->have_counts = h[kCntOverflow] == 0;
    if (h[kCntOverflow]) pl->force_sync = true;
    return SDM_OK;
}

static int build_list(sdm_ctx* c) {
    PairList* pl = c->pl;
    cudaStream_t s = c->stream;
    const int n = c->n, R = c->R, total = n * R;

    if (int rc = read_back_counts(c)) return rc;
    const bool sync = !pl->async_builds || pl->force_sync || !pl->have_counts || c->T.method != SDM_CUTOFF_PERIODIC;
    pl->force_sync = false;
    (sync ? pl->n_sync : pl->n_async)++;
    // what the captured evaluation sequence depends on
    const void* const view0[] = {pl->jent, pl->jallow, pl->runits, pl->ru_order, pl->epart, pl->cpart};
    const int ub_units0 = pl->ub_nrunits;

    if (c->T.method != SDM_CUTOFF_PERIODIC) {
        minmax_kernel<<<1, 256, 0, s>>>(total, c->d_pos, pl->minmax);
        PL_CUDA(cudaMemcpyAsync(pl->h_minmax, pl->minmax, 6 * sizeof(double), cudaMemcpyDeviceToHost, s));
        PL_CUDA(cudaStreamSynchronize(s));
        double lo[3], ext[3];
        for (int d = 0; d < 3; d++) {
            lo[d] = pl->h_minmax[d] - 1e-3;
            ext[d] = std::max(pl->h_minmax[3 + d] - pl->h_minmax[d] + 2e-3, 1e-2);
        }
        if (int rc = setup_grid(c, pl, lo, ext)) return rc;
        c->launches += 1;
    }
    const Grid& G = pl->G;
    pl->ncells = R * G.ncell;
    const int ncells = pl->ncells;
    const int noff = G.span * G.span * (G.columns ? G.kz : G.span);

    // sort by (cell, z); the cell extents found here stay valid for the two refinement rounds
    int cell_bits = 1;
    while ((1ll << cell_bits) < ncells) cell_bits++;
    const int key_end = std::min(nbl::kSubBits + cell_bits + 1, 64);
    key_kernel<<<blocks(total), 256, 0, s>>>(G, c->d_pos, pl->keys, pl->vals);
    PL_CUDA(cub::DeviceRadixSort::SortPairs(pl->cub_tmp, pl->cub_tmp_bytes, pl->keys, pl->keys_sorted,
                                            pl->vals, pl->vals_sorted, total, 0, key_end, s));
    if (G.columns) {
        // the first sort was by (column, z): cut every column into chunk cells of 64 atoms and
        // sort again by (cell, z rank)
        PL_CUDA(cudaMemsetAsync(pl->cell_count, 0, sizeof(int) * (size_t)(ncells + 1), s));
        PL_CUDA(cudaMemsetAsync(pl->cell_first, 0, sizeof(int) * (size_t)(ncells + 1), s));
        cell_bounds_kernel<<<blocks(total), 256, 0, s>>>(total, pl->keys_sorted, pl->cell_first, pl->cell_count);
        chunk_key_kernel<<<blocks(total), 256, 0, s>>>(G, total, pl->keys_sorted, pl->vals_sorted, pl->cell_first,
                                                      pl->keys, pl->vals);
        PL_CUDA(cub::DeviceRadixSort::SortPairs(pl->cub_tmp, pl->cub_tmp_bytes, pl->keys, pl->keys_sorted,
                                                pl->vals, pl->vals_sorted, total, 0, key_end, s));
        c->launches += 3;
    }
    PL_CUDA(cudaMemsetAsync(pl->cell_count, 0, sizeof(int) * (size_t)(ncells + 1), s));
    PL_CUDA(cudaMemsetAsync(pl->cell_first, 0, sizeof(int) * (size_t)(ncells + 1), s));
    cell_bounds_kernel<<<blocks(total), 256, 0, s>>>(total, pl->keys_sorted, pl->cell_first, pl->cell_count);
    cell_sizes_kernel<<<blocks(ncells), 256, 0, s>>>(ncells, pl->cell_first, pl->cell_count,
                                                    pl->cell_pcount, pl->cell_nsci);
    // kd refinement inside every cell: by y within the z halves, by x within the y halves
    const int* vals_final = pl->vals_sorted;
    if (pl->kd_in_block) {
        kd_refine_kernel<<<ncells, 64, 0, s>>>(G, c->d_pos, pl->vals_sorted, pl->cell_first, pl->cell_count, pl->keys, pl->vals);
        vals_final = pl->vals;   // the atoms only moved inside their cells: keys_sorted still names the cell
        c->launches += 1;
    } else {
        for (int level = 1; level <= 2; level++) {
            refine_key_kernel<<<blocks(total), 256, 0, s>>>(G, level, total, c->d_pos, pl->keys_sorted, pl->vals_sorted,
                                                           pl->cell_first, pl->cell_count, pl->keys, pl->vals);
            PL_CUDA(cub::DeviceRadixSort::SortPairs(pl->cub_tmp, pl->cub_tmp_bytes, pl->keys, pl->keys_sorted,
                                                    pl->vals, pl->vals_sorted, total, 0, key_end, s));
        }
        c->launches += 4;
    }
    // exclusive scans over ncells+1 elements (the extra zero element yields the totals)
    PL_CUDA(cudaMemsetAsync(pl->cell_pcount + ncells, 0, sizeof(int), s));
    PL_CUDA(cudaMemsetAsync(pl->cell_nsci + ncells, 0, sizeof(int), s));
    PL_CUDA(cub::DeviceScan::ExclusiveSum(pl->cub_tmp, pl->cub_tmp_bytes, pl->cell_pcount, pl->cell_slot, ncells + 1, s));
    PL_CUDA(cub::DeviceScan::ExclusiveSum(pl->cub_tmp, pl->cub_tmp_bytes, pl->cell_nsci, pl->cell_sci, ncells + 1, s));
    // slots and superclusters cannot exceed their capacities (n atoms + < 8 padding slots per cell)
    const int* d_nslot = pl->cell_slot + ncells;
    const int* d_nsci = pl->cell_sci + ncells;
    if (sync) {
        PL_CUDA(cudaMemcpyAsync(&pl->h_counts[kCntSlot], d_nslot, sizeof(int), cudaMemcpyDeviceToHost, s));
        PL_CUDA(cudaMemcpyAsync(&pl->h_counts[kCntSci], d_nsci, sizeof(int), cudaMemcpyDeviceToHost, s));
        PL_CUDA(cudaStreamSynchronize(s));
        pl->nslot = pl->h_counts[kCntSlot];
        pl->nsci = pl->h_counts[kCntSci];
        pl->ncl = pl->nslot / nbl::kClusterSize;
        if (pl->nslot > pl->nslot_cap || pl->nsci > pl->nsci_cap)
            return sdm_fail(SDM_ERR_CAPACITY, "internal: slot capacity exceeded");
    }
    pl->ub_nsci = sticky_bound(pl->ub_nsci, pl->nsci, 16, pl->nsci_cap);   // asynchronous build: nsci of the previous one
    // a replica holds n atoms and at most 7 padding slots per cell
    pl->scan_max = std::min(pl->nslot_cap, n + 7 * G.ncell);

    fill_slots_kernel<<<blocks(total), 256, 0, s>>>(G, c->T, total, c->d_pos, pl->keys_sorted, vals_final,
                                                   pl->cell_first, pl->cell_slot, pl->posq, pl->par,
                                                   pl->atom, pl->img, pl->slot_of);
    fill_dummies_kernel<<<blocks(ncells), 256, 0, s>>>(ncells, pl->cell_count, pl->cell_slot, pl->posq,
                                                      pl->par, pl->atom, pl->img);
    bbox_kernel<<<blocks(pl->ncl_cap), 256, 0, s>>>(d_nslot, pl->posq, pl->cl_box);
    sci_kernel<<<blocks(ncells), 256, 0, s>>>(G, ncells, pl->cell_slot, pl->cell_sci, pl->cl_box, pl->sci,
                                             pl->sci_box, pl->cl_sci);
    if (G.columns) cell_box_kernel<<<blocks(ncells), 256, 0, s>>>(ncells, pl->cell_slot, pl->cl_box, pl->cell_box);
    c->launches += 10;

    nbl::SearchView V;
    V.G = G;
    V.sci = pl->sci;
    V.sci_box = pl->sci_box;
    V.cl_box = pl->cl_box;
    V.cell_slot = pl->cell_slot;
    V.cell_box = pl->cell_box;
    V.posq4 = reinterpret_cast<const float*>(pl->posq);
    const long long nitems = (long long)pl->ub_nsci * noff;
    if ((size_t)nitems + 1 > pl->items_cap) {
        if (!sync) return sdm_fail(SDM_ERR_CUDA, "internal: item capacity changed under an asynchronous build");
        pl->items_cap = (size_t)nitems + 1;
        if (int rc = pl_realloc(pl, &pl->item_count, pl->items_cap)) return rc;
        if (int rc = pl_realloc(pl, &pl->item_off, pl->items_cap)) return rc;
    }
    search_count_kernel<<<blocks(nitems, 128), 128, 0, s>>>(V, d_nsci, (int)nitems, noff, pl->item_count);
    PL_CUDA(cudaMemsetAsync(pl->item_count + nitems, 0, sizeof(int), s));
    auto ensure_cub = [&](size_t count) -> int {
        size_t need = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, need, pl->item_count, pl->item_off, (int)count, s);
        if (need > pl->cub_tmp_bytes) {
            if (!sync) return sdm_fail(SDM_ERR_CUDA, "internal: scan scratch too small for an asynchronous build");
            if (int rc = pl_realloc(pl, (char**)&pl->cub_tmp, need)) return rc;
            pl->cub_tmp_bytes = need;
        }
        return SDM_OK;
    };
    if (int rc = ensure_cub((size_t)nitems + 1)) return rc;
    PL_CUDA(cub::DeviceScan::ExclusiveSum(pl->cub_tmp, pl->cub_tmp_bytes, pl->item_count, pl->item_off, (int)nitems + 1, s));
    const int* d_nraw = pl->item_off + nitems;
    if (sync) {
        PL_CUDA(cudaMemcpyAsync(&pl->h_counts[kCntRaw], d_nraw, sizeof(int), cudaMemcpyDeviceToHost, s));
        PL_CUDA(cudaStreamSynchronize(s));
        pl->nraw = pl->h_counts[kCntRaw];
    }
    pl->ub_nraw = sticky_bound(pl->ub_nraw, pl->nraw, 1024, 1ll << 30);
    if (sync && (size_t)pl->ub_nraw + 1 > pl->raw_cap) {
        // the pruned entries are a subset of the raw ones: one capacity for both
        pl->raw_cap = (size_t)pl->ub_nraw + 1;
        if (int rc = pl_realloc(pl, &pl->raw_entries, pl->raw_cap)) return rc;
        if (int rc = pl_realloc(pl, &pl->raw_sci, pl->raw_cap)) return rc;
        if (int rc = pl_realloc(pl, &pl->raw_c0nci, pl->raw_cap)) return rc;
        if (int rc = pl_realloc(pl, &pl->raw_jhit, pl->raw_cap)) return rc;
        if (int rc = pl_realloc(pl, &pl->raw_keep, pl->raw_cap + 1)) return rc;
        if (int rc = pl_realloc(pl, &pl->raw_pos, pl->raw_cap + 1)) return rc;
        pl->entries_cap = pl->raw_cap;
        if (int rc = pl_realloc(pl, &pl->entries, pl->entries_cap)) return rc;
        if (int rc = pl_realloc(pl, &pl->entry_sci, pl->entries_cap + 1)) return rc;
        if (int rc = pl_realloc(pl, &pl->entry_jhit, pl->entries_cap + 1)) return rc;
    }
    const int ub_nraw = (int)std::min<size_t>(pl->ub_nraw, pl->raw_cap - 1);
    search_fill_kernel<<<blocks(nitems, 128), 128, 0, s>>>(V, d_nsci, noff, pl->item_off, pl->raw_entries,
                                                          pl->raw_sci, pl->raw_c0nci, (int)pl->raw_cap);
    // exact prune (one warp per raw entry), then order-preserving compaction
    prune_kernel<<<blocks(((long long)ub_nraw + 1) * 32, 128), 128, 0, s>>>(G, d_nraw, ub_nraw, (int)pl->raw_cap,
                                                                          pl->raw_entries, pl->raw_c0nci, pl->posq,
                                                                          pl->raw_keep, pl->raw_jhit);
    if (int rc = ensure_cub((size_t)ub_nraw + 1)) return rc;
    PL_CUDA(cub::DeviceScan::ExclusiveSum(pl->cub_tmp, pl->cub_tmp_bytes, pl->raw_keep, pl->raw_pos, ub_nraw + 1, s));
    const int* d_nent = pl->raw_pos + ub_nraw;
    compact_kernel<<<blocks(ub_nraw), 256, 0, s>>>(d_nraw, (int)pl->raw_cap, pl->raw_entries, pl->raw_sci, pl->raw_keep,
                                                  pl->raw_pos, pl->entries, pl->entry_sci, (int)pl->entries_cap,
                                                  pl->raw_jhit, pl->entry_jhit);
    sci_off_kernel<<<blocks(pl->ub_nsci + 1), 256, 0, s>>>(d_nsci, noff, d_nraw, ub_nraw, pl->item_off, pl->raw_pos, pl->sci_off);
    c->launches += 7;

    // the rows the pair kernel walks: individual j-atoms per i-group, exclusions and triangle as
    // allow words, Lennard-Jones-free j-atoms last
    if (int rc = build_rows(c, sync, d_nslot, d_nsci)) return rc;
    snapshot_kernel<<<blocks(pl->nslot_cap), 256, 0, s>>>(pl->nslot_cap, pl->posq, pl->par, pl->posq_build, pl->jrec);
    // the fixed-point accumulators are indexed by slot: start from zero for the new layout
    PL_CUDA(cudaMemsetAsync(c->B.f1acc, 0, sizeof(long long) * 3 * (size_t)pl->nslot_cap, s));
    const int G_ = pl->row_group, ng = nbl::kMaxCi / G_;
    finish_kernel<<<1, 32, 0, s>>>(d_nslot, d_nsci, pl->ub_nsci, d_nraw, d_nent, pl->seg_off + 3 * pl->ub_nsci * ng,
                                  pl->row_unit_off + pl->ub_nsci * ng, (int)pl->raw_cap, ub_nraw, (int)pl->entries_cap,
                                  (int)std::min<size_t>(pl->jent_cap, 0x7fffffff), (int)pl->runits_cap,
                                  (int)std::min<size_t>(pl->ub_nrunits, pl->runits_cap), R, c->B.flags, pl->d_cnt,
                                  c->d_list_age, pl->max_disp2);
    PL_CUDA(cudaMemcpyAsync(pl->h_counts, pl->d_cnt, kCntN * sizeof(int), cudaMemcpyDeviceToHost, s));
    PL_CUDA(cudaEventRecord(pl->ev_counts, s));
    pl->counts_pending = true;
    c->launches += 8;
    PL_CUDA(cudaGetLastError());
    {   // buffers or the pair kernel's grid bound moved: the evaluation graph is captured again
        const void* const view1[] = {pl->jent, pl->jallow, pl->runits, pl->ru_order, pl->epart, pl->cpart};
        if (memcmp(view0, view1, sizeof(view0)) != 0 || ub_units0 != pl->ub_nrunits) c->graph_valid = false;
    }
    c->list_valid = true;
    c->list_age = 0;
    c->n_builds++;
    return SDM_OK;
}

}  // namespace sdm

// ---------------------------------------------------------------------------------------------
// ctx hooks
// ---------------------------------------------------------------------------------------------
using namespace sdm;

int sdm_ctx_init_pairlist(sdm_ctx* c) {
    if (c->pair_mode != SDM_PAIR_CLUSTER) return SDM_OK;
    if (c->T.method == SDM_NOCUTOFF)
        return sdm_fail(SDM_ERR_INVALID, "the cluster pair path needs a cutoff method; use SDM_PAIR_ALLPAIRS");
    const int n = c->n, R = c->R, total = n * R;
    PairList* pl = new PairList();
    c->pl = pl;
    // capacities
    const double rlist = c->T.rc + c->opt.skin;
    if (c->T.method == SDM_CUTOFF_PERIODIC) {
        for (int d = 0; d < 3; d++)
            if (c->T.box[d] < 2.0 * rlist)
                return sdm_fail(SDM_ERR_BOX, "periodic box smaller than 2*(cutoff+skin): use SDM_PAIR_ALLPAIRS or a smaller skin");
    }
    pl->cell_cap = std::max(64, n / 8 + 64);
    if (const char* e = getenv("SDMB200_LAYOUT")) pl->use_columns = std::string(e) != "cells";   // development knob
    if (c->T.method == SDM_CUTOFF_PERIODIC) {
        double lo[3] = {0, 0, 0};
        if (int rc = setup_grid(c, pl, lo, c->T.box)) return rc;
    } else {
        pl->G.ncell = pl->cell_cap;  // sized at build time from the positions
    }
    const int ncell_cap = c->T.method == SDM_CUTOFF_PERIODIC ? pl->G.ncell : pl->cell_cap;
    const int ncells_cap = R * ncell_cap;
    pl->nslot_cap = total + 7 * ncells_cap + 8;
    pl->nslot_cap = (pl->nslot_cap + 7) / 8 * 8;
    pl->ncl_cap = pl->nslot_cap / 8;
    pl->nsci_cap = pl->ncl_cap / 8 + ncells_cap + 1;

#define A(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)
    A(pl_alloc(pl, &pl->keys, total));
    A(pl_alloc(pl, &pl->keys_sorted, total));
    A(pl_alloc(pl, &pl->vals, total));
    A(pl_alloc(pl, &pl->vals_sorted, total));
    A(pl_alloc(pl, &pl->cell_first, ncells_cap + 1));
    A(pl_alloc(pl, &pl->cell_count, ncells_cap + 1));
    A(pl_alloc(pl, &pl->cell_pcount, ncells_cap + 1));
    A(pl_alloc(pl, &pl->cell_nsci, ncells_cap + 1));
    A(pl_alloc(pl, &pl->cell_slot, ncells_cap + 2));
    A(pl_alloc(pl, &pl->cell_sci, ncells_cap + 2));
    A(pl_alloc(pl, &pl->posq, pl->nslot_cap));
    A(pl_alloc(pl, &pl->posq_build, pl->nslot_cap));
    A(pl_alloc(pl, &pl->par, pl->nslot_cap));
    A(pl_alloc(pl, &pl->jrec, 2 * (size_t)pl->nslot_cap));
    A(pl_alloc(pl, &pl->atom, pl->nslot_cap));
    A(pl_alloc(pl, &pl->img, pl->nslot_cap));
    A(pl_alloc(pl, &pl->slot_of, total));
    A(pl_alloc(pl, &pl->cl_box, pl->ncl_cap));
    A(pl_alloc(pl, &pl->cl_info, pl->ncl_cap));
    A(pl_alloc(pl, &pl->cl_tiles, pl->ncl_cap));
    A(pl_alloc(pl, &pl->cell_box, ncells_cap + 1));
    A(pl_alloc(pl, &pl->sci_box, pl->nsci_cap));
    A(pl_alloc(pl, &pl->sci, pl->nsci_cap));
    A(pl_alloc(pl, &pl->cl_sci, pl->ncl_cap));
    A(pl_alloc(pl, &pl->part_off, R + 1));
    A(pl_alloc(pl, &pl->unit_counter, 1));
    A(pl_alloc(pl, &pl->max_disp2, 1));
    A(pl_alloc(pl, &pl->minmax, 6));
    pl->items_cap = (size_t)pl->nsci_cap * 64 + 1;
    A(pl_alloc(pl, &pl->item_count, pl->items_cap));
    A(pl_alloc(pl, &pl->item_off, pl->items_cap));
    // initial guesses; grown on demand at build time
    // the pruned entries are a subset of the raw ones: one capacity for both
    pl->entries_cap = (size_t)total * 12 + 1024;
    A(pl_alloc(pl, &pl->entries, pl->entries_cap));
    A(pl_alloc(pl, &pl->entry_sci, pl->entries_cap + 1));
    pl->raw_cap = pl->entries_cap;
    A(pl_alloc(pl, &pl->raw_entries, pl->raw_cap));
    A(pl_alloc(pl, &pl->raw_sci, pl->raw_cap));
    A(pl_alloc(pl, &pl->raw_c0nci, pl->raw_cap));
    A(pl_alloc(pl, &pl->raw_jhit, pl->raw_cap));
    A(pl_alloc(pl, &pl->entry_jhit, pl->entries_cap + 1));
    A(pl_alloc(pl, &pl->raw_keep, pl->raw_cap + 1));
    A(pl_alloc(pl, &pl->raw_pos, pl->raw_cap + 1));
    A(pl_alloc(pl, &pl->sci_off, pl->nsci_cap + 2));
    if (const char* e = getenv("SDMB200_ROW_GROUP")) pl->row_group = atoi(e) == 2 ? 2 : 1;            // development knob
    if (const char* e = getenv("SDMB200_ROW_CHUNK")) pl->row_chunk = std::min(nbl::kRowChunkSteps, std::max(1, atoi(e)));
    {
        const int ng = nbl::kMaxCi / pl->row_group;
        A(pl_alloc(pl, &pl->seg_total, (size_t)pl->nsci_cap * ng * 3 + 1));
        A(pl_alloc(pl, &pl->seg_off, (size_t)pl->nsci_cap * ng * 3 + 1));
        A(pl_alloc(pl, &pl->row_nunits, (size_t)pl->nsci_cap * ng + 1));
        A(pl_alloc(pl, &pl->row_unit_off, (size_t)pl->nsci_cap * ng + 1));
        // ~50 row entries per atom at 100 atoms/nm^3 and rlist 1.06 nm; grown on demand
        pl->jent_cap = (size_t)total * 64 + 4096;
        A(pl_alloc(pl, &pl->jent, pl->jent_cap));
        A(pl_alloc(pl, &pl->jallow, pl->jent_cap));
        pl->runits_cap = pl->jent_cap / 256 + (size_t)pl->nsci_cap * ng + 256;
        A(pl_alloc(pl, &pl->runits, pl->runits_cap));
        A(pl_alloc(pl, &pl->epart, pl->runits_cap));
        A(pl_alloc(pl, &pl->cpart, pl->runits_cap));
        A(pl_alloc(pl, &pl->ru_key, pl->runits_cap));
        A(pl_alloc(pl, &pl->ru_key_sorted, pl->runits_cap));
        A(pl_alloc(pl, &pl->ru_val, pl->runits_cap));
        A(pl_alloc(pl, &pl->ru_order, pl->runits_cap));
        if (const char* e = getenv("SDMB200_ROW_LPT")) pl->row_lpt = atoi(e) != 0;   // development knob
        pl->dummy_slot = pl->nslot_cap - 1;   // never a real slot: nslot <= nslot_cap - 8
        dummy_slot_kernel<<<1, 1, 0, c->stream>>>(pl->dummy_slot, pl->posq, pl->posq_build, pl->par, pl->atom, pl->img);
    }
    {
        size_t a = 0, b = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, a, pl->keys, pl->keys_sorted, pl->vals, pl->vals_sorted, total, 0, 64, c->stream);
        cub::DeviceScan::ExclusiveSum(nullptr, b, pl->item_count, pl->item_off, (int)std::max<size_t>(pl->items_cap, pl->raw_cap + 1), c->stream);
        pl->cub_tmp_bytes = std::max(a, b) + 256;
        A(pl_alloc(pl, (char**)&pl->cub_tmp, pl->cub_tmp_bytes));
    }
    PL_CUDA(cudaMallocHost((void**)&pl->h_counts, kCntN * sizeof(int)));
    memset(pl->h_counts, 0, kCntN * sizeof(int));
    A(pl_alloc(pl, &pl->d_cnt, kCntN));
    PL_CUDA(cudaEventCreateWithFlags(&pl->ev_counts, cudaEventDisableTiming));
    if (const char* e = getenv("SDMB200_KD_SORTS")) pl->kd_in_block = atoi(e) == 0;
    if (const char* e = getenv("SDMB200_ASYNC_BUILD")) pl->async_builds = atoi(e) != 0;   // development knob
    PL_CUDA(cudaMallocHost((void**)&pl->h_minmax, 6 * sizeof(double)));
    // the accumulators live in the global slot space for this path
    {
        long long* acc = nullptr;
        A(pl_alloc(pl, &acc, 3 * (size_t)pl->nslot_cap));
        c->B.f1acc = acc;
        c->B.acc_rstride = 0;
        c->B.nslot = pl->nslot_cap;
        c->B.slot_of = pl->slot_of;
    }
#undef A
    c->list_valid = false;
    return SDM_OK;
}

void sdm_ctx_free_pairlist(sdm_ctx* c) {
    if (!c->pl) return;
    for (void* p : c->pl->allocs) cudaFree(p);
    if (c->pl->h_counts) cudaFreeHost(c->pl->h_counts);
    if (c->pl->ev_counts) cudaEventDestroy(c->pl->ev_counts);
    if (c->pl->h_minmax) cudaFreeHost(c->pl->h_minmax);
    delete c->pl;
    c->pl = nullptr;
}

static PairListView make_view(const sdm_ctx* c) {
    const PairList* pl = c->pl;
    PairListView V;
    V.G = pl->G;
    V.posq = pl->posq;
    V.par = pl->par;
    V.jrec = pl->jrec;
    V.atom = pl->atom;
    V.nslot_cap = pl->nslot_cap;
    V.jent = pl->jent;
    V.jallow = pl->jallow;
    V.runits = pl->runits;
    V.runit_order = pl->row_lpt ? pl->ru_order : nullptr;
    V.nrunits = pl->d_cnt + kCntUnitsLaunch;
    V.nrunits_ub = (int)std::min<size_t>(pl->ub_nrunits, pl->runits_cap);
    V.row_group = pl->row_group;
    V.dummy_slot = pl->dummy_slot;
    return V;
}

bool sdm_ctx_pairlist_rebuild_due(const sdm_ctx* c) {
    return !c->list_valid || c->list_age >= c->opt.nstlist;
}

int sdm_ctx_pairlist_prepare(sdm_ctx* c) {
    PairList* pl = c->pl;
    if (!pl) return sdm_fail(SDM_ERR_INVALID, "cluster pair list not initialised");
    cudaStream_t s = c->stream;
    const bool rebuild = sdm_ctx_pairlist_rebuild_due(c);
    if (rebuild) {
        if (int rc = build_list(c)) return rc;
    } else {
        const float hs = 0.5f * (float)c->opt.skin;
        launch_refresh(c->T, pl->G, pl->d_cnt + kCntSlot, pl->nslot_cap, c->d_pos, pl->atom, pl->img, pl->posq_build, pl->posq, pl->jrec,
                       hs * hs, c->B.flags, c->d_list_age, pl->max_disp2,
                       // fresh state-1 accumulators: cleared slot by slot by this kernel (coalesced; cheaper than
                       // scattered stores in the mix kernel at 16 x 20 k atoms); small batches let the mix kernel
                       // clear what it reads
                       sdm_ctx_mix_clears_accumulators(c) ? nullptr : c->B.f1acc, s);
        c->launches++;
    }
    c->list_age++;
    // partial-sum buffers of this path
    c->B.epart = pl->epart;
    c->B.cpart = pl->cpart;
    c->B.part_off = pl->part_off;
    c->B.work_counter = pl->unit_counter;
    // the displaced-atom kernels scan the cell-sorted slots (spatial locality per warp)
    c->B.scan_posq = pl->posq;
    c->B.scan_atom = pl->atom;
    c->B.scan_off = pl->cell_slot;
    c->B.scan_stride = pl->G.ncell;
    c->B.scan_max = pl->scan_max;
    PL_CUDA(cudaGetLastError());
    return SDM_OK;
}

int sdm_ctx_pairlist_launch(sdm_ctx* c) {
    PairList* pl = c->pl;
    cudaStream_t s = c->stream;
    // ev[1]..ev[2] bracket the pair kernel alone (list build and refresh are outside)
    if (c->timing) PL_CUDA(cudaEventRecord(c->ev[1], s));
    // large batches: the displaced-atom rows run beside this kernel in small blocks; leave them the registers of
    // four resident blocks per SM: 20 x 2 560 + 3 x 4 608 registers fit one SM, so every pair block is resident from
    // the start and the space the row blocks free goes to the gather and exceptions kernels that follow them, not to
    // late pair blocks (with a smaller reserve those take it and the side chain ends up behind the pair kernel).
    // The same grid when the pair kernel is timed alone: what is measured is what runs.
    static const int reserve_blocks = getenv("SDMB200_PAIR_RESERVE") ? atoi(getenv("SDMB200_PAIR_RESERVE")) : 4;
    const int reserve = (sdm::ligand_rows_beside_pair_kernel(c->T, c->B, c->num_sms) && !c->timing_full_residency) ? reserve_blocks : 0;
    launch_pair_rows(c->T, make_view(c), c->d_pos, c->B.f1acc, pl->epart, pl->cpart, c->opt.exact_cutoff,
                     pl->unit_counter, c->num_sms, nullptr, reserve, s);
    if (c->timing) PL_CUDA(cudaEventRecord(c->ev[2], s));
    c->launches++;
    PL_CUDA(cudaGetLastError());
    return SDM_OK;
}

int sdm_ctx_pairlist_emit(sdm_ctx* c, int replica, int* d_counter, int* d_pairs, int cap) {
    PairList* pl = c->pl;
    if (!pl || !c->list_valid) return sdm_fail(SDM_ERR_INVALID, "no pair list built yet: call sdm_eval first");
    // the debug build of the hot kernel itself: same list, same code, plus the pair records.  It adds
    // its forces to the accumulators a second time; they are cleared before the next evaluation.
    const PairEmit em{d_counter, d_pairs, cap, replica};
    launch_pair_rows(c->T, make_view(c), c->d_pos, c->B.f1acc, pl->epart, pl->cpart, c->opt.exact_cutoff,
                     pl->unit_counter, c->num_sms, &em, 0, c->stream);
    PL_CUDA(cudaMemsetAsync(pl->unit_counter, 0, sizeof(int), c->stream));
    PL_CUDA(cudaMemsetAsync(c->B.f1acc, 0, sizeof(long long) * 3 * (size_t)pl->nslot_cap, c->stream));
    PL_CUDA(cudaStreamSynchronize(c->stream));
    return SDM_OK;
}

// After the host has seen SDM_ERR_CAPACITY: was it the last list build that ran out of room (its bounds
// or buffers)?  Then the next build sizes itself step by step on the host.
bool sdm_ctx_pairlist_overflowed(sdm_ctx* c) {
    if (!c->pl) return false;
    if (read_back_counts(c)) return false;
    return c->pl->force_sync;
}

unsigned int* sdm_ctx_pairlist_max_disp_ptr(sdm_ctx* c) { return c->pl ? c->pl->max_disp2 : nullptr; }

int sdm_ctx_pairlist_info(sdm_ctx* c, const char* key, double* value) {
    if (!c->pl) return SDM_ERR_INVALID;
    if (int rc = read_back_counts(c)) return rc;
    const PairList* pl = c->pl;
    std::string k(key);
    if (k == "n_slots") *value = pl->nslot;
    else if (k == "n_clusters") *value = pl->ncl;
    else if (k == "n_sci") *value = pl->nsci;
    else if (k == "n_entries") *value = pl->nentries;
    else if (k == "n_raw_entries") *value = pl->nraw;
    else if (k == "n_units") *value = pl->nrunits;
    else if (k == "n_row_entries") *value = pl->njent;
    else if (k == "row_group") *value = pl->row_group;
    else if (k == "n_cells") *value = pl->G.ncell;
    else if (k == "cell_span") *value = pl->G.span;
    else if (k == "layout_columns") *value = pl->G.columns;
    else if (k == "chunk_cells_per_column") *value = pl->G.kz;
    else if (k == "rlist") *value = pl->G.rlist;
    else if (k == "n_async_builds") *value = (double)pl->n_async;
    else if (k == "n_sync_builds") *value = (double)pl->n_sync;
    else return SDM_ERR_INVALID;
    return SDM_OK;
}
