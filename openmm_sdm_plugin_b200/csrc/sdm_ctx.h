// sdm_ctx.h -- the context object behind the opaque sdm_ctx handle (internal).
#pragma once

#include <string>
#include <vector>

#include "sdm_kernels.h"

namespace sdm { struct PairList; struct PmeState; struct GbState; }

struct sdm_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t side_stream = nullptr;   // displaced-atom kernels run here, next to the pair kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool own_stream = false;
    sdm_options opt{};
    int n = 0, R = 0;
    int pair_mode = SDM_PAIR_ALLPAIRS;

    sdm::Topology T{};
    sdm::EvalBuffers B{};
    std::vector<void*> allocs;          // every cudaMalloc of this ctx (freed in sdm_destroy)
    double* d_pos = nullptr;
    float* d_stage32 = nullptr;         // [R][n][3] staging of the single-precision transfer calls
    double* d_fb = nullptr;
    double* d_disp = nullptr;
    int* d_group = nullptr;
    int* d_lig_idx = nullptr;
    int* d_lig_flags = nullptr;
    uint32_t* d_hitbits = nullptr;      // displaced-atom prefilter bitmap (grown on demand)
    size_t hitbits_cap = 0;
    int* d_hitpre = nullptr;
    double* d_pairf = nullptr;
    size_t pairf_alloc = 0;             // doubles allocated for d_pairf
    int pairf_scale = 1;                // doubled whenever an eval reports SDM_ERR_CAPACITY
    int* d_cand = nullptr;              // cluster path: candidate resting atoms per displaced atom
    int* d_cand_count = nullptr;
    size_t cand_alloc = 0, cand_rows = 0;
    int64_t lig_built_for = -1;         // n_builds the candidates were built with

    std::vector<sdm_alch> h_alch;       // staging copies with ctx lifetime
    std::vector<double> h_eb;
    std::vector<int> h_group, h_lig, h_excl_start, h_excl_idx;
    sdm::ReplicaState* h_state = nullptr;  // pinned read-back buffer

    // cluster-pair list (pairlist.cu)
    sdm::PairList* pl = nullptr;
    bool list_valid = false;
    int list_age = 0;
    int64_t n_builds = 0;

    // CUDA graph of the per-eval kernel sequence (cluster path, between list rebuilds)
    cudaGraphExec_t graph_exec = nullptr;
    bool graph_valid = false;
    int graph_launches = 0;             // kernels inside the captured sequence
    int* d_list_age = nullptr;          // evals since the list was built, kept on the device

    // device-resident Langevin dynamics (sdm_md_*; no constraints)
    bool md_ready = false;
    double* d_vel = nullptr;            // [R][n][3]
    double* d_invm = nullptr;           // [n] 1/mass, 0 for massless particles
    double* d_mass = nullptr;           // [n]
    double* d_noise = nullptr;          // [R][n][3] explicit normals for the next update (test hook)
    double* d_ke = nullptr;             // [R] kinetic energies
    bool md_noise_pending = false;
    // distance constraints of the MD loop (sdm_md_set_constraints)
    std::vector<int> h_cons_pairs;
    std::vector<double> h_cons_dist;
    double cons_tol = 1e-5;
    sdm::MdConstraints mdc{};           // device tables
    double* d_xprime = nullptr;         // [R][n][3] unconstrained new positions of the clustered atoms
    unsigned long long* d_md_ctl = nullptr;   // [0] abort flag, [1] steps taken
    unsigned long long* h_md_ctl = nullptr;   // pinned read-back
    unsigned long long md_repeated = 0;
    int md_plan = 0;                    // planned list lifetime in steps (adapted to the observed motion)
    int* d_sticky = nullptr;            // [R] sticky status (see EvalBuffers::sticky)
    int* h_sticky = nullptr;            // pinned [R]
    double md_dt = 0, md_vscale = 0, md_fscale = 0, md_noisescale = 0;
    unsigned long long md_seed = 0, md_steps = 0;

    // reciprocal-space PME on the device (kernels_pme.cu; sdm_enable_reciprocal_pme): fills the external slots
    sdm::PmeState* pme = nullptr;
    std::vector<double> h_charge;       // host copy of the charges (self energy)
    double ewald_tol = 0;

    // HCT generalized Born + ACE on the device (kernels_gb.cu; sdm_enable_hct_gb): fills the external slots
    sdm::GbState* gb = nullptr;

    // external dual-state contributions (sdm_set_external_dual): reciprocal-space PME, GB, ...
    double *d_ext_f1 = nullptr, *d_ext_f2 = nullptr, *d_ext_e = nullptr;
    int* d_ext_on = nullptr;

    // restraint forces of SDMUtils (sdm_add_centroid_restraint / sdm_add_alignment_restraint)
    std::vector<sdm::RestraintTerm> h_rterms;
    std::vector<int> h_ratoms;
    std::vector<double> h_rweights;
    sdm::RestraintTables RT{};
    bool rt_dirty = false;              // host tables changed since the last upload
    double* d_erest = nullptr;          // [R] restraint energy of the last evaluation

    int64_t launches = 0;
    int64_t n_evals = 0;
    bool timing = false, timing_valid = false;
    bool timing_full_residency = false;  // sdm_set_timing(ctx, 2): the pair kernel timed with every resident block it can have
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

// api.cu
int sdm_fail(int code, const char* msg);

// kernels_pme.cu -- reciprocal-space PME
int sdm_ctx_init_pme(sdm_ctx* c, const int32_t* grid);   // grid: [3] or nullptr (OpenMM's rule)
void sdm_ctx_free_pme(sdm_ctx* c);
int sdm_ctx_pme_enqueue(sdm_ctx* c, cudaStream_t s);      // no-op without sdm_ctx_init_pme
int sdm_ctx_pme_info(sdm_ctx* c, const char* key, double* value);

// kernels_gb.cu -- HCT generalized Born + ACE surface area
int sdm_ctx_init_gb(sdm_ctx* c, const double* charge, const double* offset_radius, const double* scaled_radius,
                    double solute_dielectric, double solvent_dielectric, int sa_ace);
void sdm_ctx_free_gb(sdm_ctx* c);
int sdm_ctx_gb_enqueue(sdm_ctx* c, cudaStream_t s);       // no-op without sdm_ctx_init_gb
int sdm_ctx_gb_info(sdm_ctx* c, const char* key, double* value);
int sdm_ctx_gb_born_radii(sdm_ctx* c, int replica, int state, double* out, cudaStream_t s);

// pairlist.cu -- cluster-pair list path (SDM_PAIR_CLUSTER)
int sdm_ctx_init_pairlist(sdm_ctx* c);
void sdm_ctx_free_pairlist(sdm_ctx* c);
bool sdm_ctx_pairlist_rebuild_due(const sdm_ctx* c);
int sdm_ctx_pairlist_prepare(sdm_ctx* c);  // (re)build the list if due, else refresh the sorted positions
int sdm_ctx_pairlist_launch(sdm_ctx* c);   // the pair kernel
int sdm_ctx_pairlist_emit(sdm_ctx* c, int replica, int* d_counter, int* d_pairs, int cap);
int sdm_ctx_pairlist_info(sdm_ctx* c, const char* key, double* value);
inline bool sdm_ctx_mix_clears_accumulators(const sdm_ctx* c) { return (long long)c->R * c->n <= 65536; }
bool sdm_ctx_pairlist_overflowed(sdm_ctx* c);   // the last list build ran out of room (call after the host has seen SDM_ERR_CAPACITY)
unsigned int* sdm_ctx_pairlist_max_disp_ptr(sdm_ctx* c);   // device word: largest squared displacement since the build (float bits)
