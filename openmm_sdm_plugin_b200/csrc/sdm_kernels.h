// sdm_kernels.h -- host-callable launchers of the sm_100a kernels (C++ linkage, internal).
#pragma once

#include "sdm_internal.cuh"

namespace sdm {

// Device buffers of one context that the per-eval kernels touch.  R = replicas, n = atoms.
struct EvalBuffers {
    int R;
    const double* ext_f1;   // [R][3n] external state-1 forces (sdm_set_external_dual) or nullptr
    const double* ext_f2;   // [R][3n] external state-2 forces
    const double* ext_e;    // [R][2]  external (E1, E2); zero for a replica without the term
    const int* ext_on;      // [R]     1: the replica has external forces
    int* work_counter;      // cluster path: unit counter of the persistent pair kernel, set back to zero by the
                            // mix kernel (the last kernel of an evaluation) for the next one; nullptr otherwise
    const double* pos;      // [R][3n]   positions, System order, nm
    const double* fb;       // [R][3n]   bonded/restraint forces (zero when absent)
    float4* posq;           // [R][n]    (x,y,z wrapped into the box, q*sqrt(K)) float
    long long* f1acc;       // state-1 force accumulators, 2^32 fixed point: replica r, component c,
                            // slot s at f1acc[r*acc_rstride + c*nslot + s]
    size_t acc_rstride;     // 3*n (all-pairs, System order) or 0 (cluster path: global slot space)
    double* dF;             // [R][3n]   F2 - F1 (moved pairs only)
    double* F;              // [R][3n]   hybrid force (output)
    double* F1;             // [R][3n]   state-1 force in double (output)
    double* epart;          // pair-energy partials; replica r owns [part_off[r], part_off[r+1])
    long long* cpart;       // in-cutoff pair count partials, same ranges
    const int* part_off;    // [R+1] device
    double* eexc_part;      // [R][n_excpart] exception-energy partials
    double* uexc_part;      // [R][n_excpart] exception contribution to u
    double* upart;          // [R][n_lig] u partial per displaced atom
    long long* mcnt;        // [R][n_lig][2] moved-pair counts (x2) per displaced atom
    ReplicaState* state;    // [R]
    int* flags;             // [R] status raised by kernels of this eval (0 = ok); cleared by mix
    int* sticky;            // [R] first non-zero status since the host last read the scalars: pipelined
                            //     evaluations and multi-step MD calls cannot lose a stale-list report
    unsigned long long* md_ctl;  // device-resident MD: [0] raised on a stale list / capacity status,
                            //     [1] steps taken (kernels_md.cu); nullptr without sdm_md_init
    int n_excpart;
    int nslot;              // stride of f1acc planes (>= n)
    const int* slot_of;     // [R][n] atom -> accumulator slot, or nullptr for identity
    int n_epart_allpairs;   // blocks per replica of the all-pairs kernel
    // Scan order of the displaced-atom kernels: FP32 positions (min-image prefilter) in an order
    // with spatial locality.  Cluster path: the cell-sorted slots; all-pairs path: System order.
    const float4* scan_posq; // [scan index]
    const int* scan_atom;    // scan index -> replica*n + atom, -1 = padding; nullptr = identity
    const int* scan_off;     // replica r owns scan indices [scan_off[r*scan_stride],
    int scan_stride;         //   scan_off[(r+1)*scan_stride]); nullptr = [r*n, (r+1)*n)
    int scan_max;            // upper bound of a replica's scan length (grid sizing)
    uint32_t* hitbits;       // [R][n_lig][scan_words] prefilter hits: bit b of word w = scan index 32*w+b
    int scan_words;          // words per (replica, displaced atom) row = ceil(scan_max / 32)
    int* hitpre;             // [R][n_lig][scan_words] hits of the row before word w (probe kernel)
    double* pairf;           // [R][n_lig][pairf_cap][3] force on the resting atom of every hit
    int pairf_cap;           // hits a row can hold
    // Cluster path: the prefilter runs once per list build with the list's skin added to the
    // cutoff (filter_skin > 0) and its rows are expanded into candidate lists that stay valid
    // exactly as long as the pair list does.
    float filter_skin;       // nm added to the cutoff by the prefilter (0: per-eval prefilter)
    int* cand;               // [R][n_lig][pairf_cap] resting atom (System index) of hit h
    int* cand_count;         // [R][n_lig] hits of the row (may exceed pairf_cap: overflow)
    int* list_age;           // device: evals since the cluster-pair list was built (0: no list)
};

// ---- fused path, v0 (all-pairs tiles, System order) ------------------------------------------
void launch_prep_posq(const Topology& T, const EvalBuffers& B, cudaStream_t s);
// exact != 0: FP64 re-test in the band around the cutoff.  emit_*: optional pair dump.
void launch_allpairs(const Topology& T, const EvalBuffers& B, int exact, int* emit_counter,
                     int* emit_pairs, int emit_cap, int emit_replica, cudaStream_t s);
int allpairs_num_blocks(int n);

// ---- fused path, shared stages ----------------------------------------------------------------
void launch_ligand_probe(const Topology& T, const EvalBuffers& B, cudaStream_t s);
void launch_ligand_filter(const Topology& T, const EvalBuffers& B, cudaStream_t s);
void launch_ligand_gather(const Topology& T, const EvalBuffers& B, cudaStream_t s);
void launch_ligand_compact(const Topology& T, const EvalBuffers& B, cudaStream_t s);      // list build: bitmap -> candidates
void launch_ligand_probe_list(const Topology& T, const EvalBuffers& B, int num_sms, cudaStream_t s);   // per eval, from the candidates
// 1: the displaced-atom rows run in small blocks beside the pair kernel, which leaves them room (large batches)
int ligand_rows_beside_pair_kernel(const Topology& T, const EvalBuffers& B, int num_sms);
void launch_exceptions(const Topology& T, const EvalBuffers& B, cudaStream_t s);
int exceptions_num_blocks(int n_exceptions);
// e_scale / c_div: 0.5 / 2 when every pair was visited from both sides (all-pairs), 1 / 1 for a
// half list.
void launch_scalars(const Topology& T, const EvalBuffers& B, double e_scale, int c_div,
                    cudaStream_t s);
void launch_mix(const Topology& T, const EvalBuffers& B, int zero_acc, cudaStream_t s);

// ---- restraint forces of SDMUtils (kernels_restraints.cu) --------------------------------------
constexpr int kRestraintPoints = 8;   // points (centroids / single atoms) a term can name
struct RestraintTerm {
    int kind;                               // 0: addRestraintForce (centroid bond), 1: addAlignmentForce
    int npoints;                            // 2 or 8 (kind 0), 6 (kind 1)
    int grp_begin[kRestraintPoints + 1];    // atoms / weights of point k: [grp_begin[k], grp_begin[k+1])
    double p[16];                           // parameters, see kernels_restraints.cu
};
struct RestraintTables {
    int n_terms;
    const RestraintTerm* terms;             // device
    const int* atoms;                       // device: group members (System indices)
    const double* weights;                  // device: normalised weights (sum 1 per group)
    double control;                         // SDMRestraintControlParameter (scales kind 0)
};
struct ReplicaState;
void launch_restraints(const RestraintTables& RT, int n, int R, const double* pos_all, double* F_all,
                       ReplicaState* state, double* erest, cudaStream_t s);

// ---- literal kernel-interface operations (float4 device buffers) ------------------------------
void launch_make_state2(int n, float4* posq, const float4* displ, cudaStream_t s);
void launch_save_state1(int n, const float4* posq, const float4* force, float4* save_f,
                        float4* save_x, cudaStream_t s);
void launch_copy4(int n, const float4* src, float4* dst, cudaStream_t s);
void launch_hybrid_force(int n, const float4* f1, const float4* f2, float4* force, float sp,
                         cudaStream_t s);
void launch_langevin_part1(int n, float4* velm, const float4* force, float4* pos_delta, float vscale,
                           float fscale, float noisescale, float step_size, const float4* random,
                           unsigned random_index, cudaStream_t s);
void launch_langevin_part2(int n, float4* posq, const float4* pos_delta, float4* velm, float step_size,
                           cudaStream_t s);
// Distance constraints of the device-resident Langevin step (kernels_md.cu): device tables shared
// by all replicas.  Rigid three-site molecules (SETTLE) and small clusters (SHAKE).
struct MdConstraints {
    int n_settle, n_shake;
    const int* settle_atoms;         // [3*n_settle] apex, b, c
    const double* settle_par;        // [2*n_settle] d(apex,b) = d(apex,c), d(b,c)
    const int* shake_off;            // [n_shake+1] constraint range of every cluster
    const int* shake_ij;             // [2*ncons]
    const double* shake_d;           // [ncons]
    const int* shake_aoff;           // [n_shake+1] atom range of every cluster
    const int* shake_atoms;          // atoms of the clusters
    const unsigned char* in_cluster; // [n] 1: the atom is finished by the constraint kernel
    double tol;                      // relative tolerance (Integrator::getConstraintTolerance, 1e-5)
};
// One Langevin step of all replicas: part 1 + 2 per atom, then the constraint units.  ctl[0] != 0
// (raised by the scalar stage on SDM_ERR_STALE_LIST / SDM_ERR_CAPACITY) turns the step into a no-op;
// ctl[1] receives step + 1 when the step is really taken.
void launch_md_update(int n, int R, double* pos, double* vel, const double* force, const double* invm,
                      double vscale, double fscale, double noisescale, double dt, const double* noise,
                      unsigned long long seed, unsigned long long step, const MdConstraints* C,
                      double* xprime, unsigned long long* ctl, int* flags, cudaStream_t s);
void launch_widen(size_t count, const float* src, double* dst, cudaStream_t s);    // dst = (double)src
void launch_narrow(size_t count, const double* src, float* dst, cudaStream_t s);   // dst = (float)src
double measure_fp32_fma_tflops(int num_sms, cudaStream_t s);   // sustained packed-FMA rate (TFLOP/s), synchronises
void launch_kinetic_energy(int n, int R, const double* vel, const double* mass, double* ke, cudaStream_t s);

}  // namespace sdm
