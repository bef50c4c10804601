// api.cu -- the C ABI of libsdmb200.so (include/sdmb200.h).  Host orchestration only: owns the
// device buffers of a context, orders the sm_100a kernels on one stream, never computes a
// pair interaction or a force on the CPU.
#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "sdm_ctx.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define SDM_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            char buf_[512];                                                                   \
            snprintf(buf_, sizeof(buf_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                     __FILE__, __LINE__);                                                     \
            return fail(SDM_ERR_CUDA, buf_);                                                  \
        }                                                                                     \
    } while (0)

template <class T>
int dev_alloc(sdm_ctx* c, T** p, size_t count) {
    void* q = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    SDM_CUDA(cudaMalloc(&q, bytes));
    SDM_CUDA(cudaMemset(q, 0, bytes));
    c->allocs.push_back(q);
    *p = static_cast<T*>(q);
    return SDM_OK;
}

template <class T>
int dev_upload(sdm_ctx* c, const T** p, const std::vector<T>& v) {
    T* q = nullptr;
    int rc = dev_alloc(c, &q, v.size());
    if (rc) return rc;
    if (!v.empty()) SDM_CUDA(cudaMemcpy(q, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *p = q;
    return SDM_OK;
}

// OpenMM 7.3 NonbondedForceImpl::calcDispersionCorrection without a switching function
// (SURVEY.md Appendix B.4).  O(classes^2) host arithmetic done once at creation.
double dispersion_coefficient(const sdm_system* s) {
    if (s->method != SDM_CUTOFF_PERIODIC) return 0.0;
    std::map<std::pair<double, double>, int> classes;
    for (int i = 0; i < s->n_atoms; i++) classes[{s->sigma[i], s->epsilon[i]}]++;
    double sum1 = 0, sum2 = 0;
    for (auto& c : classes) {
        double sigma = c.first.first, eps = c.first.second, count = (double)c.second;
        count *= (count + 1) / 2;
        double s2 = sigma * sigma, s6 = s2 * s2 * s2;
        sum1 += count * eps * s6 * s6;
        sum2 += count * eps * s6;
    }
    for (auto a = classes.begin(); a != classes.end(); ++a)
        for (auto b = classes.begin(); b != a; ++b) {
            double sigma = 0.5 * (a->first.first + b->first.first);
            double eps = std::sqrt(a->first.second * b->first.second);
            double count = (double)a->second * (double)b->second;
            double s2 = sigma * sigma, s6 = s2 * s2 * s2;
            sum1 += count * eps * s6 * s6;
            sum2 += count * eps * s6;
        }
    double np = (double)s->n_atoms, ni = np * (np + 1) / 2;
    sum1 /= ni;
    sum2 /= ni;
    double rc = s->cutoff;
    return 8 * np * np * M_PI * (sum1 / (9 * std::pow(rc, 9)) - sum2 / (3 * std::pow(rc, 3)));
}

int upload_displacement(sdm_ctx* c, const double* displacement) {
    const int n = c->n;
    std::vector<double> disp(3 * (size_t)n, 0.0);
    if (displacement) std::copy(displacement, displacement + 3 * (size_t)n, disp.begin());
    std::map<std::array<double, 3>, int> ids;
    ids[{0.0, 0.0, 0.0}] = 0;
    std::vector<int> group(n, 0), lig;
    for (int i = 0; i < n; i++) {
        std::array<double, 3> d = {disp[3 * i], disp[3 * i + 1], disp[3 * i + 2]};
        // -0.0 and +0.0 displace identically
        for (auto& v : d) if (v == 0.0) v = 0.0;
        auto it = ids.find(d);
        int g;
        if (it == ids.end()) { g = (int)ids.size(); ids[d] = g; } else g = it->second;
        group[i] = g;
        if (g != 0) lig.push_back(i);
    }
    SDM_CUDA(cudaMemcpy(c->d_disp, disp.data(), disp.size() * sizeof(double), cudaMemcpyHostToDevice));
    SDM_CUDA(cudaMemcpy(c->d_group, group.data(), group.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (!lig.empty())
        SDM_CUDA(cudaMemcpy(c->d_lig_idx, lig.data(), lig.size() * sizeof(int), cudaMemcpyHostToDevice));
    c->T.n_lig = (int)lig.size();
    // displaced atoms that are excluded from (bonded to) a resting atom: only those need the
    // exclusion lookup in the displaced-vs-resting pair loops
    std::vector<int> flags(lig.size(), 0);
    for (size_t m = 0; m < lig.size(); m++)
        for (int k = c->h_excl_start[lig[m]]; k < c->h_excl_start[lig[m] + 1]; k++)
            if (group[c->h_excl_idx[k]] == 0) flags[m] |= 1;
    if (!lig.empty())
        SDM_CUDA(cudaMemcpy(c->d_lig_flags, flags.data(), flags.size() * sizeof(int), cudaMemcpyHostToDevice));
    c->h_group = group;
    c->h_lig = lig;
    return SDM_OK;
}

// Every entry point that takes a context runs on the context's device and leaves the calling
// thread's current device as it found it, so one host thread can drive contexts on several GPUs.
struct DeviceGuard {
    int prev = -1, dev = -1;
    explicit DeviceGuard(int device) : dev(device) {
        if (dev < 0) return;
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        if (dev >= 0 && prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define SDM_ON_CTX_DEVICE(c) DeviceGuard device_guard_((c) ? (c)->device : -1)

int check_ctx(sdm_ctx* c, int replica) {
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    if (replica < 0 || replica >= c->R) return fail(SDM_ERR_INVALID, "replica index out of range");
    return SDM_OK;
}

}  // namespace

int sdm_fail(int code, const char* msg) { return fail(code, msg); }

extern "C" {

int sdm_abi_version(void) { return SDM_ABI_VERSION; }

const char* sdm_last_error(void) { return g_last_error.c_str(); }

int sdm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void sdm_default_options(sdm_options* o) {
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->device = -1;
    o->pair_mode = SDM_PAIR_AUTO;
    o->skin = 0.06;
    o->nstlist = 20;
    o->exact_cutoff = 1;
    o->use_graph = 1;
}

void sdm_default_alch(sdm_alch* a) {
    if (!a) return;
    std::memset(a, 0, sizeof(*a));
    // LangevinIntegratorSDM ctor (openmmapi/src/LangevinIntegratorSDM.cpp:48-85)
    a->bias_method = SDM_BIAS_LINEAR;
    a->softcore_method = SDM_SOFTCORE_NONE;
    a->lambdac = 1.0;
    a->gammac = 0.0;
    a->wbcoeff = 1.0;
    a->w0coeff = 0.0;
    a->lambda1 = 1.0;
    a->lambda2 = 1.0;
    a->alpha = 1.0;
    a->u0 = 0.0;
    a->umax = 200.0;
    a->acore = 0.25;
    a->ubcore = 0.0;
    a->nonequilibrium = 0;
    a->noneq_tmax = 1.0;
    a->step_size = 0.001;
}

int sdm_create(const sdm_system* s_in, const sdm_options* opt_in, sdm_ctx** out) {
    if (!s_in || !out) return fail(SDM_ERR_INVALID, "null argument");
    *out = nullptr;
    // SDM_EWALD / SDM_PME: a periodic cutoff system whose Coulomb term is the direct-space Ewald one
    sdm_system sys_ = *s_in;
    const bool ewald = sys_.method == SDM_EWALD || sys_.method == SDM_PME;
    if (ewald) sys_.method = SDM_CUTOFF_PERIODIC;
    const sdm_system* s = &sys_;
    if (s->n_atoms <= 0) return fail(SDM_ERR_INVALID, "n_atoms must be positive");
    if (s->n_replicas < 1) return fail(SDM_ERR_INVALID, "n_replicas must be >= 1");
    if (!s->charge || !s->sigma || !s->epsilon) return fail(SDM_ERR_INVALID, "null parameter array");
    if (s->method < SDM_NOCUTOFF || s->method > SDM_CUTOFF_PERIODIC)
        return fail(SDM_ERR_INVALID, "unknown nonbonded method");
    if (s->method != SDM_NOCUTOFF && !(s->cutoff > 0)) return fail(SDM_ERR_INVALID, "cutoff must be positive");
    if (s->lj_combining != SDM_LJ_LORENTZ_BERTHELOT && s->lj_combining != SDM_LJ_GEOMETRIC)
        return fail(SDM_ERR_INVALID, "unknown Lennard-Jones combining rule");
    if (s->lj_combining == SDM_LJ_GEOMETRIC && s->use_dispersion_correction && s->method >= SDM_CUTOFF_PERIODIC)
        return fail(SDM_ERR_INVALID, "no dispersion correction with the geometric rule (the reference switches it off: desmonddmsfile75.py:428,438)");
    if (s->method == SDM_CUTOFF_PERIODIC)
        for (int d = 0; d < 3; d++)
            if (!(s->box[d] > 0) || 2 * s->cutoff > s->box[d])
                return fail(SDM_ERR_BOX, "The cutoff distance cannot be greater than half the periodic box size");
    if ((s->n_exclusions > 0 && !s->exclusions) || (s->n_exceptions > 0 && (!s->exceptions || !s->exception_params)))
        return fail(SDM_ERR_INVALID, "null exclusion/exception array");
    if (sdm_device_count() <= 0)
        return fail(SDM_ERR_NO_DEVICE, "no CUDA device available: libsdmb200 has no CPU fallback");

    sdm_options opt;
    sdm_default_options(&opt);
    if (opt_in) {
        opt = *opt_in;
        if (opt.skin < 0) opt.skin = 0.06;
        if (opt.nstlist <= 0) opt.nstlist = 20;
    }
    int dev = 0;
    SDM_CUDA(cudaGetDevice(&dev));
    if (opt.device >= 0) dev = opt.device;
    DeviceGuard device_guard_(dev);   // the caller's current device is restored on return
    cudaDeviceProp prop;
    SDM_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)   // the binary holds sm_100a code only: it loads on compute capability 10.x, nothing else
        return fail(SDM_ERR_NO_DEVICE, "libsdmb200 is built for sm_100a (B200, compute capability 10.x) only");

    sdm_ctx* c = new sdm_ctx();
    c->device = dev;
    c->opt = opt;
    c->n = s->n_atoms;
    c->RT.control = 1.0;   // SDMRestraintControlParameter (SDMUtils.py:97)
    c->R = s->n_replicas;
    c->num_sms = prop.multiProcessorCount;
    const int n = c->n, R = c->R;
    int rc = SDM_OK;
#define TRY(x) do { rc = (x); if (rc) { sdm_destroy(c); return rc; } } while (0)
#define TRYCUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::string m_ = std::string(#x) + ": " + cudaGetErrorString(e_); sdm_destroy(c); return fail(SDM_ERR_CUDA, m_); } } while (0)

    TRYCUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    for (int k = 0; k < 4; k++) TRYCUDA(cudaEventCreate(&c->ev[k]));
    {
        // the side stream gets the highest priority: its small kernels take every slot that frees up
        int lo_p = 0, hi_p = 0;
        cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
        TRYCUDA(cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, hi_p));
    }
    TRYCUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    TRYCUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));

    sdm::Topology& T = c->T;
    std::memset(&T, 0, sizeof(T));
    T.n = n;
    T.method = s->method;
    T.n_exceptions = s->n_exceptions;
    T.rc = s->method == SDM_NOCUTOFF ? 0.0 : s->cutoff;
    T.rc2 = T.rc * T.rc;
    if (s->method != SDM_NOCUTOFF) {
        T.krf = std::pow(T.rc, -3.0) * (s->eps_rf - 1.0) / (2.0 * s->eps_rf + 1.0);
        T.crf = (1.0 / T.rc) * (3.0 * s->eps_rf) / (2.0 * s->eps_rf + 1.0);
    }
    double cmax = 64.0;  // coordinate magnitude bound used for the FP64 re-test band
    for (int d = 0; d < 3; d++) {
        T.box[d] = s->method == SDM_CUTOFF_PERIODIC ? s->box[d] : 0.0;
        T.inv_box[d] = T.box[d] > 0 ? 1.0 / T.box[d] : 0.0;
        T.boxf[d] = (float)T.box[d];
        T.inv_boxf[d] = (float)T.inv_box[d];
    }
    if (s->method == SDM_CUTOFF_PERIODIC) cmax = std::max({T.box[0], T.box[1], T.box[2]});
    if (ewald) {
        // NonbondedForceImpl::calcPMEParameters / calcEwaldParameters: alpha = sqrt(-log(2 tol)) / cutoff
        const double tol = s->ewald_tolerance > 0 ? s->ewald_tolerance : 5e-4;
        T.ewald = 1;
        T.alpha = s->ewald_alpha > 0 ? s->ewald_alpha : std::sqrt(-std::log(2.0 * tol)) / T.rc;
        T.alphaf = (float)T.alpha;
        T.krf = T.crf = 0.0;
        c->ewald_tol = tol;
    }
    T.rc2f = (float)T.rc2;
    T.krff = (float)T.krf;
    T.crff = (float)T.crf;
    T.band = (float)(64.0 * FLT_EPSILON * T.rc * std::max(cmax, T.rc));
    if (s->method == SDM_CUTOFF_PERIODIC && s->use_dispersion_correction)
        T.e_disp = dispersion_coefficient(s) / (s->box[0] * s->box[1] * s->box[2]);
    T.lj_geom = s->lj_combining == SDM_LJ_GEOMETRIC ? 1 : 0;

    // per-atom parameters as the Reference kernel stores them: sigma/2, 2*sqrt(eps); with the geometric rule
    // sigma itself (sigma_ij^2 = sigma_i sigma_j)
    std::vector<double> q(n), hsig(n), heps(n);
    std::vector<float4> parf(n);
    const double sqrtK = std::sqrt(SDM_K_COULOMB);
    for (int i = 0; i < n; i++) {
        q[i] = s->charge[i];
        hsig[i] = T.lj_geom ? s->sigma[i] : 0.5 * s->sigma[i];
        heps[i] = 2.0 * std::sqrt(s->epsilon[i]);
        parf[i] = make_float4((float)(q[i] * sqrtK), (float)hsig[i], (float)heps[i], 0.f);
    }
    c->h_charge = q;
    TRY(dev_upload(c, &T.q, q));
    TRY(dev_upload(c, &T.hsig, hsig));
    TRY(dev_upload(c, &T.heps, heps));
    TRY(dev_upload(c, &T.parf, parf));

    // exclusions -> CSR over both directions, rows ascending, no self, no duplicates
    {
        std::vector<std::vector<int>> rows(n);
        for (int k = 0; k < s->n_exclusions; k++) {
            int a = s->exclusions[2 * k], b = s->exclusions[2 * k + 1];
            if (a < 0 || b < 0 || a >= n || b >= n) { sdm_destroy(c); return fail(SDM_ERR_INVALID, "exclusion index out of range"); }
            if (a == b) continue;
            rows[a].push_back(b);
            rows[b].push_back(a);
        }
        std::vector<int> start(n + 1, 0), idx;
        for (int i = 0; i < n; i++) {
            std::sort(rows[i].begin(), rows[i].end());
            rows[i].erase(std::unique(rows[i].begin(), rows[i].end()), rows[i].end());
            start[i + 1] = start[i] + (int)rows[i].size();
            idx.insert(idx.end(), rows[i].begin(), rows[i].end());
        }
        TRY(dev_upload(c, &T.excl_start, start));
        TRY(dev_upload(c, &T.excl_idx, idx));
        if (ewald) {   // the unique excluded pairs (i < j): each gets the erf(alpha r)/r correction once
            std::vector<int> pairs;
            for (int i = 0; i < n; i++)
                for (int j : rows[i])
                    if (j > i) { pairs.push_back(i); pairs.push_back(j); }
            T.n_excl_pairs = (int)(pairs.size() / 2);
            if (pairs.empty()) pairs.assign(2, 0);
            TRY(dev_upload(c, &T.excl_pairs, pairs));
        }
        c->h_excl_start = start;
        c->h_excl_idx = idx;
    }
    {
        std::vector<int> ep(s->exceptions, s->exceptions + 2 * (size_t)s->n_exceptions);
        for (int v : ep)
            if (v < 0 || v >= n) { sdm_destroy(c); return fail(SDM_ERR_INVALID, "exception index out of range"); }
        std::vector<double> pp(s->exception_params, s->exception_params + 3 * (size_t)s->n_exceptions);
        TRY(dev_upload(c, &T.exc_pairs, ep));
        TRY(dev_upload(c, &T.exc_params, pp));
    }
    TRY(dev_alloc(c, &c->d_disp, 3 * (size_t)n));
    TRY(dev_alloc(c, &c->d_group, (size_t)n));
    TRY(dev_alloc(c, &c->d_lig_idx, (size_t)n));
    TRY(dev_alloc(c, &c->d_lig_flags, (size_t)n));
    T.disp = c->d_disp;
    T.group = c->d_group;
    T.lig_idx = c->d_lig_idx;
    T.lig_flags = c->d_lig_flags;
    TRY(upload_displacement(c, s->displacement));

    // per-replica buffers
    sdm::EvalBuffers& B = c->B;
    std::memset(&B, 0, sizeof(B));
    B.R = R;
    B.n_epart_allpairs = sdm::allpairs_num_blocks(n);
    B.n_excpart = sdm::exceptions_num_blocks(s->n_exceptions + (T.ewald ? T.n_excl_pairs : 0));
    B.nslot = n;
    B.acc_rstride = 3 * (size_t)n;
    B.slot_of = nullptr;
    double *pos, *fb;
    TRY(dev_alloc(c, &pos, (size_t)R * 3 * n));
    TRY(dev_alloc(c, &fb, (size_t)R * 3 * n));
    B.pos = pos;
    B.fb = fb;
    c->d_pos = pos;
    c->d_fb = fb;
    TRY(dev_alloc(c, &B.posq, (size_t)R * n));
    B.scan_posq = B.posq;   // all-pairs path: System order (the cluster path installs its slots)
    B.scan_atom = nullptr;
    B.scan_off = nullptr;
    B.scan_stride = 0;
    B.scan_max = n;
    TRY(dev_alloc(c, &B.f1acc, (size_t)R * 3 * B.nslot));
    TRY(dev_alloc(c, &B.dF, (size_t)R * 3 * n));
    TRY(dev_alloc(c, &B.F, (size_t)R * 3 * n));
    TRY(dev_alloc(c, &B.F1, (size_t)R * 3 * n));
    TRY(dev_alloc(c, &B.epart, (size_t)R * B.n_epart_allpairs));
    TRY(dev_alloc(c, &B.cpart, (size_t)R * B.n_epart_allpairs));
    {
        std::vector<int> off(R + 1);
        for (int r = 0; r <= R; r++) off[r] = r * B.n_epart_allpairs;
        TRY(dev_upload(c, &B.part_off, off));
    }
    TRY(dev_alloc(c, &B.eexc_part, (size_t)R * B.n_excpart));
    TRY(dev_alloc(c, &B.uexc_part, (size_t)R * B.n_excpart));
    TRY(dev_alloc(c, &B.upart, (size_t)R * n));
    TRY(dev_alloc(c, &B.mcnt, (size_t)R * n * 2));
    TRY(dev_alloc(c, &B.state, (size_t)R));
    TRY(dev_alloc(c, &B.flags, (size_t)R));
    TRY(dev_alloc(c, &c->d_sticky, (size_t)R));
    B.sticky = c->d_sticky;
    B.md_ctl = nullptr;   // set by sdm_md_init
    TRY(dev_alloc(c, &c->d_list_age, 1));
    B.list_age = c->d_list_age;

    c->h_alch.resize(R);
    c->h_eb.assign(R, 0.0);
    for (int r = 0; r < R; r++) {
        sdm_default_alch(&c->h_alch[r]);
        TRYCUDA(cudaMemcpy(&B.state[r].alch, &c->h_alch[r], sizeof(sdm_alch), cudaMemcpyHostToDevice));
    }
    TRYCUDA(cudaMallocHost((void**)&c->h_state, sizeof(sdm::ReplicaState) * (size_t)R));

    c->pair_mode = opt.pair_mode;
    if (c->pair_mode == SDM_PAIR_AUTO) {
        // small systems and NoCutoff: all-pairs tiles; otherwise the cluster-pair list
        bool small_box = false;
        if (s->method == SDM_CUTOFF_PERIODIC)
            for (int d = 0; d < 3; d++) small_box |= s->box[d] < 2.0 * (s->cutoff + opt.skin);
        c->pair_mode = (s->method == SDM_NOCUTOFF || n < 3000 || small_box) ? SDM_PAIR_ALLPAIRS
                                                                            : SDM_PAIR_CLUSTER;
    }
    TRY(sdm_ctx_init_pairlist(c));
#undef TRY
#undef TRYCUDA
    *out = c;
    return SDM_OK;
}

void sdm_destroy(sdm_ctx* c) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return;
    if (c->stream) cudaStreamSynchronize(c->stream);
    sdm_ctx_free_pairlist(c);
    sdm_ctx_free_pme(c);
    sdm_ctx_free_gb(c);
    for (void* p : c->allocs) cudaFree(p);
    if (c->h_state) cudaFreeHost(c->h_state);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    for (int k = 0; k < 4; k++)
        if (c->ev[k]) cudaEventDestroy(c->ev[k]);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->side_stream) cudaStreamDestroy(c->side_stream);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int sdm_set_stream(sdm_ctx* c, void* cuda_stream) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    SDM_CUDA(cudaStreamSynchronize(c->stream));
    if (c->own_stream && c->stream) SDM_CUDA(cudaStreamDestroy(c->stream));
    c->own_stream = false;
    c->stream = static_cast<cudaStream_t>(cuda_stream);
    c->graph_valid = false;
    return SDM_OK;
}

int sdm_synchronize(sdm_ctx* c) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    SDM_CUDA(cudaStreamSynchronize(c->stream));
    return SDM_OK;
}

int sdm_host_alloc(void** ptr, uint64_t bytes) {
    if (!ptr) return fail(SDM_ERR_INVALID, "null argument");
    SDM_CUDA(cudaMallocHost(ptr, bytes ? bytes : 1));
    return SDM_OK;
}

int sdm_host_free(void* ptr) {
    if (ptr) SDM_CUDA(cudaFreeHost(ptr));
    return SDM_OK;
}

int sdm_set_positions(sdm_ctx* c, int replica, const double* xyz) {
    SDM_ON_CTX_DEVICE(c);
    if (int rc = check_ctx(c, replica)) return rc;
    if (!xyz) return fail(SDM_ERR_INVALID, "null positions");
    SDM_CUDA(cudaMemcpyAsync(c->d_pos + (size_t)replica * 3 * c->n, xyz, sizeof(double) * 3 * (size_t)c->n,
                             cudaMemcpyHostToDevice, c->stream));
    return SDM_OK;
}

int sdm_set_positions_device(sdm_ctx* c, int replica, const double* d_xyz) {
    SDM_ON_CTX_DEVICE(c);
    if (int rc = check_ctx(c, replica)) return rc;
    if (!d_xyz) return fail(SDM_ERR_INVALID, "null positions");
    SDM_CUDA(cudaMemcpyAsync(c->d_pos + (size_t)replica * 3 * c->n, d_xyz, sizeof(double) * 3 * (size_t)c->n,
                             cudaMemcpyDeviceToDevice, c->stream));
    return SDM_OK;
}

int sdm_set_positions_all(sdm_ctx* c, const double* xyz_all) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    if (!xyz_all) return fail(SDM_ERR_INVALID, "null positions");
    SDM_CUDA(cudaMemcpyAsync(c->d_pos, xyz_all, sizeof(double) * 3 * (size_t)c->n * c->R,
                             cudaMemcpyHostToDevice, c->stream));
    return SDM_OK;
}

static int ensure_stage32(sdm_ctx* c) {
    if (c->d_stage32) return SDM_OK;
    return dev_alloc(c, &c->d_stage32, 3 * (size_t)c->n * c->R);
}

int sdm_set_positions_all_f32(sdm_ctx* c, const float* xyz_all) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    if (!xyz_all) return fail(SDM_ERR_INVALID, "null positions");
    if (int rc = ensure_stage32(c)) return rc;
    const size_t count = 3 * (size_t)c->n * c->R;
    SDM_CUDA(cudaMemcpyAsync(c->d_stage32, xyz_all, sizeof(float) * count, cudaMemcpyHostToDevice, c->stream));
    sdm::launch_widen(count, c->d_stage32, c->d_pos, c->stream);
    c->launches++;
    SDM_CUDA(cudaGetLastError());
    return SDM_OK;
}

int sdm_positions_device_ptr(sdm_ctx* c, int replica, double** d_xyz) {
    SDM_ON_CTX_DEVICE(c);
    if (int rc = check_ctx(c, replica)) return rc;
    if (!d_xyz) return fail(SDM_ERR_INVALID, "null argument");
    *d_xyz = c->d_pos + (size_t)replica * 3 * c->n;
    return SDM_OK;
}

int sdm_set_bonded_forces(sdm_ctx* c, int replica, const double* fb, double eb) {
    SDM_ON_CTX_DEVICE(c);
    if (int rc = check_ctx(c, replica)) return rc;
    double* dst = c->d_fb + (size_t)replica * 3 * c->n;
    if (fb)
        SDM_CUDA(cudaMemcpyAsync(dst, fb, sizeof(double) * 3 * (size_t)c->n, cudaMemcpyHostToDevice, c->stream));
    else
        SDM_CUDA(cudaMemsetAsync(dst, 0, sizeof(double) * 3 * (size_t)c->n, c->stream));
    c->h_eb[replica] = eb;
    SDM_CUDA(cudaMemcpyAsync(&c->B.state[replica].sc.Eb, &c->h_eb[replica], sizeof(double),
                             cudaMemcpyHostToDevice, c->stream));
    return SDM_OK;
}

int sdm_set_alchemical(sdm_ctx* c, int replica, const sdm_alch* a) {
    SDM_ON_CTX_DEVICE(c);
    if (int rc = check_ctx(c, replica)) return rc;
    if (!a) return fail(SDM_ERR_INVALID, "null argument");
    c->h_alch[replica] = *a;
    SDM_CUDA(cudaMemcpyAsync(&c->B.state[replica].alch, &c->h_alch[replica], sizeof(sdm_alch),
                             cudaMemcpyHostToDevice, c->stream));
    return SDM_OK;
}

int sdm_get_alchemical(sdm_ctx* c, int replica, sdm_alch* a) {
    SDM_ON_CTX_DEVICE(c);
    if (int rc = check_ctx(c, replica)) return rc;
    if (!a) return fail(SDM_ERR_INVALID, "null argument");
    SDM_CUDA(cudaMemcpyAsync(&c->h_state[replica].alch, &c->B.state[replica].alch, sizeof(sdm_alch),
                             cudaMemcpyDeviceToHost, c->stream));
    SDM_CUDA(cudaStreamSynchronize(c->stream));
    *a = c->h_state[replica].alch;
    c->h_alch[replica] = *a;
    return SDM_OK;
}

int sdm_set_displacement(sdm_ctx* c, const double* displacement) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    SDM_CUDA(cudaStreamSynchronize(c->stream));
    if (int rc = upload_displacement(c, displacement)) return rc;
    c->graph_valid = false;   // the number of displaced atoms sizes two launches
    c->list_valid = false;    // the candidate lists of the displaced atoms are built with the pair list
    return SDM_OK;
}

int sdm_invalidate_list(sdm_ctx* c) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    c->list_valid = false;
    return SDM_OK;
}

// The kernels of one evaluation after the state-1 pair pass (shared by both pair modes).
// Everything that only needs the positions, not the state-1 pair pass: the displaced-atom pair
// terms (FP32 prefilter bitmap -> FP64 pair terms, once per pair -> per-atom gather) and the 1-4
// exceptions (fixed-point atomics on the state-1 accumulators commute with the pair kernel's).
// The cluster path runs this on the side stream next to the pair kernel.
static void enqueue_position_only(sdm_ctx* c, cudaStream_t s) {
    const sdm::Topology& T = c->T;
    sdm::EvalBuffers& B = c->B;
    // development knob (with SDMB200_SIDE_TIMING, outside graph capture): where the side chain spends its time
    static const bool stage_timing = getenv("SDMB200_SIDE_TIMING") != nullptr;
    static cudaEvent_t pev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (stage_timing) cudaStreamIsCapturing(s, &cap);
    const bool stages = stage_timing && cap == cudaStreamCaptureStatusNone;
    if (stages) {
        if (!pev[0]) for (auto& e : pev) cudaEventCreate(&e);
        else if (cudaEventSynchronize(pev[3]) == cudaSuccess) {
            float a = 0.f, b = 0.f, d = 0.f;
            cudaEventElapsedTime(&a, pev[0], pev[1]); cudaEventElapsedTime(&b, pev[1], pev[2]); cudaEventElapsedTime(&d, pev[2], pev[3]);
            fprintf(stderr, "  rows %.1f us  gather %.1f us  exceptions %.1f us\n", 1e3 * a, 1e3 * b, 1e3 * d);
        }
        cudaEventRecord(pev[0], s);
    }
    if (T.n_lig > 0 && c->pair_mode == SDM_PAIR_CLUSTER) {
        sdm::launch_ligand_probe_list(T, B, c->num_sms, s);   // candidates were laid down at the list build
        c->launches += 1;
    } else if (T.n_lig > 0) {
        sdm::launch_ligand_filter(T, B, s);
        sdm::launch_ligand_probe(T, B, s);
        c->launches += 2;
    }
    if (stages) cudaEventRecord(pev[1], s);
    sdm::launch_ligand_gather(T, B, s);   // per-atom gather of the displaced-atom pair forces
    if (stages) cudaEventRecord(pev[2], s);
    sdm::launch_exceptions(T, B, s);      // after the probe kernel: adds to dF of displaced 1-4 pairs
    if (stages) cudaEventRecord(pev[3], s);
    c->launches += 2;
    sdm_ctx_pme_enqueue(c, s);            // reciprocal-space PME of both states (when switched on)
    sdm_ctx_gb_enqueue(c, s);             // HCT-GB + ACE of both states (when switched on)
}

// The kernels that need both: scalar stage (soft-core, bias, bookkeeping) and the hybrid force.
static int enqueue_tail(sdm_ctx* c, double e_scale, int c_div, int zero_acc) {
    cudaStream_t s = c->stream;
    sdm::launch_scalars(c->T, c->B, e_scale, c_div, s);
    sdm::launch_mix(c->T, c->B, zero_acc, s);
    c->launches += 2;
    if (c->RT.n_terms > 0) {   // SDMUtils restraints: energy into pot_energy, forces onto the hybrid force
        sdm::launch_restraints(c->RT, c->n, c->R, c->d_pos, c->B.F, c->B.state, c->d_erest, s);
        c->launches += 1;
    }
    return SDM_OK;
}

// Device copy of the restraint tables (rebuilt whenever a term was added or removed; never inside a capture).
static int upload_restraints(sdm_ctx* c) {
    if (!c->rt_dirty) return SDM_OK;
    SDM_CUDA(cudaStreamSynchronize(c->stream));
    auto drop = [&](const void* p) {
        if (!p) return;
        c->allocs.erase(std::remove(c->allocs.begin(), c->allocs.end(), (void*)p), c->allocs.end());
        cudaFree((void*)p);
    };
    drop(c->RT.terms); drop(c->RT.atoms); drop(c->RT.weights);
    c->RT.terms = nullptr; c->RT.atoms = nullptr; c->RT.weights = nullptr;
    c->RT.n_terms = (int)c->h_rterms.size();
    if (c->RT.n_terms > 0) {
        void *t = nullptr, *a = nullptr, *w = nullptr;
        SDM_CUDA(cudaMalloc(&t, sizeof(sdm::RestraintTerm) * c->h_rterms.size()));
        SDM_CUDA(cudaMalloc(&a, sizeof(int) * std::max<size_t>(c->h_ratoms.size(), 1)));
        SDM_CUDA(cudaMalloc(&w, sizeof(double) * std::max<size_t>(c->h_rweights.size(), 1)));
        c->allocs.push_back(t); c->allocs.push_back(a); c->allocs.push_back(w);
        SDM_CUDA(cudaMemcpy(t, c->h_rterms.data(), sizeof(sdm::RestraintTerm) * c->h_rterms.size(), cudaMemcpyHostToDevice));
        SDM_CUDA(cudaMemcpy(a, c->h_ratoms.data(), sizeof(int) * c->h_ratoms.size(), cudaMemcpyHostToDevice));
        SDM_CUDA(cudaMemcpy(w, c->h_rweights.data(), sizeof(double) * c->h_rweights.size(), cudaMemcpyHostToDevice));
        c->RT.terms = (const sdm::RestraintTerm*)t;
        c->RT.atoms = (const int*)a;
        c->RT.weights = (const double*)w;
        if (!c->d_erest) {
            void* e = nullptr;
            SDM_CUDA(cudaMalloc(&e, sizeof(double) * (size_t)c->R));
            SDM_CUDA(cudaMemset(e, 0, sizeof(double) * (size_t)c->R));
            c->allocs.push_back(e);
            c->d_erest = (double*)e;
        }
    }
    c->rt_dirty = false;
    c->graph_valid = false;   // one launch more or less per evaluation
    return SDM_OK;
}

// One group of a restraint term: its atoms with normalised weights (NULL weights = equal).
static int push_group(sdm_ctx* c, sdm::RestraintTerm& t, int k, int count, const int32_t* atoms, const double* weights) {
    if (count <= 0 || !atoms) return fail(SDM_ERR_INVALID, "restraint: empty atom group");
    double sum = 0.0;
    for (int i = 0; i < count; i++) {
        if (atoms[i] < 0 || atoms[i] >= c->n) return fail(SDM_ERR_INVALID, "restraint: atom index out of range");
        const double w = weights ? weights[i] : 1.0;
        if (!(w >= 0.0)) return fail(SDM_ERR_INVALID, "restraint: negative weight");
        sum += w;
    }
    if (!(sum > 0.0)) return fail(SDM_ERR_INVALID, "restraint: weights of a group sum to zero");
    t.grp_begin[k] = (int)c->h_ratoms.size();
    for (int i = 0; i < count; i++) {
        c->h_ratoms.push_back(atoms[i]);
        c->h_rweights.push_back((weights ? weights[i] : 1.0) / sum);
    }
    t.grp_begin[k + 1] = (int)c->h_ratoms.size();
    return SDM_OK;
}

// Device buffers of the external dual-state terms ([R][3n] forces of both states, [R][2] energies, [R] flags).
static int ensure_ext_buffers(sdm_ctx* c) {
    if (c->d_ext_f1) return SDM_OK;
    const size_t n3 = 3 * (size_t)c->n;
    void *a = nullptr, *b = nullptr, *e = nullptr, *on = nullptr;
    SDM_CUDA(cudaMalloc(&a, sizeof(double) * n3 * c->R));
    SDM_CUDA(cudaMalloc(&b, sizeof(double) * n3 * c->R));
    SDM_CUDA(cudaMalloc(&e, sizeof(double) * 2 * c->R));
    SDM_CUDA(cudaMalloc(&on, sizeof(int) * c->R));
    SDM_CUDA(cudaMemset(a, 0, sizeof(double) * n3 * c->R));
    SDM_CUDA(cudaMemset(b, 0, sizeof(double) * n3 * c->R));
    SDM_CUDA(cudaMemset(e, 0, sizeof(double) * 2 * c->R));
    SDM_CUDA(cudaMemset(on, 0, sizeof(int) * c->R));
    c->allocs.push_back(a); c->allocs.push_back(b); c->allocs.push_back(e); c->allocs.push_back(on);
    c->d_ext_f1 = (double*)a; c->d_ext_f2 = (double*)b; c->d_ext_e = (double*)e; c->d_ext_on = (int*)on;
    c->B.ext_f1 = c->d_ext_f1; c->B.ext_f2 = c->d_ext_f2; c->B.ext_e = c->d_ext_e; c->B.ext_on = c->d_ext_on;
    c->graph_valid = false;   // kernel arguments of the captured sequence changed
    return SDM_OK;
}

int sdm_enable_reciprocal_pme(sdm_ctx* c, const int32_t* grid) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    SDM_CUDA(cudaStreamSynchronize(c->stream));
    if (int rc = ensure_ext_buffers(c)) return rc;
    if (int rc = sdm_ctx_init_pme(c, grid)) return rc;
    c->graph_valid = false;
    return SDM_OK;
}

int sdm_enable_hct_gb(sdm_ctx* c, const double* charge, const double* offset_radius, const double* scaled_radius,
                      double solute_dielectric, double solvent_dielectric, int sa_ace) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    if (!offset_radius || !scaled_radius) return fail(SDM_ERR_INVALID, "null radius array");
    if (c->T.method == SDM_CUTOFF_PERIODIC || c->T.ewald)
        return fail(SDM_ERR_INVALID, "HCT-GB is evaluated without a cutoff on plain distances: not with a periodic method");
    if (c->pme) return fail(SDM_ERR_INVALID, "the external slots are filled by the library's reciprocal-space PME");
    if (!(solute_dielectric > 0.0) || !(solvent_dielectric > 0.0)) return fail(SDM_ERR_INVALID, "dielectric constants must be positive");
    for (int a = 0; a < c->n; a++)
        if (!(offset_radius[a] > 0.0) || !(scaled_radius[a] >= 0.0)) return fail(SDM_ERR_INVALID, "offset radii must be positive, scaled radii non-negative");
    SDM_CUDA(cudaStreamSynchronize(c->stream));
    if (int rc = ensure_ext_buffers(c)) return rc;
    if (int rc = sdm_ctx_init_gb(c, charge, offset_radius, scaled_radius, solute_dielectric, solvent_dielectric, sa_ace)) return rc;
    c->graph_valid = false;
    return SDM_OK;
}

int sdm_get_born_radii(sdm_ctx* c, int replica, int state, double* radii) {
    SDM_ON_CTX_DEVICE(c);
    if (!c || !radii) return fail(SDM_ERR_INVALID, "null argument");
    if (replica < 0 || replica >= c->R || state < 1 || state > 2) return fail(SDM_ERR_INVALID, "replica or state out of range");
    return sdm_ctx_gb_born_radii(c, replica, state - 1, radii, c->stream);
}

int sdm_set_external_dual(sdm_ctx* c, int replica, const double* f1_ext, const double* f2_ext, double e1_ext,
                          double e2_ext) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    if (replica < 0 || replica >= c->R) return fail(SDM_ERR_INVALID, "replica out of range");
    if ((f1_ext == nullptr) != (f2_ext == nullptr)) return fail(SDM_ERR_INVALID, "give both force arrays or neither");
    if (c->pme) return fail(SDM_ERR_INVALID, "the external slots are filled by the library's reciprocal-space PME");
    if (c->gb) return fail(SDM_ERR_INVALID, "the external slots are filled by the library's HCT-GB model");
    const size_t n3 = 3 * (size_t)c->n;
    if (!c->d_ext_f1 && !f1_ext) return SDM_OK;   // nothing to remove
    if (int rc = ensure_ext_buffers(c)) return rc;
    // staged through the stream (pageable host memory: the copies return when the data has been taken)
    const double e[2] = {f1_ext ? e1_ext : 0.0, f1_ext ? e2_ext : 0.0};
    const int on = f1_ext ? 1 : 0;
    if (f1_ext) {
        SDM_CUDA(cudaMemcpyAsync(c->d_ext_f1 + n3 * replica, f1_ext, sizeof(double) * n3, cudaMemcpyHostToDevice, c->stream));
        SDM_CUDA(cudaMemcpyAsync(c->d_ext_f2 + n3 * replica, f2_ext, sizeof(double) * n3, cudaMemcpyHostToDevice, c->stream));
    }
    SDM_CUDA(cudaMemcpyAsync(c->d_ext_e + 2 * replica, e, sizeof(e), cudaMemcpyHostToDevice, c->stream));
    SDM_CUDA(cudaMemcpyAsync(c->d_ext_on + replica, &on, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    SDM_CUDA(cudaStreamSynchronize(c->stream));
    return SDM_OK;
}

int sdm_add_centroid_restraint(sdm_ctx* c, const sdm_centroid_restraint* r) {
    SDM_ON_CTX_DEVICE(c);
    if (!c || !r) return fail(SDM_ERR_INVALID, "null argument");
    sdm::RestraintTerm t{};
    t.kind = 0;
    t.npoints = r->do_angles ? 8 : 2;
    const size_t keep_a = c->h_ratoms.size();
    int rc = push_group(c, t, 0, r->n_lig_cm, r->lig_cm_atoms, r->lig_cm_weights);
    if (!rc) rc = push_group(c, t, 1, r->n_rcpt_cm, r->rcpt_cm_atoms, r->rcpt_cm_weights);
    for (int k = 0; k < 3 && !rc && r->do_angles; k++) rc = push_group(c, t, 2 + k, 1, &r->rcpt_ref[k], nullptr);
    for (int k = 0; k < 3 && !rc && r->do_angles; k++) rc = push_group(c, t, 5 + k, 1, &r->lig_ref[k], nullptr);
    if (rc) { c->h_ratoms.resize(keep_a); c->h_rweights.resize(keep_a); return rc; }
    t.p[0] = r->kfcm; t.p[1] = r->tolcm; t.p[2] = r->offset[0]; t.p[3] = r->offset[1]; t.p[4] = r->offset[2];
    for (int k = 0; k < 3; k++) { t.p[5 + 3 * k] = r->kfcd[k]; t.p[6 + 3 * k] = r->a[k]; t.p[7 + 3 * k] = r->b[k]; }
    c->h_rterms.push_back(t);
    c->rt_dirty = true;
    return SDM_OK;
}

int sdm_add_alignment_restraint(sdm_ctx* c, const sdm_alignment_restraint* r) {
    SDM_ON_CTX_DEVICE(c);
    if (!c || !r) return fail(SDM_ERR_INVALID, "null argument");
    sdm::RestraintTerm t{};
    t.kind = 1;
    t.npoints = 6;
    const size_t keep_a = c->h_ratoms.size();
    int rc = SDM_OK;
    for (int k = 0; k < 3 && !rc; k++) rc = push_group(c, t, k, 1, &r->ligb_ref[k], nullptr);
    for (int k = 0; k < 3 && !rc; k++) rc = push_group(c, t, 3 + k, 1, &r->liga_ref[k], nullptr);
    if (rc) { c->h_ratoms.resize(keep_a); c->h_rweights.resize(keep_a); return rc; }
    t.p[0] = r->kfdispl; t.p[1] = r->ktheta; t.p[2] = r->kpsi;
    t.p[3] = r->offset[0]; t.p[4] = r->offset[1]; t.p[5] = r->offset[2];
    c->h_rterms.push_back(t);
    c->rt_dirty = true;
    return SDM_OK;
}

int sdm_clear_restraints(sdm_ctx* c) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    c->h_rterms.clear(); c->h_ratoms.clear(); c->h_rweights.clear();
    c->rt_dirty = true;
    return SDM_OK;
}

int sdm_set_restraint_control(sdm_ctx* c, double value) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    c->RT.control = value;
    c->graph_valid = false;   // a launch parameter of the captured sequence
    return SDM_OK;
}

int sdm_get_restraint_energy(sdm_ctx* c, int replica, double* energy) {
    SDM_ON_CTX_DEVICE(c);
    if (!c || !energy) return fail(SDM_ERR_INVALID, "null argument");
    if (replica < 0 || replica >= c->R) return fail(SDM_ERR_INVALID, "replica out of range");
    *energy = 0.0;
    if (!c->d_erest || c->h_rterms.empty()) return SDM_OK;
    SDM_CUDA(cudaStreamSynchronize(c->stream));
    SDM_CUDA(cudaMemcpy(energy, c->d_erest + replica, sizeof(double), cudaMemcpyDeviceToHost));
    return SDM_OK;
}

// Scratch of the displaced-atom kernels, grown on demand (never shrinks): the prefilter bitmap,
// its per-word prefix counts and the per-hit pair forces.
static int regrow_bytes(sdm_ctx* c, void** p, size_t bytes) {
    if (*p) {
        SDM_CUDA(cudaStreamSynchronize(c->stream));
        c->allocs.erase(std::remove(c->allocs.begin(), c->allocs.end(), *p), c->allocs.end());
        SDM_CUDA(cudaFree(*p));
        *p = nullptr;
    }
    c->graph_valid = false;
    SDM_CUDA(cudaMalloc(p, std::max<size_t>(bytes, 16)));
    SDM_CUDA(cudaMemset(*p, 0, std::max<size_t>(bytes, 16)));
    c->allocs.push_back(*p);
    return SDM_OK;
}

static int ensure_hitbits(sdm_ctx* c) {
    sdm::EvalBuffers& B = c->B;
    const sdm::Topology& T = c->T;
    B.scan_words = (B.scan_max + 31) / 32;
    const size_t rows = (size_t)c->R * std::max(T.n_lig, 1);
    const size_t need = rows * B.scan_words;
    if (need > c->hitbits_cap) {
        c->hitbits_cap = need + need / 4;
        if (int rc = regrow_bytes(c, (void**)&c->d_hitbits, c->hitbits_cap * sizeof(uint32_t))) return rc;
        if (int rc = regrow_bytes(c, (void**)&c->d_hitpre, c->hitbits_cap * sizeof(int))) return rc;
    }
    // hits per row: atoms within the cutoff of a displaced atom in either state, at up to ~2x
    // liquid-water number density; everything when there is no cutoff
    int cap = c->n;
    const bool static_cand = c->pair_mode == SDM_PAIR_CLUSTER;
    B.filter_skin = static_cand ? (float)c->opt.skin : 0.f;
    if (T.method != SDM_NOCUTOFF) {
        const double r = T.rc + B.filter_skin + 0.05;
        cap = (int)std::min<double>(c->n, (2.0 * 4.19 * r * r * r * 200.0 + 64.0) * c->pairf_scale);
    }
    cap = std::max(cap, 32);
    const size_t need_f = rows * (size_t)cap * 3;
    if (need_f > c->pairf_alloc) {
        c->pairf_alloc = need_f;
        if (int rc = regrow_bytes(c, (void**)&c->d_pairf, c->pairf_alloc * sizeof(double))) return rc;
    }
    if (static_cand) {
        const size_t need_c = rows * (size_t)cap;
        if (need_c > c->cand_alloc) {
            c->cand_alloc = need_c;
            if (int rc = regrow_bytes(c, (void**)&c->d_cand, c->cand_alloc * sizeof(int))) return rc;
            c->lig_built_for = -1;
        }
        if (rows > c->cand_rows) {
            c->cand_rows = rows;
            if (int rc = regrow_bytes(c, (void**)&c->d_cand_count, c->cand_rows * sizeof(int))) return rc;
            c->lig_built_for = -1;
        }
    }
    if (cap != B.pairf_cap) c->lig_built_for = -1;   // the candidate rows are laid out with this stride
    B.hitbits = c->d_hitbits;
    B.hitpre = c->d_hitpre;
    B.pairf = c->d_pairf;
    B.pairf_cap = cap;
    B.cand = c->d_cand;
    B.cand_count = c->d_cand_count;
    return SDM_OK;
}

int sdm_eval(sdm_ctx* c) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    if (int rc = upload_restraints(c)) return rc;
    cudaStream_t s = c->stream;
    const sdm::Topology& T = c->T;
    sdm::EvalBuffers& B = c->B;
    if (c->pair_mode == SDM_PAIR_ALLPAIRS) {
        if (c->timing) SDM_CUDA(cudaEventRecord(c->ev[0], s));
        if (int rc = ensure_hitbits(c)) return rc;
        sdm::launch_prep_posq(T, B, s);
        if (c->timing) SDM_CUDA(cudaEventRecord(c->ev[1], s));
        sdm::launch_allpairs(T, B, c->opt.exact_cutoff, nullptr, nullptr, 0, -1, s);
        if (c->timing) SDM_CUDA(cudaEventRecord(c->ev[2], s));
        c->launches += 2;
        enqueue_position_only(c, s);
        if (int rc = enqueue_tail(c, 0.5, 2, 0)) return rc;
    } else {
        // Between list rebuilds the kernel sequence is identical from one evaluation to the next
        // (same grids, same pointers): it is captured once and replayed as a CUDA graph, which
        // removes the per-launch CPU cost (8 launches per evaluation).
        const bool graphable = c->opt.use_graph && !c->timing && s != nullptr &&
                               !sdm_ctx_pairlist_rebuild_due(c);
        if (graphable && c->graph_valid) {
            SDM_CUDA(cudaGraphLaunch(c->graph_exec, s));
            c->list_age++;
            c->launches += c->graph_launches;
            c->timing_valid = false;
            c->n_evals++;
            return SDM_OK;
        }
        if (c->timing) SDM_CUDA(cudaEventRecord(c->ev[0], s));
        bool capturing = false;
        int64_t launches0 = c->launches;
        if (graphable) {
            if (int rc = ensure_hitbits(c)) return rc;
            if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) capturing = true;
            else cudaGetLastError();
        }
        int rc = sdm_ctx_pairlist_prepare(c);   // list (re)build or refresh of the sorted positions
        if (!rc) rc = ensure_hitbits(c);         // no-op while capturing (sized before the capture began)
        if (!rc && T.n_lig > 0 && c->lig_built_for != c->n_builds) {
            // new list (never inside a capture: a rebuild is not graphable): prefilter with the
            // list's skin, rows expanded into the candidate lists the evaluations walk
            sdm::launch_ligand_filter(T, B, s);
            sdm::launch_ligand_compact(T, B, s);
            c->launches += 2;
            c->lig_built_for = c->n_builds;
        }
        if (!rc) {
            // fork: the displaced-atom pair terms only need the positions; they fill the tail of
            // the (persistent) pair kernel instead of waiting for it.  Serial when the pair kernel
            // is being timed on its own.
            const bool fork = !c->timing;
            if (fork) {
                cudaEventRecord(c->ev_fork, s);
                cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0);
                // side work first: its (small) grids are dispatched ahead of the persistent pair
                // kernel, whose blocks then fill every slot that is or becomes free
                static const bool side_timing = getenv("SDMB200_SIDE_TIMING") != nullptr;   // development knob (no graph)
                static cudaEvent_t sev[4] = {nullptr, nullptr, nullptr, nullptr};
                if (side_timing && !capturing) {
                    if (!sev[0]) for (auto& e : sev) cudaEventCreate(&e);
                    else {
                        float side_ms = 0.f, pair_ms = 0.f, lag = 0.f;
                        cudaEventSynchronize(sev[1]); cudaEventSynchronize(sev[3]);
                        cudaEventElapsedTime(&side_ms, sev[0], sev[1]);
                        cudaEventElapsedTime(&pair_ms, sev[2], sev[3]);
                        cudaEventElapsedTime(&lag, sev[3], sev[1]);
                        fprintf(stderr, "side chain %.1f us  pair kernel %.1f us  side ends %.1f us after the pair kernel\n",
                                1e3 * side_ms, 1e3 * pair_ms, 1e3 * lag);
                    }
                    cudaEventRecord(sev[0], c->side_stream);
                }
                enqueue_position_only(c, c->side_stream);
                if (side_timing && !capturing) cudaEventRecord(sev[1], c->side_stream);
                cudaEventRecord(c->ev_join, c->side_stream);
                if (side_timing && !capturing) cudaEventRecord(sev[2], s);
                rc = sdm_ctx_pairlist_launch(c);
                if (side_timing && !capturing) cudaEventRecord(sev[3], s);
                cudaStreamWaitEvent(s, c->ev_join, 0);
            } else {
                rc = sdm_ctx_pairlist_launch(c);  // records ev[1], ev[2] around the pair kernel when timing
                enqueue_position_only(c, s);
            }
        }
        // the accumulators are cleared by a memset before the pair pass, or (small batches) by the mix kernel
        if (!rc) rc = enqueue_tail(c, 1.0, 1, sdm_ctx_mix_clears_accumulators(c) ? 1 : 0);
        if (capturing) {
            cudaGraph_t g = nullptr;
            cudaError_t e = cudaStreamEndCapture(s, &g);
            if (rc) { if (g) cudaGraphDestroy(g); return rc; }
            if (e != cudaSuccess || !g) return fail(SDM_ERR_CUDA, std::string("graph capture failed: ") + cudaGetErrorString(e));
            if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
            e = cudaGraphInstantiate(&c->graph_exec, g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) return fail(SDM_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
            c->graph_valid = true;
            c->graph_launches = (int)(c->launches - launches0);
            SDM_CUDA(cudaGraphLaunch(c->graph_exec, s));
        } else if (rc) {
            return rc;
        }
    }
    if (c->timing) SDM_CUDA(cudaEventRecord(c->ev[3], s));
    c->timing_valid = c->timing;
    SDM_CUDA(cudaGetLastError());
    c->n_evals++;
    return SDM_OK;
}

// An evaluation that ran out of per-hit scratch says so in its status; the next one gets twice
// the room (the caller repeats the evaluation, as the header documents for SDM_ERR_CAPACITY).
static void note_status(sdm_ctx* c, int status) {
    if (status == SDM_ERR_CAPACITY && sdm_ctx_pairlist_overflowed(c)) {
        c->list_valid = false;   // the list build ran past its bounds: the next one is sized on the host
        return;
    }
    if (status == SDM_ERR_CAPACITY && c->pairf_scale < (1 << 16)) {
        c->pairf_scale *= 2;
        c->graph_valid = false;
        c->list_valid = false;   // cluster path: the candidate lists are laid down with the list
    }
    // an atom outran the list buffer: the next evaluation rebuilds the list (the caller repeats
    // the evaluation; forces of the flagged one may miss pairs)
    if (status == SDM_ERR_STALE_LIST) c->list_valid = false;
}

int sdm_get_scalars(sdm_ctx* c, int replica, sdm_scalars* out) {
    SDM_ON_CTX_DEVICE(c);
    if (int rc = check_ctx(c, replica)) return rc;
    if (!out) return fail(SDM_ERR_INVALID, "null argument");
    SDM_CUDA(cudaMemcpyAsync(&c->h_state[replica].sc, &c->B.state[replica].sc, sizeof(sdm_scalars),
                             cudaMemcpyDeviceToHost, c->stream));
    SDM_CUDA(cudaMemsetAsync(c->d_sticky + replica, 0, sizeof(int), c->stream));   // the host has it now
    SDM_CUDA(cudaStreamSynchronize(c->stream));
    *out = c->h_state[replica].sc;
    note_status(c, out->status);
    return SDM_OK;
}

int sdm_get_forces(sdm_ctx* c, int replica, int which, double* out) {
    SDM_ON_CTX_DEVICE(c);
    if (int rc = check_ctx(c, replica)) return rc;
    if (!out) return fail(SDM_ERR_INVALID, "null argument");
    const size_t n3 = 3 * (size_t)c->n, off = (size_t)replica * n3;
    const double* src = nullptr;
    switch (which) {
        case SDM_FORCE_HYBRID: src = c->B.F + off; break;
        case SDM_FORCE_STATE1: src = c->B.F1 + off; break;
        case SDM_FORCE_DELTA: src = c->B.dF + off; break;
        case SDM_FORCE_STATE2: src = c->B.F1 + off; break;
        default: return fail(SDM_ERR_INVALID, "unknown force selector");
    }
    SDM_CUDA(cudaMemcpyAsync(out, src, sizeof(double) * n3, cudaMemcpyDeviceToHost, c->stream));
    SDM_CUDA(cudaStreamSynchronize(c->stream));
    if (which == SDM_FORCE_STATE2) {
        std::vector<double> d(n3);
        SDM_CUDA(cudaMemcpy(d.data(), c->B.dF + off, sizeof(double) * n3, cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < n3; k++) out[k] += d[k];  // debug accessor: F2 = F1 + dF
    }
    return SDM_OK;
}

int sdm_read_results(sdm_ctx* c, double* forces_all, sdm_scalars* scalars_all) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    if (forces_all)
        SDM_CUDA(cudaMemcpyAsync(forces_all, c->B.F, sizeof(double) * 3 * (size_t)c->n * c->R,
                                 cudaMemcpyDeviceToHost, c->stream));
    if (scalars_all) {
        SDM_CUDA(cudaMemcpyAsync(c->h_state, c->B.state, sizeof(sdm::ReplicaState) * (size_t)c->R,
                                 cudaMemcpyDeviceToHost, c->stream));
        SDM_CUDA(cudaMemsetAsync(c->d_sticky, 0, sizeof(int) * (size_t)c->R, c->stream));
    }
    SDM_CUDA(cudaStreamSynchronize(c->stream));
    if (scalars_all)
        for (int r = 0; r < c->R; r++) {
            scalars_all[r] = c->h_state[r].sc;
            note_status(c, scalars_all[r].status);
        }
    return SDM_OK;
}

int sdm_enqueue_results(sdm_ctx* c, double* forces_all) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    if (forces_all)
        SDM_CUDA(cudaMemcpyAsync(forces_all, c->B.F, sizeof(double) * 3 * (size_t)c->n * c->R,
                                 cudaMemcpyDeviceToHost, c->stream));
    SDM_CUDA(cudaMemcpyAsync(c->h_state, c->B.state, sizeof(sdm::ReplicaState) * (size_t)c->R,
                             cudaMemcpyDeviceToHost, c->stream));
    // the copy above carries the sticky status to the host; later evaluations start clean (an
    // evaluation that still runs on the stale list raises it again)
    SDM_CUDA(cudaMemsetAsync(c->d_sticky, 0, sizeof(int) * (size_t)c->R, c->stream));
    return SDM_OK;
}

int sdm_enqueue_results_f32(sdm_ctx* c, float* forces_all) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    if (forces_all) {
        if (int rc = ensure_stage32(c)) return rc;
        const size_t count = 3 * (size_t)c->n * c->R;
        sdm::launch_narrow(count, c->B.F, c->d_stage32, c->stream);
        c->launches++;
        SDM_CUDA(cudaMemcpyAsync(forces_all, c->d_stage32, sizeof(float) * count, cudaMemcpyDeviceToHost, c->stream));
    }
    SDM_CUDA(cudaMemcpyAsync(c->h_state, c->B.state, sizeof(sdm::ReplicaState) * (size_t)c->R,
                             cudaMemcpyDeviceToHost, c->stream));
    SDM_CUDA(cudaMemsetAsync(c->d_sticky, 0, sizeof(int) * (size_t)c->R, c->stream));
    return SDM_OK;
}

int sdm_collect_scalars(sdm_ctx* c, sdm_scalars* scalars_all) {
    SDM_ON_CTX_DEVICE(c);
    if (!c || !scalars_all) return fail(SDM_ERR_INVALID, "null argument");
    for (int r = 0; r < c->R; r++) {
        scalars_all[r] = c->h_state[r].sc;
        note_status(c, scalars_all[r].status);
    }
    return SDM_OK;
}

int sdm_forces_device_ptr(sdm_ctx* c, int replica, double** d_f) {
    SDM_ON_CTX_DEVICE(c);
    if (int rc = check_ctx(c, replica)) return rc;
    if (!d_f) return fail(SDM_ERR_INVALID, "null argument");
    *d_f = c->B.F + (size_t)replica * 3 * c->n;
    return SDM_OK;
}

int sdm_get_pairs(sdm_ctx* c, int replica, int32_t* pairs, int64_t max_pairs, int64_t* n_out) {
    SDM_ON_CTX_DEVICE(c);
    if (int rc = check_ctx(c, replica)) return rc;
    if (!n_out) return fail(SDM_ERR_INVALID, "null argument");
    cudaStream_t s = c->stream;
    int* d_counter = nullptr;
    SDM_CUDA(cudaMalloc((void**)&d_counter, sizeof(int)));
    int h_count = 0;
    int* d_pairs = nullptr;
    int cap = 0;
    for (int pass = 0; pass < 2; pass++) {
        SDM_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(int), s));
        int rc = SDM_OK;
        if (c->pair_mode == SDM_PAIR_ALLPAIRS) {
            sdm::launch_prep_posq(c->T, c->B, s);
            // emit_pairs must be non-null to enable emission; capacity 0 only counts
            sdm::launch_allpairs(c->T, c->B, c->opt.exact_cutoff, d_counter,
                                 d_pairs ? d_pairs : d_counter, cap, replica, s);
        } else {
            rc = sdm_ctx_pairlist_emit(c, replica, d_counter, d_pairs ? d_pairs : d_counter, cap);
        }
        if (rc) { cudaFree(d_counter); if (d_pairs) cudaFree(d_pairs); return rc; }
        SDM_CUDA(cudaMemcpyAsync(&h_count, d_counter, sizeof(int), cudaMemcpyDeviceToHost, s));
        SDM_CUDA(cudaStreamSynchronize(s));
        if (pass == 0) {
            if (!pairs || h_count == 0) break;
            cap = h_count;
            SDM_CUDA(cudaMalloc((void**)&d_pairs, sizeof(int) * 2 * (size_t)cap));
        }
    }
    *n_out = h_count;
    if (pairs && d_pairs) {
        std::vector<std::pair<int, int>> v((size_t)h_count);
        SDM_CUDA(cudaMemcpy(v.data(), d_pairs, sizeof(int) * 2 * (size_t)h_count, cudaMemcpyDeviceToHost));
        std::sort(v.begin(), v.end());
        int64_t m = std::min<int64_t>(h_count, max_pairs);
        for (int64_t k = 0; k < m; k++) { pairs[2 * k] = v[k].first; pairs[2 * k + 1] = v[k].second; }
    }
    cudaFree(d_counter);
    if (d_pairs) cudaFree(d_pairs);
    return SDM_OK;
}

int sdm_get_launch_count(sdm_ctx* c, int64_t* n) {
    SDM_ON_CTX_DEVICE(c);
    if (!c || !n) return fail(SDM_ERR_INVALID, "null argument");
    *n = c->launches;
    return SDM_OK;
}

int sdm_set_timing(sdm_ctx* c, int enabled) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    c->timing = enabled != 0;
    c->timing_full_residency = enabled == 2;
    c->timing_valid = false;
    c->graph_valid = false;
    return SDM_OK;
}

int sdm_get_last_timing(sdm_ctx* c, float* pair_ms, float* total_ms) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    if (!c->timing_valid) return fail(SDM_ERR_INVALID, "timing not enabled for the last eval");
    SDM_CUDA(cudaEventSynchronize(c->ev[3]));
    float a = 0, b = 0;
    SDM_CUDA(cudaEventElapsedTime(&a, c->ev[1], c->ev[2]));
    SDM_CUDA(cudaEventElapsedTime(&b, c->ev[0], c->ev[3]));
    if (pair_ms) *pair_ms = a;
    if (total_ms) *total_ms = b;
    return SDM_OK;
}

int sdm_get_info(sdm_ctx* c, const char* key, double* value) {
    SDM_ON_CTX_DEVICE(c);
    if (!c || !key || !value) return fail(SDM_ERR_INVALID, "null argument");
    std::string k(key);
    if (k == "n_atoms") *value = c->n;
    else if (k == "n_replicas") *value = c->R;
    else if (k == "n_displaced") *value = c->T.n_lig;
    else if (k == "pair_mode") *value = c->pair_mode;
    else if (k == "n_evals") *value = (double)c->n_evals;
    else if (k == "n_list_builds") *value = (double)c->n_builds;
    else if (k == "list_age") *value = c->list_valid ? (double)c->list_age : 0.0;   // evaluations since the list was built
    else if (k == "e_dispersion") *value = c->T.e_disp;
    else if (k == "num_sms") *value = c->num_sms;
    else if (k == "fp32_fma_tflops_measured") *value = sdm::measure_fp32_fma_tflops(c->num_sms, c->stream);   // ~1 ms of FMAs
    else if (k == "ewald_alpha") *value = c->T.ewald ? c->T.alpha : 0.0;
    else if (c->pme && sdm_ctx_pme_info(c, k.c_str(), value) == SDM_OK) return SDM_OK;
    else if (c->gb && sdm_ctx_gb_info(c, k.c_str(), value) == SDM_OK) return SDM_OK;
    else if (sdm_ctx_pairlist_info(c, k.c_str(), value) == SDM_OK) return SDM_OK;
    else return fail(SDM_ERR_INVALID, "unknown info key: " + k);
    return SDM_OK;
}

// ---- (B) literal kernel-interface operations ---------------------------------------------------
int sdm_k_make_state2(void* stream, int n, void* posq, const void* displ) {
    if (n < 0 || (n > 0 && (!posq || !displ))) return fail(SDM_ERR_INVALID, "bad argument");
    sdm::launch_make_state2(n, (float4*)posq, (const float4*)displ, (cudaStream_t)stream);
    SDM_CUDA(cudaGetLastError());
    return SDM_OK;
}

int sdm_k_save_state1(void* stream, int n, const void* posq, const void* force, void* save_f, void* save_x) {
    if (n < 0 || (n > 0 && (!posq || !force || !save_f || !save_x))) return fail(SDM_ERR_INVALID, "bad argument");
    sdm::launch_save_state1(n, (const float4*)posq, (const float4*)force, (float4*)save_f, (float4*)save_x,
                            (cudaStream_t)stream);
    SDM_CUDA(cudaGetLastError());
    return SDM_OK;
}

int sdm_k_save_state2(void* stream, int n, const void* force, void* save_f) {
    if (n < 0 || (n > 0 && (!force || !save_f))) return fail(SDM_ERR_INVALID, "bad argument");
    sdm::launch_copy4(n, (const float4*)force, (float4*)save_f, (cudaStream_t)stream);
    SDM_CUDA(cudaGetLastError());
    return SDM_OK;
}

int sdm_k_restore_state1(void* stream, int n, void* posq, const void* saved) {
    if (n < 0 || (n > 0 && (!posq || !saved))) return fail(SDM_ERR_INVALID, "bad argument");
    sdm::launch_copy4(n, (const float4*)saved, (float4*)posq, (cudaStream_t)stream);
    SDM_CUDA(cudaGetLastError());
    return SDM_OK;
}

int sdm_k_hybrid_force(void* stream, int n, const void* f1, const void* f2, void* force, float sp) {
    if (n < 0 || (n > 0 && (!f1 || !f2 || !force))) return fail(SDM_ERR_INVALID, "bad argument");
    sdm::launch_hybrid_force(n, (const float4*)f1, (const float4*)f2, (float4*)force, sp, (cudaStream_t)stream);
    SDM_CUDA(cudaGetLastError());
    return SDM_OK;
}

int sdm_k_langevin_part1(void* stream, int n, void* velm, const void* force, void* pos_delta, float vscale,
                         float fscale, float noisescale, float step_size, const void* random,
                         uint32_t random_index) {
    if (n < 0 || (n > 0 && (!velm || !force || !pos_delta || !random))) return fail(SDM_ERR_INVALID, "bad argument");
    sdm::launch_langevin_part1(n, (float4*)velm, (const float4*)force, (float4*)pos_delta, vscale, fscale,
                               noisescale, step_size, (const float4*)random, random_index, (cudaStream_t)stream);
    SDM_CUDA(cudaGetLastError());
    return SDM_OK;
}

int sdm_k_langevin_part2(void* stream, int n, void* posq, const void* pos_delta, void* velm, float step_size) {
    if (n < 0 || (n > 0 && (!posq || !pos_delta || !velm))) return fail(SDM_ERR_INVALID, "bad argument");
    if (!(step_size > 0.f)) return fail(SDM_ERR_INVALID, "step size must be positive");
    sdm::launch_langevin_part2(n, (float4*)posq, (const float4*)pos_delta, (float4*)velm, step_size,
                               (cudaStream_t)stream);
    SDM_CUDA(cudaGetLastError());
    return SDM_OK;
}

int sdm_langevin_params(double temperature, double friction, double step_size, double* vscale,
                        double* fscale, double* noisescale) {
    if (!vscale || !fscale || !noisescale) return fail(SDM_ERR_INVALID, "null argument");
    const double BOLTZ = 1.380658e-23 * 6.0221367e23 / 1000.0;   // kJ/mol/K, OpenCLSDMKernels.cpp:57-60
    const double kT = BOLTZ * temperature;
    *vscale = std::exp(-step_size * friction);
    *fscale = friction == 0 ? step_size : (1 - *vscale) / friction;
    *noisescale = std::sqrt(kT * (1 - *vscale * *vscale));
    return SDM_OK;
}

// ---- device-resident Langevin dynamics (no constraints) -----------------------------

int sdm_md_init(sdm_ctx* c, const double* masses, double temperature, double friction, double step_size,
                uint64_t seed) {
    SDM_ON_CTX_DEVICE(c);
    if (!c || !masses) return fail(SDM_ERR_INVALID, "null argument");
    if (!(friction > 0.0) || !(step_size > 0.0) || !(temperature >= 0.0))
        return fail(SDM_ERR_INVALID, "sdm_md_init needs friction > 0, step_size > 0, temperature >= 0 "
                                     "(the reference's update divides by the friction)");
    const int n = c->n, R = c->R;
    if (!c->md_ready) {
        if (int rc = dev_alloc(c, &c->d_vel, 3 * (size_t)n * R)) return rc;
        if (int rc = dev_alloc(c, &c->d_noise, 3 * (size_t)n * R)) return rc;
        if (int rc = dev_alloc(c, &c->d_invm, (size_t)n)) return rc;
        if (int rc = dev_alloc(c, &c->d_mass, (size_t)n)) return rc;
        if (int rc = dev_alloc(c, &c->d_ke, (size_t)R)) return rc;
    }
    std::vector<double> invm(n);
    for (int i = 0; i < n; i++) invm[i] = masses[i] == 0.0 ? 0.0 : 1.0 / masses[i];   // ReferenceStochasticDynamicsSDM.cpp:233-240
    SDM_CUDA(cudaMemcpyAsync(c->d_invm, invm.data(), sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    SDM_CUDA(cudaMemcpyAsync(c->d_mass, masses, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    SDM_CUDA(cudaMemsetAsync(c->d_vel, 0, sizeof(double) * 3 * (size_t)n * R, c->stream));
    SDM_CUDA(cudaStreamSynchronize(c->stream));   // invm is a host temporary
    // the reference's constants, same expressions (ReferenceStochasticDynamicsSDM.cpp:144-148)
    const double tau = 1.0 / friction;
    const double BOLTZ = 1.380658e-23 * 6.0221367e23 / 1000.0;
    const double kT = BOLTZ * temperature;
    c->md_dt = step_size;
    c->md_vscale = std::exp(-step_size / tau);
    c->md_fscale = (1 - c->md_vscale) * tau;
    c->md_noisescale = std::sqrt(2 * kT / tau) * std::sqrt(0.5 * (1 - c->md_vscale * c->md_vscale) * tau);
    // The noise of step k is Philox(seed, atom, offset 8*k).  Re-initialising with new parameters
    // (the mirrors do that when temperature, friction or step size change: "dynamics object
    // recreated", LangevinIntegratorSDM.cpp:160-168) must not replay the stream from step 0 -- the
    // reference's SimTK generator is global and carries on -- so the step counter survives unless
    // the seed changes.
    if (!c->md_ready || seed != c->md_seed) c->md_steps = 0;
    c->md_seed = seed;
    c->md_noise_pending = false;
    if (!c->d_md_ctl) {
        if (int rc = dev_alloc(c, &c->d_md_ctl, 4)) return rc;
        SDM_CUDA(cudaMallocHost((void**)&c->h_md_ctl, 8 * sizeof(unsigned long long)));
        c->B.md_ctl = c->d_md_ctl;
        c->graph_valid = false;   // the captured scalar stage must see the control words
    }
    {
        const unsigned long long ctl0[4] = {0ull, c->md_steps, 0ull, 0ull};
        SDM_CUDA(cudaMemcpy(c->d_md_ctl, ctl0, sizeof(ctl0), cudaMemcpyHostToDevice));
    }
    c->md_ready = true;
    return SDM_OK;
}

// Distance constraints: connected clusters of the constraint graph; rigid three-site molecules get
// SETTLE, everything else an in-thread SHAKE (kernels_md.cu).
int sdm_md_set_constraints(sdm_ctx* c, int32_t n_constraints, const int32_t* pairs, const double* distances,
                           double tolerance) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    if (!c->md_ready) return fail(SDM_ERR_INVALID, "call sdm_md_init first");
    if (n_constraints < 0 || (n_constraints > 0 && (!pairs || !distances)))
        return fail(SDM_ERR_INVALID, "bad constraint arguments");
    const int n = c->n, R = c->R;
    SDM_CUDA(cudaStreamSynchronize(c->stream));
    c->cons_tol = tolerance > 0 ? tolerance : 1e-5;
    c->h_cons_pairs.assign(pairs, pairs + 2 * (size_t)n_constraints);
    c->h_cons_dist.assign(distances, distances + n_constraints);
    c->mdc = sdm::MdConstraints{};
    c->mdc.tol = c->cons_tol;
    if (n_constraints == 0) return SDM_OK;
    std::vector<double> mass(n);
    SDM_CUDA(cudaMemcpy(mass.data(), c->d_mass, sizeof(double) * n, cudaMemcpyDeviceToHost));
    // union-find over the constraint graph
    std::vector<int> parent(n);
    for (int i = 0; i < n; i++) parent[i] = i;
    auto find = [&](int a) { while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; } return a; };
    for (int k = 0; k < n_constraints; k++) {
        const int a = pairs[2 * k], b = pairs[2 * k + 1];
        if (a < 0 || b < 0 || a >= n || b >= n || a == b || !(distances[k] > 0))
            return fail(SDM_ERR_INVALID, "constraint with a bad particle index or distance");
        parent[find(a)] = find(b);
    }
    std::vector<std::vector<int>> cons_of(n), atoms_of(n);   // keyed by cluster root
    for (int k = 0; k < n_constraints; k++) cons_of[find(pairs[2 * k])].push_back(k);
    std::vector<unsigned char> in_cluster(n, 0);
    for (int i = 0; i < n; i++)
        if (!cons_of[find(i)].empty()) { atoms_of[find(i)].push_back(i); in_cluster[i] = 1; }
    std::vector<int> settle_atoms, shake_off{0}, shake_ij, shake_aoff{0}, shake_atoms;
    std::vector<double> settle_par, shake_d;
    for (int root = 0; root < n; root++) {
        const std::vector<int>& ks = cons_of[root];
        const std::vector<int>& as = atoms_of[root];
        if (ks.empty()) continue;
        bool settled = false;
        if (as.size() == 3 && ks.size() == 3) {
            // three mutual constraints: find the apex (two equal legs to atoms of equal, non-zero mass)
            auto dist = [&](int a, int b) {
                for (int k : ks)
                    if ((pairs[2 * k] == a && pairs[2 * k + 1] == b) || (pairs[2 * k] == b && pairs[2 * k + 1] == a))
                        return distances[k];
                return -1.0;
            };
            for (int t = 0; t < 3 && !settled; t++) {
                const int a0 = as[t], a1 = as[(t + 1) % 3], a2 = as[(t + 2) % 3];
                const double d01 = dist(a0, a1), d02 = dist(a0, a2), d12 = dist(a1, a2);
                if (d01 > 0 && d02 > 0 && d12 > 0 && std::fabs(d01 - d02) <= 1e-12 * d01 && mass[a1] == mass[a2] &&
                    mass[a0] > 0 && mass[a1] > 0 && d12 < 2 * d01) {
                    settle_atoms.insert(settle_atoms.end(), {a0, a1, a2});
                    settle_par.insert(settle_par.end(), {d01, d12});
                    settled = true;
                }
            }
        }
        if (settled) continue;
        for (int k : ks) {
            shake_ij.push_back(pairs[2 * k]);
            shake_ij.push_back(pairs[2 * k + 1]);
            shake_d.push_back(distances[k]);
        }
        shake_off.push_back((int)shake_d.size());
        shake_atoms.insert(shake_atoms.end(), as.begin(), as.end());
        shake_aoff.push_back((int)shake_atoms.size());
    }
    sdm::MdConstraints& M = c->mdc;
    M.n_settle = (int)settle_atoms.size() / 3;
    M.n_shake = (int)shake_off.size() - 1;
    if (int rc = dev_upload(c, &M.settle_atoms, settle_atoms)) return rc;
    if (int rc = dev_upload(c, &M.settle_par, settle_par)) return rc;
    if (int rc = dev_upload(c, &M.shake_off, shake_off)) return rc;
    if (int rc = dev_upload(c, &M.shake_ij, shake_ij)) return rc;
    if (int rc = dev_upload(c, &M.shake_d, shake_d)) return rc;
    if (int rc = dev_upload(c, &M.shake_aoff, shake_aoff)) return rc;
    if (int rc = dev_upload(c, &M.shake_atoms, shake_atoms)) return rc;
    if (int rc = dev_upload(c, &M.in_cluster, in_cluster)) return rc;
    if (!c->d_xprime)
        if (int rc = dev_alloc(c, &c->d_xprime, 3 * (size_t)n * R)) return rc;
    return SDM_OK;
}

int sdm_md_get_counters(sdm_ctx* c, uint64_t* steps_taken, uint64_t* steps_repeated) {
    SDM_ON_CTX_DEVICE(c);
    if (!c) return fail(SDM_ERR_INVALID, "null context");
    if (!c->md_ready) return fail(SDM_ERR_INVALID, "call sdm_md_init first");
    if (steps_taken) *steps_taken = c->md_steps;
    if (steps_repeated) *steps_repeated = c->md_repeated;
    return SDM_OK;
}

static int md_check(sdm_ctx* c, int replica) {
    if (int rc = check_ctx(c, replica)) return rc;
    if (!c->md_ready) return fail(SDM_ERR_INVALID, "call sdm_md_init first");
    return SDM_OK;
}

int sdm_md_set_velocities(sdm_ctx* c, int replica, const double* v) {
    SDM_ON_CTX_DEVICE(c);
    if (int rc = md_check(c, replica)) return rc;
    if (!v) return fail(SDM_ERR_INVALID, "null velocities");
    SDM_CUDA(cudaMemcpyAsync(c->d_vel + (size_t)replica * 3 * c->n, v, sizeof(double) * 3 * (size_t)c->n,
                             cudaMemcpyHostToDevice, c->stream));
    SDM_CUDA(cudaStreamSynchronize(c->stream));
    return SDM_OK;
}

int sdm_md_get_velocities(sdm_ctx* c, int replica, double* v) {
    SDM_ON_CTX_DEVICE(c);
    if (int rc = md_check(c, replica)) return rc;
    if (!v) return fail(SDM_ERR_INVALID, "null argument");
    SDM_CUDA(cudaMemcpyAsync(v, c->d_vel + (size_t)replica * 3 * c->n, sizeof(double) * 3 * (size_t)c->n,
                             cudaMemcpyDeviceToHost, c->stream));
    SDM_CUDA(cudaStreamSynchronize(c->stream));
    return SDM_OK;
}

int sdm_get_positions(sdm_ctx* c, int replica, double* xyz) {
    SDM_ON_CTX_DEVICE(c);
    if (int rc = check_ctx(c, replica)) return rc;
    if (!xyz) return fail(SDM_ERR_INVALID, "null argument");
    SDM_CUDA(cudaMemcpyAsync(xyz, c->d_pos + (size_t)replica * 3 * c->n, sizeof(double) * 3 * (size_t)c->n,
                             cudaMemcpyDeviceToHost, c->stream));
    SDM_CUDA(cudaStreamSynchronize(c->stream));
    return SDM_OK;
}

int sdm_md_set_noise(sdm_ctx* c, const double* xi_all) {
    SDM_ON_CTX_DEVICE(c);
    if (int rc = md_check(c, 0)) return rc;
    c->md_noise_pending = xi_all != nullptr;
    if (xi_all) {
        SDM_CUDA(cudaMemcpyAsync(c->d_noise, xi_all, sizeof(double) * 3 * (size_t)c->n * c->R,
                                 cudaMemcpyHostToDevice, c->stream));
        SDM_CUDA(cudaStreamSynchronize(c->stream));
    }
    return SDM_OK;
}

// One Langevin update (+ constraints) of all replicas at step number c->md_steps.  guarded: the
// kernels obey the control words (a stale list reported by this step's evaluation turns the update
// into a no-op); sdm_md_update passes false, it integrates whatever force it is given.
static void md_enqueue_update(sdm_ctx* c, bool guarded) {
    sdm::launch_md_update(c->n, c->R, c->d_pos, c->d_vel, c->B.F, c->d_invm, c->md_vscale, c->md_fscale,
                          c->md_noisescale, c->md_dt, c->md_noise_pending ? c->d_noise : nullptr, c->md_seed,
                          c->md_steps, &c->mdc, c->d_xprime, guarded ? c->d_md_ctl : nullptr, c->B.flags, c->stream);
    c->md_noise_pending = false;
    c->md_steps++;
    c->launches += (c->mdc.n_settle + c->mdc.n_shake) > 0 ? 2 : 1;
}

int sdm_md_update(sdm_ctx* c, const double* forces_all) {
    SDM_ON_CTX_DEVICE(c);
    if (int rc = md_check(c, 0)) return rc;
    if (forces_all)
        SDM_CUDA(cudaMemcpyAsync(c->B.F, forces_all, sizeof(double) * 3 * (size_t)c->n * c->R,
                                 cudaMemcpyHostToDevice, c->stream));
    md_enqueue_update(c, false);
    SDM_CUDA(cudaGetLastError());
    return SDM_OK;
}

int sdm_md_step(sdm_ctx* c, int nsteps) {
    SDM_ON_CTX_DEVICE(c);
    if (int rc = md_check(c, 0)) return rc;
    if (nsteps <= 0) return SDM_OK;
    const unsigned long long target = c->md_steps + (unsigned long long)nsteps;
    const int nst = std::max(1, c->opt.nstlist > 0 ? c->opt.nstlist : 20);
    unsigned int* d_disp = sdm_ctx_pairlist_max_disp_ptr(c);
    // Planned list lifetime: the list is rebuilt BEFORE its fastest atom has used up half the skin --
    // an evaluation on a stale list is wasted work (it is never integrated) -- so the host watches
    // how fast the largest displacement grows and schedules the rebuild at 95 % of the predicted
    // lifetime (measured on the 20 k-atom fixture, 16 replicas, 400 steps: 0.8 / 0.9 / 0.95 / 1.0 -> 0.511 /
    // 0.494 / 0.482 / 0.481 ms per step with 0 / 0 / 0 / 3 repeated steps), never later than opt.nstlist.  A
    // stale list that slips through is still caught, and shortens the plan.
    if (c->md_plan <= 0 || c->md_plan > nst) c->md_plan = nst;
    int futile = 0;   // consecutive chunks that did not advance at all
    while (c->md_steps < target) {
        const unsigned long long start = c->md_steps;
        if (d_disp && c->list_valid && c->list_age >= c->md_plan) c->list_valid = false;
        const int age = (d_disp && c->list_valid) ? c->list_age : 0;
        const int todo = (int)std::min<unsigned long long>(target - start, (unsigned long long)std::max(1, (d_disp ? c->md_plan : nst) - age));
        for (int k = 0; k < todo; k++) {
            if (int rc = sdm_eval(c)) return rc;   // hybrid force of every replica at the current positions
            md_enqueue_update(c, true);            // positions and velocities advance on the device
        }
        // how far did the device get?  ([0] stale list / capacity, [1] steps taken, [2] constraints)
        SDM_CUDA(cudaMemcpyAsync(c->h_md_ctl, c->d_md_ctl, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                 c->stream));
        if (d_disp)
            SDM_CUDA(cudaMemcpyAsync(&c->h_md_ctl[4], d_disp, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
        SDM_CUDA(cudaMemcpyAsync(c->h_state, c->B.state, sizeof(sdm::ReplicaState) * (size_t)c->R,
                                 cudaMemcpyDeviceToHost, c->stream));
        SDM_CUDA(cudaStreamSynchronize(c->stream));
        if (c->h_md_ctl[2] != 0ull) {
            c->md_steps = c->h_md_ctl[1];
            return fail(SDM_ERR_CONSTRAINT, "a constraint cluster did not converge (step size too large?)");
        }
        if (c->h_md_ctl[0] == 0ull) {   // all of them taken
            if (d_disp && c->list_valid && c->list_age >= 2) {
                float d2;
                const unsigned int bits = (unsigned int)c->h_md_ctl[4];
                std::memcpy(&d2, &bits, sizeof(d2));
                const double rate = std::sqrt((double)d2) / (double)(c->list_age - 1);   // nm per step, so far
                if (rate > 0) {
                    const double life = 0.5 * c->opt.skin / rate;
                    static const double safety = getenv("SDMB200_MD_SAFETY") ? atof(getenv("SDMB200_MD_SAFETY")) : 0.95;   // development knob
                    c->md_plan = std::max(2, std::min(nst, (int)std::floor(safety * life)));
                }
            }
            continue;
        }
        // the evaluation of step h_md_ctl[1] reported a stale list / full scratch: that step and the
        // ones behind it were not taken.  Rebuild / grow (note_status), clear the flags, go again.
        const unsigned long long taken = c->h_md_ctl[1];
        c->md_repeated += start + (unsigned long long)todo - taken;
        c->md_steps = taken;
        for (int r = 0; r < c->R; r++) note_status(c, c->h_state[r].sc.status);
        c->list_valid = false;
        c->md_plan = std::max(2, (int)(0.7 * c->md_plan));
        SDM_CUDA(cudaMemsetAsync(c->d_md_ctl, 0, sizeof(unsigned long long), c->stream));
        SDM_CUDA(cudaMemsetAsync(c->d_sticky, 0, sizeof(int) * (size_t)c->R, c->stream));
        futile = taken == start ? futile + 1 : 0;
        if (futile >= 3)
            return fail(SDM_ERR_STALE_LIST, "a freshly built pair list goes stale within one step: "
                                            "skin too small for this step size");
    }
    SDM_CUDA(cudaGetLastError());
    return SDM_OK;
}

int sdm_md_kinetic_energy(sdm_ctx* c, int replica, double* ke) {
    SDM_ON_CTX_DEVICE(c);
    if (int rc = md_check(c, replica)) return rc;
    if (!ke) return fail(SDM_ERR_INVALID, "null argument");
    sdm::launch_kinetic_energy(c->n, c->R, c->d_vel, c->d_mass, c->d_ke, c->stream);
    c->launches++;
    SDM_CUDA(cudaMemcpyAsync(ke, c->d_ke + replica, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SDM_CUDA(cudaStreamSynchronize(c->stream));
    return SDM_OK;
}

int sdm_execute_scalars(sdm_alch* alch, double E1, double E2, double Eb, sdm_scalars* out) {
    if (!alch || !out) return fail(SDM_ERR_INVALID, "null argument");
    std::memset(out, 0, sizeof(*out));
    out->E1 = E1;
    out->E2 = E2;
    out->Eb = Eb;
    out->u = E2 - E1;
    sdm::execute_scalars(alch, out);
    if (out->status == SDM_ERR_SOFTCORE) return fail(SDM_ERR_SOFTCORE, "Unknown soft core method");
    return SDM_OK;
}

}  // extern "C"
