// gm_ctx.h -- the context object behind the C ABI and the helpers shared by gm_api.cu / gm_dist.cu (internal)
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gm_b200.h"
#include "gm_kernels.h"

using namespace gm;

inline std::string& create_error() { static thread_local std::string e; return e; }

// a patch pool and the stream its kernels are ordered on, shared by a context and its clones (gm_clone): the clones' patch
// directories hold references into the same pool, copy-on-write keeps them apart
struct PoolShare {
    int users = 1;
    std::vector<struct gm_ctx*> members;
};

struct gm_ctx {
    PoolShare* share = nullptr;  // null: the context owns its pool alone
    int device = 0;
    cudaStream_t st = nullptr;
    cudaStream_t st_w = nullptr;  // weights (normalize / resample indexes): independent of the maps, may overtake a registration
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evw0 = nullptr, evw1 = nullptr;
    cudaEvent_t evr0 = nullptr, evr1 = nullptr;  // the registration's own pair: it may run while a scan match is being timed
    cudaEvent_t ev_sm = nullptr;                 // k_scanmatch done (st_w waits for it before normalising)
    uint32_t max_particles = 0, n = 0, threads = 256;
    DevParams P{};
    bool laser_set = false, match_set = false, inited = false;
    double angles_h[GM_MAXBEAMS];

    Pool pool{};
    int* refcnt_new = nullptr;
    int *dir = nullptr, *dir_alt = nullptr;
    int dir_cap = 0;
    Geom *geom = nullptr, *geom_alt = nullptr, *geom_new = nullptr;
    int4* shift = nullptr;

    double *d_ranges = nullptr, *d_angles = nullptr, *d_poses = nullptr;
    double *d_score = nullptr, *d_s = nullptr, *d_l = nullptr;
    unsigned *d_c = nullptr, *d_evals = nullptr, *d_idx = nullptr, *d_emitted = nullptr, *d_widx = nullptr;
    double *d_w = nullptr, *d_logw = nullptr, *d_cum = nullptr, *d_tgt = nullptr, *d_neff = nullptr;
    // query staging (gm_score / gm_likelihood_and_score)
    unsigned* d_qidx = nullptr; double *d_qposes = nullptr, *d_qposes_out = nullptr, *d_qs = nullptr, *d_ql = nullptr; unsigned* d_qc = nullptr;
    uint32_t q_cap = 0;
    Counters* d_counters = nullptr;
    Counters* d_counters_sm = nullptr;  // gm_scanmatch_weights' own: a registration queued right behind the scan match zeroes the others
    bool fused_register = false;        // gm_set_fused_register
    // launches of at most this many particles use k_scanmatch's wide shape (GM_B200_SM_WIDE_BELOW).  Measured on a B200 at 1081 beams:
    // 512 particles 1.32 -> 1.10 ms, 1024: 2.04 -> 1.87, 2048: 3.56 -> 3.51, 4096: 6.61 -> 6.77 (profiles/experiments_r02.json)
    uint32_t sm_wide_below = 2048;
    // directory buffers a sharded context outgrew: its peers may still have them mapped (CUDA IPC), and freeing exported memory
    // before every importer has closed it is undefined -- they are freed with the context
    std::vector<void*> retired;
    bool sw_pending = false; uint32_t sw_total = 0; double sw_gain = 1;  // between gm_scanmatch_weights_begin and _end
    double* d_inv_n = nullptr;
    size_t scanmatch_smem = 0;
    unsigned long long* d_digest = nullptr;
    int *d_list = nullptr, *d_count = nullptr; int4* d_export = nullptr; size_t export_cap = 0; size_t list_cap = 0;

    gm_stats stats{};
    std::string err;
    size_t register_smem = 0;
    signed char* d_occ = nullptr; size_t occ_cap = 0;  // gm_export_occupancy staging
    bool async_register = false, reg_pending = false;  // gm_set_async_register: a registration may still be running
    int reg_tile_cap = 0;
    int reg_active_hint = 0;  // largest active-patch count seen by the last k_register (sizes its tile allocation)
    uint32_t weights_cap = 0;  // capacity of the weight buffers: max(max_particles, global particle count)
    struct gm_dist* dist = nullptr;  // multi-GPU state (gm_dist.cu), null on a single GPU
    // gm_scanmatch_weights: result rows {x, y, theta, score, l} of all (global) particles, the CTA-completion counter of the
    // kernel's epilogue, and pinned host staging for the one upload and the one download of a scan
    double* d_rows = nullptr;
    unsigned* d_done = nullptr;
    unsigned char *h_in = nullptr, *h_out = nullptr; size_t h_in_bytes = 0, h_out_bytes = 0;
    unsigned dir_gen = 0;  // bumped whenever dir / dir_alt are re-allocated (peers re-open their IPC mappings)
    void (*fail_hook)(gm_ctx*) = nullptr;  // sharded contexts: tell the other ranks that this one failed (they stop waiting)
};


extern "C" int gmi_register_finish(gm_ctx* ctx);

inline int fail(gm_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf; else create_error() = buf;
    if (c && c->fail_hook && code != GM_ERR_ARG) c->fail_hook(c);
    return code;
}

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) return fail(ctx, GM_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <class T> inline cudaError_t dalloc(T*& p, size_t count) {
    if (p) { cudaFree(p); p = nullptr; }
    return cudaMalloc(reinterpret_cast<void**>(&p), std::max<size_t>(count, 1) * sizeof(T));
}

// free a directory buffer -- later, if peers may have it mapped
inline void release_exported(gm_ctx* ctx, void* p) {
    if (!p) return;
    if (ctx->dist) ctx->retired.push_back(p); else cudaFree(p);
}

inline int bind(gm_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
    return GM_OK;
}

inline int read_counters(gm_ctx* ctx, Counters& h) {
    CK(cudaMemcpyAsync(&h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return GM_OK;
}

inline int zero_counters(gm_ctx* ctx) {
    CK(cudaMemsetAsync(ctx->d_counters, 0, sizeof(Counters), ctx->st));
    return GM_OK;
}

inline int check_launch(gm_ctx* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, GM_ERR_CUDA, "launch %s: %s", what, cudaGetErrorString(e));
    return GM_OK;
}

inline unsigned count_score_beams(const gm_ctx* ctx, const double* ranges) {
    const DevParams& P = ctx->P;
    unsigned v = 0;
    for (unsigned b = P.beam_skip; b < P.nbeams; b++) {
        if (((b - P.beam_skip + 1u) % (P.lskip + 1u)) != 0u) continue;
        double r = ranges[b];
        if (r > P.usable_range || r == 0.0) continue;
        v++;
    }
    return v;
}

inline int ensure_query(gm_ctx* ctx, uint32_t count) {
    if (count <= ctx->q_cap) return GM_OK;
    CK(dalloc(ctx->d_qidx, count)); CK(dalloc(ctx->d_qposes, (size_t)count * 3)); CK(dalloc(ctx->d_qposes_out, (size_t)count * 3));
    CK(dalloc(ctx->d_qs, count)); CK(dalloc(ctx->d_ql, count)); CK(dalloc(ctx->d_qc, count));
    ctx->q_cap = count;
    return GM_OK;
}

// constants of score()/likelihoodAndScore() evaluated once on the host with the reference's expressions
inline void derive_params(DevParams& P) {
    volatile double delta = P.delta, ratio = P.free_cell_ratio, gs = P.gaussian_sigma, ls = P.likelihood_sigma;
    volatile double free_delta = delta * ratio;
    P.inv_delta = 1. / delta;
    P.free_off_score = delta * free_delta;
    P.free_off_lik = free_delta;
    P.cexp = -1. / gs;
    P.clik = -1. / ls;
    P.no_hit = -.5 / ls;
    const double offs[2] = {P.free_off_score, P.free_off_lik}, cen[2] = {P.cx, P.cy};
    for (int m = 0; m < 2; m++) {
        bool constant = true;
        int q[2] = {0, 0};
        for (int ax = 0; ax < 2; ax++) {
            // the computed difference pfree - phit can exceed free_off by rounding errors of the two sums: 1e-9 m of slack
            const volatile double lo = (-(offs[m] + 1e-9) - cen[ax]) / delta, hi = ((offs[m] + 1e-9) - cen[ax]) / delta;
            const double rl = round(lo), rh = round(hi);
            if (rl != rh || !(fabs(rl) < 1e9)) constant = false;
            q[ax] = (int)rl;
        }
        P.free_q[m][0] = constant ? q[0] : INT_MIN;
        P.free_q[m][1] = constant ? q[1] : 0;
    }
}

inline int prepare_scanmatch(gm_ctx* ctx) {
    const size_t smem = scanmatch_smem_bytes(ctx->P.nbeams);
    if (smem != ctx->scanmatch_smem) { CK(scanmatch_prepare(smem)); ctx->scanmatch_smem = smem; }
    return GM_OK;
}

inline int ready(gm_ctx* ctx, bool need_particles) {
    if (!ctx) return GM_ERR_ARG;
    if (!ctx->laser_set) return fail(ctx, GM_ERR_ARG, "gm_set_laser has not been called");
    if (!ctx->match_set) return fail(ctx, GM_ERR_ARG, "gm_set_matching has not been called");
    if (need_particles && !ctx->inited) return fail(ctx, GM_ERR_ARG, "gm_init_particles has not been called");
    if (int r = bind(ctx)) return r;
    if (ctx->sw_pending) return fail(ctx, GM_ERR_ARG, "a gm_scanmatch_weights_begin is waiting for its gm_scanmatch_weights_end");
    return ctx->reg_pending ? gmi_register_finish(ctx) : GM_OK;  // an asynchronous registration reports here
}
// entry points that touch weights only: they run on their own stream and leave a registration in flight alone
inline int ready_weights(gm_ctx* ctx) {
    if (!ctx) return GM_ERR_ARG;
    return bind(ctx);
}


// phases of gm_register (gm_api.cu), reused by the multi-GPU path (gm_dist.cu)
extern "C" {
int gmi_stage(gm_ctx* ctx, const double* ranges, const double* poses);
int gmi_remap(gm_ctx* ctx, const uint32_t* idx, uint32_t src_limit);
int gmi_bounds_and_resize(gm_ctx* ctx);
int gmi_register_update(gm_ctx* ctx);
int gmi_register_finish(gm_ctx* ctx);
int gmi_ensure_dir_cap(gm_ctx* ctx, int need);
void gmi_dist_destroy(gm_ctx* ctx);
// sharded contexts: where k_scanmatch's epilogue writes the result rows (every rank's gather buffer) and the flag it raises;
// then the wait for every rank's flag.  Single GPU: the context's own d_rows, no flag.
int gmi_dist_rows(gm_ctx* ctx, ScanMatchArgs* a);
int gmi_dist_wait_rows(gm_ctx* ctx, const double** rows);
}
