// gm_api.cu -- the C ABI (include/gm_b200.h) over the sm_100a kernels of gm_kernels.cu.
//
// One gm_ctx owns, on one GPU: the patch pool, the per-particle patch directories (double buffered),
// per-particle map geometry, the beam table and all staging buffers.  Every entry point copies the
// caller's host buffers in, launches on the context's stream, copies results out and synchronises.
// There is no CPU path: without a device gm_create fails with GM_ERR_NO_DEVICE.
#include "gm_ctx.h"

extern "C" {

const char* gm_last_error(const gm_ctx* ctx) { return ctx ? ctx->err.c_str() : create_error().c_str(); }

void gm_default_matching(gm_match_params* p) {  // scanmatcher.cpp:17-56
    memset(p, 0, sizeof *p);
    p->laser_max_range = 80; p->usable_range = 80; p->gaussian_sigma = 1; p->likelihood_sigma = 1;
    p->kernel_size = 1; p->generate_map = 1;
    p->opt_linear_delta = 0.05; p->opt_angular_delta = 0.1;
    p->opt_recursive_iterations = 3; p->likelihood_skip = 0;
    p->enlarge_step = 10.; p->fullness_threshold = 0.1;
    p->angular_odometry_reliability = 0.; p->linear_odometry_reliability = 0.;
    p->free_cell_ratio = sqrt(2.);
    p->initial_beams_skip = 0;
}

// share_with: take that context's patch pool and stream instead of allocating (gm_clone)
static int create_ctx(gm_ctx** out, const gm_config* cfg, gm_ctx* share_with) {
    gm_ctx* ctx = nullptr;  // CK() reports into g_create_error while ctx is null
    if (!out || !cfg) return fail(nullptr, GM_ERR_ARG, "gm_create: null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(nullptr, GM_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, GM_ERR_ARG, "gm_create: device %d of %d", cfg->device, ndev);
    if (!cfg->max_particles) return fail(nullptr, GM_ERR_ARG, "gm_create: max_particles == 0");
    CK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) return fail(nullptr, GM_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);

    gm_ctx* c = new gm_ctx();
    c->device = cfg->device;
    c->max_particles = cfg->max_particles;
    c->threads = cfg->block_threads ? cfg->block_threads : 256;
    if (c->threads % 32 || c->threads > 256 || c->threads < 32) { delete c; return fail(nullptr, GM_ERR_ARG, "block_threads must be a multiple of 32 in [32,256]"); }
    ctx = c;
#define CKC(call) do { cudaError_t e2_ = (call); if (e2_ != cudaSuccess) { fail(nullptr, GM_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e2_)); gm_destroy(c); return GM_ERR_CUDA; } } while (0)
    if (share_with) c->st = share_with->st;
    else CKC(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;  // the weight kernels are one CTA each: let them in ahead of the registration's grid
        CKC(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CKC(cudaStreamCreateWithPriority(&c->st_w, cudaStreamNonBlocking, hi));
    }
    if (const char* e = getenv("GM_B200_SM_WIDE_BELOW")) c->sm_wide_below = (uint32_t)std::max(0, atoi(e));
    CKC(cudaEventCreate(&c->ev0)); CKC(cudaEventCreate(&c->ev1)); CKC(cudaEventCreate(&c->evw0)); CKC(cudaEventCreate(&c->evw1));
    CKC(cudaEventCreate(&c->evr0)); CKC(cudaEventCreate(&c->evr1)); CKC(cudaEventCreateWithFlags(&c->ev_sm, cudaEventDisableTiming));
    size_t cap = cfg->pool_patches;
    if (share_with) {
        c->pool = share_with->pool;
        c->refcnt_new = share_with->refcnt_new;
        cap = (size_t)c->pool.cap;
    } else {
        // 0 = half of the free HBM; an explicit request is honoured up to 80 % of it (a larger one would only fail in cudaMalloc)
        size_t free_b = 0, total_b = 0;
        CKC(cudaMemGetInfo(&free_b, &total_b));
        // hit-capable patches own a slice of every plane (36 KB), free-space patches only their visit counters (4 KB); the
        // share of hit-capable ids is GM_B200_HIT_FRACTION (default 0.6: about half of the patches of an indoor map hold end
        // points, and a hit-capable id can stand in for a free-space one but not the other way round)
        double hit_fraction = 0.6;
        if (const char* e = getenv("GM_B200_HIT_FRACTION")) hit_fraction = std::min(1.0, std::max(0.05, atof(e)));
        const double per_hit = kPatchCells * (sizeof(int4) + sizeof(double2)) + kPatch * sizeof(unsigned), per_any = kPatchCells * sizeof(int) + 4 * sizeof(int);
        const double per_patch = per_any + hit_fraction * per_hit;
        const size_t half = (size_t)((double)free_b * 0.5 / per_patch), most = (size_t)((double)free_b * 0.8 / per_patch);
        if (!cap) cap = std::min<size_t>(half, (size_t)1 << 22);  // 2^22 patches at most by default
        else if (cap > most) cap = most;
        if (cap < 64) cap = 64;
    if (cap > ((size_t)1 << 22) - 2) cap = ((size_t)1 << 22) - 2;  // cell indexes (patch * 1024 + cell) stay below 2^32
    size_t cap_h = std::min(cap, std::max<size_t>(32, (size_t)std::ceil((double)cap * hit_fraction)));
    c->pool.cap = (int)cap;
    c->pool.cap_h = (int)cap_h;
    CKC(dalloc(c->pool.cells, (cap_h + 1) * kPatchCells));  // + the all-zero patch at index cap_h (unknown cells, free-space patches)
    CKC(cudaMemsetAsync(c->pool.cells + cap_h * kPatchCells, 0, kPatchCells * sizeof(int4), c->st));
    CKC(dalloc(c->pool.visits, (cap + 1) * kPatchCells));
    CKC(cudaMemsetAsync(c->pool.visits + cap * kPatchCells, 0, kPatchCells * sizeof(int), c->st));
    CKC(dalloc(c->pool.means, (cap_h + 1) * kPatchCells));
    CKC(cudaMemsetAsync(c->pool.means + cap_h * kPatchCells, 0, kPatchCells * sizeof(double2), c->st));
    CKC(dalloc(c->pool.occ, (cap_h + 1) * kPatch));
    CKC(cudaMemsetAsync(c->pool.occ, 0, (cap_h + 1) * kPatch * sizeof(unsigned), c->st));
    CKC(dalloc(c->pool.refcnt, cap)); CKC(dalloc(c->pool.free_list, cap)); CKC(dalloc(c->pool.pending, cap));
    CKC(dalloc(c->pool.free_top, 2)); CKC(dalloc(c->pool.pending_n, 1));
    CKC(dalloc(c->refcnt_new, cap));
    CKC(cudaMemsetAsync(c->refcnt_new, 0, cap * sizeof(int), c->st));
    }
    const uint32_t N = c->max_particles;
    c->weights_cap = std::max(cfg->max_particles, cfg->global_particles);
    CKC(dalloc(c->geom, N)); CKC(dalloc(c->geom_alt, N)); CKC(dalloc(c->geom_new, N)); CKC(dalloc(c->shift, N));
    CKC(dalloc(c->d_ranges, GM_MAXBEAMS)); CKC(dalloc(c->d_angles, GM_MAXBEAMS)); CKC(dalloc(c->d_poses, (size_t)N * 3));
    CKC(dalloc(c->d_score, N)); CKC(dalloc(c->d_s, N)); CKC(dalloc(c->d_l, N));
    CKC(dalloc(c->d_c, N)); CKC(dalloc(c->d_evals, N)); CKC(dalloc(c->d_idx, N)); CKC(dalloc(c->d_emitted, 1));
    CKC(dalloc(c->d_counters, 1));
    CKC(dalloc(c->d_counters_sm, 1));
    CKC(dalloc(c->d_rows, (size_t)c->weights_cap * 5));
    CKC(dalloc(c->d_done, 1));
    CKC(cudaMemsetAsync(c->d_done, 0, sizeof(unsigned), c->st));
    // pinned staging: ranges + local poses + all log-weights in; rows + weights + Neff + counters out
    c->h_in_bytes = sizeof(double) * (GM_MAXBEAMS + (size_t)N * 3 + c->weights_cap);
    c->h_out_bytes = sizeof(double) * ((size_t)c->weights_cap * 6 + 2) + sizeof(Counters);
    CKC(cudaHostAlloc(reinterpret_cast<void**>(&c->h_in), c->h_in_bytes, cudaHostAllocDefault));
    CKC(cudaHostAlloc(reinterpret_cast<void**>(&c->h_out), c->h_out_bytes, cudaHostAllocDefault));
    CKC(dalloc(c->d_digest, (size_t)N * 7));
    CKC(dalloc(c->d_count, 1));
    {
        // 1./n exactly as the host FPU divides (IEEE): PointAccumulator::mean() = (1./n) * acc, smmap.h:24
        // followed by 2^(j/128) for exp_neg()
        std::vector<double> inv(kInvTable + kExpTable);
        for (int i = 0; i < kInvTable; i++) { volatile double one = 1., den = (double)i; inv[i] = one / den; }
        for (int j = 0; j < kExpTable; j++) inv[kInvTable + j] = exp2((double)j / kExpTable);
        CKC(dalloc(c->d_inv_n, inv.size()));
        CKC(cudaMemcpy(c->d_inv_n, inv.data(), sizeof(double) * inv.size(), cudaMemcpyHostToDevice));
    }
    if (!share_with) launch_pool_init(c->pool, c->st);
    CKC(cudaGetLastError());
    CKC(cudaStreamSynchronize(c->st));
#undef CKC
    c->stats.pool_capacity = cap;
    c->stats.pool_hit_capacity = (uint64_t)c->pool.cap_h;
    c->stats.launches = 1;
    *out = c;
    return GM_OK;
}

int gm_create(gm_ctx** out, const gm_config* cfg) { return create_ctx(out, cfg, nullptr); }

int gm_destroy(gm_ctx* c) {
    if (!c) return GM_OK;
    cudaSetDevice(c->device);
    if (c->st) cudaStreamSynchronize(c->st);
    c->reg_pending = false;
    gmi_dist_destroy(c);
    if (c->st) cudaStreamSynchronize(c->st);
    for (void* p : c->retired) cudaFree(p);
    c->retired.clear();
    bool last = true;
    if (c->share) {
        // a context that shares its pool gives its references back; the pool and the stream go with the last user
        if (c->inited && c->dir && c->share->users > 1) {
            launch_release_slots(c->geom, c->dir, c->dir_cap, 0, c->n, c->pool, c->st);
            launch_merge_free(c->pool, c->st);
            cudaStreamSynchronize(c->st);
        }
        std::vector<gm_ctx*>& m = c->share->members;
        m.erase(std::remove(m.begin(), m.end(), c), m.end());
        last = --c->share->users == 0;
        if (last) delete c->share;
        c->share = nullptr;
    }
    if (last) {
        cudaFree(c->pool.cells); cudaFree(c->pool.visits); cudaFree(c->pool.means); cudaFree(c->pool.occ); cudaFree(c->pool.refcnt); cudaFree(c->pool.free_list); cudaFree(c->pool.pending);
        cudaFree(c->pool.free_top); cudaFree(c->pool.pending_n); cudaFree(c->refcnt_new);
    }
    cudaFree(c->dir); cudaFree(c->dir_alt); cudaFree(c->geom); cudaFree(c->geom_alt); cudaFree(c->geom_new); cudaFree(c->shift);
    cudaFree(c->d_ranges); cudaFree(c->d_angles); cudaFree(c->d_poses); cudaFree(c->d_score); cudaFree(c->d_s); cudaFree(c->d_l);
    cudaFree(c->d_c); cudaFree(c->d_evals); cudaFree(c->d_idx); cudaFree(c->d_emitted);
    cudaFree(c->d_w); cudaFree(c->d_widx); cudaFree(c->d_logw); cudaFree(c->d_cum); cudaFree(c->d_tgt); cudaFree(c->d_neff);
    cudaFree(c->d_qidx); cudaFree(c->d_qposes); cudaFree(c->d_qposes_out); cudaFree(c->d_qs); cudaFree(c->d_ql); cudaFree(c->d_qc);
    cudaFree(c->d_rows); cudaFree(c->d_done); cudaFree(c->d_counters_sm);
    if (c->h_in) cudaFreeHost(c->h_in);
    if (c->h_out) cudaFreeHost(c->h_out);
    cudaFree(c->d_occ); cudaFree(c->d_counters); cudaFree(c->d_inv_n); cudaFree(c->d_digest); cudaFree(c->d_list); cudaFree(c->d_count); cudaFree(c->d_export);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->evr0) cudaEventDestroy(c->evr0);
    if (c->evr1) cudaEventDestroy(c->evr1);
    if (c->ev_sm) cudaEventDestroy(c->ev_sm);
    if (c->st && last) cudaStreamDestroy(c->st);
    if (c->st_w) cudaStreamDestroy(c->st_w);
    if (c->evw0) cudaEventDestroy(c->evw0);
    if (c->evw1) cudaEventDestroy(c->evw1);
    delete c;
    return GM_OK;
}

int gm_set_laser(gm_ctx* ctx, uint32_t nbeams, const double* angles, const double laser_pose[3]) {
    if (!ctx) return GM_ERR_ARG;
    if (!angles || !laser_pose) return fail(ctx, GM_ERR_ARG, "gm_set_laser: null argument");
    if (nbeams == 0 || nbeams >= GM_MAXBEAMS)  // the reference asserts beams < LASER_MAXBEAMS (scanmatcher.cpp:704)
        return fail(ctx, GM_ERR_ARG, "gm_set_laser: %u beams (must be in [1, %d))", nbeams, GM_MAXBEAMS);
    if (int r = bind(ctx)) return r;
    ctx->P.nbeams = nbeams;
    memcpy(ctx->angles_h, angles, sizeof(double) * nbeams);
    for (int i = 0; i < 3; i++) ctx->P.laser_pose[i] = laser_pose[i];
    CK(cudaMemcpyAsync(ctx->d_angles, angles, sizeof(double) * nbeams, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->laser_set = true;
    return GM_OK;
}

int gm_set_matching(gm_ctx* ctx, const gm_match_params* p) {
    if (!ctx) return GM_ERR_ARG;
    if (!p) return fail(ctx, GM_ERR_ARG, "gm_set_matching: null argument");
    if (p->kernel_size < 0 || p->kernel_size > 8) return fail(ctx, GM_ERR_ARG, "gm_set_matching: kernel_size %d outside [0,8]", p->kernel_size);
    DevParams& P = ctx->P;
    P.laser_max_range = p->laser_max_range; P.usable_range = p->usable_range;
    P.gaussian_sigma = p->gaussian_sigma; P.likelihood_sigma = p->likelihood_sigma;
    P.kernel_size = p->kernel_size; P.generate_map = p->generate_map;
    P.opt_linear_delta = p->opt_linear_delta; P.opt_angular_delta = p->opt_angular_delta;
    P.opt_iters = p->opt_recursive_iterations; P.lskip = p->likelihood_skip;
    P.enlarge_step = p->enlarge_step; P.fullness_threshold = p->fullness_threshold;
    P.ang_rel = p->angular_odometry_reliability; P.lin_rel = p->linear_odometry_reliability;
    P.free_cell_ratio = p->free_cell_ratio; P.beam_skip = p->initial_beams_skip;
    P.thr_is_tenth = (p->fullness_threshold == 0.1) ? 1 : 0;
    ctx->match_set = true;
    return GM_OK;
}

int gm_init_particles(gm_ctx* ctx, uint32_t n, const double center[2], double world_w, double world_h, double delta) {
    if (!ctx) return GM_ERR_ARG;
    if (!center || !n || n > ctx->max_particles || !(delta > 0)) return fail(ctx, GM_ERR_ARG, "gm_init_particles: bad argument (n=%u, max=%u)", n, ctx->max_particles);
    if (int r = bind(ctx)) return r;
    if (ctx->share && ctx->share->users > 1) return fail(ctx, GM_ERR_ARG, "gm_init_particles: this context shares its patch pool with %d clone(s)", ctx->share->users - 1);
    // Map(center, worldSizeX, worldSizeY, delta): map.h:171-184; HierarchicalArray2D(xsize>>5, ysize>>5): harray2d.h:54-59
    const int xs = (int)ceil(world_w / delta), ys = (int)ceil(world_h / delta);
    Geom g;
    g.npx = xs >> kPatchMag; g.npy = ys >> kPatchMag;
    if (g.npx <= 0 || g.npy <= 0) return fail(ctx, GM_ERR_ARG, "gm_init_particles: map smaller than one patch");
    g.sx2 = (g.npx << kPatchMag) >> 1; g.sy2 = (g.npy << kPatchMag) >> 1;
    ctx->P.cx = center[0]; ctx->P.cy = center[1]; ctx->P.delta = delta;
    ctx->n = n;
    ctx->dir_cap = g.npx * g.npy;
    const size_t dn = (size_t)ctx->max_particles * ctx->dir_cap;
    CK(dalloc(ctx->dir, dn)); CK(dalloc(ctx->dir_alt, dn));
    launch_pool_init(ctx->pool, ctx->st);
    launch_fill_int(ctx->dir, dn, -1, ctx->st);
    launch_fill_int(ctx->dir_alt, dn, -1, ctx->st);
    launch_fill_geom(ctx->geom, ctx->max_particles, g, ctx->st);
    ctx->stats.launches += 4;
    if (int r = check_launch(ctx, "init")) return r;
    CK(cudaMemsetAsync(ctx->refcnt_new, 0, (size_t)ctx->pool.cap * sizeof(int), ctx->st));
    CK(dalloc(ctx->d_w, ctx->weights_cap)); CK(dalloc(ctx->d_logw, ctx->weights_cap));
    CK(dalloc(ctx->d_cum, ctx->weights_cap)); CK(dalloc(ctx->d_tgt, ctx->weights_cap)); CK(dalloc(ctx->d_neff, 1));
    CK(dalloc(ctx->d_widx, ctx->weights_cap));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->stats.pool_in_use = 0;
    ctx->inited = true;
    return GM_OK;
}

uint32_t gm_particles(const gm_ctx* ctx) { return ctx ? ctx->n : 0; }

static int run_eval(gm_ctx* ctx, int mode, uint32_t count, const uint32_t* particle_idx, const double* poses,
                    const double* ranges, double* s_out, double* l_out, uint32_t* c_out, double* poses_out) {
    if (int r = ready(ctx, true)) return r;
    if (!count) return GM_OK;
    if (!particle_idx || !poses || !ranges) return fail(ctx, GM_ERR_ARG, "eval: null argument");
    for (uint32_t i = 0; i < count; i++)
        if (particle_idx[i] >= ctx->n) return fail(ctx, GM_ERR_ARG, "eval: particle index %u >= %u", particle_idx[i], ctx->n);
    if (int r = ensure_query(ctx, count)) return r;
    CK(cudaMemcpyAsync(ctx->d_ranges, ranges, sizeof(double) * ctx->P.nbeams, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->d_qidx, particle_idx, sizeof(uint32_t) * count, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->d_qposes, poses, sizeof(double) * 3 * count, cudaMemcpyHostToDevice, ctx->st));
    ScanMatchArgs a{};
    a.P = ctx->P; a.ranges = ctx->d_ranges; a.angles = ctx->d_angles; a.poses = ctx->d_qposes;
    a.geom = ctx->geom; a.dir = ctx->dir; a.dir_cap = ctx->dir_cap; a.pool = ctx->pool.cells; a.visits = ctx->pool.visits; a.means = ctx->pool.means; a.occ = ctx->pool.occ; a.inv_n = ctx->d_inv_n;
    a.s = ctx->d_qs; a.l = ctx->d_ql; a.c = ctx->d_qc; a.score = ctx->d_qs; a.poses_out = ctx->d_qposes_out;
    a.query = mode; a.particle_idx = ctx->d_qidx; a.zero_patch = ctx->pool.cap_h;
    derive_params(a.P);
    if (int r = prepare_scanmatch(ctx)) return r;
    launch_scanmatch(a, count, false, ctx->st);
    ctx->stats.launches++;
    if (int r = check_launch(ctx, "k_scanmatch (query)")) return r;
    if (s_out) CK(cudaMemcpyAsync(s_out, ctx->d_qs, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->st));
    if (mode == 1 && l_out) CK(cudaMemcpyAsync(l_out, ctx->d_ql, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->st));
    if (mode == 1 && c_out) CK(cudaMemcpyAsync(c_out, ctx->d_qc, sizeof(uint32_t) * count, cudaMemcpyDeviceToHost, ctx->st));
    if (mode == 2 && poses_out) CK(cudaMemcpyAsync(poses_out, ctx->d_qposes_out, sizeof(double) * 3 * count, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return GM_OK;
}

int gm_score(gm_ctx* ctx, uint32_t count, const uint32_t* particle_idx, const double* poses, const double* ranges, double* score_out) {
    return run_eval(ctx, 0, count, particle_idx, poses, ranges, score_out, nullptr, nullptr, nullptr);
}

int gm_likelihood_and_score(gm_ctx* ctx, uint32_t count, const uint32_t* particle_idx, const double* poses, const double* ranges,
                            double* s_out, double* l_out, uint32_t* c_out) {
    return run_eval(ctx, 1, count, particle_idx, poses, ranges, s_out, l_out, c_out, nullptr);
}

int gm_optimize(gm_ctx* ctx, uint32_t count, const uint32_t* particle_idx, const double* poses_in, const double* ranges,
                double* poses_out, double* score_out) {
    return run_eval(ctx, 2, count, particle_idx, poses_in, ranges, score_out, nullptr, nullptr, poses_out);
}

int gm_scanmatch(gm_ctx* ctx, const double* ranges, double* poses_inout, double minimum_score, double* score_out, double* s_out,
                 double* l_out, uint32_t* c_out, uint32_t* evals_out) {
    if (int r = ready(ctx, true)) return r;
    if (!ranges || !poses_inout) return fail(ctx, GM_ERR_ARG, "gm_scanmatch: null argument");
    const uint32_t n = ctx->n;
    CK(cudaMemcpyAsync(ctx->d_ranges, ranges, sizeof(double) * ctx->P.nbeams, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->d_poses, poses_inout, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->st));
    if (int r = zero_counters(ctx)) return r;
    ScanMatchArgs a{};
    a.query = -1;
    a.P = ctx->P; a.ranges = ctx->d_ranges; a.angles = ctx->d_angles; a.poses = ctx->d_poses;
    a.geom = ctx->geom; a.dir = ctx->dir; a.dir_cap = ctx->dir_cap; a.pool = ctx->pool.cells; a.visits = ctx->pool.visits; a.means = ctx->pool.means; a.occ = ctx->pool.occ; a.inv_n = ctx->d_inv_n;
    a.minimum_score = minimum_score;
    a.score = ctx->d_score; a.s = ctx->d_s; a.l = ctx->d_l; a.c = ctx->d_c; a.evals = ctx->d_evals;
    a.counters = ctx->d_counters;
    a.zero_patch = ctx->pool.cap_h;
    derive_params(a.P);
    a.valid_score_beams = count_score_beams(ctx, ranges);
    if (int r = prepare_scanmatch(ctx)) return r;
    CK(cudaEventRecord(ctx->ev0, ctx->st));
    launch_scanmatch(a, n, n <= ctx->sm_wide_below, ctx->st);
    CK(cudaEventRecord(ctx->ev1, ctx->st));
    ctx->stats.launches++;
    if (int r = check_launch(ctx, "k_scanmatch")) return r;
    CK(cudaMemcpyAsync(poses_inout, ctx->d_poses, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->st));
    if (score_out) CK(cudaMemcpyAsync(score_out, ctx->d_score, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->st));
    if (s_out) CK(cudaMemcpyAsync(s_out, ctx->d_s, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->st));
    if (l_out) CK(cudaMemcpyAsync(l_out, ctx->d_l, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->st));
    if (c_out) CK(cudaMemcpyAsync(c_out, ctx->d_c, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->st));
    if (evals_out) CK(cudaMemcpyAsync(evals_out, ctx->d_evals, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->st));
    Counters h;
    if (int r = read_counters(ctx, h)) return r;
    ctx->stats.score_evals = h.score_evals; ctx->stats.beam_evals = h.beam_evals;
    CK(cudaEventElapsedTime(&ctx->stats.ms_scanmatch, ctx->ev0, ctx->ev1));
    ctx->stats.total_ms[0] += ctx->stats.ms_scanmatch; ctx->stats.total_beam_evals += h.beam_evals; ctx->stats.total_score_evals += h.score_evals;
    ctx->stats.total_scanmatches++;
    return GM_OK;
}

// scanMatch + weight += l + normalize as one stream-ordered sequence: one upload (pinned), k_scanmatch, [sharded: the rows of
// every rank land in every rank's HBM from the kernel's epilogue], k_normalize, one download, ONE synchronisation.
int gm_scanmatch_weights(gm_ctx* ctx, const double* ranges, const double* poses_local, double minimum_score, const double* log_weights,
                         double obs_sigma_gain, double* rows_out, double* weights_out, double* neff_out) {
    if (int r = gm_scanmatch_weights_begin(ctx, ranges, poses_local, minimum_score, log_weights, obs_sigma_gain)) return r;
    return gm_scanmatch_weights_end(ctx, rows_out, weights_out, neff_out);
}

// the launch half: uploads, k_scanmatch (and, fused, the registration behind it) are queued; the caller's host work overlaps them
int gm_scanmatch_weights_begin(gm_ctx* ctx, const double* ranges, const double* poses_local, double minimum_score, const double* log_weights,
                               double obs_sigma_gain) {
    if (int r = ready(ctx, true)) return r;
    if (ctx->sw_pending) return fail(ctx, GM_ERR_ARG, "gm_scanmatch_weights_begin: the previous one has not been collected (gm_scanmatch_weights_end)");
    if (!ranges || !poses_local || !log_weights) return fail(ctx, GM_ERR_ARG, "gm_scanmatch_weights: null argument");
    const uint32_t n = ctx->n, nb = ctx->P.nbeams;
    uint32_t N = n, first = 0;
    ScanMatchArgs a{};
    a.query = -1;
    a.nranks = 1; a.rows[0] = ctx->d_rows; a.first_global = 0; a.done_counter = ctx->d_done; a.flag = nullptr; a.seq = 0;
    if (ctx->dist) {
        if (int r = gmi_dist_rows(ctx, &a)) return r;
        N = n * (uint32_t)a.nranks; first = a.first_global;
    }
    if (N > ctx->weights_cap) return fail(ctx, GM_ERR_ARG, "gm_scanmatch_weights: %u particles over all ranks, the context was created for %u", N, ctx->weights_cap);
    double* hin = reinterpret_cast<double*>(ctx->h_in);
    memcpy(hin, ranges, sizeof(double) * nb);
    memcpy(hin + nb, poses_local, sizeof(double) * 3 * n);
    memcpy(hin + nb + 3 * (size_t)n, log_weights, sizeof(double) * N);
    CK(cudaMemcpyAsync(ctx->d_ranges, hin, sizeof(double) * nb, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->d_poses, hin + nb, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->d_logw, hin + nb + 3 * (size_t)n, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemsetAsync(ctx->d_counters_sm, 0, sizeof(Counters), ctx->st));
    a.P = ctx->P; a.ranges = ctx->d_ranges; a.angles = ctx->d_angles; a.poses = ctx->d_poses;
    a.geom = ctx->geom; a.dir = ctx->dir; a.dir_cap = ctx->dir_cap; a.pool = ctx->pool.cells; a.visits = ctx->pool.visits; a.means = ctx->pool.means; a.occ = ctx->pool.occ; a.inv_n = ctx->d_inv_n;
    a.minimum_score = minimum_score;
    a.score = ctx->d_score; a.s = ctx->d_s; a.l = ctx->d_l; a.c = ctx->d_c; a.evals = ctx->d_evals;
    a.counters = ctx->d_counters_sm;
    a.zero_patch = ctx->pool.cap_h;
    derive_params(a.P);
    a.valid_score_beams = count_score_beams(ctx, ranges);
    if (int r = prepare_scanmatch(ctx)) return r;
    CK(cudaEventRecord(ctx->ev0, ctx->st));
    launch_scanmatch(a, n, n <= ctx->sm_wide_below, ctx->st);
    CK(cudaEventRecord(ctx->ev1, ctx->st));
    ctx->stats.launches++;
    if (int r = check_launch(ctx, "k_scanmatch")) return r;
    // the weights take the other stream from here: with gm_set_fused_register the registration of this scan is queued right
    // behind the scan match (the matched poses and the scan are on the device already) and runs while this rank waits for the
    // other ranks' rows, normalises and hands the results to the host
    cudaStream_t sw = ctx->st_w;
    CK(cudaEventRecord(ctx->ev_sm, ctx->st));
    CK(cudaStreamWaitEvent(sw, ctx->ev_sm, 0));
    if (ctx->fused_register) {
        if (int r = zero_counters(ctx)) return r;
        ctx->stats.ms_remap = 0.f; ctx->stats.ms_migrate = 0.f; ctx->stats.migrated_particles = 0; ctx->stats.migrated_bytes = 0;
        CK(cudaEventRecord(ctx->evr0, ctx->st));
        if (int r = gmi_bounds_and_resize(ctx)) return r;
        const bool was_async = ctx->async_register;
        ctx->async_register = true;  // collected by the next entry point, whatever the mode: the point is not to wait here
        const int r = gmi_register_update(ctx);
        ctx->async_register = was_async;
        if (r) return r;
    }
    ctx->sw_pending = true; ctx->sw_total = N; ctx->sw_gain = obs_sigma_gain;
    (void)first;
    return GM_OK;
}

// the collecting half: [sharded: wait for every rank's rows,] normalise on the second stream, one download, one synchronisation
int gm_scanmatch_weights_end(gm_ctx* ctx, double* rows_out, double* weights_out, double* neff_out) {
    if (!ctx) return GM_ERR_ARG;
    if (int r = bind(ctx)) return r;
    if (!ctx->sw_pending) return fail(ctx, GM_ERR_ARG, "gm_scanmatch_weights_end: nothing to collect (gm_scanmatch_weights_begin)");
    if (!rows_out) return fail(ctx, GM_ERR_ARG, "gm_scanmatch_weights_end: null argument");
    ctx->sw_pending = false;
    const uint32_t N = ctx->sw_total;
    const double obs_sigma_gain = ctx->sw_gain;
    cudaStream_t sw = ctx->st_w;
    const double* rows = ctx->d_rows;
    if (ctx->dist)
        if (int r = gmi_dist_wait_rows(ctx, &rows)) return r;  // every rank's slice has landed in this rank's gather buffer
    CK(cudaEventRecord(ctx->evw0, sw));
    launch_normalize(ctx->d_logw, rows, N, obs_sigma_gain, ctx->d_w, ctx->d_neff, sw);
    CK(cudaEventRecord(ctx->evw1, sw));
    ctx->stats.launches++;
    if (int r = check_launch(ctx, "k_normalize")) return r;
    double* hout = reinterpret_cast<double*>(ctx->h_out);
    CK(cudaMemcpyAsync(hout, rows, sizeof(double) * 5 * N, cudaMemcpyDeviceToHost, sw));
    CK(cudaMemcpyAsync(hout + 5 * (size_t)N, ctx->d_w, sizeof(double) * N, cudaMemcpyDeviceToHost, sw));
    CK(cudaMemcpyAsync(hout + 6 * (size_t)N, ctx->d_neff, sizeof(double), cudaMemcpyDeviceToHost, sw));
    CK(cudaMemcpyAsync(hout + 6 * (size_t)N + 2, ctx->d_counters_sm, sizeof(Counters), cudaMemcpyDeviceToHost, sw));
    CK(cudaStreamSynchronize(sw));
    memcpy(rows_out, hout, sizeof(double) * 5 * N);
    if (weights_out) memcpy(weights_out, hout + 5 * (size_t)N, sizeof(double) * N);
    if (neff_out) *neff_out = hout[6 * (size_t)N];
    Counters h;
    memcpy(&h, hout + 6 * (size_t)N + 2, sizeof h);
    ctx->stats.score_evals = h.score_evals; ctx->stats.beam_evals = h.beam_evals;
    CK(cudaEventElapsedTime(&ctx->stats.ms_scanmatch, ctx->ev0, ctx->ev1));
    CK(cudaEventElapsedTime(&ctx->stats.ms_weights, ctx->evw0, ctx->evw1));
    ctx->stats.total_ms[0] += ctx->stats.ms_scanmatch; ctx->stats.total_ms[1] += ctx->stats.ms_weights;
    ctx->stats.total_beam_evals += h.beam_evals; ctx->stats.total_score_evals += h.score_evals;
    ctx->stats.total_scanmatches++;
    return GM_OK;
}

int gm_normalize_neff(gm_ctx* ctx, const double* log_weights, uint32_t n, double obs_sigma_gain, double* weights_out, double* neff_out) {
    if (int r = ready_weights(ctx)) return r;
    if (!ctx->inited) return fail(ctx, GM_ERR_ARG, "gm_init_particles has not been called");
    if (!log_weights || !n || n > ctx->weights_cap) return fail(ctx, GM_ERR_ARG, "gm_normalize_neff: bad argument");
    cudaStream_t sw = ctx->st_w;  // not ordered after a registration in flight on ctx->st
    CK(cudaMemcpyAsync(ctx->d_logw, log_weights, sizeof(double) * n, cudaMemcpyHostToDevice, sw));
    CK(cudaEventRecord(ctx->evw0, sw));
    launch_normalize(ctx->d_logw, nullptr, n, obs_sigma_gain, ctx->d_w, ctx->d_neff, sw);
    CK(cudaEventRecord(ctx->evw1, sw));
    ctx->stats.launches++;
    if (int r = check_launch(ctx, "k_normalize")) return r;
    if (weights_out) CK(cudaMemcpyAsync(weights_out, ctx->d_w, sizeof(double) * n, cudaMemcpyDeviceToHost, sw));
    if (neff_out) CK(cudaMemcpyAsync(neff_out, ctx->d_neff, sizeof(double), cudaMemcpyDeviceToHost, sw));
    CK(cudaStreamSynchronize(sw));
    CK(cudaEventElapsedTime(&ctx->stats.ms_weights, ctx->evw0, ctx->evw1));
    ctx->stats.total_ms[1] += ctx->stats.ms_weights;
    return GM_OK;
}

int gm_resample_indexes(gm_ctx* ctx, const double* weights, uint32_t n, double u01, uint32_t* idx_out, uint32_t* emitted_out) {
    return gm_resample_indexes_n(ctx, weights, n, n, u01, idx_out, emitted_out);
}

int gm_resample_indexes_n(gm_ctx* ctx, const double* weights, uint32_t n, uint32_t n_out, double u01, uint32_t* idx_out, uint32_t* emitted_out) {
    if (int r = ready_weights(ctx)) return r;
    if (!ctx->inited) return fail(ctx, GM_ERR_ARG, "gm_init_particles has not been called");
    if (!weights || !idx_out || !n || n > ctx->weights_cap || !n_out || n_out > ctx->weights_cap)
        return fail(ctx, GM_ERR_ARG, "gm_resample_indexes: bad argument (n=%u, n_out=%u, capacity %u)", n, n_out, ctx->weights_cap);
    cudaStream_t sw = ctx->st_w;
    CK(cudaMemcpyAsync(ctx->d_w, weights, sizeof(double) * n, cudaMemcpyHostToDevice, sw));
    CK(cudaMemsetAsync(ctx->d_emitted, 0, sizeof(unsigned), sw));
    launch_resample(ctx->d_w, n, n_out, u01, ctx->d_cum, ctx->d_tgt, ctx->d_widx, ctx->d_emitted, sw);
    ctx->stats.launches++;
    if (int r = check_launch(ctx, "k_resample")) return r;
    CK(cudaMemcpyAsync(idx_out, ctx->d_widx, sizeof(uint32_t) * n_out, cudaMemcpyDeviceToHost, sw));
    uint32_t em = 0;
    CK(cudaMemcpyAsync(&em, ctx->d_emitted, sizeof(uint32_t), cudaMemcpyDeviceToHost, sw));
    CK(cudaStreamSynchronize(sw));
    if (emitted_out) *emitted_out = em;
    return GM_OK;
}

// computeActiveArea's observable part for all particles: bounding box of the scan at the pose in d_poses and the
// Map::resize it triggers (scanmatcher.cpp:96-166, map.h:218-241, harray2d.h:80-109).  d_ranges / d_poses are staged.
// Nothing here waits for the device: k_relayout* look at k_bounds' counters themselves (no-ops unless a particle left its
// map), and a resize that needs a longer directory stride than the allocation has makes them -- and k_register -- stand
// down; gmi_register_finish then grows the directories and runs the sequence again (gmi_grow_and_redo).
static void launch_bounds_and_relayout(gm_ctx* ctx, int new_cap, int* dir_out) {
    const uint32_t n = ctx->n;
    BoundsArgs b;
    b.P = ctx->P; b.ranges = ctx->d_ranges; b.angles = ctx->d_angles; b.poses = ctx->d_poses;
    b.geom = ctx->geom; b.geom_new = ctx->geom_new; b.shift = ctx->shift; b.counters = ctx->d_counters;
    launch_bounds(b, n, ctx->threads, ctx->st);
    RelayoutArgs r;
    r.counters = ctx->d_counters;
    r.geom = ctx->geom; r.geom_new = ctx->geom_new; r.shift = ctx->shift;
    r.dir = ctx->dir; r.dir_cap_old = ctx->dir_cap; r.pool = ctx->pool;
    r.dir_alt = ctx->dir_alt; r.dir_cap_new = new_cap; r.dir_out = dir_out ? dir_out : ctx->dir;
    launch_relayout(r, n, ctx->st);
    launch_relayout_commit(r, n, ctx->st);
    launch_merge_free(ctx->pool, ctx->st);
    ctx->stats.launches += 4;
}

int gmi_bounds_and_resize(gm_ctx* ctx) {
    launch_bounds_and_relayout(ctx, ctx->dir_cap, nullptr);
    return check_launch(ctx, "k_bounds / k_relayout");
}

// a Map::resize outgrew the directory stride: allocate longer directories for everybody, then bounds + relayout again into
// them (the first attempt changed nothing: every kernel after k_bounds stood down)
static int gmi_grow_dirs_and_relayout(gm_ctx* ctx, int new_cap) {
    const size_t dn = (size_t)ctx->max_particles * new_cap;
    int* grown = nullptr;
    release_exported(ctx, ctx->dir_alt); ctx->dir_alt = nullptr;
    CK(dalloc(ctx->dir_alt, dn));
    CK(cudaMalloc(reinterpret_cast<void**>(&grown), dn * sizeof(int)));
    launch_fill_int(grown, dn, -1, ctx->st);
    ctx->stats.launches++;
    if (int rc = zero_counters(ctx)) { cudaFree(grown); return rc; }
    launch_bounds_and_relayout(ctx, new_cap, grown);
    if (int rc = check_launch(ctx, "relayout")) { cudaFree(grown); return rc; }
    CK(cudaStreamSynchronize(ctx->st));
    release_exported(ctx, ctx->dir);
    ctx->dir = grown;
    ctx->dir_cap = new_cap;
    ctx->dir_gen++;
    launch_fill_int(ctx->dir_alt, dn, -1, ctx->st);
    ctx->stats.launches++;
    return GM_OK;
}

// ---- the phases of gm_register, shared with the multi-GPU path (gm_dist.cu)

// stage the scan and the poses of the ctx->n local particles
int gmi_stage(gm_ctx* ctx, const double* ranges, const double* poses) {
    CK(cudaMemcpyAsync(ctx->d_ranges, ranges, sizeof(double) * ctx->P.nbeams, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->d_poses, poses, sizeof(double) * 3 * ctx->n, cudaMemcpyHostToDevice, ctx->st));
    return zero_counters(ctx);
}

// particle copy by patch reference (gridslamprocessor.hxx:184-213): slot j <- slot idx[j] for the ctx->n local
// particles; idx may name any slot below src_limit (imported parents sit above ctx->n)
int gmi_remap(gm_ctx* ctx, const uint32_t* idx, uint32_t src_limit) {
    const uint32_t n = ctx->n;
    for (uint32_t i = 0; i < n; i++)
        if (idx[i] >= src_limit) return fail(ctx, GM_ERR_ARG, "gm_register: parent index %u >= %u", idx[i], src_limit);
    CK(cudaMemcpyAsync(ctx->d_idx, idx, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaEventRecord(ctx->ev0, ctx->st));
    RemapArgs r;
    r.idx = ctx->d_idx; r.geom = ctx->geom; r.geom_alt = ctx->geom_alt; r.dir = ctx->dir; r.dir_alt = ctx->dir_alt;
    r.dir_cap = ctx->dir_cap; r.refcnt = ctx->pool.refcnt;
    launch_remap(r, n, ctx->st);  // the new generation takes its references first: a shared patch never drops to zero in between
    launch_release_slots(ctx->geom, ctx->dir, ctx->dir_cap, 0, src_limit, ctx->pool, ctx->st);  // the old one, imported parents included
    launch_merge_free(ctx->pool, ctx->st);
    CK(cudaEventRecord(ctx->ev1, ctx->st));
    ctx->stats.launches += 3;
    if (int rc = check_launch(ctx, "remap")) return rc;
    std::swap(ctx->dir, ctx->dir_alt);
    std::swap(ctx->geom, ctx->geom_alt);
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaEventElapsedTime(&ctx->stats.ms_remap, ctx->ev0, ctx->ev1));
    return GM_OK;
}

// active area + copy-on-write + ray/endpoint updates (scanmatcher.cpp:168-354) for the ctx->n local particles
int gmi_register_update(gm_ctx* ctx) {
    const uint32_t n = ctx->n;
    RegisterArgs a;
    a.P = ctx->P; a.ranges = ctx->d_ranges; a.angles = ctx->d_angles; a.poses = ctx->d_poses;
    a.geom = ctx->geom; a.dir = ctx->dir; a.dir_cap = ctx->dir_cap; a.pool = ctx->pool; a.counters = ctx->d_counters;
    a.bitmap_words = (ctx->dir_cap + 31) / 32;
    int sort_n = 64;  // power of two >= nbeams: sizes the hit table of k_register
    while (sort_n < (int)ctx->P.nbeams) sort_n <<= 1;
    a.sort_n = sort_n;
    // visit-count tiles: as many as fit (one pass over the rays when the active area has at most that many patches);
    // GM_B200_REGISTER_TILES caps them (tests of the multi-pass path)
    const size_t fixed = register_smem_bytes(a.bitmap_words, 0, (int)ctx->P.nbeams, 0);
    const size_t budget = 220 * 1024;  // of the 227 KB a CTA may use
    if (fixed + (size_t)a.sort_n * 36 > budget) return fail(ctx, GM_ERR_ARG, "gm_register: patch directory of %d entries needs %zu B of shared memory", ctx->dir_cap, fixed);
    int tile_cap = std::min(112, (int)((budget - fixed) / 2048));  // 112 = kMaxTiles of k_register
    if (const char* e = getenv("GM_B200_REGISTER_TILES")) tile_cap = std::max(1, std::min(tile_cap, atoi(e)));
    a.tile_cap = tile_cap;
    const size_t smem = register_smem_bytes(a.bitmap_words, a.sort_n, (int)ctx->P.nbeams, tile_cap);
    if (smem != ctx->register_smem) { CK(register_prepare(smem)); ctx->register_smem = smem; }
    launch_register(a, n, smem, ctx->st);
    launch_merge_free(ctx->pool, ctx->st);
    CK(cudaEventRecord(ctx->evr1, ctx->st));
    ctx->stats.launches += 2;
    if (int rc = check_launch(ctx, "k_register")) return rc;
    ctx->reg_pending = true;
    ctx->reg_tile_cap = tile_cap;
    // asynchronous registration: the kernel runs on while the caller does its host-side bookkeeping; counters, timing
    // and device-side errors are collected by the next entry point (or gm_synchronize)
    return ctx->async_register ? GM_OK : gmi_register_finish(ctx);
}

// collect what a launched registration left behind: counters, timing, pool level, device-side errors
int gmi_register_finish(gm_ctx* ctx) {
    if (!ctx->reg_pending) return GM_OK;
    ctx->reg_pending = false;
    Counters h;
    if (int rc = read_counters(ctx, h)) return rc;
    if (h.error == GM_ERR_LINE) return fail(ctx, h.error, "gm_register: a ray exceeds the reference's %d-point line buffer", kLineCap);
    if (!h.error && h.need_dir_cap > ctx->dir_cap) {
        // the rare slow path: the scan is still staged in d_ranges / d_poses, nothing but k_bounds has run
        if (int rc = gmi_grow_dirs_and_relayout(ctx, h.need_dir_cap)) return rc;
        const bool was_async = ctx->async_register;
        ctx->async_register = true;
        const int rc = gmi_register_update(ctx);
        ctx->async_register = was_async;
        if (rc) return rc;
        ctx->reg_pending = false;
        Counters h2;
        if (int rc2 = read_counters(ctx, h2)) return rc2;
        h2.resizes = h.resizes;
        h = h2;
    }
    int tops[2] = {0, 0};  // free hit-capable / free-space patch ids
    CK(cudaMemcpyAsync(tops, ctx->pool.free_top, sizeof tops, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    const int top = std::max(0, tops[0]) + std::max(0, tops[1]);
    ctx->stats.pool_hit_in_use = (uint64_t)(ctx->pool.cap_h - std::max(0, tops[0]));
    CK(cudaEventElapsedTime(&ctx->stats.ms_register, ctx->evr0, ctx->evr1));
    ctx->stats.total_ms[3] += ctx->stats.ms_register; ctx->stats.total_ms[2] += ctx->stats.ms_remap; ctx->stats.total_ms[4] += ctx->stats.ms_migrate;
    ctx->stats.total_registers++;
    ctx->stats.cow_copies = h.cow_copies; ctx->stats.patch_allocs = h.patch_allocs;
    ctx->stats.free_updates = h.free_updates; ctx->stats.hit_updates = h.hit_updates; ctx->stats.resizes = h.resizes;
    ctx->stats.pool_in_use = (uint64_t)(ctx->pool.cap - top);
    ctx->reg_active_hint = h.max_active;
    ctx->stats.max_active_patches = (uint32_t)h.max_active; ctx->stats.register_tile_cap = (uint32_t)ctx->reg_tile_cap;
    for (int k = 0; k < 8; k++) ctx->stats.register_phase_cycles[k] = h.reg_phase[k];
    if (h.error == GM_ERR_POOL)
        return fail(ctx, GM_ERR_POOL, "gm_register: patch pool exhausted (%d patch ids, %d of them able to hold end points: %d and %d free; "
                    "GM_B200_POOL_PATCHES / gm_config.pool_patches, GM_B200_HIT_FRACTION)", ctx->pool.cap, ctx->pool.cap_h, std::max(0, tops[0]) + std::max(0, tops[1]), std::max(0, tops[0]));
    if (h.error) return fail(ctx, h.error, "gm_register: device-side error %d", h.error);
    return GM_OK;
}

// make every directory slot at least `need` entries long (a Map::resize or an imported map that outgrows the stride)
int gmi_ensure_dir_cap(gm_ctx* ctx, int need) {
    if (need <= ctx->dir_cap) return GM_OK;
    const size_t dn = (size_t)ctx->max_particles * need;
    int* grown = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&grown), dn * sizeof(int)));
    launch_fill_int(grown, dn, -1, ctx->st);
    CK(cudaMemcpy2DAsync(grown, sizeof(int) * need, ctx->dir, sizeof(int) * ctx->dir_cap, sizeof(int) * ctx->dir_cap, ctx->max_particles,
                         cudaMemcpyDeviceToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    release_exported(ctx, ctx->dir);
    ctx->dir = grown;
    release_exported(ctx, ctx->dir_alt); ctx->dir_alt = nullptr;
    CK(dalloc(ctx->dir_alt, dn));
    launch_fill_int(ctx->dir_alt, dn, -1, ctx->st);
    ctx->dir_cap = need;
    ctx->dir_gen++;
    ctx->stats.launches += 2;
    return check_launch(ctx, "directory growth");
}

int gm_register(gm_ctx* ctx, const double* ranges, const double* poses, const uint32_t* idx) {
    if (int r = ready(ctx, true)) return r;
    if (!ranges || !poses) return fail(ctx, GM_ERR_ARG, "gm_register: null argument");
    if (int r = gmi_stage(ctx, ranges, poses)) return r;
    ctx->stats.ms_remap = 0.f;
    if (idx)
        if (int r = gmi_remap(ctx, idx, ctx->n)) return r;
    CK(cudaEventRecord(ctx->evr0, ctx->st));
    if (int rc = gmi_bounds_and_resize(ctx)) return rc;
    return gmi_register_update(ctx);
}

int gm_copy_particles(gm_ctx* ctx, const uint32_t* idx) {
    if (int r = ready(ctx, true)) return r;  // completes a registration in flight: the copies take the registered maps
    if (!idx) return fail(ctx, GM_ERR_ARG, "gm_copy_particles: null argument");
    if (ctx->dist) return fail(ctx, GM_ERR_ARG, "gm_copy_particles: sharded context, use gm_dist_copy_particles");
    if (int r = gmi_remap(ctx, idx, ctx->n)) return r;
    ctx->stats.total_ms[2] += ctx->stats.ms_remap;
    return GM_OK;
}

int gm_copy_particles_n(gm_ctx* ctx, const uint32_t* idx, uint32_t n_new) {
    if (int r = ready(ctx, true)) return r;
    if (!idx) return fail(ctx, GM_ERR_ARG, "gm_copy_particles_n: null argument");
    if (ctx->dist) return fail(ctx, GM_ERR_ARG, "gm_copy_particles_n: the particle count of a sharded context is fixed");
    if (!n_new || n_new > ctx->max_particles)
        return fail(ctx, GM_ERR_ARG, "gm_copy_particles_n: %u particles, the context was created for at most %u", n_new, ctx->max_particles);
    const uint32_t old = ctx->n;
    ctx->n = n_new;  // the new generation has n_new slots, its parents are among the old one's
    if (int r = gmi_remap(ctx, idx, old)) { ctx->n = old; return r; }
    ctx->stats.total_ms[2] += ctx->stats.ms_remap;
    return GM_OK;
}

int gm_compute_active_area(gm_ctx* ctx, const double* ranges, const double* poses) {
    if (int r = ready(ctx, true)) return r;
    if (!ranges || !poses) return fail(ctx, GM_ERR_ARG, "gm_compute_active_area: null argument");
    CK(cudaMemcpyAsync(ctx->d_ranges, ranges, sizeof(double) * ctx->P.nbeams, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->d_poses, poses, sizeof(double) * 3 * ctx->n, cudaMemcpyHostToDevice, ctx->st));
    if (int r = zero_counters(ctx)) return r;
    if (int r = gmi_bounds_and_resize(ctx)) return r;
    Counters h;
    if (int r = read_counters(ctx, h)) return r;
    if (h.error == GM_ERR_LINE) return fail(ctx, h.error, "gm_compute_active_area: a ray exceeds the reference's %d-point line buffer", kLineCap);
    if (h.need_dir_cap > ctx->dir_cap)
        if (int r = gmi_grow_dirs_and_relayout(ctx, h.need_dir_cap)) return r;
    CK(cudaStreamSynchronize(ctx->st));
    return GM_OK;
}

int gm_upload_map(gm_ctx* ctx, uint32_t particle, const gm_map_info* info, const int32_t* patch_coords, const gm_cell* cells,
                  uint32_t npatches) {
    if (int r = ready(ctx, true)) return r;
    if (particle >= ctx->n || !info || (npatches && (!patch_coords || !cells))) return fail(ctx, GM_ERR_ARG, "gm_upload_map: bad argument");
    Geom g;
    g.sx2 = info->size_x2; g.sy2 = info->size_y2; g.npx = info->patches_x; g.npy = info->patches_y;
    if (g.npx <= 0 || g.npy <= 0) return fail(ctx, GM_ERR_ARG, "gm_upload_map: empty directory");
    for (uint32_t k = 0; k < npatches; k++)
        if (patch_coords[2 * k] < 0 || patch_coords[2 * k] >= g.npx || patch_coords[2 * k + 1] < 0 || patch_coords[2 * k + 1] >= g.npy)
            return fail(ctx, GM_ERR_ARG, "gm_upload_map: patch (%d,%d) outside the %dx%d directory", patch_coords[2 * k], patch_coords[2 * k + 1], g.npx, g.npy);
    Geom old;
    CK(cudaMemcpyAsync(&old, ctx->geom + particle, sizeof old, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    if (int r = zero_counters(ctx)) return r;
    launch_release_particle(ctx->dir + (size_t)particle * ctx->dir_cap, old.npx * old.npy, ctx->pool, ctx->st);
    launch_merge_free(ctx->pool, ctx->st);
    ctx->stats.launches += 2;
    if (int r = gmi_ensure_dir_cap(ctx, g.npx * g.npy)) return r;
    CK(cudaMemcpyAsync(ctx->geom + particle, &g, sizeof g, cudaMemcpyHostToDevice, ctx->st));
    if (npatches) {
        if (2 * (size_t)npatches > 2 * ctx->list_cap) { CK(dalloc(ctx->d_list, (size_t)npatches * 2)); ctx->list_cap = npatches; }
        if ((size_t)npatches > ctx->export_cap) { CK(dalloc(ctx->d_export, (size_t)npatches * kPatchCells)); ctx->export_cap = npatches; }
        CK(cudaMemcpyAsync(ctx->d_list, patch_coords, sizeof(int) * 2 * npatches, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemcpyAsync(ctx->d_export, cells, sizeof(int4) * kPatchCells * (size_t)npatches, cudaMemcpyHostToDevice, ctx->st));
        launch_install_patches(ctx->dir + (size_t)particle * ctx->dir_cap, g, ctx->d_list, ctx->d_export, ctx->pool, ctx->d_counters, (int)npatches, ctx->st);
        ctx->stats.launches++;
    }
    if (int r = check_launch(ctx, "gm_upload_map")) return r;
    Counters h;
    if (int r = read_counters(ctx, h)) return r;
    if (h.error) return fail(ctx, h.error, "gm_upload_map: patch pool of %d patches exhausted", ctx->pool.cap);
    return GM_OK;
}

int gm_map_info_get(gm_ctx* ctx, uint32_t particle, gm_map_info* info) {
    return gm_download_map(ctx, particle, info, nullptr, nullptr, 0);
}

int gm_download_map(gm_ctx* ctx, uint32_t particle, gm_map_info* info, int32_t* patch_coords, gm_cell* cells, uint32_t max_patches) {
    if (int r = ready(ctx, true)) return r;
    if (particle >= ctx->n || !info) return fail(ctx, GM_ERR_ARG, "gm_download_map: bad argument");
    Geom g;
    CK(cudaMemcpyAsync(&g, ctx->geom + particle, sizeof g, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    const size_t nent = (size_t)g.npx * g.npy;
    if (nent > ctx->list_cap) { CK(dalloc(ctx->d_list, nent * 2)); ctx->list_cap = nent; }
    CK(cudaMemsetAsync(ctx->d_count, 0, sizeof(int), ctx->st));
    launch_export_list(ctx->dir + (size_t)particle * ctx->dir_cap, g, ctx->d_list, ctx->d_count, ctx->st);
    ctx->stats.launches++;
    if (int r = check_launch(ctx, "k_export_list")) return r;
    int cnt = 0;
    CK(cudaMemcpyAsync(&cnt, ctx->d_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    info->center_x = ctx->P.cx; info->center_y = ctx->P.cy; info->delta = ctx->P.delta;
    info->size_x2 = g.sx2; info->size_y2 = g.sy2; info->patches_x = g.npx; info->patches_y = g.npy;
    info->allocated_patches = cnt; info->reserved = 0;
    if (!patch_coords || !cells || !max_patches || !cnt) return GM_OK;
    std::vector<int> list((size_t)cnt * 2);
    CK(cudaMemcpyAsync(list.data(), ctx->d_list, sizeof(int) * 2 * cnt, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    std::vector<std::pair<int, int>> ent(cnt);
    for (int k = 0; k < cnt; k++) ent[k] = {list[2 * k], list[2 * k + 1]};
    std::sort(ent.begin(), ent.end());  // directory order: px-major, like the reference's storage walk
    const int m = std::min<int>(cnt, (int)max_patches);
    for (int k = 0; k < m; k++) {
        list[2 * k] = ent[k].first; list[2 * k + 1] = ent[k].second;
        patch_coords[2 * k] = ent[k].first / g.npy; patch_coords[2 * k + 1] = ent[k].first % g.npy;
    }
    CK(cudaMemcpyAsync(ctx->d_list, list.data(), sizeof(int) * 2 * m, cudaMemcpyHostToDevice, ctx->st));
    if ((size_t)m > ctx->export_cap) { CK(dalloc(ctx->d_export, (size_t)m * kPatchCells)); ctx->export_cap = m; }
    launch_export_gather(ctx->pool.cells, ctx->pool.visits, ctx->pool.cap_h, ctx->d_list, ctx->d_export, m, ctx->st);
    ctx->stats.launches++;
    if (int r = check_launch(ctx, "k_export_gather")) return r;
    CK(cudaMemcpyAsync(cells, ctx->d_export, sizeof(int4) * kPatchCells * (size_t)m, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return GM_OK;
}

int gm_export_occupancy(gm_ctx* ctx, uint32_t particle, double occ_thresh, int8_t* out, uint32_t width, uint32_t height) {
    if (int r = ready(ctx, true)) return r;
    if (particle >= ctx->n || !out || !width || !height) return fail(ctx, GM_ERR_ARG, "gm_export_occupancy: bad argument");
    Geom g;
    CK(cudaMemcpyAsync(&g, ctx->geom + particle, sizeof g, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    if ((int)width != (g.npx << kPatchMag) || (int)height != (g.npy << kPatchMag))
        return fail(ctx, GM_ERR_ARG, "gm_export_occupancy: the map is %d x %d cells, the buffer %u x %u", g.npx << kPatchMag, g.npy << kPatchMag, width, height);
    const size_t bytes = (size_t)width * height;
    if (bytes > ctx->occ_cap) { CK(dalloc(ctx->d_occ, bytes)); ctx->occ_cap = bytes; }
    launch_export_occupancy(ctx->dir + (size_t)particle * ctx->dir_cap, g, ctx->pool.cells, ctx->pool.visits, ctx->pool.cap_h, occ_thresh, ctx->d_occ, (int)width, (int)height, ctx->st);
    ctx->stats.launches++;
    if (int r = check_launch(ctx, "k_export_occupancy")) return r;
    CK(cudaMemcpyAsync(out, ctx->d_occ, bytes, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return GM_OK;
}

int gm_map_digest(gm_ctx* ctx, uint32_t first, uint32_t count, gm_digest* out) {
    if (int r = ready(ctx, true)) return r;
    if (!out || first + count > ctx->n) return fail(ctx, GM_ERR_ARG, "gm_map_digest: bad argument");
    if (!count) return GM_OK;
    DigestArgs a;
    a.geom = ctx->geom; a.dir = ctx->dir; a.dir_cap = ctx->dir_cap; a.pool = ctx->pool.cells; a.visits = ctx->pool.visits; a.means = ctx->pool.means; a.occ = ctx->pool.occ; a.first = first; a.out = ctx->d_digest; a.cap_h = ctx->pool.cap_h;
    launch_digest(a, count, ctx->st);
    ctx->stats.launches++;
    if (int r = check_launch(ctx, "k_digest")) return r;
    static_assert(sizeof(gm_digest) == 48, "gm_digest is six 64-bit words");
    std::vector<unsigned long long> words((size_t)count * 7);
    CK(cudaMemcpyAsync(words.data(), ctx->d_digest, sizeof(unsigned long long) * words.size(), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    unsigned long long bad = 0;
    for (uint32_t i = 0; i < count; i++) { memcpy(&out[i], &words[(size_t)i * 7], sizeof(gm_digest)); bad += words[(size_t)i * 7 + 6]; }
    if (bad) return fail(ctx, GM_ERR_INTERNAL, "gm_map_digest: %llu occupancy-bitmap rows disagree with their cells", bad);
    return GM_OK;
}

int gm_map_entropy(gm_ctx* ctx, uint32_t first, uint32_t count, double* out) {
    if (int r = ready(ctx, true)) return r;
    if (!out || first + count > ctx->n) return fail(ctx, GM_ERR_ARG, "gm_map_entropy: bad argument");
    if (!count) return GM_OK;
    DigestArgs a;
    a.geom = ctx->geom; a.dir = ctx->dir; a.dir_cap = ctx->dir_cap; a.pool = ctx->pool.cells; a.visits = ctx->pool.visits; a.means = ctx->pool.means; a.occ = ctx->pool.occ; a.first = first; a.out = nullptr; a.cap_h = ctx->pool.cap_h;
    double* d = reinterpret_cast<double*>(ctx->d_digest);  // scratch: seven words per particle
    launch_entropy(a, d, count, ctx->st);
    ctx->stats.launches++;
    if (int r = check_launch(ctx, "k_entropy")) return r;
    CK(cudaMemcpyAsync(out, d, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return GM_OK;
}

int gm_check_refcounts(gm_ctx* ctx, uint64_t* mismatches_out, uint64_t* referenced_out) {
    if (int r = ready(ctx, true)) return r;
    unsigned long long* out = ctx->d_digest;  // scratch: three words
    CK(cudaMemsetAsync(out, 0, 3 * sizeof(unsigned long long), ctx->st));
    // every context that shares the pool holds references: all their directories are counted (refcnt_new is all-zero scratch)
    if (ctx->share) {
        for (gm_ctx* m : ctx->share->members) {
            if (m != ctx && m->reg_pending) { gm_ctx* self = ctx; ctx = m; const int r = gmi_register_finish(m); ctx = self; if (r) return fail(ctx, r, "gm_check_refcounts: %s", m->err.c_str()); }
            if (m->inited) launch_recount(m->geom, m->dir, m->dir_cap, ctx->pool, ctx->refcnt_new, out, m->n, ctx->st);
        }
    } else launch_recount(ctx->geom, ctx->dir, ctx->dir_cap, ctx->pool, ctx->refcnt_new, out, ctx->n, ctx->st);
    launch_recount_compare(ctx->pool, ctx->refcnt_new, out, ctx->st);
    ctx->stats.launches += 2;
    if (int r = check_launch(ctx, "k_recount")) return r;
    unsigned long long h[3] = {0, 0, 0};
    int tops[2] = {0, 0};
    CK(cudaMemcpyAsync(h, out, sizeof h, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(tops, ctx->pool.free_top, sizeof tops, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    const int top = std::max(0, tops[0]) + std::max(0, tops[1]);
    if (mismatches_out) *mismatches_out = h[0] + h[2];
    if (referenced_out) *referenced_out = h[1];
    if (h[0] || h[2]) return fail(ctx, GM_ERR_INTERNAL, "gm_check_refcounts: %llu patches whose reference count differs from the directories, %llu dangling references", h[0], h[2]);
    // every patch is either referenced or on the free stack
    if ((unsigned long long)(ctx->pool.cap - top) != h[1])
        return fail(ctx, GM_ERR_INTERNAL, "gm_check_refcounts: %llu patches are referenced but %d are off the free stack", h[1], ctx->pool.cap - top);
    return GM_OK;
}

int gm_clone(gm_ctx* src, gm_ctx** out) {
    if (!src || !out) return GM_ERR_ARG;
    *out = nullptr;
    if (src->dist) return fail(src, GM_ERR_ARG, "gm_clone: a sharded context cannot be cloned");
    if (int r = bind(src)) return r;
    { gm_ctx* ctx = src; if (int r = gmi_register_finish(ctx)) return r; }
    gm_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.device = src->device; cfg.max_particles = src->max_particles; cfg.pool_patches = (uint64_t)src->pool.cap;
    cfg.block_threads = src->threads; cfg.global_particles = src->weights_cap;
    gm_ctx* dst = nullptr;
    if (int r = create_ctx(&dst, &cfg, src)) return fail(src, r, "gm_clone: %s", gm_last_error(nullptr));
    // the clone is a second set of patch directories over the SAME pool: a reference more per patch, no patch is copied
    // (the reference's copy: HierarchicalArray2D copy by autoptr shares, harray2d.h:63-77); its kernels run on the same stream
    if (!src->share) { src->share = new PoolShare(); src->share->members.push_back(src); }
    dst->share = src->share;
    dst->share->users++;
    dst->share->members.push_back(dst);
    dst->P = src->P; dst->laser_set = src->laser_set; dst->match_set = src->match_set;
    dst->async_register = src->async_register; dst->fused_register = src->fused_register;
    memcpy(dst->angles_h, src->angles_h, sizeof src->angles_h);
    auto bail = [&](cudaError_t e, const char* what) {
        dst->inited = false;
        gm_destroy(dst);
        return fail(src, GM_ERR_CUDA, "gm_clone: %s: %s", what, cudaGetErrorString(e));
    };
    cudaError_t e;
    if ((e = cudaMemcpyAsync(dst->d_angles, src->d_angles, sizeof(double) * GM_MAXBEAMS, cudaMemcpyDeviceToDevice, src->st)) != cudaSuccess) return bail(e, "angles");
    if (src->inited) {
        dst->n = src->n; dst->dir_cap = src->dir_cap; dst->reg_active_hint = src->reg_active_hint;
        const size_t dn = (size_t)src->max_particles * src->dir_cap;
        if ((e = dalloc(dst->dir, dn)) != cudaSuccess || (e = dalloc(dst->dir_alt, dn)) != cudaSuccess) return bail(e, "directories");
        if ((e = dalloc(dst->d_w, dst->weights_cap)) != cudaSuccess || (e = dalloc(dst->d_logw, dst->weights_cap)) != cudaSuccess ||
            (e = dalloc(dst->d_cum, dst->weights_cap)) != cudaSuccess || (e = dalloc(dst->d_tgt, dst->weights_cap)) != cudaSuccess ||
            (e = dalloc(dst->d_neff, 1)) != cudaSuccess || (e = dalloc(dst->d_widx, dst->weights_cap)) != cudaSuccess) return bail(e, "weight buffers");
        if ((e = cudaMemcpyAsync(dst->dir, src->dir, dn * sizeof(int), cudaMemcpyDeviceToDevice, src->st)) != cudaSuccess ||
            (e = cudaMemcpyAsync(dst->geom, src->geom, sizeof(Geom) * src->max_particles, cudaMemcpyDeviceToDevice, src->st)) != cudaSuccess) return bail(e, "device copy");
        launch_fill_int(dst->dir_alt, dn, -1, src->st);
        launch_addref_slots(dst->geom, dst->dir, dst->dir_cap, dst->n, dst->pool, src->st);
        if ((e = cudaGetLastError()) != cudaSuccess) return bail(e, "launch");
        dst->inited = true;
    }
    if ((e = cudaStreamSynchronize(src->st)) != cudaSuccess) return bail(e, "synchronize");
    dst->stats = src->stats;
    *out = dst;
    return GM_OK;
}

int gm_set_async_register(gm_ctx* ctx, int on) {
    if (!ctx) return GM_ERR_ARG;
    if (int r = bind(ctx)) return r;
    if (int r = gmi_register_finish(ctx)) return r;
    ctx->async_register = on != 0;
    return GM_OK;
}

int gm_set_fused_register(gm_ctx* ctx, int on) {
    if (!ctx) return GM_ERR_ARG;
    if (int r = bind(ctx)) return r;
    if (int r = gmi_register_finish(ctx)) return r;
    ctx->fused_register = on != 0;
    return GM_OK;
}

int gm_synchronize(gm_ctx* ctx) {
    if (!ctx) return GM_ERR_ARG;
    if (int r = bind(ctx)) return r;
    if (int r = gmi_register_finish(ctx)) return r;
    CK(cudaStreamSynchronize(ctx->st));
    return GM_OK;
}

int gm_get_stats(gm_ctx* ctx, gm_stats* out) {
    if (!ctx || !out) return GM_ERR_ARG;
    if (ctx->reg_pending) { if (int r = bind(ctx)) return r; if (int r = gmi_register_finish(ctx)) return r; }
    *out = ctx->stats;
    return GM_OK;
}

}  // extern "C"
