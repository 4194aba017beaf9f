/* tests/mock/gm_mock.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A stand-in for libgm_b200.so that implements the C ABI of include/gm_b200.h on the CPU with the oracle
 * (oracle/gm_oracle.c), so that the C++ host layer (slam-gmapping-learn_b200/host: GridSlamProcessor::processScan's
 * orchestration, the trajectory tree, the host map mirror, the .gfs trace, ScanMatcher's standalone entry points) can be
 * exercised by the `-m "not gpu"` tests on a machine without a GPU.  It is built into tests/_build/libgm_mock.so and
 * injected with LD_PRELOAD by tests/test_host_layer_cpu.py only; nothing under slam-gmapping-learn_b200/ knows it exists
 * and the product library still refuses to run without a CUDA device.
 *
 * Semantics follow the comments of include/gm_b200.h.  The multi-rank entry points (gm_dist_*) exchange through files in
 * $GM_MOCK_DIST_DIR, one process per rank: enough to run the sharded host logic of two ranks on one CPU box.
 */
#define _DEFAULT_SOURCE /* usleep */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include "gm_b200.h"
#include "gm_oracle.h"

struct gm_ctx {
    uint32_t max_particles, n, weights_cap;
    orc_matcher sm;
    int laser_set, match_set, inited;
    orc_map** maps; /* [max_particles] */
    gm_stats stats;
    char err[256];
    struct dist_state* dist; /* multi-rank runs: see the end of this file */
    int fused_register;
    double *sw_rows, *sw_w, sw_neff; uint32_t sw_total; /* between gm_scanmatch_weights_begin and _end */
    uint8_t *dist_sent_prev, *dist_sent_prev2;
    unsigned long dist_seq_prev, dist_seq_prev2;
};

static char g_create_err[256] = "";

static int fail(gm_ctx* c, int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c ? c->err : g_create_err, 256, fmt, ap);
    va_end(ap);
    return code;
}

const char* gm_last_error(const gm_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

void gm_default_matching(gm_match_params* p) {
    orc_matcher m;
    orc_matcher_defaults(&m);
    memset(p, 0, sizeof *p);
    p->laser_max_range = m.laser_max_range; p->usable_range = m.usable_range;
    p->gaussian_sigma = m.gaussian_sigma; p->likelihood_sigma = m.likelihood_sigma;
    p->kernel_size = m.kernel_size; p->generate_map = m.generate_map;
    p->opt_linear_delta = m.opt_linear_delta; p->opt_angular_delta = m.opt_angular_delta;
    p->opt_recursive_iterations = m.opt_recursive_iterations; p->likelihood_skip = m.likelihood_skip;
    p->enlarge_step = m.enlarge_step; p->fullness_threshold = m.fullness_threshold;
    p->angular_odometry_reliability = m.angular_odometry_reliability; p->linear_odometry_reliability = m.linear_odometry_reliability;
    p->free_cell_ratio = m.free_cell_ratio; p->initial_beams_skip = m.initial_beams_skip;
}

int gm_create(gm_ctx** out, const gm_config* cfg) {
    if (!out || !cfg || !cfg->max_particles) return fail(NULL, GM_ERR_ARG, "gm_create: bad argument");
    gm_ctx* c = (gm_ctx*)calloc(1, sizeof *c);
    c->max_particles = cfg->max_particles;
    c->weights_cap = cfg->global_particles > cfg->max_particles ? cfg->global_particles : cfg->max_particles;
    c->maps = (orc_map**)calloc(c->max_particles, sizeof(orc_map*));
    orc_matcher_defaults(&c->sm);
    c->stats.pool_capacity = cfg->pool_patches;
    *out = c;
    return GM_OK;
}

static void drop_maps(gm_ctx* c) {
    for (uint32_t i = 0; i < c->max_particles; i++) { orc_map_free(c->maps[i]); c->maps[i] = NULL; }
}

int gm_destroy(gm_ctx* c) {
    if (!c) return GM_ERR_ARG;
    drop_maps(c);
    free(c->maps);
    free(c);
    return GM_OK;
}

int gm_set_laser(gm_ctx* c, uint32_t nbeams, const double* angles, const double laser_pose[3]) {
    if (!c) return GM_ERR_ARG;
    if (!angles || !laser_pose) return fail(c, GM_ERR_ARG, "gm_set_laser: null argument");
    if (nbeams == 0 || nbeams >= GM_MAXBEAMS) return fail(c, GM_ERR_ARG, "gm_set_laser: %u beams (must be in [1, %d))", nbeams, GM_MAXBEAMS);
    c->sm.beams = nbeams;
    memcpy(c->sm.angles, angles, sizeof(double) * nbeams);
    memcpy(c->sm.laser_pose, laser_pose, sizeof(double) * 3);
    c->laser_set = 1;
    return GM_OK;
}

int gm_set_matching(gm_ctx* c, const gm_match_params* p) {
    if (!c) return GM_ERR_ARG;
    if (!p) return fail(c, GM_ERR_ARG, "gm_set_matching: null argument");
    if (p->kernel_size < 0 || p->kernel_size > 8) return fail(c, GM_ERR_ARG, "gm_set_matching: kernel_size %d outside [0,8]", p->kernel_size);
    orc_matcher* m = &c->sm;
    m->laser_max_range = p->laser_max_range; m->usable_range = p->usable_range;
    m->gaussian_sigma = p->gaussian_sigma; m->likelihood_sigma = p->likelihood_sigma;
    m->kernel_size = p->kernel_size; m->generate_map = p->generate_map;
    m->opt_linear_delta = p->opt_linear_delta; m->opt_angular_delta = p->opt_angular_delta;
    m->opt_recursive_iterations = p->opt_recursive_iterations; m->likelihood_skip = p->likelihood_skip;
    m->enlarge_step = p->enlarge_step; m->fullness_threshold = p->fullness_threshold;
    m->angular_odometry_reliability = p->angular_odometry_reliability; m->linear_odometry_reliability = p->linear_odometry_reliability;
    m->free_cell_ratio = p->free_cell_ratio; m->initial_beams_skip = p->initial_beams_skip;
    c->match_set = 1;
    return GM_OK;
}

int gm_init_particles(gm_ctx* c, uint32_t n, const double center[2], double world_w, double world_h, double delta) {
    if (!c) return GM_ERR_ARG;
    if (!center || !n || n > c->max_particles || !(delta > 0)) return fail(c, GM_ERR_ARG, "gm_init_particles: bad argument (n=%u, max=%u)", n, c->max_particles);
    drop_maps(c);
    for (uint32_t i = 0; i < n; i++) c->maps[i] = orc_map_new(center[0], center[1], world_w, world_h, delta);
    if (c->maps[0]->npx <= 0 || c->maps[0]->npy <= 0) { drop_maps(c); return fail(c, GM_ERR_ARG, "gm_init_particles: map smaller than one patch"); }
    c->n = n;
    c->inited = 1;
    return GM_OK;
}

uint32_t gm_particles(const gm_ctx* c) { return c ? c->n : 0; }

static int ready(gm_ctx* c) {
    if (!c) return GM_ERR_ARG;
    if (!c->inited) return fail(c, GM_ERR_ARG, "gm_init_particles has not been called");
    if (!c->laser_set) return fail(c, GM_ERR_ARG, "gm_set_laser has not been called");
    if (!c->match_set) return fail(c, GM_ERR_ARG, "gm_set_matching has not been called");
    return GM_OK;
}

static int check_idx(gm_ctx* c, uint32_t count, const uint32_t* idx) {
    for (uint32_t i = 0; i < count; i++)
        if (idx[i] >= c->n) return fail(c, GM_ERR_ARG, "eval: particle index %u >= %u", idx[i], c->n);
    return GM_OK;
}

int gm_score(gm_ctx* c, uint32_t count, const uint32_t* idx, const double* poses, const double* ranges, double* score_out) {
    int r = ready(c);
    if (r) return r;
    if (!count) return GM_OK;
    if (!idx || !poses || !ranges) return fail(c, GM_ERR_ARG, "eval: null argument");
    if ((r = check_idx(c, count, idx))) return r;
    for (uint32_t i = 0; i < count; i++) {
        double s = orc_score(&c->sm, c->maps[idx[i]], poses + 3 * i, ranges);
        if (score_out) score_out[i] = s;
    }
    return GM_OK;
}

int gm_likelihood_and_score(gm_ctx* c, uint32_t count, const uint32_t* idx, const double* poses, const double* ranges,
                            double* s_out, double* l_out, uint32_t* c_out) {
    int r = ready(c);
    if (r) return r;
    if (!count) return GM_OK;
    if (!idx || !poses || !ranges) return fail(c, GM_ERR_ARG, "eval: null argument");
    if ((r = check_idx(c, count, idx))) return r;
    for (uint32_t i = 0; i < count; i++) {
        double s, l;
        uint32_t k = orc_likelihood_and_score(&c->sm, c->maps[idx[i]], poses + 3 * i, ranges, &s, &l);
        if (s_out) s_out[i] = s;
        if (l_out) l_out[i] = l;
        if (c_out) c_out[i] = k;
    }
    return GM_OK;
}

int gm_optimize(gm_ctx* c, uint32_t count, const uint32_t* idx, const double* poses_in, const double* ranges, double* poses_out,
                double* score_out) {
    int r = ready(c);
    if (r) return r;
    if (!count) return GM_OK;
    if (!idx || !poses_in || !ranges) return fail(c, GM_ERR_ARG, "eval: null argument");
    if ((r = check_idx(c, count, idx))) return r;
    for (uint32_t i = 0; i < count; i++) {
        double out[3];
        double s = orc_optimize(&c->sm, c->maps[idx[i]], poses_in + 3 * i, ranges, out);
        if (poses_out) memcpy(poses_out + 3 * i, out, sizeof out);
        if (score_out) score_out[i] = s;
    }
    return GM_OK;
}

int gm_scanmatch(gm_ctx* c, const double* ranges, double* poses, double minimum_score, double* score_out, double* s_out,
                 double* l_out, uint32_t* c_out, uint32_t* evals_out) {
    int r = ready(c);
    if (r) return r;
    if (!ranges || !poses) return fail(c, GM_ERR_ARG, "gm_scanmatch: null argument");
    const uint64_t calls0 = c->sm.score_calls, beams0 = c->sm.beam_evals;
    for (uint32_t i = 0; i < c->n; i++) {
        double corrected[3], s, l;
        const uint64_t k0 = c->sm.score_calls;
        const double score = orc_optimize(&c->sm, c->maps[i], poses + 3 * i, ranges, corrected);
        if (evals_out) evals_out[i] = (uint32_t)(c->sm.score_calls - k0);
        if (score > minimum_score) memcpy(poses + 3 * i, corrected, sizeof corrected);
        const uint32_t k = orc_likelihood_and_score(&c->sm, c->maps[i], poses + 3 * i, ranges, &s, &l);
        if (score_out) score_out[i] = score;
        if (s_out) s_out[i] = s;
        if (l_out) l_out[i] = l;
        if (c_out) c_out[i] = k;
    }
    c->stats.score_evals = c->sm.score_calls - calls0; c->stats.beam_evals = c->sm.beam_evals - beams0;
    c->stats.total_score_evals += c->stats.score_evals; c->stats.total_beam_evals += c->stats.beam_evals;
    c->stats.total_scanmatches++;
    return GM_OK;
}

int gm_normalize_neff(gm_ctx* c, const double* logw, uint32_t n, double gain, double* w_out, double* neff_out) {
    if (!c) return GM_ERR_ARG;
    if (!c->inited) return fail(c, GM_ERR_ARG, "gm_init_particles has not been called");
    if (!logw || !n || n > c->weights_cap) return fail(c, GM_ERR_ARG, "gm_normalize_neff: bad argument");
    double* w = (double*)malloc(sizeof(double) * n);
    const double neff = orc_normalize(logw, n, gain, w);
    if (w_out) memcpy(w_out, w, sizeof(double) * n);
    if (neff_out) *neff_out = neff;
    free(w);
    return GM_OK;
}

int gm_resample_indexes_n(gm_ctx* c, const double* w, uint32_t n, uint32_t n_out, double u01, uint32_t* idx_out, uint32_t* emitted_out) {
    if (!c) return GM_ERR_ARG;
    if (!c->inited) return fail(c, GM_ERR_ARG, "gm_init_particles has not been called");
    if (!w || !idx_out || !n || n > c->weights_cap || !n_out || n_out > c->weights_cap) return fail(c, GM_ERR_ARG, "gm_resample_indexes: bad argument");
    const uint32_t k = orc_resample_indexes_n(w, n, n_out, u01, idx_out);
    if (emitted_out) *emitted_out = k;
    return GM_OK;
}

int gm_resample_indexes(gm_ctx* c, const double* w, uint32_t n, double u01, uint32_t* idx_out, uint32_t* emitted_out) {
    return gm_resample_indexes_n(c, w, n, n, u01, idx_out, emitted_out);
}

/* the new generation has n_new maps, map j a copy of the old generation's idx[j] */
static int remap_n(gm_ctx* c, const uint32_t* idx, uint32_t n_new) {
    for (uint32_t j = 0; j < n_new; j++)
        if (idx[j] >= c->n) return fail(c, GM_ERR_ARG, "remap: parent index %u >= %u", idx[j], c->n);
    orc_map** next = (orc_map**)calloc(n_new, sizeof(orc_map*));
    for (uint32_t j = 0; j < n_new; j++) next[j] = orc_map_clone(c->maps[idx[j]]);
    for (uint32_t j = 0; j < c->n; j++) { orc_map_free(c->maps[j]); c->maps[j] = NULL; }
    for (uint32_t j = 0; j < n_new; j++) c->maps[j] = next[j];
    c->n = n_new;
    free(next);
    return GM_OK;
}

static int remap(gm_ctx* c, const uint32_t* idx) {
    return remap_n(c, idx, c->n);
}

int gm_copy_particles(gm_ctx* c, const uint32_t* idx) {
    int r = ready(c);
    if (r) return r;
    if (!idx) return fail(c, GM_ERR_ARG, "gm_copy_particles: null argument");
    return remap(c, idx);
}

int gm_copy_particles_n(gm_ctx* c, const uint32_t* idx, uint32_t n_new) {
    int r = ready(c);
    if (r) return r;
    if (!idx) return fail(c, GM_ERR_ARG, "gm_copy_particles_n: null argument");
    if (c->dist) return fail(c, GM_ERR_ARG, "gm_copy_particles_n: the particle count of a sharded context is fixed");
    if (!n_new || n_new > c->max_particles) return fail(c, GM_ERR_ARG, "gm_copy_particles_n: %u particles, the context was created for at most %u", n_new, c->max_particles);
    return remap_n(c, idx, n_new);
}

int gm_register(gm_ctx* c, const double* ranges, const double* poses, const uint32_t* idx) {
    int r = ready(c);
    if (r) return r;
    if (!ranges || !poses) return fail(c, GM_ERR_ARG, "gm_register: null argument");
    if (idx && (r = remap(c, idx))) return r;
    for (uint32_t i = 0; i < c->n; i++) {
        orc_compute_active_area(&c->sm, c->maps[i], poses + 3 * i, ranges);
        orc_register_scan(&c->sm, c->maps[i], poses + 3 * i, ranges);
    }
    c->stats.total_registers++;
    return GM_OK;
}

int gm_compute_active_area(gm_ctx* c, const double* ranges, const double* poses) {
    int r = ready(c);
    if (r) return r;
    if (!ranges || !poses) return fail(c, GM_ERR_ARG, "gm_compute_active_area: null argument");
    for (uint32_t i = 0; i < c->n; i++) orc_compute_active_area(&c->sm, c->maps[i], poses + 3 * i, ranges);
    return GM_OK;
}

int gm_upload_map(gm_ctx* c, uint32_t particle, const gm_map_info* info, const int32_t* coords, const gm_cell* cells, uint32_t npatches) {
    int r = ready(c);
    if (r) return r;
    if (particle >= c->n || !info || (npatches && (!coords || !cells))) return fail(c, GM_ERR_ARG, "gm_upload_map: bad argument");
    if (info->patches_x <= 0 || info->patches_y <= 0) return fail(c, GM_ERR_ARG, "gm_upload_map: empty directory");
    for (uint32_t k = 0; k < npatches; k++)
        if (coords[2 * k] < 0 || coords[2 * k] >= info->patches_x || coords[2 * k + 1] < 0 || coords[2 * k + 1] >= info->patches_y)
            return fail(c, GM_ERR_ARG, "gm_upload_map: patch (%d,%d) outside the %dx%d directory", coords[2 * k], coords[2 * k + 1], info->patches_x, info->patches_y);
    orc_map* old = c->maps[particle];
    orc_map* m = (orc_map*)calloc(1, sizeof *m);
    m->cx = old->cx; m->cy = old->cy; m->delta = old->delta;
    m->npx = info->patches_x; m->npy = info->patches_y;
    m->map_w = m->npx << ORC_PATCH_MAG; m->map_h = m->npy << ORC_PATCH_MAG;
    m->world_w = m->map_w * m->delta; m->world_h = m->map_h * m->delta;
    m->size_x2 = info->size_x2; m->size_y2 = info->size_y2;
    const size_t nd = (size_t)m->npx * m->npy;
    m->dir = (orc_cell**)calloc(nd, sizeof(orc_cell*));
    m->active = (uint8_t*)calloc(nd, 1);
    for (uint32_t k = 0; k < npatches; k++) {
        orc_cell** slot = &m->dir[(size_t)coords[2 * k] * m->npy + coords[2 * k + 1]];
        if (!*slot) *slot = (orc_cell*)malloc(sizeof(orc_cell) * 1024);
        memcpy(*slot, cells + (size_t)1024 * k, sizeof(orc_cell) * 1024);
    }
    orc_map_free(old);
    c->maps[particle] = m;
    return GM_OK;
}

int gm_download_map(gm_ctx* c, uint32_t particle, gm_map_info* info, int32_t* coords, gm_cell* cells, uint32_t max_patches) {
    int r = ready(c);
    if (r) return r;
    if (particle >= c->n || !info) return fail(c, GM_ERR_ARG, "gm_download_map: bad argument");
    const orc_map* m = c->maps[particle];
    info->center_x = m->cx; info->center_y = m->cy; info->delta = m->delta;
    info->size_x2 = m->size_x2; info->size_y2 = m->size_y2; info->patches_x = m->npx; info->patches_y = m->npy;
    info->allocated_patches = (int32_t)orc_map_patch_count(m); info->reserved = 0;
    if (!coords || !cells || !max_patches || !info->allocated_patches) return GM_OK;
    orc_map_export(m, coords, (orc_cell*)cells, max_patches); /* px-major directory order, gm_cell == orc_cell */
    return GM_OK;
}

int gm_map_info_get(gm_ctx* c, uint32_t particle, gm_map_info* info) { return gm_download_map(c, particle, info, NULL, NULL, 0); }

int gm_export_occupancy(gm_ctx* c, uint32_t particle, double occ_thresh, int8_t* out, uint32_t width, uint32_t height) {
    int r = ready(c);
    if (r) return r;
    if (particle >= c->n || !out || !width || !height) return fail(c, GM_ERR_ARG, "gm_export_occupancy: bad argument");
    const orc_map* m = c->maps[particle];
    if ((int)width != m->map_w || (int)height != m->map_h)
        return fail(c, GM_ERR_ARG, "gm_export_occupancy: the map is %d x %d cells, the buffer %u x %u", m->map_w, m->map_h, width, height);
    for (uint32_t y = 0; y < height; y++)
        for (uint32_t x = 0; x < width; x++) {
            const orc_cell* p = m->dir[(size_t)(x >> ORC_PATCH_MAG) * m->npy + (y >> ORC_PATCH_MAG)];
            int8_t v = -1;
            if (p) {
                const orc_cell* a = &p[((x & 31) << ORC_PATCH_MAG) + (y & 31)];
                if (a->visits) v = ((double)a->n / (double)a->visits > occ_thresh) ? 100 : 0;
            }
            out[(size_t)y * width + x] = v;
        }
    return GM_OK;
}

int gm_map_digest(gm_ctx* c, uint32_t first, uint32_t count, gm_digest* out) {
    int r = ready(c);
    if (r) return r;
    if (!out || first + count > c->n) return fail(c, GM_ERR_ARG, "gm_map_digest: bad argument");
    for (uint32_t i = 0; i < count; i++) {
        orc_digest d;
        orc_map_digest(c->maps[first + i], &d);
        out[i].patches = d.patches; out[i].cells = d.cells; out[i].sum_n = d.sum_n; out[i].sum_visits = d.sum_visits;
        out[i].hash_counts = d.hash_counts; out[i].hash_acc = d.hash_acc;
    }
    return GM_OK;
}

int gm_map_entropy(gm_ctx* c, uint32_t first, uint32_t count, double* out) {
    int r = ready(c);
    if (r) return r;
    if (!out || first + count > c->n) return fail(c, GM_ERR_ARG, "gm_map_entropy: bad argument");
    for (uint32_t k = 0; k < count; k++) {
        const orc_map* m = c->maps[first + k];
        const int64_t np = orc_map_patch_count(m);
        int32_t* coords = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)(np ? np : 1));
        orc_cell* cells = (orc_cell*)malloc(sizeof(orc_cell) * 1024 * (size_t)(np ? np : 1));
        orc_map_export(m, coords, cells, np);
        double acc = 0;
        for (int64_t i = 0; i < np * 1024; i++) {
            const int n = cells[i].n, v = cells[i].visits;
            if (!v) continue;
            double h = 0;
            if (n != v && n != 0) { const double x = (double)n / (double)v; h = -(x * log(x) + (1. - x) * log(1. - x)); }
            acc += h - 0.6931471805599453;
        }
        out[k] = acc;
        free(coords); free(cells);
    }
    return GM_OK;
}

/* the mock's maps own their patches outright (orc_map_clone copies): there are no shared patches to miscount */
int gm_check_refcounts(gm_ctx* c, uint64_t* mismatches_out, uint64_t* referenced_out) {
    int r = ready(c);
    if (r) return r;
    uint64_t used = 0;
    for (uint32_t i = 0; i < c->n; i++) used += (uint64_t)orc_map_patch_count(c->maps[i]);
    if (mismatches_out) *mismatches_out = 0;
    if (referenced_out) *referenced_out = used;
    return GM_OK;
}

int gm_clone(gm_ctx* src, gm_ctx** out) {
    if (!src || !out) return GM_ERR_ARG;
    if (src->dist) return fail(src, GM_ERR_ARG, "gm_clone: a sharded context cannot be cloned");
    gm_ctx* d = (gm_ctx*)calloc(1, sizeof *d);
    *d = *src;
    d->maps = (orc_map**)calloc(src->max_particles, sizeof(orc_map*));
    for (uint32_t i = 0; i < src->n; i++) d->maps[i] = orc_map_clone(src->maps[i]);
    *out = d;
    return GM_OK;
}

int gm_set_async_register(gm_ctx* c, int on) { (void)on; return c ? GM_OK : GM_ERR_ARG; }
int gm_set_fused_register(gm_ctx* c, int on) { if (!c) return GM_ERR_ARG; c->fused_register = on != 0; return GM_OK; }
int gm_synchronize(gm_ctx* c) { return c ? GM_OK : GM_ERR_ARG; }

int gm_get_stats(gm_ctx* c, gm_stats* out) {
    if (!c || !out) return GM_ERR_ARG;
    uint64_t used = 0;
    for (uint32_t i = 0; i < c->n; i++) used += (uint64_t)orc_map_patch_count(c->maps[i]);
    c->stats.pool_in_use = used;
    *out = c->stats;
    return GM_OK;
}

/* ---- multi-rank: one process per rank, exchange through files in $GM_MOCK_DIST_DIR/<communicator id> (written under a
 * temporary name and renamed, so a reader that sees a file sees all of it).  Sequence numbers keep the calls of the ranks in
 * step; a rank deletes what it wrote two exchanges ago (no peer can still need it: a peer is at most one exchange behind). */
struct dist_state { int rank, world; unsigned long seq; char dir[400]; };

int gm_dist_unique_id(void* id_out) {
    unsigned char* b = (unsigned char*)id_out;
    FILE* f = fopen("/dev/urandom", "rb");
    if (!f || fread(b, 1, GM_DIST_ID_BYTES, f) != GM_DIST_ID_BYTES) { if (f) fclose(f); return fail(NULL, GM_ERR_ARG, "gm_mock: /dev/urandom"); }
    fclose(f);
    return GM_OK;
}

int gm_dist_init(gm_ctx* c, int rank, int world, const void* id) {
    if (!c || !id || world < 1 || rank < 0 || rank >= world) return fail(c, GM_ERR_ARG, "gm_dist_init: bad argument");
    const char* base = getenv("GM_MOCK_DIST_DIR");
    if (!base) return fail(c, GM_ERR_NO_DEVICE, "gm_mock: GM_MOCK_DIST_DIR is not set (multi-rank runs of the mock exchange through files)");
    struct dist_state* d = (struct dist_state*)calloc(1, sizeof *d);
    d->rank = rank; d->world = world;
    const unsigned char* b = (const unsigned char*)id;
    int n = snprintf(d->dir, sizeof d->dir, "%s/comm_", base);
    for (int i = 0; i < 12; i++) n += snprintf(d->dir + n, sizeof d->dir - n, "%02x", b[i]);
    mkdir(d->dir, 0700); /* EEXIST from the other ranks is fine */
    c->dist = d;
    return GM_OK;
}

static void dist_path(const struct dist_state* d, char* out, size_t cap, const char* kind, unsigned long seq, unsigned who, int tmp) {
    snprintf(out, cap, "%s/%s_%lu_%u%s", d->dir, kind, seq, who, tmp ? ".tmp" : ".bin");
}

static int dist_write(gm_ctx* c, const char* kind, unsigned long seq, unsigned who, const void* data, size_t bytes) {
    struct dist_state* d = c->dist;
    char tmp[480], fin[480];
    dist_path(d, tmp, sizeof tmp, kind, seq, who, 1); dist_path(d, fin, sizeof fin, kind, seq, who, 0);
    FILE* f = fopen(tmp, "wb");
    if (!f || fwrite(data, 1, bytes, f) != bytes) { if (f) fclose(f); return fail(c, GM_ERR_INTERNAL, "gm_mock: cannot write %s", tmp); }
    fclose(f);
    if (rename(tmp, fin)) return fail(c, GM_ERR_INTERNAL, "gm_mock: cannot rename %s", tmp);
    return GM_OK;
}

/* wait (up to 120 s) for a peer's file, return its content (malloc) */
static int dist_read(gm_ctx* c, const char* kind, unsigned long seq, unsigned who, void** data, size_t* bytes) {
    char fin[480];
    dist_path(c->dist, fin, sizeof fin, kind, seq, who, 0);
    FILE* f = NULL;
    for (int spin = 0; spin < 600000 && !(f = fopen(fin, "rb")); spin++) usleep(200);
    if (!f) return fail(c, GM_ERR_INTERNAL, "gm_mock: timed out waiting for %s", fin);
    fseek(f, 0, SEEK_END);
    *bytes = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    *data = malloc(*bytes ? *bytes : 1);
    const size_t got = fread(*data, 1, *bytes, f);
    fclose(f);
    if (got != *bytes) { free(*data); return fail(c, GM_ERR_INTERNAL, "gm_mock: short read of %s", fin); }
    return GM_OK;
}

static void dist_unlink(const struct dist_state* d, const char* kind, unsigned long seq, unsigned who) {
    char fin[480];
    dist_path(d, fin, sizeof fin, kind, seq, who, 0);
    unlink(fin);
}

static int dist_allgather(gm_ctx* c, const double* local, uint32_t count, double* all) {
    if (!c || !c->dist) return fail(c, GM_ERR_ARG, "gm_dist_allgather: gm_dist_init has not been called");
    if (!local || !all) return fail(c, GM_ERR_ARG, "gm_dist_allgather: null argument");
    struct dist_state* d = c->dist;
    const unsigned long seq = ++d->seq;
    int r = dist_write(c, "ag", seq, (unsigned)d->rank, local, sizeof(double) * count);
    if (r) return r;
    for (int g = 0; g < d->world; g++) {
        void* buf; size_t bytes;
        if ((r = dist_read(c, "ag", seq, (unsigned)g, &buf, &bytes))) return r;
        if (bytes != sizeof(double) * count) { free(buf); return fail(c, GM_ERR_ARG, "gm_dist_allgather: rank %d sent %zu bytes, expected %zu", g, bytes, sizeof(double) * count); }
        memcpy(all + (size_t)g * count, buf, bytes);
        free(buf);
    }
    if (seq > 2) dist_unlink(d, "ag", seq - 2, (unsigned)d->rank);
    return GM_OK;
}

/* a map on the wire: 5 doubles (cx cy world_w world_h delta), 7 ints (map_w map_h size_x2 size_y2 npx npy patches), coords, cells */
static void* map_pack(const orc_map* m, size_t* bytes) {
    const int64_t np = orc_map_patch_count(m);
    *bytes = 5 * sizeof(double) + 7 * sizeof(int32_t) + (size_t)np * (2 * sizeof(int32_t) + 1024 * sizeof(orc_cell));
    char* buf = (char*)malloc(*bytes);
    const double h[5] = {m->cx, m->cy, m->world_w, m->world_h, m->delta};
    const int32_t g[7] = {m->map_w, m->map_h, m->size_x2, m->size_y2, m->npx, m->npy, (int32_t)np};
    memcpy(buf, h, sizeof h); memcpy(buf + sizeof h, g, sizeof g);
    int32_t* coords = (int32_t*)(buf + sizeof h + sizeof g);
    orc_map_export(m, coords, (orc_cell*)(coords + 2 * np), np);
    return buf;
}

static orc_map* map_unpack(const void* data) {
    const char* buf = (const char*)data;
    double h[5]; int32_t g[7];
    memcpy(h, buf, sizeof h); memcpy(g, buf + sizeof h, sizeof g);
    orc_map* m = (orc_map*)calloc(1, sizeof *m);
    m->cx = h[0]; m->cy = h[1]; m->world_w = h[2]; m->world_h = h[3]; m->delta = h[4];
    m->map_w = g[0]; m->map_h = g[1]; m->size_x2 = g[2]; m->size_y2 = g[3]; m->npx = g[4]; m->npy = g[5];
    const size_t nd = (size_t)m->npx * m->npy;
    m->dir = (orc_cell**)calloc(nd, sizeof(orc_cell*));
    m->active = (uint8_t*)calloc(nd, 1);
    const int32_t* coords = (const int32_t*)(buf + sizeof h + sizeof g);
    const orc_cell* cells = (const orc_cell*)(coords + 2 * (size_t)g[6]);
    for (int32_t k = 0; k < g[6]; k++) {
        orc_cell* p = (orc_cell*)malloc(sizeof(orc_cell) * 1024);
        memcpy(p, cells + (size_t)1024 * k, sizeof(orc_cell) * 1024);
        m->dir[(size_t)coords[2 * k] * m->npy + coords[2 * k + 1]] = p;
    }
    return m;
}

/* particle copy across ranks: child j of this rank continues the GLOBAL parent idx_global[first + j].  Parents that live here
 * and have children elsewhere are published once each; parents that live elsewhere are fetched once each. */
static int dist_remap(gm_ctx* c, const uint32_t* idx_global) {
    struct dist_state* d = c->dist;
    const uint32_t nl = c->n, first = (uint32_t)d->rank * nl, total = nl * (uint32_t)d->world;
    const unsigned long seq = ++d->seq;
    int r;
    for (uint32_t j = 0; j < total; j++)
        if (idx_global[j] >= total) return fail(c, GM_ERR_ARG, "gm_dist: parent index %u >= %u", idx_global[j], total);
    /* publish */
    uint8_t* sent = (uint8_t*)calloc(nl, 1);
    for (uint32_t j = 0; j < total; j++) {
        const uint32_t p = idx_global[j];
        if (j / nl == (uint32_t)d->rank || p / nl != (uint32_t)d->rank || sent[p - first]) continue;
        size_t bytes; void* buf = map_pack(c->maps[p - first], &bytes);
        r = dist_write(c, "map", seq, p, buf, bytes);
        free(buf);
        if (r) { free(sent); return r; }
        sent[p - first] = 1;
    }
    /* fetch + local clones */
    orc_map** next = (orc_map**)calloc(nl, sizeof(orc_map*));
    uint64_t imported = 0, imported_bytes = 0;
    for (uint32_t j = 0; j < nl; j++) {
        const uint32_t p = idx_global[first + j];
        if (p / nl == (uint32_t)d->rank) { next[j] = orc_map_clone(c->maps[p - first]); continue; }
        if (j && idx_global[first + j - 1] == p) { next[j] = orc_map_clone(next[j - 1]); continue; } /* systematic resampling: equal parents are adjacent */
        void* buf; size_t bytes;
        if ((r = dist_read(c, "map", seq, p, &buf, &bytes))) { free(sent); return r; }
        next[j] = map_unpack(buf);
        free(buf);
        imported++; imported_bytes += bytes;
    }
    for (uint32_t j = 0; j < nl; j++) { orc_map_free(c->maps[j]); c->maps[j] = next[j]; }
    free(next);
    /* what this rank published two exchanges ago is no longer needed by anybody */
    if (c->dist_sent_prev2) { for (uint32_t k = 0; k < nl; k++) if (c->dist_sent_prev2[k]) dist_unlink(d, "map", c->dist_seq_prev2, first + k); free(c->dist_sent_prev2); }
    c->dist_sent_prev2 = c->dist_sent_prev; c->dist_seq_prev2 = c->dist_seq_prev;
    c->dist_sent_prev = sent; c->dist_seq_prev = seq;
    c->stats.migrated_particles = imported; c->stats.migrated_bytes = imported_bytes;
    c->stats.total_migrated_particles += imported; c->stats.total_migrated_bytes += imported_bytes;
    return GM_OK;
}

int gm_dist_copy_particles(gm_ctx* c, const uint32_t* idx_global) {
    int r = ready(c);
    if (r) return r;
    if (!c->dist) return fail(c, GM_ERR_ARG, "gm_dist_copy_particles: gm_dist_init has not been called");
    if (!idx_global) return fail(c, GM_ERR_ARG, "gm_dist_copy_particles: null argument");
    return dist_remap(c, idx_global);
}

int gm_dist_register(gm_ctx* c, const double* ranges, const double* poses, const uint32_t* idx_global) {
    int r = ready(c);
    if (r) return r;
    if (!c->dist) return fail(c, GM_ERR_ARG, "gm_dist_register: gm_dist_init has not been called");
    if (!ranges || !poses) return fail(c, GM_ERR_ARG, "gm_dist_register: null argument");
    c->stats.migrated_particles = 0; c->stats.migrated_bytes = 0;
    if (idx_global && (r = dist_remap(c, idx_global))) return r;
    return gm_register(c, ranges, poses, NULL);
}

/* scanMatch + weight += l + normalize in one call; on a sharded context the rows of all ranks (exchanged through files here,
 * through peer memory in the product) in global particle order */
int gm_scanmatch_weights_begin(gm_ctx* c, const double* ranges, const double* poses_local, double minimum_score, const double* log_weights,
                               double gain) {
    int r = ready(c);
    if (r) return r;
    if (c->sw_rows) return fail(c, GM_ERR_ARG, "gm_scanmatch_weights_begin: the previous one has not been collected");
    if (!ranges || !poses_local || !log_weights) return fail(c, GM_ERR_ARG, "gm_scanmatch_weights: null argument");
    const uint32_t n = c->n, world = c->dist ? (uint32_t)c->dist->world : 1u, N = n * world;
    if (N > c->weights_cap) return fail(c, GM_ERR_ARG, "gm_scanmatch_weights: %u particles over all ranks, the context was created for %u", N, c->weights_cap);
    /* the mock computes everything here and keeps it for _end */
    double* rows_out = (double*)malloc(sizeof(double) * 5 * N);
    double* weights_out = (double*)malloc(sizeof(double) * N);
    double neff_keep = 0, *neff_out = &neff_keep;
    c->sw_total = N;
    if (N > c->weights_cap) return fail(c, GM_ERR_ARG, "gm_scanmatch_weights: %u particles over all ranks, the context was created for %u", N, c->weights_cap);
    double* poses = (double*)malloc(sizeof(double) * 3 * n);
    double* score = (double*)malloc(sizeof(double) * n);
    double* l = (double*)malloc(sizeof(double) * n);
    double* mine = (double*)malloc(sizeof(double) * 5 * n);
    memcpy(poses, poses_local, sizeof(double) * 3 * n);
    r = gm_scanmatch(c, ranges, poses, minimum_score, score, NULL, l, NULL, NULL);
    for (uint32_t i = 0; i < n && !r; i++) {
        mine[5 * i] = poses[3 * i]; mine[5 * i + 1] = poses[3 * i + 1]; mine[5 * i + 2] = poses[3 * i + 2];
        mine[5 * i + 3] = score[i]; mine[5 * i + 4] = l[i];
    }
    if (!r) {
        if (c->dist) r = dist_allgather(c, mine, 5 * n, rows_out);
        else memcpy(rows_out, mine, sizeof(double) * 5 * n);
    }
    if (!r && c->fused_register) r = gm_register(c, ranges, poses, NULL); /* the scan goes into every local map at its matched pose */
    free(poses); free(score); free(l); free(mine);
    if (r) { free(rows_out); free(weights_out); return r; }
    double* logw = (double*)malloc(sizeof(double) * N);
    double* w = (double*)malloc(sizeof(double) * N);
    for (uint32_t i = 0; i < N; i++) logw[i] = log_weights[i] + rows_out[5 * (size_t)i + 4];
    const double neff = orc_normalize(logw, N, gain, w);
    if (weights_out) memcpy(weights_out, w, sizeof(double) * N);
    if (neff_out) *neff_out = neff;
    free(logw); free(w);
    c->sw_rows = rows_out; c->sw_w = weights_out; c->sw_neff = neff_keep;
    return GM_OK;
}

int gm_scanmatch_weights_end(gm_ctx* c, double* rows_out, double* weights_out, double* neff_out) {
    if (!c) return GM_ERR_ARG;
    if (!c->sw_rows) return fail(c, GM_ERR_ARG, "gm_scanmatch_weights_end: nothing to collect");
    if (!rows_out) return fail(c, GM_ERR_ARG, "gm_scanmatch_weights_end: null argument");
    memcpy(rows_out, c->sw_rows, sizeof(double) * 5 * c->sw_total);
    if (weights_out) memcpy(weights_out, c->sw_w, sizeof(double) * c->sw_total);
    if (neff_out) *neff_out = c->sw_neff;
    free(c->sw_rows); free(c->sw_w); c->sw_rows = NULL; c->sw_w = NULL;
    return GM_OK;
}

int gm_scanmatch_weights(gm_ctx* c, const double* ranges, const double* poses_local, double minimum_score, const double* log_weights,
                         double gain, double* rows_out, double* weights_out, double* neff_out) {
    int r = gm_scanmatch_weights_begin(c, ranges, poses_local, minimum_score, log_weights, gain);
    if (r) return r;
    return gm_scanmatch_weights_end(c, rows_out, weights_out, neff_out);
}
