/* gm_b200.h -- C ABI of the B200-native GMapping per-scan particle update.
 *
 * One gm_ctx per GMapping::GridSlamProcessor (per GPU).  The context owns, in HBM: the patch pool
 * (32x32-cell patches: 16-byte cells {acc, n}, a visit-counter plane, and two derived planes -- per-cell means and an
 * occupancy bitmap -- 36.1 KiB per patch), per-particle patch directories + map geometry, refcounts,
 * the beam table and the matcher parameters.  The host owns poses, weights, the trajectory tree and
 * the drand48 stream.
 *
 * All entry points: plain pointers and sizes, host buffers owned by the caller and copied during the
 * call, return 0 on success or a negative gm_status; gm_last_error() gives the message.  Calls on one
 * context are serialised by the caller (the reference is single-threaded per processor) and are
 * synchronous at return.  There is NO CPU fallback: without a CUDA device gm_create fails.
 *
 * "replaces:" cites the reference code each entry point stands in for (paths relative to
 * openslam_gmapping/ in the reference tree).
 */
#ifndef GM_B200_H
#define GM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GM_MAXBEAMS 2048 /* LASER_MAXBEAMS, include/gmapping/scanmatcher/scanmatcher.h:11 */
#define GM_PATCH 32      /* 2^5 cells per patch side, include/gmapping/grid/harray2d.h:28 */

typedef enum gm_status {
    GM_OK = 0,
    GM_ERR_CUDA = -1,       /* a CUDA runtime call failed */
    GM_ERR_ARG = -2,        /* bad argument / call order */
    GM_ERR_POOL = -3,       /* patch pool exhausted */
    GM_ERR_NO_DEVICE = -4,  /* no usable CUDA device (there is no CPU path) */
    GM_ERR_LINE = -5,       /* a ray longer than the reference's 20000-point buffer (scanmatcher.cpp:55) */
    GM_ERR_INTERNAL = -6    /* gm_map_digest: the occupancy bitmap disagrees with the cells (a bug, never expected) */
} gm_status;

typedef struct gm_ctx gm_ctx;

typedef struct gm_config {
    int32_t device;         /* CUDA device ordinal */
    uint32_t max_particles; /* capacity; gm_init_particles may use fewer */
    uint64_t pool_patches;  /* patch-pool capacity (36.1 KiB each); 0 = half of the free HBM */
    uint32_t block_threads; /* threads per particle CTA, 0 = default (256) */
    uint32_t flags;         /* reserved, 0 */
    uint32_t global_particles; /* multi-GPU: particles over all ranks (sizes the weight buffers); 0 = max_particles */
    uint32_t reserved;
} gm_config;

/* replaces: the ScanMatcher parameter block, include/gmapping/scanmatcher/scanmatcher.h:61-82,
 * defaults scanmatcher/scanmatcher.cpp:17-56, setMatchingParameters scanmatcher.cpp:872-883 */
typedef struct gm_match_params {
    double laser_max_range, usable_range, gaussian_sigma, likelihood_sigma;
    int32_t kernel_size;
    int32_t generate_map;
    double opt_linear_delta, opt_angular_delta;
    uint32_t opt_recursive_iterations, likelihood_skip;
    double enlarge_step, fullness_threshold, angular_odometry_reliability, linear_odometry_reliability, free_cell_ratio;
    uint32_t initial_beams_skip;
    uint32_t reserved;
} gm_match_params;

/* the reference's cell, include/gmapping/scanmatcher/smmap.h:35-36 (sizeof == 16) */
typedef struct gm_cell { float acc_x, acc_y; int32_t n, visits; } gm_cell;

/* per-particle map geometry, include/gmapping/grid/map.h:137-141 */
typedef struct gm_map_info {
    double center_x, center_y, delta;
    int32_t size_x2, size_y2;     /* m_sizeX2 / m_sizeY2 */
    int32_t patches_x, patches_y; /* directory dims (storage xsize/ysize); map size = patches << 5 */
    int32_t allocated_patches;
    int32_t reserved;
} gm_map_info;

/* order-independent map digest in world-cell coordinates (test / consistency aid); gm_map_digest also verifies the derived
 * occupancy bitmap the scan matcher reads against the cells and returns GM_ERR_INTERNAL when they disagree */
typedef struct gm_digest {
    int64_t patches, cells, sum_n, sum_visits;
    uint64_t hash_counts, hash_acc;
} gm_digest;

typedef struct gm_stats {
    uint64_t pool_capacity, pool_in_use;
    uint64_t score_evals;     /* sum over particles of score() evaluations in the last gm_scanmatch (incl. the initial one) */
    uint64_t beam_evals;      /* (particle, pose, valid beam) window searches in the last gm_scanmatch incl. likelihoodAndScore */
    uint64_t cow_copies, patch_allocs; /* last gm_register */
    uint64_t free_updates, hit_updates; /* last gm_register: visits++ along rays / endpoint updates */
    uint64_t resizes;         /* particles whose map was resized in the last gm_register */
    float ms_scanmatch, ms_register, ms_remap, ms_weights; /* CUDA-event device time of the last calls */
    float ms_migrate;   /* multi-GPU: device time of the last patch migration (pull kernels) */
    float ms_allgather; /* multi-GPU: HOST ms the last gm_scanmatch_weights waited, after its own kernel, for the other ranks' rows */
    uint64_t migrated_particles, migrated_bytes; /* last migration: imported parents, bytes read from peers */
    uint32_t launches;        /* kernels launched by this context since gm_create */
    uint32_t reserved;
    /* last gm_register: SM cycles summed over CTAs per phase of k_register (beam setup, ray marking, ranks, copy-on-write,
     * ray counting, tile flush, hit sort, hit accumulation) */
    uint64_t register_phase_cycles[8];
    uint32_t max_active_patches; /* last gm_register: largest active area (patches) of a particle */
    uint32_t register_tile_cap;  /* last gm_register: visit-count tiles per pass (shared memory) */
    /* running totals since gm_create: device ms of scanmatch, weights, remap, register, migrate, allgather; work; call counts */
    double total_ms[6];
    uint64_t total_beam_evals, total_score_evals, total_scanmatches, total_registers;
    uint64_t total_migrated_particles, total_migrated_bytes; /* multi-GPU, since gm_create */
    /* of the pool's patch ids, those that can hold end points (all planes, 36 KB each; the others are free-space patches: visit
     * counters only, 4 KB) and how many of them were in use after the last gm_register */
    uint64_t pool_hit_capacity, pool_hit_in_use;
} gm_stats;

int gm_create(gm_ctx** out, const gm_config* cfg);
int gm_destroy(gm_ctx* ctx);
const char* gm_last_error(const gm_ctx* ctx); /* ctx may be NULL: last gm_create failure */

/* replaces: ScanMatcher::setLaserParameters, scanmatcher/scanmatcher.cpp:699-709 */
int gm_set_laser(gm_ctx* ctx, uint32_t nbeams, const double* angles, const double laser_pose[3]);
/* replaces: ScanMatcher::setMatchingParameters + the PARAM_SET_GET setters */
int gm_set_matching(gm_ctx* ctx, const gm_match_params* p);
void gm_default_matching(gm_match_params* p);

/* replaces: GridSlamProcessor::init's map/particle construction, gridfastslam/gridslamprocessor.cpp:354-403
 * (Map(center, worldSizeX, worldSizeY, delta), include/gmapping/grid/map.h:171-184) */
int gm_init_particles(gm_ctx* ctx, uint32_t n, const double center[2], double world_w, double world_h, double delta);

/* replaces: ScanMatcher::score, include/gmapping/scanmatcher/scanmatcher.h:171-259, for `count`
 * (particle map, pose) queries.  particle_idx[i] selects whose map is read. */
int gm_score(gm_ctx* ctx, uint32_t count, const uint32_t* particle_idx, const double* poses /*count*3*/,
             const double* ranges, double* score_out);
/* replaces: ScanMatcher::likelihoodAndScore, scanmatcher.h:271-380 */
int gm_likelihood_and_score(gm_ctx* ctx, uint32_t count, const uint32_t* particle_idx, const double* poses,
                            const double* ranges, double* s_out, double* l_out, uint32_t* c_out);

/* replaces: ScanMatcher::optimize(pnew, map, init, readings), scanmatcher/scanmatcher.cpp:386-529, for `count`
 * (particle map, initial pose) queries: poses_out[i] = the hill-climbed pose, score_out[i] = its score. */
int gm_optimize(gm_ctx* ctx, uint32_t count, const uint32_t* particle_idx, const double* poses_in, const double* ranges,
                double* poses_out, double* score_out);

/* replaces: GridSlamProcessor::scanMatch, include/gmapping/gridfastslam/gridslamprocessor.hxx:15-75
 * (ScanMatcher::optimize scanmatcher.cpp:386-529, then likelihoodAndScore at the accepted pose).
 * poses_inout: N*3 motion-model poses in, scan-matched poses out (kept when score <= minimum_score).
 * score_out: optimize()'s return; s_out/l_out/c_out: likelihoodAndScore at the final pose;
 * evals_out: number of score() evaluations per particle.  Any *_out may be NULL. */
int gm_scanmatch(gm_ctx* ctx, const double* ranges, double* poses_inout, double minimum_score,
                 double* score_out, double* s_out, double* l_out, uint32_t* c_out, uint32_t* evals_out);

/* replaces: GridSlamProcessor::scanMatch (gridslamprocessor.hxx:15-75) together with the normalize() of the updateTreeWeights
 * that follows it (gridslamprocessor.cpp:586, hxx:82-121), as ONE stream-ordered sequence with a single host synchronisation:
 * scan matching of the n local particles at poses_local (n*3), log_weights[i] + l[i] for all N particles (the reference's
 * `weight += l`, hxx:50-52; log_weights = Particle::weight before this scan), the normalised weights and Neff.
 * rows_out: N * 5 doubles = x, y, theta, optimize()'s score, l per particle.  On a sharded context (gm_dist_init) N = n * world
 * in global particle order: every rank's rows arrive in every rank's HBM from the scan-match kernel's epilogue (peer stores
 * over NVLink, no separate collective), and every rank computes the same weights.  weights_out (N) / neff_out may be NULL. */
int gm_scanmatch_weights(gm_ctx* ctx, const double* ranges, const double* poses_local, double minimum_score,
                         const double* log_weights /*N*/, double obs_sigma_gain, double* rows_out /*N*5*/,
                         double* weights_out /*N*/, double* neff_out);

/* gm_scanmatch_weights in two halves: _begin uploads and queues the kernels and returns at once, _end waits and hands out the
 * results.  Host work that does not need the results (GridSlamProcessor grows the trajectory tree there) overlaps the scan
 * match; no other entry point may be called on the context in between. */
int gm_scanmatch_weights_begin(gm_ctx* ctx, const double* ranges, const double* poses_local, double minimum_score,
                               const double* log_weights /*N*/, double obs_sigma_gain);
int gm_scanmatch_weights_end(gm_ctx* ctx, double* rows_out /*N*5*/, double* weights_out /*N*/, double* neff_out);

/* replaces: GridSlamProcessor::normalize, gridslamprocessor.hxx:82-121 (sequential sums kept) */
int gm_normalize_neff(gm_ctx* ctx, const double* log_weights, uint32_t n, double obs_sigma_gain,
                      double* weights_out, double* neff_out);
/* replaces: uniform_resampler<double,double>::resampleIndexes, include/gmapping/particlefilter/particlefilter.h:180-229;
 * u01 is the caller's drand48() draw.  emitted_out: how many indexes the walk produced (== n normally). */
int gm_resample_indexes(gm_ctx* ctx, const double* weights, uint32_t n, double u01, uint32_t* idx_out, uint32_t* emitted_out);
/* replaces: the same with its `nparticles` argument set -- GridSlamProcessor::processScan(reading, adaptParticles),
 * include/gmapping/gridfastslam/gridslamprocessor.hxx:153-156 -> particlefilter.h:196-203: the interval is the total weight over
 * n_out and n_out indexes come out (idx_out has n_out entries). */
int gm_resample_indexes_n(gm_ctx* ctx, const double* weights, uint32_t n, uint32_t n_out, double u01, uint32_t* idx_out, uint32_t* emitted_out);

/* replaces: the particle-copy + registerScan loops of GridSlamProcessor::resample, gridslamprocessor.hxx:184-296,
 * and the first-scan loop gridslamprocessor.cpp:617-633: (optional) remap particle j <- parent idx[j] by
 * patch-reference copy, then computeActiveArea (scanmatcher.cpp:83-254) + registerScan (:266-354) for all. */
int gm_register(gm_ctx* ctx, const double* ranges, const double* poses /*N*3*/, const uint32_t* idx_or_null);
/* replaces: the particle copy alone of GridSlamProcessor::resample, gridslamprocessor.hxx:184-213 (Particle copy =
 * HierarchicalArray2D copy by patch reference): particle j <- parent idx[j].  For callers that register FIRST:
 * gm_register(ranges, poses, NULL) for every particle at its scan-matched pose, then gm_copy_particles(idx), gives the maps
 * of gm_register(ranges, poses_of_parents, idx) -- a child starts from its parent's pose and map, so registering the scan
 * before or after the copy writes the same cells -- and lets the registration start before Neff is known. */
int gm_copy_particles(gm_ctx* ctx, const uint32_t* idx /*N*/);
/* the same when the resampling changes the particle count (adaptParticles): the new generation has n_new particles
 * (1 <= n_new <= gm_config.max_particles), particle j continues parent idx[j] of the old one.  Single-GPU contexts only. */
int gm_copy_particles_n(gm_ctx* ctx, const uint32_t* idx /*n_new*/, uint32_t n_new);

/* replaces: ScanMatcher::computeActiveArea, scanmatcher/scanmatcher.cpp:83-254, as far as it is observable: the
 * bounding box of the scan at each particle's pose and the Map::resize it triggers.  (The active patch set itself
 * is recomputed inside gm_register, like registerScan does after invalidateActiveArea.) */
int gm_compute_active_area(gm_ctx* ctx, const double* ranges, const double* poses /*N*3*/);

/* the inverse of gm_download_map: replace one particle's map by host content (geometry from info->size_x2/size_y2/
 * patches_x/patches_y, `npatches` patches at patch_coords with 1024 cells each).  Used to make a host-built
 * ScanMatcherMap device-resident (standalone ScanMatcher users such as slam_gmapping.cpp:866-927). */
int gm_upload_map(gm_ctx* ctx, uint32_t particle, const gm_map_info* info, const int32_t* patch_coords, const gm_cell* cells,
                  uint32_t npatches);

/* replaces: host access to Particle::map (ScanMatcherMap cells / geometry).  Fills up to max_patches patches:
 * patch_coords[2k..2k+1] = (px,py) and cells[k*1024 + lx*32 + ly].  Returns patch count via info. */
int gm_map_info_get(gm_ctx* ctx, uint32_t particle, gm_map_info* info);
int gm_download_map(gm_ctx* ctx, uint32_t particle, gm_map_info* info, int32_t* patch_coords, gm_cell* cells, uint32_t max_patches);
/* replaces: the occupancy-grid fill of the ROS wrapper, gmapping/src/slam_gmapping.cpp:956-993:
 * out[y * width + x] = -1 never visited, 100 if n/visits > occ_thresh, else 0; width/height = the particle's map size in cells
 * (gm_map_info: patches_x * 32, patches_y * 32). */
int gm_export_occupancy(gm_ctx* ctx, uint32_t particle, double occ_thresh, int8_t* out, uint32_t width, uint32_t height);
int gm_map_digest(gm_ctx* ctx, uint32_t first, uint32_t count, gm_digest* out /*count*/);
/* replaces: the bookkeeping behind ScanMatcher::registerScan's return value (scanmatcher/scanmatcher.cpp:317-341: the summed
 * change of PointAccumulator::entropy(), include/gmapping/scanmatcher/smmap.h:68-79, over every cell update of the scan).  The sum
 * telescopes per cell: out[i] = sum over the visited cells of map first + i of (entropy - log 2); the return value of a
 * registration is this figure after it minus before it (equal to the reference's sequential sum up to rounding). */
int gm_map_entropy(gm_ctx* ctx, uint32_t first, uint32_t count, double* out /*count*/);
/* replaces: the MAP_CONSISTENCY_CHECK block of GridSlamProcessor::processScan, gridfastslam/gridslamprocessor.cpp:82-143 (every
 * patch's share count against the number of particle maps that point at it): the references are recounted from the patch
 * directories of the live particles and compared with the pool's reference counts, and "referenced" with "off the free stack".
 * GM_ERR_INTERNAL on any difference.  mismatches_out / referenced_out may be NULL. */
int gm_check_refcounts(gm_ctx* ctx, uint64_t* mismatches_out, uint64_t* referenced_out);

/* ---- multi-GPU (one process and one gm_ctx per GPU of one node; peer memory over NVLink / NVSwitch).  Rank r owns the global
 * particles [r*n, (r+1)*n), n = gm_particles(ctx); the host state (poses, weights, trajectory tree, drand48 stream) is
 * replicated.  Every rank's patch pool, directories and result buffers are mapped into every other rank with CUDA IPC; a small
 * POSIX shared-memory segment named after `id` carries the handles and the flags the ranks synchronise on.
 * replaces: nothing in the reference (it is single-threaded); see DESIGN.md "multi-GPU". */
#define GM_DIST_ID_BYTES 128
int gm_dist_unique_id(void* id_out /* GM_DIST_ID_BYTES */);  /* random session id: call on one rank, hand the bytes to all */
/* after gm_init_particles, collectively on all ranks: exports this rank's allocations, maps everybody else's */
int gm_dist_init(gm_ctx* ctx, int rank, int world, const void* id /* GM_DIST_ID_BYTES */);
/* gm_register across ranks: idx_global[N] are the resampled parents as GLOBAL particle indexes (identical on every
 * rank) or NULL.  Parents that live on another rank are pulled here first: a kernel on this GPU reads the owner's patch
 * directory and patches out of the owner's HBM (peer loads) into spare particle slots, copying a patch that several imported
 * parents share once; then the local patch-reference remap and the registration run as in gm_register.
 * poses: the n local particles. */
int gm_dist_register(gm_ctx* ctx, const double* ranges, const double* poses, const uint32_t* idx_global);
/* gm_copy_particles across ranks: the migration and the remap of gm_dist_register without the registration */
int gm_dist_copy_particles(gm_ctx* ctx, const uint32_t* idx_global /*N global*/);
/* planning only (no device work): which global parents this rank must import for idx_global, in import order;
 * returns the count, fills parents_out (capacity n) -- used by the tests of the host-side logic */
int gm_dist_plan(uint32_t n_local, int rank, int world, const uint32_t* idx_global, uint32_t* parents_out, uint32_t* local_idx_out);

/* replaces: the map part of GridSlamProcessor's copy constructor / clone(), gridfastslam/gridslamprocessor.cpp:38-97
 * (SURVEY.md §8 row f3): a second context on the same GPU with identical particle maps (the whole pool is copied). */
int gm_clone(gm_ctx* src, gm_ctx** out);

/* Asynchronous registration: with on != 0, gm_register / gm_dist_register return as soon as the kernels are launched so
 * that the caller's host-side bookkeeping (trajectory tree, next motion-model draws) overlaps them.  Counters, timing and
 * device-side errors (pool exhausted) of that registration are reported by the NEXT call on the context -- every entry
 * point first collects a pending registration -- or by gm_synchronize.  Default: off (every call is complete at return). */
int gm_set_async_register(gm_ctx* ctx, int on);
/* With on != 0, gm_scanmatch_weights also registers the scan in every local particle at its scan-matched pose -- what
 * gm_register(ranges, matched poses, NULL) would do, i.e. the caller registers FIRST and copies resampled particles afterwards
 * (gm_copy_particles) -- queued on the device right behind the scan match, while the weights are normalised and handed to the
 * host on a second stream: on a sharded context the wait for the other ranks' rows is then hidden behind the registration.
 * The registration is collected like an asynchronous one (by the next entry point or gm_synchronize).  Default: off. */
int gm_set_fused_register(gm_ctx* ctx, int on);
int gm_synchronize(gm_ctx* ctx);

int gm_get_stats(gm_ctx* ctx, gm_stats* out);
uint32_t gm_particles(const gm_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
