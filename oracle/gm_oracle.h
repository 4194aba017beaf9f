/* oracle/gm_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Plain-C CPU restatement of GMapping's per-scan particle update, used ONLY as the checker in
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The product path
 * (slam-gmapping-learn_b200/csrc) never links, loads or calls this.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this restatement bit-for-bit
 * against traces produced by the unmodified reference build (oracle/_ref/ref_driver, committed as
 * tests/golden/ (npz) together with the tools that made them).
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/openslam_gmapping).
 */
#ifndef GM_ORACLE_H
#define GM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAXBEAMS 2048 /* include/gmapping/scanmatcher/scanmatcher.h:11 */
#define ORC_PATCH_MAG 5
#define ORC_PATCH 32

typedef struct { float ax, ay; int32_t n, visits; } orc_cell; /* smmap.h:35-36 (16 B) */

typedef struct orc_map {
    double cx, cy, world_w, world_h, delta; /* map.h:137-141 */
    int map_w, map_h, size_x2, size_y2;     /* cells */
    int npx, npy;                           /* patch directory dims */
    orc_cell** dir;                         /* [npx*npy], index px*npy+py, NULL = unallocated */
    uint8_t* active;                        /* [npx*npy] active-area mask (harray2d.h:137-149) */
} orc_map;

typedef struct orc_matcher { /* scanmatcher.h:53-86, scanmatcher.cpp:17-56 */
    uint32_t beams;
    double angles[ORC_MAXBEAMS];
    double laser_pose[3];
    double laser_max_range, usable_range, gaussian_sigma, likelihood_sigma;
    int32_t kernel_size;
    double opt_angular_delta, opt_linear_delta;
    uint32_t opt_recursive_iterations, likelihood_skip;
    int32_t generate_map;
    double enlarge_step, fullness_threshold, angular_odometry_reliability, linear_odometry_reliability, free_cell_ratio;
    uint32_t initial_beams_skip;
    /* statistics (ours) */
    uint64_t score_calls, beam_evals;
} orc_matcher;

typedef struct { int64_t patches, cells, sum_n, sum_visits; uint64_t hash_counts, hash_acc; double sum_accx, sum_accy; } orc_digest;

void orc_matcher_defaults(orc_matcher* m);

orc_map* orc_map_new(double cx, double cy, double world_w, double world_h, double delta);
orc_map* orc_map_clone(const orc_map* m);
void orc_map_free(orc_map* m);
void orc_map_world2map(const orc_map* m, double x, double y, int* ix, int* iy);
void orc_map_resize(orc_map* m, double xmin, double ymin, double xmax, double ymax);
int64_t orc_map_patch_count(const orc_map* m);
/* export allocated patches: coords[2*k] (patch px,py) and cells[k*1024] (lx*32+ly); returns count */
int64_t orc_map_export(const orc_map* m, int32_t* coords, orc_cell* cells, int64_t max_patches);
void orc_map_digest(const orc_map* m, orc_digest* out);

double orc_score(orc_matcher* sm, const orc_map* map, const double pose[3], const double* ranges);
uint32_t orc_likelihood_and_score(orc_matcher* sm, const orc_map* map, const double pose[3], const double* ranges,
                                  double* s, double* l);
double orc_optimize(orc_matcher* sm, const orc_map* map, const double init[3], const double* ranges, double pnew[3]);
void orc_compute_active_area(const orc_matcher* sm, orc_map* map, const double pose[3], const double* ranges);
void orc_register_scan(const orc_matcher* sm, orc_map* map, const double pose[3], const double* ranges);
int orc_grid_line(int x0, int y0, int x1, int y1, int32_t* pts /* 2*n */, int max_pts);

double orc_normalize(const double* logw, uint32_t n, double obs_sigma_gain, double* w_out);
uint32_t orc_resample_indexes(const double* w, uint32_t n, double u01, uint32_t* idx_out);
uint32_t orc_resample_indexes_n(const double* w, uint32_t n, uint32_t n_out, double u01, uint32_t* idx_out);

void orc_seed(long seed);
double orc_sample_gaussian(double sigma);
void orc_draw_from_motion(const double p[3], const double pnew[3], const double pold[3], double srr, double srt,
                          double str, double stt, double out[3]);

/* ---- whole filter (gridslamprocessor.cpp:425-662) ---- */
typedef struct orc_filter orc_filter;
orc_filter* orc_filter_new(uint32_t particles, double xmin, double ymin, double xmax, double ymax, double delta,
                           const double init_pose[3], const orc_matcher* sm);
void orc_filter_free(orc_filter* f);
void orc_filter_set_motion(orc_filter* f, double srr, double srt, double str, double stt);
void orc_filter_set_update(orc_filter* f, double linear, double angular, double resample_thr, double period,
                           double minimum_score, double obs_sigma_gain);
int orc_filter_process(orc_filter* f, const double* ranges, const double odom_pose[3], double stamp);
uint32_t orc_filter_size(const orc_filter* f);
/* rows of 6: x y theta weight weightSum previousIndex */
void orc_filter_state(const orc_filter* f, double* rows);
void orc_filter_pre_poses(const orc_filter* f, double* rows3); /* poses right after the motion model */
/* rows of 8 from the last scanMatch: optimize score, corrected x y th, s, l, c, evals */
void orc_filter_match_rows(const orc_filter* f, double* rows8);
double orc_filter_neff(const orc_filter* f);
void orc_filter_weights(const orc_filter* f, double* w);
int orc_filter_last_resampled(const orc_filter* f, uint32_t* idx_out, double* neff_at_resample, double* w_at_resample);
const orc_map* orc_filter_map(const orc_filter* f, uint32_t i);
int orc_filter_best(const orc_filter* f);
orc_matcher* orc_filter_matcher(orc_filter* f);
/* the drand48() draw consumed by the last resample (particlefilter.h:199) */
double orc_filter_last_u01(const orc_filter* f);

#ifdef __cplusplus
}
#endif
#endif
