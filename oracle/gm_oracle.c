/* oracle/gm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT.  See gm_oracle.h.
 *
 * Plain-C restatement of the reference's CPU algorithm for the per-scan particle update.
 * The floating-point operation ORDER of every expression follows the reference so that, built
 * with -O2 -ffp-contract=off on x86-64 against the same libm, results are bit-identical to the
 * reference build (checked in tests/test_oracle_vs_reference.py against golden traces).
 *
 * The data structures are ours (flat patch directory, owning maps, no refcounted pointers);
 * only the arithmetic and the control flow are the reference's.
 */
#define _XOPEN_SOURCE 700
#define _DEFAULT_SOURCE
#include "gm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ map */

static const orc_cell k_unknown = {0.f, 0.f, 0, 0}; /* map.h:142, smmap.h:19 (Cell(-1): all zero) */

/* Map(center, worldSizeX, worldSizeY, delta): map.h:171-184 + harray2d.h:54-59 */
orc_map* orc_map_new(double cx, double cy, double world_w, double world_h, double delta) {
    orc_map* m = (orc_map*)calloc(1, sizeof *m);
    m->cx = cx; m->cy = cy; m->world_w = world_w; m->world_h = world_h; m->delta = delta;
    int xs = (int)ceil(world_w / delta), ys = (int)ceil(world_h / delta);
    m->npx = xs >> ORC_PATCH_MAG; m->npy = ys >> ORC_PATCH_MAG;
    m->map_w = m->npx << ORC_PATCH_MAG; m->map_h = m->npy << ORC_PATCH_MAG;
    m->size_x2 = m->map_w >> 1; m->size_y2 = m->map_h >> 1;
    m->dir = (orc_cell**)calloc((size_t)m->npx * m->npy, sizeof(orc_cell*));
    m->active = (uint8_t*)calloc((size_t)m->npx * m->npy, 1);
    return m;
}

/* particle copy: harray2d.h:63-77 shares patches by refcount and allocActiveArea (harray2d.h:169-185)
 * copies before any write, so a deep copy is observationally identical.  The copy constructor does
 * NOT carry m_activeArea over. */
orc_map* orc_map_clone(const orc_map* s) {
    orc_map* m = (orc_map*)malloc(sizeof *m);
    *m = *s;
    size_t n = (size_t)s->npx * s->npy;
    m->dir = (orc_cell**)calloc(n, sizeof(orc_cell*));
    m->active = (uint8_t*)calloc(n, 1);
    for (size_t i = 0; i < n; i++)
        if (s->dir[i]) {
            m->dir[i] = (orc_cell*)malloc(sizeof(orc_cell) * ORC_PATCH * ORC_PATCH);
            memcpy(m->dir[i], s->dir[i], sizeof(orc_cell) * ORC_PATCH * ORC_PATCH);
        }
    return m;
}

void orc_map_free(orc_map* m) {
    if (!m) return;
    size_t n = (size_t)m->npx * m->npy;
    for (size_t i = 0; i < n; i++) free(m->dir[i]);
    free(m->dir); free(m->active); free(m);
}

/* map.h:283-287 -- round() is half-away-from-zero */
void orc_map_world2map(const orc_map* m, double x, double y, int* ix, int* iy) {
    *ix = (int)round((x - m->cx) / m->delta) + m->size_x2;
    *iy = (int)round((y - m->cy) / m->delta) + m->size_y2;
}

/* map.h:295-298 */
static void map2world(const orc_map* m, int ix, int iy, double* x, double* y) {
    *x = (ix - m->size_x2) * m->delta + m->cx;
    *y = (iy - m->size_y2) * m->delta + m->cy;
}

/* harray2d.h:203-211 + array2d.h isInside */
static int patch_of(const orc_map* m, int x, int y, int* px, int* py) {
    if (x >= 0 && y >= 0) { *px = x >> ORC_PATCH_MAG; *py = y >> ORC_PATCH_MAG; }
    else { *px = -1; *py = -1; }
    return *px >= 0 && *py >= 0 && *px < m->npx && *py < m->npy;
}

static int is_inside_cell(const orc_map* m, int x, int y) { int px, py; return patch_of(m, x, y, &px, &py); }

static int is_inside_world(const orc_map* m, double x, double y) { /* map.h:111-114 */
    int ix, iy; orc_map_world2map(m, x, y, &ix, &iy); return is_inside_cell(m, ix, iy);
}

/* const Map::cell(IntPoint): map.h:347-354 -- unknown when outside or unallocated */
static const orc_cell* cell_const(const orc_map* m, int x, int y) {
    int px, py;
    if (!patch_of(m, x, y, &px, &py)) return &k_unknown;
    const orc_cell* p = m->dir[(size_t)px * m->npy + py];
    if (!p) return &k_unknown;
    return &p[((x - (px << ORC_PATCH_MAG)) << ORC_PATCH_MAG) + (y - (py << ORC_PATCH_MAG))];
}

/* non-const HierarchicalArray2D::cell: harray2d.h:221-238 -- allocates the patch on demand */
static orc_cell* cell_mut(orc_map* m, int x, int y) {
    int px, py;
    if (!patch_of(m, x, y, &px, &py)) return NULL; /* reference asserts */
    orc_cell** slot = &m->dir[(size_t)px * m->npy + py];
    if (!*slot) *slot = (orc_cell*)calloc(ORC_PATCH * ORC_PATCH, sizeof(orc_cell));
    return &(*slot)[((x - (px << ORC_PATCH_MAG)) << ORC_PATCH_MAG) + (y - (py << ORC_PATCH_MAG))];
}

/* Map::resize: map.h:218-241 + HierarchicalArray2D::resize: harray2d.h:80-109 */
void orc_map_resize(orc_map* m, double xmin, double ymin, double xmax, double ymax) {
    int iminx, iminy, imaxx, imaxy;
    orc_map_world2map(m, xmin, ymin, &iminx, &iminy);
    orc_map_world2map(m, xmax, ymax, &imaxx, &imaxy);
    int pxmin = (int)floor((float)iminx / (1 << ORC_PATCH_MAG));
    int pxmax = (int)ceil((float)imaxx / (1 << ORC_PATCH_MAG));
    int pymin = (int)floor((float)iminy / (1 << ORC_PATCH_MAG));
    int pymax = (int)ceil((float)imaxy / (1 << ORC_PATCH_MAG));
    int nx = pxmax - pxmin, ny = pymax - pymin;
    orc_cell** nd = (orc_cell**)calloc((size_t)nx * ny, sizeof(orc_cell*));
    int dx = pxmin < 0 ? 0 : pxmin, dy = pymin < 0 ? 0 : pymin;
    int Dx = pxmax < m->npx ? pxmax : m->npx, Dy = pymax < m->npy ? pymax : m->npy;
    for (int x = 0; x < m->npx; x++)
        for (int y = 0; y < m->npy; y++) {
            orc_cell* p = m->dir[(size_t)x * m->npy + y];
            if (x >= dx && x < Dx && y >= dy && y < Dy) nd[(size_t)(x - pxmin) * ny + (y - pymin)] = p;
            else free(p); /* patches cropped away lose their last reference */
        }
    free(m->dir); free(m->active);
    m->dir = nd; m->npx = nx; m->npy = ny;
    m->active = (uint8_t*)calloc((size_t)nx * ny, 1);
    m->map_w = nx << ORC_PATCH_MAG; m->map_h = ny << ORC_PATCH_MAG;
    m->world_w = xmax - xmin; m->world_h = ymax - ymin;
    m->size_x2 -= pxmin * (1 << ORC_PATCH_MAG);
    m->size_y2 -= pymin * (1 << ORC_PATCH_MAG);
}

int64_t orc_map_patch_count(const orc_map* m) {
    int64_t c = 0; size_t n = (size_t)m->npx * m->npy;
    for (size_t i = 0; i < n; i++) c += m->dir[i] != NULL;
    return c;
}

int64_t orc_map_export(const orc_map* m, int32_t* coords, orc_cell* cells, int64_t max_patches) {
    int64_t k = 0;
    for (int px = 0; px < m->npx; px++)
        for (int py = 0; py < m->npy; py++) {
            const orc_cell* p = m->dir[(size_t)px * m->npy + py];
            if (!p) continue;
            if (k < max_patches) {
                coords[2 * k] = px; coords[2 * k + 1] = py;
                memcpy(cells + k * ORC_PATCH * ORC_PATCH, p, sizeof(orc_cell) * ORC_PATCH * ORC_PATCH);
            }
            k++;
        }
    return k;
}

static uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

/* order-independent digest in world-cell coordinates (same definition as oracle/ref_driver.cpp) */
void orc_map_digest(const orc_map* m, orc_digest* d) {
    memset(d, 0, sizeof *d);
    for (int px = 0; px < m->npx; px++)
        for (int py = 0; py < m->npy; py++) {
            const orc_cell* p = m->dir[(size_t)px * m->npy + py];
            if (!p) continue;
            d->patches++;
            for (int lx = 0; lx < ORC_PATCH; lx++)
                for (int ly = 0; ly < ORC_PATCH; ly++) {
                    const orc_cell* a = &p[lx * ORC_PATCH + ly];
                    if (!a->visits && !a->n) continue;
                    int32_t wx = px * ORC_PATCH + lx - m->size_x2, wy = py * ORC_PATCH + ly - m->size_y2;
                    d->cells++; d->sum_n += a->n; d->sum_visits += a->visits; d->sum_accx += a->ax; d->sum_accy += a->ay;
                    uint64_t key = ((uint64_t)(uint32_t)wx << 32) | (uint32_t)wy;
                    d->hash_counts += mix64(key ^ mix64(((uint64_t)(uint32_t)a->n << 32) | (uint32_t)a->visits));
                    uint32_t ax, ay; memcpy(&ax, &a->ax, 4); memcpy(&ay, &a->ay, 4);
                    d->hash_acc += mix64(key ^ mix64(((uint64_t)ax << 32) | ay) ^ 0x5555);
                }
        }
}

/* -------------------------------------------------------------------------------------- matcher */

void orc_matcher_defaults(orc_matcher* m) { /* scanmatcher.cpp:17-56 */
    memset(m, 0, sizeof *m);
    m->opt_recursive_iterations = 3;
    m->enlarge_step = 10.;
    m->fullness_threshold = 0.1;
    m->free_cell_ratio = sqrt(2.);
    m->kernel_size = 1;
    m->generate_map = 1;
}

/* PointAccumulator::operator double: smmap.h:25 */
static inline double occ(const orc_cell* c) { return c->visits ? (double)c->n / (double)c->visits : -1; }

/* laser pose in the map frame: scanmatcher.h:177-180 */
static void laser_pose(const orc_matcher* sm, const double p[3], double lp[3]) {
    lp[0] = p[0]; lp[1] = p[1]; lp[2] = p[2];
    lp[0] += cos(p[2]) * sm->laser_pose[0] - sin(p[2]) * sm->laser_pose[1];
    lp[1] += sin(p[2]) * sm->laser_pose[0] + cos(p[2]) * sm->laser_pose[1];
    lp[2] += sm->laser_pose[2];
}

/* window search shared by score / likelihoodAndScore: scanmatcher.h:213-252 and :325-365.
 * free_off is `delta*freeDelta` for score (h:208) and `freeDelta` for likelihoodAndScore (h:320). */
static int window_match(const orc_matcher* sm, const orc_map* map, const double lp[3], double r, double angle,
                        double free_off, double* mux, double* muy) {
    double phx = lp[0], phy = lp[1];
    phx += r * cos(lp[2] + angle);
    phy += r * sin(lp[2] + angle);
    int ihx, ihy; orc_map_world2map(map, phx, phy, &ihx, &ihy);
    double pfx = lp[0], pfy = lp[1];
    pfx += (r - free_off) * cos(lp[2] + angle);
    pfy += (r - free_off) * sin(lp[2] + angle);
    pfx = pfx - phx; pfy = pfy - phy;
    int ifx, ify; orc_map_world2map(map, pfx, pfy, &ifx, &ify); /* the world2map-of-a-difference quirk */
    int found = 0; double bx = 0., by = 0.;
    for (int xx = -sm->kernel_size; xx <= sm->kernel_size; xx++)
        for (int yy = -sm->kernel_size; yy <= sm->kernel_size; yy++) {
            int prx = ihx + xx, pry = ihy + yy;
            int pfxx = prx + ifx, pfyy = pry + ify;
            const orc_cell* cell = cell_const(map, prx, pry);
            const orc_cell* fcell = cell_const(map, pfxx, pfyy);
            if (occ(cell) > sm->fullness_threshold && occ(fcell) < sm->fullness_threshold) {
                /* mean(): smmap.h:24 = (1./n) * Point(acc.x, acc.y) */
                double inv = 1. / cell->n;
                double mx = phx - (double)cell->ax * inv, my = phy - (double)cell->ay * inv;
                if (!found) { bx = mx; by = my; found = 1; }
                else if ((mx * mx + my * my) < (bx * bx + by * by)) { bx = mx; by = my; }
            }
        }
    *mux = bx; *muy = by;
    return found;
}

/* ScanMatcher::score: scanmatcher.h:171-259 */
double orc_score(orc_matcher* sm, const orc_map* map, const double p[3], const double* readings) {
    double s = 0;
    double lp[3]; laser_pose(sm, p, lp);
    unsigned int skip = 0;
    double freeDelta = map->delta * sm->free_cell_ratio;
    sm->score_calls++;
    for (uint32_t i = sm->initial_beams_skip; i < sm->beams; i++) {
        double r = readings[i];
        skip++;
        skip = skip > sm->likelihood_skip ? 0 : skip;
        if (skip || r > sm->usable_range || r == 0.0) continue;
        double mx, my;
        sm->beam_evals++;
        if (window_match(sm, map, lp, r, sm->angles[i], map->delta * freeDelta, &mx, &my)) {
            /* exp(-1./sigma * bestMu * bestMu) == ((c*mu.x)*mu.x + (c*mu.y)*mu.y), point.h:34-49 */
            double c = -1. / sm->gaussian_sigma;
            s += exp((mx * c) * mx + (my * c) * my);
        }
    }
    return s;
}

/* ScanMatcher::likelihoodAndScore: scanmatcher.h:271-380 */
uint32_t orc_likelihood_and_score(orc_matcher* sm, const orc_map* map, const double p[3], const double* readings,
                                  double* s_out, double* l_out) {
    double l = 0, s = 0;
    double lp[3]; laser_pose(sm, p, lp);
    double noHit = -.5 / (sm->likelihood_sigma); /* nullLikelihood: scanmatcher.cpp:15 */
    unsigned int skip = 0, c = 0;
    double freeDelta = map->delta * sm->free_cell_ratio;
    for (uint32_t i = sm->initial_beams_skip; i < sm->beams; i++) {
        double r = readings[i];
        skip++;
        skip = skip > sm->likelihood_skip ? 0 : skip;
        if (r > sm->usable_range) continue;
        if (skip) continue;
        double mx, my;
        sm->beam_evals++;
        int found = window_match(sm, map, lp, r, sm->angles[i], freeDelta, &mx, &my);
        if (found) {
            double cc = -1. / sm->gaussian_sigma;
            s += exp((mx * cc) * mx + (my * cc) * my);
            c++;
        }
        {
            double f = (-1. / sm->likelihood_sigma) * (mx * mx + my * my);
            l += found ? f : noHit;
        }
    }
    *s_out = s; *l_out = l;
    return c;
}

/* ScanMatcher::optimize(pnew, map, init, readings): scanmatcher.cpp:386-529 */
double orc_optimize(orc_matcher* sm, const orc_map* map, const double init[3], const double* readings, double pnew[3]) {
    double bestScore = -1;
    double cur[3] = {init[0], init[1], init[2]};
    double currentScore = orc_score(sm, map, cur, readings);
    double adelta = sm->opt_angular_delta, ldelta = sm->opt_linear_delta;
    unsigned int refinement = 0;
    do {
        if (bestScore >= currentScore) { refinement++; adelta *= .5; ldelta *= .5; }
        bestScore = currentScore;
        double bestLocal[3] = {cur[0], cur[1], cur[2]};
        for (int move = 0; move < 6; move++) { /* Front, Back, Left, Right, TurnLeft, TurnRight */
            double lpz[3] = {cur[0], cur[1], cur[2]};
            switch (move) {
                case 0: lpz[0] += ldelta; break;
                case 1: lpz[0] -= ldelta; break;
                case 2: lpz[1] -= ldelta; break;
                case 3: lpz[1] += ldelta; break;
                case 4: lpz[2] += adelta; break;
                default: lpz[2] -= adelta; break;
            }
            double odo_gain = 1;
            if (sm->angular_odometry_reliability > 0.) {
                double dth = init[2] - lpz[2];
                dth = atan2(sin(dth), cos(dth));
                dth *= dth;
                odo_gain *= exp(-sm->angular_odometry_reliability * dth);
            }
            if (sm->linear_odometry_reliability > 0.) {
                double dx = init[0] - lpz[0], dy = init[1] - lpz[1];
                double drho = dx * dx + dy * dy;
                odo_gain *= exp(-sm->linear_odometry_reliability * drho);
            }
            double localScore = odo_gain * orc_score(sm, map, lpz, readings);
            if (localScore > currentScore) {
                currentScore = localScore;
                bestLocal[0] = lpz[0]; bestLocal[1] = lpz[1]; bestLocal[2] = lpz[2];
            }
        }
        cur[0] = bestLocal[0]; cur[1] = bestLocal[1]; cur[2] = bestLocal[2];
    } while (currentScore > bestScore || refinement < sm->opt_recursive_iterations);
    pnew[0] = cur[0]; pnew[1] = cur[1]; pnew[2] = cur[2];
    return bestScore;
}

/* GridLineTraversal::gridLine: gridlinetraversal.h:20-124.  Integer Bresenham that always walks
 * the major axis upward and is then put into start->end order. */
int orc_grid_line(int x0, int y0, int x1, int y1, int32_t* pts, int max_pts) {
    int dx = abs(x1 - x0), dy = abs(y1 - y0);
    int steep = !(dy <= dx);
    int a0 = steep ? y0 : x0, b0 = steep ? x0 : y0, a1 = steep ? y1 : x1, b1 = steep ? x1 : y1;
    int da = steep ? dy : dx, db = steep ? dx : dy;
    int from_end = a0 > a1;
    int a = from_end ? a1 : a0, b = from_end ? b1 : b0, aend = from_end ? a0 : a1;
    int dirflag = from_end ? -1 : 1;
    int bstep = ((b1 - b0) * dirflag) > 0 ? 1 : -1;
    int d = 2 * db - da, incr1 = 2 * db, incr2 = 2 * (db - da);
    int n = 0;
#define EMIT() do { if (n < max_pts) { pts[2 * n] = steep ? b : a; pts[2 * n + 1] = steep ? a : b; } n++; } while (0)
    EMIT();
    while (a < aend) {
        a++;
        if (d < 0) d += incr1; else { b += bstep; d += incr2; }
        EMIT();
    }
#undef EMIT
    if (from_end && n <= max_pts)
        for (int i = 0, j = n - 1; i < n / 2; i++, j--) {
            int32_t tx = pts[2 * i], ty = pts[2 * i + 1];
            pts[2 * i] = pts[2 * j]; pts[2 * i + 1] = pts[2 * j + 1];
            pts[2 * j] = tx; pts[2 * j + 1] = ty;
        }
    return n;
}

#define LINE_CAP 20000 /* scanmatcher.cpp:55 */

/* ScanMatcher::computeActiveArea: scanmatcher.cpp:83-254 (the m_activeAreaComputed guard is the
 * caller's business here: the filter invalidates before every call) */
void orc_compute_active_area(const orc_matcher* sm, orc_map* map, const double p[3], const double* readings) {
    double lp[3]; laser_pose(sm, p, lp);
    double minx, miny, maxx, maxy;
    map2world(map, 0, 0, &minx, &miny);
    map2world(map, map->map_w - 1, map->map_h - 1, &maxx, &maxy);
    if (lp[0] < minx) minx = lp[0];
    if (lp[1] < miny) miny = lp[1];
    if (lp[0] > maxx) maxx = lp[0];
    if (lp[1] > maxy) maxy = lp[1];
    for (uint32_t i = sm->initial_beams_skip; i < sm->beams; i++) {
        double r = readings[i];
        if (r > sm->laser_max_range || r == 0.0 || isnan(r)) continue;
        double d = r > sm->usable_range ? sm->usable_range : r;
        double phx = lp[0], phy = lp[1];
        phx += d * cos(lp[2] + sm->angles[i]);
        phy += d * sin(lp[2] + sm->angles[i]);
        if (phx < minx) minx = phx;
        if (phy < miny) miny = phy;
        if (phx > maxx) maxx = phx;
        if (phy > maxy) maxy = phy;
    }
    if (!is_inside_world(map, minx, miny) || !is_inside_world(map, maxx, maxy)) {
        double lminx, lminy, lmaxx, lmaxy;
        map2world(map, 0, 0, &lminx, &lminy);
        map2world(map, map->map_w - 1, map->map_h - 1, &lmaxx, &lmaxy);
        minx = (minx >= lminx) ? lminx : minx - sm->enlarge_step;
        maxx = (maxx <= lmaxx) ? lmaxx : maxx + sm->enlarge_step;
        miny = (miny >= lminy) ? lminy : miny - sm->enlarge_step;
        maxy = (maxy <= lmaxy) ? lmaxy : maxy + sm->enlarge_step;
        orc_map_resize(map, minx, miny, maxx, maxy);
    }
    memset(map->active, 0, (size_t)map->npx * map->npy);
    int32_t* line = (int32_t*)malloc(sizeof(int32_t) * 2 * LINE_CAP);
    for (uint32_t i = sm->initial_beams_skip; i < sm->beams; i++) {
        if (sm->generate_map) {
            double d = readings[i];
            if (d > sm->laser_max_range || d == 0.0 || isnan(d)) continue;
            if (d > sm->usable_range) d = sm->usable_range;
            double phx = lp[0] + d * cos(lp[2] + sm->angles[i]);
            double phy = lp[1] + d * sin(lp[2] + sm->angles[i]);
            int p0x, p0y, p1x, p1y;
            orc_map_world2map(map, lp[0], lp[1], &p0x, &p0y);
            orc_map_world2map(map, phx, phy, &p1x, &p1y);
            int n = orc_grid_line(p0x, p0y, p1x, p1y, line, LINE_CAP);
            for (int k = 0; k < n - 1; k++) {
                int px, py;
                if (patch_of(map, line[2 * k], line[2 * k + 1], &px, &py)) map->active[(size_t)px * map->npy + py] = 1;
            }
            if (d < sm->usable_range) {
                int px, py;
                if (patch_of(map, p1x, p1y, &px, &py)) map->active[(size_t)px * map->npy + py] = 1;
            }
        } else {
            double r = readings[i];
            if (r > sm->laser_max_range || r > sm->usable_range || r == 0.0 || isnan(r)) continue;
            double phx = lp[0], phy = lp[1];
            phx += r * cos(lp[2] + sm->angles[i]);
            phy += r * sin(lp[2] + sm->angles[i]);
            int p1x, p1y, px, py;
            orc_map_world2map(map, phx, phy, &p1x, &p1y);
            if (patch_of(map, p1x, p1y, &px, &py)) map->active[(size_t)px * map->npy + py] = 1;
        }
    }
    free(line);
}

/* ScanMatcher::registerScan: scanmatcher.cpp:266-354 (entropy bookkeeping `esum` is unused by every
 * hot-path caller and is omitted).  Must be preceded by orc_compute_active_area on the same pose. */
void orc_register_scan(const orc_matcher* sm, orc_map* map, const double p[3], const double* readings) {
    /* allocActiveArea: harray2d.h:169-185 (copy-or-create; with owning maps: create if missing) */
    size_t nd = (size_t)map->npx * map->npy;
    for (size_t i = 0; i < nd; i++)
        if (map->active[i] && !map->dir[i]) map->dir[i] = (orc_cell*)calloc(ORC_PATCH * ORC_PATCH, sizeof(orc_cell));
    double lp[3]; laser_pose(sm, p, lp);
    int p0x, p0y; orc_map_world2map(map, lp[0], lp[1], &p0x, &p0y);
    int32_t* line = (int32_t*)malloc(sizeof(int32_t) * 2 * LINE_CAP);
    for (uint32_t i = sm->initial_beams_skip; i < sm->beams; i++) {
        if (sm->generate_map) {
            double d = readings[i];
            if (d > sm->laser_max_range || d == 0.0 || isnan(d)) continue;
            if (d > sm->usable_range) d = sm->usable_range;
            double phx = lp[0] + d * cos(lp[2] + sm->angles[i]);
            double phy = lp[1] + d * sin(lp[2] + sm->angles[i]);
            int p1x, p1y; orc_map_world2map(map, phx, phy, &p1x, &p1y);
            int n = orc_grid_line(p0x, p0y, p1x, p1y, line, LINE_CAP);
            for (int k = 0; k < n - 1; k++) {
                orc_cell* c = cell_mut(map, line[2 * k], line[2 * k + 1]);
                if (c) c->visits++; /* update(false): smmap.h:56-59 */
            }
            if (d < sm->usable_range) {
                orc_cell* c = cell_mut(map, p1x, p1y);
                if (c) { c->ax += (float)phx; c->ay += (float)phy; c->n++; c->visits += 1; } /* smmap.h:47-54 */
            }
        } else {
            double r = readings[i];
            if (r > sm->laser_max_range || r > sm->usable_range || r == 0.0 || isnan(r)) continue;
            double phx = lp[0], phy = lp[1];
            phx += r * cos(lp[2] + sm->angles[i]);
            phy += r * sin(lp[2] + sm->angles[i]);
            int p1x, p1y; orc_map_world2map(map, phx, phy, &p1x, &p1y);
            orc_cell* c = cell_mut(map, p1x, p1y);
            if (c) { c->ax += (float)phx; c->ay += (float)phy; c->n++; c->visits += 1; }
        }
    }
    free(line);
}

/* ------------------------------------------------------------------------------ particle filter */

/* GridSlamProcessor::normalize: gridslamprocessor.hxx:82-121; returns neff */
double orc_normalize(const double* logw, uint32_t n, double obs_sigma_gain, double* w) {
    double gain = 1. / (obs_sigma_gain * n);
    double lmax = -1.7976931348623157e308;
    for (uint32_t i = 0; i < n; i++) lmax = logw[i] > lmax ? logw[i] : lmax;
    double wcum = 0;
    for (uint32_t i = 0; i < n; i++) { w[i] = exp(gain * (logw[i] - lmax)); wcum += w[i]; }
    double neff = 0;
    for (uint32_t i = 0; i < n; i++) { w[i] = w[i] / wcum; double x = w[i]; neff += x * x; }
    return 1. / neff;
}

/* uniform_resampler::resampleIndexes: particlefilter.h:180-229 with the drand48() draw passed in
 * as u01.  indexes(n) is zero-initialised; emission beyond n (reference: out-of-bounds write) is
 * dropped.  Returns the number of emitted indexes (== n except for FP corner cases). */
uint32_t orc_resample_indexes(const double* w, uint32_t n, double u01, uint32_t* idx) { return orc_resample_indexes_n(w, n, n, u01, idx); }

/* the same with the `nparticles` argument (particlefilter.h:196-203: `if (nparticles > 0) n = nparticles;` -- the interval is
 * the total weight over n_out, indexes(n_out)) */
uint32_t orc_resample_indexes_n(const double* w, uint32_t n, uint32_t n_out, double u01, uint32_t* idx) {
    double cweight = 0;
    for (uint32_t i = 0; i < n; i++) cweight += w[i];
    double interval = cweight / n_out;
    double target = interval * u01;
    cweight = 0;
    for (uint32_t i = 0; i < n_out; i++) idx[i] = 0;
    uint32_t k = 0;
    for (uint32_t i = 0; i < n; i++) {
        cweight += w[i];
        while (cweight > target) {
            if (k < n_out) idx[k] = i;
            k++;
            target += interval;
        }
    }
    return k;
}

/* pf_ran_gaussian: stat.cpp:31-47 (polar Box-Muller; note the discarded draw before x2) */
static double ran_gaussian(double sigma) {
    double x1, x2, w, r;
    do {
        do { r = drand48(); } while (r == 0.0);
        x1 = 2.0 * r - 1.0;
        do { r = drand48(); } while (r == 0.0);
        x2 = 2.0 * drand48() - 1.0;
        w = x1 * x1 + x2 * x2;
    } while (w > 1.0 || w == 0.0);
    return (sigma * x2 * sqrt(-2.0 * log(w) / w));
}

double orc_sample_gaussian(double sigma) { /* stat.cpp:49-65 */
    if (sigma == 0) return 0;
    return ran_gaussian(sigma);
}

/* the drivers seed with sampleGaussian(1, S) (gsp_thread.cpp:393-394, slam_gmapping.cpp:676): that is
 * srand48(S) FOLLOWED by one N(0,1) draw, which advances the stream (stat.cpp:56-65). */
void orc_seed(long seed) { srand48(seed); (void)ran_gaussian(1.0); }

/* MotionModel::drawFromMotion(p, pnew, pold): motionmodel.cpp:40-60 with absoluteDifference /
 * absoluteSum from point.h:117-134 */
void orc_draw_from_motion(const double p[3], const double pnew[3], const double pold[3], double srr, double srt,
                          double str, double stt, double out[3]) {
    double sxy = 0.3 * srr;
    double ddx = pnew[0] - pold[0], ddy = pnew[1] - pold[1], dth = pnew[2] - pold[2];
    dth = atan2(sin(dth), cos(dth));
    double s = sin(pold[2]), c = cos(pold[2]);
    double dx = c * ddx + s * ddy, dy = -s * ddx + c * ddy;
    double nx = dx, ny = dy, nth = dth;
    nx += orc_sample_gaussian(srr * fabs(dx) + str * fabs(dth) + sxy * fabs(dy));
    ny += orc_sample_gaussian(srr * fabs(dy) + str * fabs(dth) + sxy * fabs(dx));
    nth += orc_sample_gaussian(stt * fabs(dth) + srt * sqrt(dx * dx + dy * dy));
    nth = fmod(nth, 2 * M_PI);
    if (nth > M_PI) nth -= 2 * M_PI;
    double s1 = sin(p[2]), c1 = cos(p[2]);
    out[0] = (c1 * nx - s1 * ny) + p[0];
    out[1] = (s1 * nx + c1 * ny) + p[1];
    out[2] = nth + p[2];
}

typedef struct { orc_map* map; double pose[3], prev_pose[3], weight, weight_sum; int prev_index; } orc_particle;

struct orc_filter {
    uint32_t n;
    orc_particle* p;
    orc_matcher sm;
    double srr, srt, str, stt;
    double lin_thr, ang_thr, resample_thr, period, minimum_score, obs_sigma_gain;
    int count, reading_count;
    double last_part_pose[3], odo_pose[3], lin_dist, ang_dist, last_update_time, neff;
    double *weights, *pre_poses, *match_rows;
    uint32_t* indexes;
    int last_resampled;
    double neff_at_resample, *w_at_resample, last_u01;
};

/* GridSlamProcessor::init: gridslamprocessor.cpp:354-403 */
orc_filter* orc_filter_new(uint32_t particles, double xmin, double ymin, double xmax, double ymax, double delta,
                           const double init_pose[3], const orc_matcher* sm) {
    orc_filter* f = (orc_filter*)calloc(1, sizeof *f);
    f->n = particles; f->sm = *sm;
    f->p = (orc_particle*)calloc(particles, sizeof(orc_particle));
    for (uint32_t i = 0; i < particles; i++) {
        f->p[i].map = orc_map_new((xmin + xmax) * .5, (ymin + ymax) * .5, xmax - xmin, ymax - ymin, delta);
        memcpy(f->p[i].pose, init_pose, sizeof(double) * 3);
        memcpy(f->p[i].prev_pose, init_pose, sizeof(double) * 3);
    }
    f->neff = (double)particles;
    f->period = 5.0; f->obs_sigma_gain = 1; f->resample_thr = 0.5; f->minimum_score = 0.; /* ctor: cpp:25-31 */
    f->weights = (double*)calloc(particles, sizeof(double));
    f->pre_poses = (double*)calloc(3 * (size_t)particles, sizeof(double));
    f->match_rows = (double*)calloc(8 * (size_t)particles, sizeof(double));
    f->w_at_resample = (double*)calloc(particles, sizeof(double));
    f->indexes = (uint32_t*)calloc(particles, sizeof(uint32_t));
    return f;
}

void orc_filter_free(orc_filter* f) {
    if (!f) return;
    for (uint32_t i = 0; i < f->n; i++) orc_map_free(f->p[i].map);
    free(f->p); free(f->weights); free(f->pre_poses); free(f->match_rows); free(f->w_at_resample); free(f->indexes); free(f);
}

void orc_filter_set_motion(orc_filter* f, double srr, double srt, double str, double stt) { f->srr = srr; f->srt = srt; f->str = str; f->stt = stt; }
void orc_filter_set_update(orc_filter* f, double linear, double angular, double resample_thr, double period,
                           double minimum_score, double obs_sigma_gain) {
    f->lin_thr = linear; f->ang_thr = angular; f->resample_thr = resample_thr; f->period = period;
    f->minimum_score = minimum_score; f->obs_sigma_gain = obs_sigma_gain;
}

static void filter_normalize(orc_filter* f) { /* updateTreeWeights -> normalize (tree itself is host bookkeeping) */
    double* lw = (double*)malloc(sizeof(double) * f->n);
    for (uint32_t i = 0; i < f->n; i++) lw[i] = f->p[i].weight;
    f->neff = orc_normalize(lw, f->n, f->obs_sigma_gain, f->weights);
    free(lw);
}

/* GridSlamProcessor::processScan: gridslamprocessor.cpp:425-662 with scanMatch (hxx:15-75) and
 * resample (hxx:131-302) inlined.  The trajectory tree and the .gfs trace are not part of the oracle. */
int orc_filter_process(orc_filter* f, const double* ranges, const double rel[3], double stamp) {
    if (!f->count) { memcpy(f->last_part_pose, rel, 24); memcpy(f->odo_pose, rel, 24); }
    for (uint32_t i = 0; i < f->n; i++) {
        double out[3];
        orc_draw_from_motion(f->p[i].pose, rel, f->odo_pose, f->srr, f->srt, f->str, f->stt, out);
        memcpy(f->p[i].pose, out, 24);
        memcpy(f->pre_poses + 3 * i, out, 24);
    }
    double mx = rel[0] - f->odo_pose[0], my = rel[1] - f->odo_pose[1], mth = rel[2] - f->odo_pose[2];
    mth = atan2(sin(mth), cos(mth));
    f->lin_dist += sqrt(mx * mx + my * my);
    f->ang_dist += fabs(mth);
    memcpy(f->odo_pose, rel, 24);
    int processed = 0;
    f->last_resampled = 0;
    if (!f->count || f->lin_dist >= f->lin_thr || f->ang_dist >= f->ang_thr ||
        (f->period >= 0.0 && (stamp - f->last_update_time) > f->period)) {
        f->last_update_time = stamp;
        if (f->count > 0) {
            /* scanMatch */
            for (uint32_t i = 0; i < f->n; i++) {
                orc_particle* p = &f->p[i];
                double corrected[3], s, l;
                uint64_t calls0 = f->sm.score_calls;
                double score = orc_optimize(&f->sm, p->map, p->pose, ranges, corrected);
                double* row = f->match_rows + 8 * i;
                row[0] = score; row[1] = corrected[0]; row[2] = corrected[1]; row[3] = corrected[2];
                row[7] = (double)(f->sm.score_calls - calls0);
                if (score > f->minimum_score) memcpy(p->pose, corrected, 24);
                uint32_t c = orc_likelihood_and_score(&f->sm, p->map, p->pose, ranges, &s, &l);
                row[4] = s; row[5] = l; row[6] = (double)c;
                p->weight += l; p->weight_sum += l;
                orc_compute_active_area(&f->sm, p->map, p->pose, ranges);
            }
            filter_normalize(f);
            /* resample */
            if (f->neff < f->resample_thr * f->n) {
                f->neff_at_resample = f->neff;
                memcpy(f->w_at_resample, f->weights, sizeof(double) * f->n);
                double u = drand48();
                f->last_u01 = u;
                orc_resample_indexes(f->weights, f->n, u, f->indexes);
                orc_particle* np = (orc_particle*)calloc(f->n, sizeof(orc_particle));
                for (uint32_t i = 0; i < f->n; i++) {
                    np[i] = f->p[f->indexes[i]];
                    np[i].map = orc_map_clone(f->p[f->indexes[i]].map);
                    np[i].prev_index = (int)f->indexes[i];
                }
                for (uint32_t i = 0; i < f->n; i++) orc_map_free(f->p[i].map);
                free(f->p); f->p = np;
                for (uint32_t i = 0; i < f->n; i++) {
                    f->p[i].weight = 0;
                    orc_compute_active_area(&f->sm, f->p[i].map, f->p[i].pose, ranges);
                    orc_register_scan(&f->sm, f->p[i].map, f->p[i].pose, ranges);
                }
                f->last_resampled = 1;
            } else {
                for (uint32_t i = 0; i < f->n; i++) {
                    orc_compute_active_area(&f->sm, f->p[i].map, f->p[i].pose, ranges);
                    orc_register_scan(&f->sm, f->p[i].map, f->p[i].pose, ranges);
                    f->p[i].prev_index = (int)i;
                }
            }
        } else {
            for (uint32_t i = 0; i < f->n; i++) {
                orc_compute_active_area(&f->sm, f->p[i].map, f->p[i].pose, ranges);
                orc_register_scan(&f->sm, f->p[i].map, f->p[i].pose, ranges);
            }
        }
        filter_normalize(f);
        memcpy(f->last_part_pose, f->odo_pose, 24);
        f->lin_dist = 0; f->ang_dist = 0; f->count++;
        processed = 1;
        for (uint32_t i = 0; i < f->n; i++) memcpy(f->p[i].prev_pose, f->p[i].pose, 24);
    }
    f->reading_count++;
    return processed;
}

uint32_t orc_filter_size(const orc_filter* f) { return f->n; }
void orc_filter_state(const orc_filter* f, double* rows) {
    for (uint32_t i = 0; i < f->n; i++) {
        double* r = rows + 6 * i; const orc_particle* p = &f->p[i];
        r[0] = p->pose[0]; r[1] = p->pose[1]; r[2] = p->pose[2]; r[3] = p->weight; r[4] = p->weight_sum; r[5] = (double)p->prev_index;
    }
}
void orc_filter_pre_poses(const orc_filter* f, double* rows3) { memcpy(rows3, f->pre_poses, sizeof(double) * 3 * f->n); }
void orc_filter_match_rows(const orc_filter* f, double* rows8) { memcpy(rows8, f->match_rows, sizeof(double) * 8 * f->n); }
double orc_filter_neff(const orc_filter* f) { return f->neff; }
void orc_filter_weights(const orc_filter* f, double* w) { memcpy(w, f->weights, sizeof(double) * f->n); }
int orc_filter_last_resampled(const orc_filter* f, uint32_t* idx_out, double* neff_at, double* w_at) {
    if (f->last_resampled) {
        if (idx_out) memcpy(idx_out, f->indexes, sizeof(uint32_t) * f->n);
        if (neff_at) *neff_at = f->neff_at_resample;
        if (w_at) memcpy(w_at, f->w_at_resample, sizeof(double) * f->n);
    }
    return f->last_resampled;
}
const orc_map* orc_filter_map(const orc_filter* f, uint32_t i) { return f->p[i].map; }
int orc_filter_best(const orc_filter* f) { /* gridslamprocessor.cpp:677-688 */
    unsigned bi = 0; double bw = -1.7976931348623157e308;
    for (uint32_t i = 0; i < f->n; i++) if (bw < f->p[i].weight_sum) { bw = f->p[i].weight_sum; bi = i; }
    return (int)bi;
}
orc_matcher* orc_filter_matcher(orc_filter* f) { return &f->sm; }
double orc_filter_last_u01(const orc_filter* f) { return f->last_u01; }
