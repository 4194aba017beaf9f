// gm_kernels.cu -- hand-written sm_100a kernels for GMapping's per-scan particle update.
// Compiled with --fmad=false: the reference x86-64 build has no FMA contraction and the hill-climb's
// decisions depend on the exact FP64 evaluation order (SURVEY.md §7 "hard parts").
//
// Kernels (one CTA per particle unless noted):
//   k_scanmatch   ScanMatcher::optimize + likelihoodAndScore, and the    (scanmatcher.cpp:386-529, scanmatcher.h:171-380)
//                 standalone score / likelihoodAndScore / optimize queries (A.query)
//   k_bounds      computeActiveArea's bounding box + Map::resize maths  (scanmatcher.cpp:83-166, map.h:218-241)
//   k_relayout    HierarchicalArray2D::resize                           (harray2d.h:80-109)
//   k_register    active area + copy-on-write + ray update + hits       (scanmatcher.cpp:168-354, harray2d.h:169-185)
//   k_remap/k_release_slots/k_merge_free   particle copy by patch reference     (gridslamprocessor.hxx:184-213)
//   k_normalize / k_resample       (gridslamprocessor.hxx:82-121, particlefilter.h:180-229)
//   k_digest / k_export_* / k_export_occupancy / k_install_patches   map read-back, occupancy grid, upload
#include <algorithm>
#include <climits>

#include "gm_device.cuh"
#include "gm_kernels.h"

namespace gm {

// ------------------------------------------------------------------------------------------------
// Scan matching.  What the reference computes per (pose, beam) -- a "beam evaluation" -- is restated so that
// the result is unchanged but the instruction count and, above all, the code size drop (the literal
// transcription was 226 KB of SASS and stalled on instruction fetch):
//   * ONE copy of the beam evaluation: the kernel is a small state machine (initial score -> hill-climb
//     rounds -> likelihoodAndScore) whose steps all go through the same evaluation loop; thread 0 plans each
//     step in shared memory (candidate laser poses, which trig table each candidate uses), everybody
//     evaluates, warps reduce, thread 0 decides.  The hill-climb state lives in shared memory, not registers;
//   * sin/cos of (laser theta + beam angle) live in shared memory, one table per distinct theta: the four
//     translation moves of a round share the current theta, the two turns have their own, and a table is
//     recomputed only when its theta changes (a few times per optimize(), not 6x per round);
//   * world2map's division by delta is a multiply by 1/delta with an exactness guard: whenever the product is
//     within 1e-6 of a rounding boundary the reference's division + round() is evaluated (out of line), so the
//     integer cell index is always the reference's;
//   * the 3x3 window: its occupancy (n / visits > 0.1) is read from the pool's occupancy bitmap -- three words, one per
//     window row, kept in step by every kernel that writes cells -- so free and unknown cells are never loaded; only
//     occupied cells load {acc, n} and look at the free cell.  The free-cell window is resolved once per beam: the
//     reference's world2map-of-a-difference quirk puts it half a map away, where it is almost always wholly unknown;
//   * mean() = acc * (1./n): 1./n comes from a table of exactly rounded reciprocals (n < 4096).
// Generic code paths (negative fullnessThreshold, any kernelSize, windows across patch edges) remain.

// (int)round(d / delta) the way the reference's x86-64 build evaluates it: cvttsd2si yields INT_MIN for NaN and for
// values outside int range (a NaN range reading is not skipped by score(), scanmatcher.h:195)
__device__ __noinline__ int cell_index_slow(double d, double delta) {
    const double q = round(d / delta);
    return fabs(q) < 2147483648.0 ? (int)q : (int)0x80000000;
}
__device__ __noinline__ double recip_slow(int n) { return 1. / n; }
__device__ __noinline__ double exp_cold(double x) { return exp(x); }

// exp(x) for the x <= 0 of the score sums (scanmatcher.h:257,368): 2^(k/128) from a table of rounded powers times a
// degree-5 polynomial on |r| <= ln2/256, < 1 ulp like the library function it replaces (checked against expl() on the
// host over [-700, 0]), about half its instructions and a third of its dependent chain.  Arguments below -700 (results
// near the subnormal range), positive ones and NaN take the library call.
__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ pow2) {
    if (!(x > -700.0 && x <= 0.0)) return exp_cold(x);
    const double shift = 6755399441055744.0;  // 1.5 * 2^52: the sum's low word is rint(x * 128/ln2)
    const double t = __fma_rn(x, 0x1.71547652b82fep+7, shift);
    const int ki = __double2loint(t);
    const double kd = t - shift;
    double r = __fma_rn(kd, -0x1.62e42fefa0000p-8, x);  // ln2/128 in two pieces
    r = __fma_rn(kd, -0x1.cf79abc9e3b3ap-47, r);
    const double r2 = r * r;
    const double p1 = __fma_rn(r, 0x1.5555555555555p-3, 0.5);
    const double p2 = __fma_rn(r, 0x1.1111111111111p-7, 0x1.5555555555555p-5);
    double tail = __fma_rn(r2, p1, r);
    tail = __fma_rn(r2 * r2, p2, tail);
    const double m = __ldg(pow2 + (ki & (kExpTable - 1)));
    const double scale = __hiloint2double(__double2hiint(m) + ((ki >> 7) << 20), __double2loint(m));
    return __fma_rn(scale, tail, scale);
}

// reference-exact (int)round(d / delta), d = world coordinate minus map centre (map.h:283-287): multiply by
// 1/delta, and fall back to the division whenever the product is within 1e-6 of a rounding boundary (this
// also catches NaN and out-of-range values, for which |q - rn(q)| is not <= 0.499999)
__device__ __forceinline__ int cell_index(double d, double delta, double inv_delta) {
    const double q = d * inv_delta;
    int i = __double2int_rn(q);
    if (!(fabs(q - (double)i) <= 0.499999)) i = cell_index_slow(d, delta);
    return i;
}

// both coordinates of a point at once: one guard, one (cold) branch
__device__ __forceinline__ void cell_index2(double dx, double dy, double delta, double inv_delta, int& ix, int& iy) {
    const double qx = dx * inv_delta, qy = dy * inv_delta;
    ix = __double2int_rn(qx); iy = __double2int_rn(qy);
    const bool okx = fabs(qx - (double)ix) <= 0.499999, oky = fabs(qy - (double)iy) <= 0.499999;
    if (!(okx && oky)) { ix = cell_index_slow(dx, delta); iy = cell_index_slow(dy, delta); }
}

__device__ __forceinline__ double2 ldg_d2(const double2* p) {
    double2 v;
    asm("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// predicated load of a directory entry: no branch, no access when `pred` is false
__device__ __forceinline__ int ldg_if(const int* p, bool pred, int dflt) {
    int v;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\tmov.s32 %0, %3;\n\t@q ld.global.nc.s32 %0, [%1];\n\t}" : "=&r"(v) : "l"(p), "r"((int)pred), "r"(dflt));
    return v;
}

#ifndef GM_SM_MIN_CTAS
#define GM_SM_MIN_CTAS 3  // resident CTAs per SM k_scanmatch is compiled for (72 registers); 4 (56 registers, spills) measured slower
#endif
#ifndef GM_SM_WIDE_CTAS
#define GM_SM_WIDE_CTAS 2  // resident CTAs per SM of the wide (2 x 288 threads) shape
#endif
constexpr int kSmThreads = 288;
constexpr int kSmGroups = kSmThreads / 6;  // 48 beam groups x 6 candidate poses: every thread works

struct MatchCtx {
    // the particle's directory, read through a shared-memory slot at every use: kept in registers the compiler spends ten
    // instructions per lookup rebuilding the 64-bit address from the kernel parameters instead
    const int* const* dirp_s;
    const int4* bbox_s;  // shared: cell bounding box of the hit-capable patches {x_lo, y_lo, x_hi, y_hi} (inclusive; empty: lo > hi)
    const int4* __restrict__ pool;
    const int* __restrict__ visits;
    const double2* __restrict__ means;
    const unsigned* __restrict__ occ;
    const double* __restrict__ inv_n;
    Geom g;
    int zero_patch;  // = Pool::cap_h, the id of the all-zero patch of the hit planes: stands in for unallocated / outside patches (const
                     // Map::cell, map.h:347-354) and for free-space patches (ids >= cap_h: n == 0 in every cell)
};

__device__ __forceinline__ int dir_at(const MatchCtx& M, int px, int py) {
    return ((unsigned)px < (unsigned)M.g.npx && (unsigned)py < (unsigned)M.g.npy) ? __ldg(*M.dirp_s + (px * M.g.npy + py)) : -1;
}

// (n, visits) of map cell (x, y): unknown (0, 0) outside the map or in an unallocated patch; n = 0 in a free-space patch
__device__ __forceinline__ void cell_nv(const MatchCtx& M, int x, int y, int& n, int& v) {
    long long ci;
    n = 0; v = 0;
    if (cell_addr(M.g, *M.dirp_s, x, y, ci)) {
        v = __ldg(M.visits + ci);
        if (ci < ((long long)M.zero_patch << 10)) n = __ldg(M.pool + ci).z;
    }
}

// occupancy tests (smmap.h:25 against fullnessThreshold); DEF = the default threshold 0.1, decided in exact integers
template <bool DEF> __device__ __forceinline__ bool is_occ(int n, int v, const DevParams& P) {
    return DEF ? (v != 0 && 10ll * n > (long long)v) : occ_gt(n, v, P);
}
template <bool DEF> __device__ __forceinline__ bool is_free(int n, int v, const DevParams& P) {
    return DEF ? (v == 0 || 10ll * n < (long long)v) : occ_lt(n, v, P);
}

// window search for one beam (scanmatcher.h:198-252 / :311-365).  (c, s) = cos/sin(lp.theta + angle);
// `free_off` is delta*freeDelta for score() and freeDelta for likelihoodAndScore().
// K1: kernelSize 1 and fullnessThreshold 0.1 (the defaults) -> specialised, branch-light window code
template <bool K1>
// (fqx, fqy): the beam-independent free-cell index offset when it exists (see derive_params), else fqx == INT_MIN
// (phx, phy): the beam end point lp + r * (c, s).  r, c and s themselves are only needed for the free point when its
// cell offset is not a constant (fqx == INT_MIN): they are re-read then (`ranges`, and the plain trig table at tab_s).
__device__ __forceinline__ bool beam_match(const DevParams& P, const MatchCtx& M, double lpx, double lpy, double phx, double phy, unsigned b,
                                           unsigned tab_s, const double* __restrict__ ranges, double free_off, int fqx, int fqy,
                                           double& bmx, double& bmy) {
    const Geom& g = M.g;
    int ihx, ihy;
    cell_index2(phx - P.cx, phy - P.cy, P.delta, P.inv_delta, ihx, ihy);
    ihx += g.sx2; ihy += g.sy2;
    const int k = K1 ? 1 : P.kernel_size;
    const int x0 = ihx - k, y0 = ihy - k, x1 = ihx + k, y1 = ihy + k;
    bool found = false;
    double bx = 0., by = 0., bn = 0.;

    if (K1) {
        // ---- 3x3 window.  It spans at most 2x2 patches (a: low x / low y, b: high y, c: high x, d: both).  Its occupancy
        // comes from the pool's bitmap: one word per window row (two when the row crosses a patch edge in y), funnel-
        // shifted to the window's first column; no cell is read unless it is occupied.
        const int pxa = x0 >> kPatchMag, pya = y0 >> kPatchMag, pxb = x1 >> kPatchMag, pyb = y1 >> kPatchMag;
        const bool twoy = pyb != pya, twox = pxb != pxa;
        const bool xa_in = (unsigned)pxa < (unsigned)g.npx, ya_in = (unsigned)pya < (unsigned)g.npy;
        const int* da = *M.dirp_s + (pxa * g.npy + pya);
        const unsigned ysh = (unsigned)(y0 & 31);
        unsigned mask, ea, eb, ec, ed, base = 0;
        int nxa, nya;
        // a warp works on the six candidate poses of five or six neighbouring beams: its windows lie within a few cells of
        // each other, and most of the time none of them crosses a patch edge
        const bool fast = __all_sync(__activemask(), !twox && !twoy);
        if (fast) {
            const int ida = ldg_if(da, xa_in && ya_in, -1);
            if ((unsigned)ida >= (unsigned)M.zero_patch) return false;  // unknown, or a free-space patch: nothing occupied
            const unsigned* w = M.occ + (((unsigned)ida << kPatchMag) + (unsigned)(x0 & 31));  // rows x0 .. x0 + 2: consecutive words
            const unsigned w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
            mask = ((w0 >> ysh) & 7u) | (((w1 >> ysh) & 7u) << 8) | (((w2 >> ysh) & 7u) << 16);
            base = ((unsigned)ida << 10) + ((unsigned)(x0 & 31) << kPatchMag) + ysh;
            ea = eb = ec = ed = (unsigned)ida; nxa = nya = kPatch;
        } else {
            // four directory entries (-1 outside the directory) with predicated loads: a branch would run its body for the
            // whole warp on behalf of the one lane that needs it
            const bool xb_in = (unsigned)pxb < (unsigned)g.npx, yb_in = (unsigned)pyb < (unsigned)g.npy;
            const int* dc = da + (twox ? g.npy : 0);
            const int ida = ldg_if(da, xa_in && ya_in, -1);
            const int idb = ldg_if(da + 1, twoy && xa_in && yb_in, twoy ? -1 : ida);
            const int idc = ldg_if(dc, twox && xb_in && ya_in, twox ? -1 : ida);
            const int idd = ldg_if(dc + 1, twox && twoy && xb_in && yb_in, twox ? (twoy ? -1 : idc) : idb);
            if ((ida & idb & idc & idd) < 0) return false;  // every patch unknown: nothing can match
            // unknown patches read the all-zero patch
            const unsigned zp = (unsigned)M.zero_patch;  // = pool capacity > every patch id; -1 as unsigned is larger still
            ea = min((unsigned)ida, zp); eb = min((unsigned)idb, zp); ec = min((unsigned)idc, zp); ed = min((unsigned)idd, zp);
            nxa = ((pxa + 1) << kPatchMag) - x0;  // window rows inside patches a / b
            nya = ((pya + 1) << kPatchMag) - y0;
            mask = 0;
#pragma unroll
            for (int xx = 0; xx < 3; xx++) {
                const unsigned row = (unsigned)((x0 + xx) & 31);
                const unsigned lo = __ldg(M.occ + (((xx < nxa ? ea : ec) << kPatchMag) + row));
                const unsigned hi = twoy ? __ldg(M.occ + (((xx < nxa ? eb : ed) << kPatchMag) + row)) : 0u;
                mask |= (__funnelshift_r(lo, hi, ysh) & 7u) << (8 * xx);  // bit 8 * xx + yy: ascending bits == scan order
            }
        }
        if (!mask) return false;
        // cell index (patch * 1024 + cell) of the window cell at mask bit i = 8 * xx + yy
        auto cell_of = [&](int i) -> unsigned {
            const int xx = i >> 3, yy = i & 7;
            const unsigned e = xx < nxa ? (yy < nya ? ea : eb) : (yy < nya ? ec : ed);
            return (e << 10) + ((unsigned)((x0 + xx) & 31) << kPatchMag) + (unsigned)((y0 + yy) & 31);
        };
        auto cell_of_fast = [&](int i) -> unsigned { return base + (((unsigned)i & 0x18u) << 2) + ((unsigned)i & 7u); };
        // free window (the reference's `ipfree` quirk: world2map applied to the difference pfree - phit).  |pfree - phit|
        // <= free_off, so for score() the index is the same for every beam and comes precomputed in (fqx, fqy)
        int ifx, ify;
        if (fqx != INT_MIN) { ifx = fqx + g.sx2; ify = fqy + g.sy2; }
        else {
            const double r = __ldg(ranges + b);
            double c, s;
            asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c), "=d"(s) : "r"(tab_s + b * 16u));
            const double pfx = (lpx + (r - free_off) * c) - phx, pfy = (lpy + (r - free_off) * s) - phy;
            ifx = cell_index(pfx - P.cx, P.delta, P.inv_delta) + g.sx2;
            ify = cell_index(pfy - P.cy, P.delta, P.inv_delta) + g.sy2;
        }
        const int fx0 = x0 + ifx, fy0 = y0 + ify, fx1 = x1 + ifx, fy1 = y1 + ify;
        bool fcheck = true;  // false: every free cell of this window is unknown, i.e. passes
        const int4 bb = *M.bbox_s;
        if (fx1 < bb.x || fy1 < bb.y || fx0 > bb.z || fy0 > bb.w) fcheck = false;  // clear of every allocated patch (the usual case)
        else if (fx1 < 0 || fy1 < 0 || (fx0 >> kPatchMag) >= g.npx || (fy0 >> kPatchMag) >= g.npy) fcheck = false;
        else if (fx0 >= 0 && fy0 >= 0 && (fx0 >> kPatchMag) == (fx1 >> kPatchMag) && (fy0 >> kPatchMag) == (fy1 >> kPatchMag) &&
                 (unsigned)dir_at(M, fx0 >> kPatchMag, fy0 >> kPatchMag) >= (unsigned)M.zero_patch) fcheck = false;  // unknown or free-space patch: every cell passes
        // occupied cells in ascending bit order == the reference's xx-outer / yy-inner scan order
        if (!fcheck) {
            // three cells per trip (a wall in the window is usually three to six cells): their load -> distance chains are
            // independent and overlap; the comparisons still run in scan order, so ties resolve as in the reference
            auto pair_loop = [&](auto cell) {
                while (mask) {
                    const int i = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const bool two = mask != 0;
                    const int j = two ? __ffs(mask) - 1 : i;
                    mask &= mask - 1;  // no-op when mask is already 0
                    const bool three = mask != 0;
                    const int k = three ? __ffs(mask) - 1 : i;
                    mask &= mask - 1;
                    const double2 ma = ldg_d2(M.means + cell(i)), mb = ldg_d2(M.means + cell(j)), mc = ldg_d2(M.means + cell(k));  // mean(), smmap.h:24
                    const double ax = phx - ma.x, ay = phy - ma.y, bxx = phx - mb.x, byy = phy - mb.y, cx = phx - mc.x, cy = phy - mc.y;
                    const double na = ax * ax + ay * ay, nb = bxx * bxx + byy * byy, nc2 = cx * cx + cy * cy;
                    if (!found || na < bn) { bx = ax; by = ay; bn = na; found = true; }
                    if (two && nb < bn) { bx = bxx; by = byy; bn = nb; }
                    if (three && nc2 < bn) { bx = cx; by = cy; bn = nc2; }
                }
            };
            if (fast) pair_loop(cell_of_fast); else pair_loop(cell_of);
        } else {
            while (mask) {
                const int i = __ffs(mask) - 1;
                mask &= mask - 1;
                const int xx = i >> 3, yy = i & 7;
                int fn, fv;
                cell_nv(M, x0 + xx + ifx, y0 + yy + ify, fn, fv);
                if (!is_free<true>(fn, fv, P)) continue;
                const double2 mean = ldg_d2(M.means + cell_of(i));
                const double mx = phx - mean.x, my = phy - mean.y;
                const double nn = mx * mx + my * my;
                if (!found || nn < bn) { bx = mx; by = my; bn = nn; found = true; }
            }
        }
        bmx = bx; bmy = by;
        return found;
    }

    // ---- generic path: per-cell directory lookups, any kernel size / threshold
    int ifx = 0, ify = 0;
    bool have_free = false;
    for (int xx = 0; xx <= 2 * k; xx++)
        for (int yy = 0; yy <= 2 * k; yy++) {
            long long ci;
            int4 cell = make_int4(0, 0, 0, 0);
            if (cell_addr(g, *M.dirp_s, x0 + xx, y0 + yy, ci)) {
                if (ci < ((long long)M.zero_patch << 10)) cell = __ldg(M.pool + ci);
                cell.w = __ldg(M.visits + ci);
            }
            if (!is_occ<false>(cell.z, cell.w, P)) continue;
            if (!have_free) {
                const double r = __ldg(ranges + b);
            double c, s;
            asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c), "=d"(s) : "r"(tab_s + b * 16u));
            const double pfx = (lpx + (r - free_off) * c) - phx, pfy = (lpy + (r - free_off) * s) - phy;
                ifx = cell_index(pfx - P.cx, P.delta, P.inv_delta) + g.sx2;
                ify = cell_index(pfy - P.cy, P.delta, P.inv_delta) + g.sy2;
                have_free = true;
            }
            int fn, fv;
            cell_nv(M, x0 + xx + ifx, y0 + yy + ify, fn, fv);
            if (!is_free<false>(fn, fv, P)) continue;
            const double inv = (cell.z >= 0 && cell.z < kInvTable) ? __ldg(M.inv_n + cell.z) : recip_slow(cell.z);
            const double mx = phx - (double)__int_as_float(cell.x) * inv, my = phy - (double)__int_as_float(cell.y) * inv;
            const double nn = mx * mx + my * my;
            if (!found || nn < bn) { bx = mx; by = my; bn = nn; found = true; }
        }
    bmx = bx; bmy = by;
    return found;
}

__device__ __noinline__ double2 sincos_cold(double t) { double s, c; sincos(t, &s, &c); return make_double2(s, c); }  // (sin, cos), by value

// the odometry prior of optimize() (scanmatcher.cpp:497-509); 1 unless the odometry reliabilities are set.
// (arguments by value: a pointer to a local array or to the kernel's parameter block would force either into local memory)
__device__ __noinline__ double odo_gain_of(double ang_rel, double lin_rel, double ix, double iy, double it, double cx, double cy, double ct) {
    double odo_gain = 1;
    if (ang_rel > 0.) {
        double dth = it - ct;
        dth = atan2(sin(dth), cos(dth));
        dth *= dth;
        odo_gain *= exp(-ang_rel * dth);
    }
    if (lin_rel > 0.) {
        double dx = ix - cx, dy = iy - cy;
        double drho = dx * dx + dy * dy;
        odo_gain *= exp(-lin_rel * drho);
    }
    return odo_gain;
}

// laser pose in the map frame (scanmatcher.h:177-180)
__device__ __forceinline__ void laser_pose3(double lpx, double lpy, double lpt, double x, double y, double t, double& lx, double& ly, double& lt) {
    if (lpx == 0. && lpy == 0.) {  // x + (c*0 - s*0) == x exactly
        lx = x; ly = y;
    } else {
        const double2 sc = sincos_cold(t);
        lx = x + (sc.y * lpx - sc.x * lpy);
        ly = y + (sc.x * lpx + sc.y * lpy);
    }
    lt = t + lpt;
}
__device__ __forceinline__ void laser_pose(const DevParams& P, double x, double y, double t, double& lx, double& ly, double& lt) {
    laser_pose3(P.laser_pose[0], P.laser_pose[1], P.laser_pose[2], x, y, t, lx, ly, lt);
}

// beam i (counted from initialBeamsSkip) is evaluated iff the reference's `skip` counter wraps to 0
// there: skip++ ; skip = skip > likelihoodSkip ? 0 : skip  (scanmatcher.h:189-196)
__device__ __forceinline__ bool beam_selected(const DevParams& P, unsigned b) {
    return P.lskip == 0u || ((b - P.beam_skip + 1u) % (P.lskip + 1u)) == 0u;
}


struct SmShared {
    // plan of the next evaluation step (thread 0 writes, everybody reads after a barrier)
    double lx[6], ly[6];   // laser position of each candidate pose
    double fill_theta[3];  // theta for the trig tables listed in fill[]
    int fill[3];           // physical trig tables to recompute before evaluating (-1: none)
    int tab_of[3];         // physical table of: the current theta, theta + adelta, theta - adelta
    const int* dirp;       // directory of this CTA's particle
    int4 bbox;             // see MatchCtx::bbox_s
    int bb_lo_x, bb_lo_y, bb_hi_x, bb_hi_y;  // its reduction: patch coordinates
    int ncand;             // poses to evaluate: 1, or the 6 moves of a round, or 5 of them (see plan_round)
    int cand[6];           // which moves
    int skip;              // the move left out (-1: none)
    int mode;              // 0 score(), 1 likelihoodAndScore()
    int done;
    // reduction scratch
    double part[kSmThreads];
    double red[16], red_l[16];
    unsigned red_c[16], red_v[16];
    // hill-climb state (thread 0 only; shared memory keeps it out of everybody's registers)
    double cur[3], init[3], key[3], currentScore, bestScore, adelta, ldelta;
    double prev[3];        // the pose the last winning move started from
    unsigned refinement, evals, evaluated, valid_lik;
    int phase, key_ok[3], key_kind[3], fill_kind[3], last_move;  // table kinds: 0 (cos, sin), 1 end-point offsets r * (cos, sin)
};

// the wide shape (two half-CTAs per particle, see k_scanmatch): what the upper half hands to the lower one
template <int WIDE> struct SmWide { double s[kSmThreads], l[kSmThreads]; unsigned c[kSmThreads], v[kSmThreads]; };
template <> struct SmWide<0> {};

// candidate pose m of a hill-climb round around `cur` (scanmatcher.cpp:441-494: Front, Back, Left, Right, TurnLeft, TurnRight)
__device__ __forceinline__ void candidate(const double* cur, int m, double ldelta, double adelta, double (&out)[3]) {
    out[0] = cur[0]; out[1] = cur[1]; out[2] = cur[2];
    switch (m) {
        case 0: out[0] += ldelta; break;
        case 1: out[0] -= ldelta; break;
        case 2: out[1] -= ldelta; break;
        case 3: out[1] += ldelta; break;
        case 4: out[2] += adelta; break;
        default: out[2] -= adelta; break;
    }
}

// thread 0: make logical table `which` hold theta (recompute only when its key differs)
__device__ __forceinline__ void want_table(SmShared& S, int which, double theta, int kind, int& nfill) {
    const int phys = S.tab_of[which];
    if (!S.key_ok[phys] || S.key[phys] != theta || S.key_kind[phys] != kind) {
        S.key[phys] = theta; S.key_ok[phys] = 1; S.key_kind[phys] = kind;
        S.fill[nfill] = phys; S.fill_theta[nfill] = theta; S.fill_kind[nfill] = kind; nfill++;
    }
}

// thread 0: plan one hill-climb round (scanmatcher.cpp:418-440).
// The round after a winning move re-evaluates, as the move opposite to it, the pose it came from -- whose score (odometry
// gain included) is the previous currentScore, strictly below the present one, so it cannot win the `localScore >
// currentScore` test of scanmatcher.cpp:510.  When that candidate is bit for bit the previous pose ((x + d) - d == x almost
// always) it is counted but not evaluated.
__device__ __noinline__ void plan_round(double lpx, double lpy, double lpt, SmShared& S, int kind) {
    const bool halved = S.bestScore >= S.currentScore;
    if (halved) { S.refinement++; S.adelta *= .5; S.ldelta *= .5; }
    S.bestScore = S.currentScore;
    S.skip = -1;
    if (S.last_move >= 0 && !halved) {
        double back[3];
        candidate(S.cur, S.last_move ^ 1, S.ldelta, S.adelta, back);
        if (back[0] == S.prev[0] && back[1] == S.prev[1] && back[2] == S.prev[2]) S.skip = S.last_move ^ 1;
    }
    int nfill = 0, nc = 0;
    double turn_theta[2] = {0., 0.};
    for (int m = 0; m < 6; m++) {
        if (m == S.skip) continue;
        double cm[3], lx, ly, lt;
        candidate(S.cur, m, S.ldelta, S.adelta, cm);
        laser_pose3(lpx, lpy, lpt, cm[0], cm[1], cm[2], lx, ly, lt);
        S.lx[m] = lx; S.ly[m] = ly;
        if (m >= 4) turn_theta[m - 4] = lt;
        S.cand[nc++] = m;
    }
    // the two tables beside the current one: after a winning turn the theta it came from is the new theta -+ adelta (bit for bit
    // whenever (t + a) - a == t, i.e. almost always) and its table is still there -- in the other slot; swap rather than refill
    {
        const bool want1 = S.skip != 4, want2 = S.skip != 5;
        auto holds = [&](int phys, double theta) { return S.key_ok[phys] && S.key[phys] == theta && S.key_kind[phys] == kind; };
        const int a = S.tab_of[1], b = S.tab_of[2];
        const bool a1 = want1 && holds(a, turn_theta[0]), b2 = want2 && holds(b, turn_theta[1]);
        if (!a1 && !b2 && ((want1 && holds(b, turn_theta[0])) || (want2 && holds(a, turn_theta[1])))) { S.tab_of[1] = b; S.tab_of[2] = a; }
        if (want1) want_table(S, 1, turn_theta[0], kind, nfill);
        if (want2) want_table(S, 2, turn_theta[1], kind, nfill);
    }
    for (int j = nfill; j < 3; j++) S.fill[j] = -1;
    S.ncand = nc; S.mode = 0; S.phase = 1;
}

// scanMatch for one particle per CTA: hill-climb (optimize), accept/reject, likelihoodAndScore; or, with
// A.query >= 0, one standalone score()/likelihoodAndScore() per CTA.
// dynamic shared memory: three (cos, sin) tables of nbeams double2 each.
// QUERY = (A.query >= 0); without it the particle is the block index, which keeps the per-particle pointers in uniform registers
//
// Every thread sums the beams it is dealt in TWO chains -- the even and the odd ones of its sequence -- and adds the two at the
// end.  WIDE = 1 is the shape for few particles per GPU (a launch that does not fill the machine is as long as its longest
// hill climb, one dependent chain of evaluations): 2 x 288 threads per particle, the upper half takes the odd beams of the
// lower half's sequences, hands its chains over through shared memory, and from there on everything is the narrow shape's.
// Same sums in the same order: the two shapes give the same bits.
template <bool K1, bool QUERY, int WIDE>
__global__ void __launch_bounds__(kSmThreads * (1 + WIDE), WIDE ? GM_SM_WIDE_CTAS : GM_SM_MIN_CTAS) k_scanmatch(ScanMatchArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ SmShared S;
    __shared__ SmWide<WIDE> W;
    constexpr int kThreads = kSmThreads * (1 + WIDE);
    const DevParams& P = A.P;
    double2* const tabs = reinterpret_cast<double2*>(smem_raw);
    const unsigned q = blockIdx.x, p = QUERY ? A.particle_idx[q] : q;
    // score() with the default window and a beam-independent free-cell offset: the tables hold the end-point offsets
    // r * (cos, sin), NaN for the beams score() skips, and the beam loop reads nothing else
    const int premul = (K1 && A.P.free_q[0][0] != INT_MIN) ? 1 : 0;
    const int t = threadIdx.x;
    const int half = WIDE ? (t >= kSmThreads ? 1 : 0) : 0, tt = t - half * kSmThreads;  // tt: the thread of the narrow shape this one works for
    MatchCtx M;
    M.dirp_s = &S.dirp; M.bbox_s = &S.bbox; M.pool = A.pool; M.visits = A.visits; M.means = A.means; M.occ = A.occ; M.inv_n = A.inv_n; M.g = A.geom[p];
    M.zero_patch = A.zero_patch;

    if (t == 0) {
        S.dirp = A.dir + (size_t)p * A.dir_cap;
        S.init[0] = S.cur[0] = A.poses[3 * q]; S.init[1] = S.cur[1] = A.poses[3 * q + 1]; S.init[2] = S.cur[2] = A.poses[3 * q + 2];
        S.adelta = P.opt_angular_delta; S.ldelta = P.opt_linear_delta;
        S.refinement = 0; S.evals = 1; S.evaluated = 1; S.bestScore = -1; S.done = 0; S.last_move = -1; S.skip = -1; S.cand[0] = 0;
        S.tab_of[0] = 0; S.tab_of[1] = 1; S.tab_of[2] = 2;
        S.key_ok[0] = S.key_ok[1] = S.key_ok[2] = 0;
        double lx, ly, lt;
        laser_pose(P, S.cur[0], S.cur[1], S.cur[2], lx, ly, lt);
        S.lx[0] = lx; S.ly[0] = ly;
        int nfill = 0;
        want_table(S, 0, lt, A.query == 1 ? 0 : premul, nfill);
        S.fill[1] = S.fill[2] = -1;
        S.ncand = 1; S.mode = A.query == 1 ? 1 : 0; S.phase = (A.query == 0 || A.query == 1) ? 3 : 0;
    }
    if (K1) {
        // bounding box of the patches that can hold occupied cells (a cell of any other patch passes the free-cell test): the
        // free-cell windows (half a map away, see beam_match) normally miss it
        if (t == 0) { S.bb_lo_x = INT_MAX; S.bb_lo_y = INT_MAX; S.bb_hi_x = -1; S.bb_hi_y = -1; }
        __syncthreads();
        const int* dirp = A.dir + (size_t)p * A.dir_cap;
        int lox = INT_MAX, loy = INT_MAX, hix = -1, hiy = -1;
        for (int e = t; e < M.g.npx * M.g.npy; e += kThreads)
            if ((unsigned)__ldg(dirp + e) < (unsigned)A.zero_patch) { const int px = e / M.g.npy, py = e - px * M.g.npy; lox = min(lox, px); hix = max(hix, px); loy = min(loy, py); hiy = max(hiy, py); }
        lox = __reduce_min_sync(0xffffffffu, lox); loy = __reduce_min_sync(0xffffffffu, loy);
        hix = __reduce_max_sync(0xffffffffu, hix); hiy = __reduce_max_sync(0xffffffffu, hiy);
        if ((t & 31) == 0 && hix >= 0) { atomicMin(&S.bb_lo_x, lox); atomicMin(&S.bb_lo_y, loy); atomicMax(&S.bb_hi_x, hix); atomicMax(&S.bb_hi_y, hiy); }
        __syncthreads();
        if (t == 0) S.bbox = S.bb_hi_x < 0 ? make_int4(1, 1, 0, 0) : make_int4(S.bb_lo_x << kPatchMag, S.bb_lo_y << kPatchMag, (S.bb_hi_x << kPatchMag) + kPatch - 1, (S.bb_hi_y << kPatchMag) + kPatch - 1);
    } else if (t == 0) S.bbox = make_int4(INT_MIN, INT_MIN, INT_MAX, INT_MAX);
    for (;;) {
        __syncthreads();  // the plan is visible
        if (S.done) break;
        // ---- trig tables that changed
        bool filled = false;
        for (int j = 0; j < 3; j++) {
            const int phys = S.fill[j];
            if (phys < 0) break;
            const double theta = S.fill_theta[j];
            const int kind = S.fill_kind[j];
            double2* tab = tabs + (size_t)phys * P.nbeams;
            for (unsigned b = P.beam_skip + t; b < P.nbeams; b += kThreads) {
                double sn, cs;
                sincos(theta + __ldg(A.angles + b), &sn, &cs);
                if (kind) {
                    const double r = __ldg(A.ranges + b);
                    const bool skip = !beam_selected(P, b) || r > P.usable_range || r == 0.0;  // scanmatcher.h:189-196
                    const double nan = __longlong_as_double(0x7ff8000000000000ll);
                    tab[b] = skip ? make_double2(nan, nan) : make_double2(r * cs, r * sn);
                } else tab[b] = make_double2(cs, sn);
            }
            filled = true;
        }
        if (filled) __syncthreads();
        // ---- evaluate: thread t works for candidate cand[t % ncand] on beams t / ncand, + ngroups, ...
        const int mode = S.mode, nc = S.ncand;
        const bool multi = nc > 1;
        const int ngroups = nc == 6 ? kSmGroups : kSmThreads / nc, grp = nc == 6 ? tt / 6 : tt / nc;
        const int m = multi ? S.cand[nc == 6 ? tt % 6 : tt - grp * nc] : 0;
        double acc_s = 0., acc_l = 0., acc_s1 = 0., acc_l1 = 0.;  // chains of the even / odd beams of this thread's sequence
        unsigned cnt = 0, valid = 0;
        if (grp < ngroups) {
            const double mlx = S.lx[m], mly = S.ly[m];
            const unsigned tab_s = (unsigned)__cvta_generic_to_shared(tabs + (size_t)S.tab_of[m < 4 ? 0 : m - 3] * P.nbeams);
            // pfree offset: map.getDelta()*freeDelta in score() (scanmatcher.h:208), freeDelta in likelihoodAndScore() (:320)
            const double free_off = mode ? P.free_off_lik : P.free_off_score;
            const int fqx = mode ? P.free_q[1][0] : P.free_q[0][0], fqy = mode ? P.free_q[1][1] : P.free_q[0][1];  // (constant indexes)
            const double cexp = P.cexp, clik = P.clik, noHit = P.no_hit;
            const bool offsets = premul && mode == 0;  // the table in use holds end-point offsets
            unsigned odd = (unsigned)half;  // (two passes, even beams then odd beams, measured slower than the predicated adds below)
            for (unsigned b = P.beam_skip + grp + half * ngroups; b < P.nbeams; b += (WIDE ? 2 * ngroups : ngroups), odd ^= (WIDE ? 0u : 1u)) {
                double2 cs;
                asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(cs.x), "=d"(cs.y) : "r"(tab_s + b * 16u));
                double phx, phy;
                if (offsets) {
                    if (cs.x != cs.x) continue;  // a beam score() skips
                    phx = mlx + cs.x; phy = mly + cs.y;
                } else {
                    const double r = __ldg(A.ranges + b);
                    // score() also skips r == 0 (scanmatcher.h:195); likelihoodAndScore() does not (:303-308)
                    if (!beam_selected(P, b) || r > P.usable_range || (mode == 0 && r == 0.0)) continue;
                    phx = mlx + r * cs.x; phy = mly + r * cs.y;
                }
                double mx, my;
                valid++;
                const bool found = beam_match<K1>(P, M, mlx, mly, phx, phy, b, tab_s, A.ranges, free_off, fqx, fqy, mx, my);
                if (found) {  // ((-1/sigma)*mu)*mu, point.h:34-49
                    const double e = exp_neg((mx * cexp) * mx + (my * cexp) * my, A.inv_n + kInvTable);
                    if (odd) acc_s1 += e; else acc_s += e;
                    cnt++;
                }
                if (mode) { const double f = clik * (mx * mx + my * my), a = found ? f : noHit; if (odd) acc_l1 += a; else acc_l += a; }
            }
        }
        if constexpr (WIDE != 0) {  // the upper half's chains go to the thread of the lower half they belong to
            if (half) { W.s[tt] = acc_s1; W.l[tt] = acc_l1; W.c[tt] = cnt; W.v[tt] = valid; }
            __syncthreads();
            if (!half) { acc_s1 = W.s[tt]; acc_l1 = W.l[tt]; cnt += W.c[tt]; valid += W.v[tt]; }
        }
        acc_s = acc_s + acc_s1; acc_l = acc_l + acc_l1;
        // ---- reduce (fixed shape: the sums do not depend on scheduling)
        const int warp = t >> 5, lane = t & 31;
        if (multi) {
            if (!half) S.part[tt] = acc_s;
            __syncthreads();
            if (warp < nc) {  // warp w sums the partials of candidate cand[w]: lane j takes groups j and j + 32 (ngroups <= 64)
                double v = lane < ngroups ? S.part[nc * lane + warp] : 0.;
                if (lane + 32 < ngroups) v += S.part[nc * (lane + 32) + warp];
                v = warp_sum(v);
                if (lane == 0) S.red[S.cand[warp]] = v;
            }
        } else {
            acc_s = warp_sum(acc_s); acc_l = warp_sum(acc_l);
            cnt = __reduce_add_sync(0xffffffffu, cnt); valid = __reduce_add_sync(0xffffffffu, valid);
            if (lane == 0 && !half) { S.red[warp] = acc_s; S.red_l[warp] = acc_l; S.red_c[warp] = cnt; S.red_v[warp] = valid; }
        }
        __syncthreads();
        if (t != 0) continue;

        // ---- thread 0: the reference's sequential decision logic
        if (S.phase == 1) {
            // scanmatcher.cpp:441-517: pick the best of the six moves (strict >, first wins)
            S.evals += 6;
            S.evaluated += (unsigned)S.ncand;
            int bm = -1;
            for (int mm = 0; mm < 6; mm++) {
                if (mm == S.skip) continue;
                double odo_gain = 1;
                if (P.ang_rel > 0. || P.lin_rel > 0.) {
                    double cm[3];
                    candidate(S.cur, mm, S.ldelta, S.adelta, cm);
                    odo_gain = odo_gain_of(P.ang_rel, P.lin_rel, S.init[0], S.init[1], S.init[2], cm[0], cm[1], cm[2]);
                }
                const double localScore = odo_gain * S.red[mm];
                if (localScore > S.currentScore) { S.currentScore = localScore; bm = mm; }
            }
            S.last_move = bm;
            if (bm >= 0) {
                double nb[3];
                candidate(S.cur, bm, S.ldelta, S.adelta, nb);
                S.prev[0] = S.cur[0]; S.prev[1] = S.cur[1]; S.prev[2] = S.cur[2];
                S.cur[0] = nb[0]; S.cur[1] = nb[1]; S.cur[2] = nb[2];
                if (bm >= 4) {  // a turn won: its table now holds the current theta
                    const int o = S.tab_of[0];
                    S.tab_of[0] = S.tab_of[bm - 3]; S.tab_of[bm - 3] = o;
                }
            }
            if (S.currentScore > S.bestScore || S.refinement < P.opt_iters) { plan_round(P.laser_pose[0], P.laser_pose[1], P.laser_pose[2], S, premul); continue; }
        } else if (S.phase == 0) {
            double v = 0.;
            for (int w = 0; w < kSmThreads / 32; w++) v += S.red[w];
            S.currentScore = v;
            plan_round(P.laser_pose[0], P.laser_pose[1], P.laser_pose[2], S, premul);
            continue;
        } else {
            // phase 2: likelihoodAndScore of the accepted pose; phase 3: standalone query
            double ss = 0., ll = 0.; unsigned cc = 0, vv = 0;
            for (int w = 0; w < kSmThreads / 32; w++) { ss += S.red[w]; ll += S.red_l[w]; cc += S.red_c[w]; vv += S.red_v[w]; }
            if (S.phase == 3) {
                A.s[q] = ss;
                if (A.query == 1) { A.l[q] = ll; A.c[q] = cc; }
            } else {
                A.poses[3 * p] = S.cur[0]; A.poses[3 * p + 1] = S.cur[1]; A.poses[3 * p + 2] = S.cur[2];
                A.score[p] = S.bestScore; A.s[p] = ss; A.l[p] = ll; A.c[p] = cc; A.evals[p] = S.evals;
                // the all-gather of the sharded filter happens here: every rank's gather buffer gets this particle's row
                // (its own HBM, or a peer's over NVLink); the flag below tells the ranks when a whole slice has landed
#pragma unroll
                for (int r = 0; r < kMaxRanks; r++) {  // (constant indexes: the parameter block is not copied to local memory)
                    if (r < A.nranks) {
                        double* row = A.rows[r] + 5 * (size_t)(A.first_global + p);
                        row[0] = S.cur[0]; row[1] = S.cur[1]; row[2] = S.cur[2]; row[3] = S.bestScore; row[4] = ll;
                    }
                }
                atomicAdd(&A.counters->score_evals, (unsigned long long)S.evals);
                atomicAdd(&A.counters->beam_evals, (unsigned long long)S.evaluated * A.valid_score_beams + vv);
            }
            S.done = 1;
            continue;
        }
        if (A.query == 2) {  // standalone optimize(): report the climbed pose and its score
            A.poses_out[3 * q] = S.cur[0]; A.poses_out[3 * q + 1] = S.cur[1]; A.poses_out[3 * q + 2] = S.cur[2];
            A.score[q] = S.bestScore;
            S.done = 1;
            continue;
        }
        // ---- optimize() finished (gridslamprocessor.hxx:34-48): keep the motion-model pose when the match is
        // not good enough, then likelihoodAndScore at the accepted pose
        if (!(S.bestScore > A.minimum_score)) { S.cur[0] = S.init[0]; S.cur[1] = S.init[1]; S.cur[2] = S.init[2]; }
        double lx, ly, lt;
        laser_pose(P, S.cur[0], S.cur[1], S.cur[2], lx, ly, lt);
        S.lx[0] = lx; S.ly[0] = ly;
        int nfill = 0;
        want_table(S, 0, lt, 0, nfill);
        for (int j = nfill; j < 3; j++) S.fill[j] = -1;
        S.ncand = 1; S.mode = 1; S.phase = 2;
    }
    // the last CTA publishes the sequence number of this scan where every rank's host can see it (release pattern: rows,
    // system-wide fence, count; the CTA that counts last fences again and writes the flag)
    if (!QUERY && A.flag && t == 0) {
        __threadfence_system();
        if (atomicAdd(A.done_counter, 1u) == gridDim.x - 1) {
            *A.done_counter = 0;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned*>(A.flag) = A.seq;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// computeActiveArea part 1: bounding box of the laser pose and the clipped endpoints, and the
// Map::resize arithmetic when it leaves the map (scanmatcher.cpp:96-166, map.h:218-241).
__global__ void __launch_bounds__(256) k_bounds(BoundsArgs A) {
    __shared__ double red[4 * 32];
    const DevParams& P = A.P;
    const unsigned p = blockIdx.x;
    const Geom g = A.geom[p];
    double lx, ly, lt;
    laser_pose(P, A.poses[3 * p], A.poses[3 * p + 1], A.poses[3 * p + 2], lx, ly, lt);
    double minx = lx, miny = ly, maxx = lx, maxy = ly;
    int too_long = 0;
    const int p0x = w2m(lx, P.cx, P.delta, g.sx2), p0y = w2m(ly, P.cy, P.delta, g.sy2);
    for (unsigned b = P.beam_skip + threadIdx.x; b < P.nbeams; b += blockDim.x) {
        double r = __ldg(A.ranges + b);
        if (r > P.laser_max_range || r == 0.0 || isnan(r)) continue;
        double d = r > P.usable_range ? P.usable_range : r;
        double s, c;
        sincos(lt + __ldg(A.angles + b), &s, &c);
        double phx = lx + d * c, phy = ly + d * s;
        minx = fmin(minx, phx); miny = fmin(miny, phy); maxx = fmax(maxx, phx); maxy = fmax(maxy, phy);
        int p1x = w2m(phx, P.cx, P.delta, g.sx2), p1y = w2m(phy, P.cy, P.delta, g.sy2);
        if (max(abs(p1x - p0x), abs(p1y - p0y)) + 1 > kLineCap) too_long = 1;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    minx = warp_min(minx); miny = warp_min(miny); maxx = warp_max(maxx); maxy = warp_max(maxy);
    too_long = __any_sync(0xffffffffu, too_long);
    if (lane == 0) { red[warp] = minx; red[32 + warp] = miny; red[64 + warp] = maxx; red[96 + warp] = maxy; if (too_long) atomicCAS(&A.counters->error, 0, -5); }
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int w = 1; w < nw; w++) { minx = fmin(minx, red[w]); miny = fmin(miny, red[32 + w]); maxx = fmax(maxx, red[64 + w]); maxy = fmax(maxy, red[96 + w]); }
    // the box starts from the current map extent (scanmatcher.cpp:102-114)
    const int mw = g.npx << kPatchMag, mh = g.npy << kPatchMag;
    const double lminx = m2w(0, P.cx, P.delta, g.sx2), lminy = m2w(0, P.cy, P.delta, g.sy2);
    const double lmaxx = m2w(mw - 1, P.cx, P.delta, g.sx2), lmaxy = m2w(mh - 1, P.cy, P.delta, g.sy2);
    minx = fmin(minx, lminx); miny = fmin(miny, lminy); maxx = fmax(maxx, lmaxx); maxy = fmax(maxy, lmaxy);
    auto inside = [&](double x, double y) {
        int ix = w2m(x, P.cx, P.delta, g.sx2), iy = w2m(y, P.cy, P.delta, g.sy2);
        return ix >= 0 && iy >= 0 && (ix >> kPatchMag) < g.npx && (iy >> kPatchMag) < g.npy;
    };
    Geom ng = g;
    int4 shift = make_int4(0, 0, 0, 0);
    if (!inside(minx, miny) || !inside(maxx, maxy)) {
        minx = (minx >= lminx) ? lminx : minx - P.enlarge_step;
        maxx = (maxx <= lmaxx) ? lmaxx : maxx + P.enlarge_step;
        miny = (miny >= lminy) ? lminy : miny - P.enlarge_step;
        maxy = (maxy <= lmaxy) ? lmaxy : maxy + P.enlarge_step;
        int iminx = w2m(minx, P.cx, P.delta, g.sx2), iminy = w2m(miny, P.cy, P.delta, g.sy2);
        int imaxx = w2m(maxx, P.cx, P.delta, g.sx2), imaxy = w2m(maxy, P.cy, P.delta, g.sy2);
        int pxmin = (int)floorf((float)iminx / (float)kPatch), pxmax = (int)ceilf((float)imaxx / (float)kPatch);
        int pymin = (int)floorf((float)iminy / (float)kPatch), pymax = (int)ceilf((float)imaxy / (float)kPatch);
        ng.npx = pxmax - pxmin; ng.npy = pymax - pymin;
        ng.sx2 = g.sx2 - pxmin * kPatch; ng.sy2 = g.sy2 - pymin * kPatch;
        shift = make_int4(pxmin, pymin, 1, 0);
        atomicAdd(&A.counters->resizes, 1ull);
        atomicMax(&A.counters->need_dir_cap, ng.npx * ng.npy);
        atomicExch(&A.counters->any_resize, 1);
    }
    A.geom_new[p] = ng;
    A.shift[p] = shift;
}

// HierarchicalArray2D::resize for the flagged particles: dir_alt[p] <- re-based dir[p]; patches that fall
// outside the new directory lose this particle's reference (harray2d.h:80-109).
__global__ void k_relayout(RelayoutArgs A) {
    const unsigned p = blockIdx.x;
    if (A.counters->need_dir_cap > A.dir_cap_old && A.dir_cap_new == A.dir_cap_old) return;  // the host grows the stride first
    const int4 sh = A.shift[p];
    if (!sh.z) return;
    const Geom og = A.geom[p], ng = A.geom_new[p];
    const int* src = A.dir + (size_t)p * A.dir_cap_old;
    int* dst = A.dir_alt + (size_t)p * A.dir_cap_new;
    for (int e = threadIdx.x; e < ng.npx * ng.npy; e += blockDim.x) {
        int x = e / ng.npy + sh.x, y = e % ng.npy + sh.y;
        dst[e] = (x >= 0 && y >= 0 && x < og.npx && y < og.npy) ? src[x * og.npy + y] : -1;
    }
    for (int e = threadIdx.x; e < og.npx * og.npy; e += blockDim.x) {
        int x = e / og.npy - sh.x, y = e % og.npy - sh.y;
        int id = src[e];
        if (id >= 0 && !(x >= 0 && y >= 0 && x < ng.npx && y < ng.npy))
            if (atomicSub(&A.pool.refcnt[id], 1) == 1) A.pool.pending[atomicAdd(A.pool.pending_n, 1)] = id;
    }
}
__global__ void k_relayout_commit(RelayoutArgs A) {
    const unsigned p = blockIdx.x;
    if (A.counters->need_dir_cap > A.dir_cap_old && A.dir_cap_new == A.dir_cap_old) return;
    const int4 sh = A.shift[p];
    const Geom ng = A.geom_new[p];
    // after a capacity growth every particle moves to the new stride; otherwise only resized ones change
    const int* src = sh.z ? A.dir_alt + (size_t)p * A.dir_cap_new : A.dir + (size_t)p * A.dir_cap_old;
    int* dst = A.dir_out + (size_t)p * A.dir_cap_new;
    if (src != dst)
        for (int e = threadIdx.x; e < ng.npx * ng.npy; e += blockDim.x) dst[e] = src[e];
    if (threadIdx.x == 0) A.geom[p] = ng;
}

// ------------------------------------------------------------------------------------------------

// computeActiveArea part 2 + allocActiveArea + registerScan for one particle per CTA.
//   0. one thread per beam: endpoint cell, (float) endpoint, validity -- sin/cos and the reference's divide+round
//      once per beam, kept in shared memory for the later phases;
//   1. mark the patches under every ray cell (all but the last point) and under the endpoint when
//      d < usableRange (scanmatcher.cpp:181-218), in a shared-memory bitmap over the directory;
//   2. make every active patch private: create it, or copy it when another particle still references
//      it (harray2d.h:169-185 copies always; copying only shared patches yields identical cells);
//   3. visits++ along the rays (scanmatcher.cpp:317-325); integer atomics commute, so the counts are bit-exact;
//   4. endpoint hits (scanmatcher.cpp:327-334): float `acc` is order dependent, so hits are grouped by cell (a hash
//      table in shared memory) and each cell's beams are accumulated in ascending order by its first beam.
// Ray walks (1 and 3): rays are cut into work items of 32 cells; a thread starts its item from the closed form of the
// reference's Bresenham state and steps incrementally, looking the patch up again only when the cell leaves it.
struct BeamRec { int p1x, p1y; float hx, hy; int flags; };  // flags: 1 = ray is traced, 2 = endpoint is a hit, 4 = ... inside the map

constexpr int kRayChunk = 32;  // cells per work item of a ray walk

// number of cells visited on the line p0 -> p1: every point but p1 (scanmatcher.cpp:317-325)
__device__ __forceinline__ int ray_cells(int p0x, int p0y, int p1x, int p1y) { return max(abs(p1x - p0x), abs(p1y - p0y)); }

// visit chunk `chunk` (kRayChunk cells) of the line p0 -> p1 except p1 itself, following the reference's Bresenham
// (gridlinetraversal.h:20-108: major axis upward from the lower endpoint): the chunk starts from the closed form of the
// walk's state and then steps incrementally; f(x, y, px, py, enteredNewPatch)
template <class F>
__device__ __forceinline__ void walk_ray(int p0x, int p0y, int p1x, int p1y, int chunk, F f) {
    const Line L = make_line(p0x, p0y, p1x, p1y);
    const int m = L.n - 1;
    const int jlo = L.from_end ? 1 : 0;  // p1 is the LOW end when the walk runs from the end: skip it
    int j = jlo + chunk * kRayChunk;
    const int jend = min(jlo + m, j + kRayChunk);
    if (j >= jend) return;
    // state after j steps of the reference's loop: a = a_lo + j, b = b_lo + bstep * inc(j), d = 2db(j+1) - da - 2da*inc(j),
    // inc(j) = floor((2 db j + da) / (2 da))
    int inc = 0;
    if (j && L.da) {
        const unsigned num = 2u * (unsigned)L.db * (unsigned)j + (unsigned)L.da, den = 2u * (unsigned)L.da;
        int q = (int)((double)num * (1.0 / (double)den));  // exact after the one-step correction below
        const long long rem = (long long)num - (long long)q * den;
        q += rem >= (long long)den ? 1 : (rem < 0 ? -1 : 0);
        inc = q;
    }
    int a = L.a_lo + j, bb = L.b_lo + L.bstep * inc;
    int d = 2 * L.db * (j + 1) - L.da - 2 * L.da * inc;
    const int incr1 = 2 * L.db, incr2 = 2 * (L.db - L.da);
    int lpx = INT_MIN, lpy = INT_MIN;
    for (; j < jend; j++) {
        const int x = L.steep ? bb : a, y = L.steep ? a : bb;
        const int px = x >> kPatchMag, py = y >> kPatchMag;
        const bool np = px != lpx || py != lpy;
        lpx = px; lpy = py;
        f(x, y, px, py, np);
        a++;
        if (d < 0) d += incr1; else { bb += L.bstep; d += incr2; }
    }
}

// minor coordinate of the line after j steps from its low end: the closed form of the reference's Bresenham (gm_device.cuh)
__device__ __forceinline__ int line_minor(const Line& L, int j) {
    int inc = 0;
    if (j && L.da) {
        const unsigned num = 2u * (unsigned)L.db * (unsigned)j + (unsigned)L.da, den = 2u * (unsigned)L.da;
        int q = (int)((double)num * (1.0 / (double)den));
        const long long rem = (long long)num - (long long)q * den;
        q += rem >= (long long)den ? 1 : (rem < 0 ? -1 : 0);
        inc = q;
    }
    return L.b_lo + L.bstep * inc;
}

// The patches (as px, py pairs) under chunk `chunk` of the line p0 -> p1 (p1 excluded), without walking it: a chunk of
// kRayChunk <= 32 cells crosses at most one patch boundary along the major axis, and between two cells the minor
// coordinate moves monotonically by at most one per step, so each of the (at most two) major-axis segments lies in the
// patch rows of its first and last cell.  Returns the number of pairs written (0 or 4, with duplicates).
__device__ __forceinline__ int item_patches(int p0x, int p0y, int p1x, int p1y, int chunk, int (&px)[4], int (&py)[4]) {
    const Line L = make_line(p0x, p0y, p1x, p1y);
    const int m = L.n - 1, jlo = L.from_end ? 1 : 0;
    const int j0 = jlo + chunk * kRayChunk, j1 = min(jlo + m, j0 + kRayChunk) - 1;  // first / last step of the chunk
    if (j0 > j1) return 0;
    const int a0 = L.a_lo + j0, a1 = L.a_lo + j1;
    const int aB = ((a0 >> kPatchMag) + 1) << kPatchMag;  // first major coordinate of the next patch
    // four fixed slots (a slot that is not needed repeats an earlier one): indexing the arrays with a running count would
    // put them in local memory
    const int e1 = min(a1, aB - 1);
    const bool second = aB <= a1;
    const int b0 = line_minor(L, j0);
    const int b1 = e1 != a0 ? line_minor(L, e1 - L.a_lo) : b0;
    const int a2 = second ? aB : a0, a3 = second ? a1 : a0;
    const int b2 = second ? line_minor(L, aB - L.a_lo) : b0;
    const int b3 = second ? (a1 != aB ? line_minor(L, j1) : b2) : b0;
    px[0] = (L.steep ? b0 : a0) >> kPatchMag; py[0] = (L.steep ? a0 : b0) >> kPatchMag;
    px[1] = (L.steep ? b1 : e1) >> kPatchMag; py[1] = (L.steep ? e1 : b1) >> kPatchMag;
    px[2] = (L.steep ? b2 : a2) >> kPatchMag; py[2] = (L.steep ? a2 : b2) >> kPatchMag;
    px[3] = (L.steep ? b3 : a3) >> kPatchMag; py[3] = (L.steep ? a3 : b3) >> kPatchMag;
    return 4;
}

constexpr int kRegThreads = 1024;
constexpr int kTileWords = kPatchCells / 2;  // a tile counts visits of one patch: 1024 cells x 16 bit, two per word
// hit grouping (k_register phase 4), per unit of sort_n: 2 table keys (8 B) + 2 first + 2 last (4 B each) + 1 slot (4 B)
constexpr int kHitBytes = 36;
constexpr int kMaxTiles = 112;  // tiles that fit the 227 KB of shared memory a CTA can have

// dynamic shared memory of k_register (see register_smem_bytes): active-patch bitmap over the directory, its
// running popcount per word, per-beam records, visit-count tiles (reused by the hit table afterwards)
struct RegSmem {
    unsigned* active; unsigned short* prefix; BeamRec* rec; int* items; unsigned* tiles; unsigned long long* keys;  // keys alias tiles
};
__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }
__device__ __forceinline__ RegSmem carve(unsigned char* raw, int bitmap_words, int sort_n, int nbeams) {
    RegSmem m;
    size_t o = 0;
    m.active = reinterpret_cast<unsigned*>(raw + o); o += align16((size_t)bitmap_words * 4);
    m.prefix = reinterpret_cast<unsigned short*>(raw + o); o += align16((size_t)bitmap_words * 2);
    m.rec = reinterpret_cast<BeamRec*>(raw + o); o += align16((size_t)nbeams * sizeof(BeamRec));
    m.items = reinterpret_cast<int*>(raw + o); o += align16((size_t)(nbeams + 1) * 4);
    m.tiles = reinterpret_cast<unsigned*>(raw + o);
    m.keys = reinterpret_cast<unsigned long long*>(raw + o);  // the hit table is built after the last tile flush
    return m;
}

__global__ void __launch_bounds__(kRegThreads, 1) k_register(RegisterArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevParams& P = A.P;
    const unsigned p = blockIdx.x;
    // launched without waiting for k_bounds' verdict: stand down when a ray is too long, or when a resize needs a longer
    // directory stride than the allocation has (the host grows it and registers again, see gmi_register_finish)
    if (A.counters->error != 0 || A.counters->need_dir_cap > A.dir_cap) return;
    const Geom g = A.geom[p];
    if (g.npx > 2047 || g.npy > 2047) {  // the ray walk packs a cell into one word, 16 bits a coordinate (3.2 km a side at 5 cm)
        if (threadIdx.x == 0) atomicCAS(&A.counters->error, 0, -2);
        return;
    }
    int* dirp = A.dir + (size_t)p * A.dir_cap;
    const int nent = g.npx * g.npy, nwords = (nent + 31) >> 5;
    const RegSmem M = carve(smem_raw, A.bitmap_words, A.sort_n, (int)P.nbeams);
    unsigned* active = M.active;
    BeamRec* rec = M.rec;
    __shared__ unsigned long long s_free, s_hit, s_cow, s_alloc;
    __shared__ int s_part[kRegThreads / 32];
    __shared__ int s_nactive, s_nwork;
    __shared__ int s_tile_id[kMaxTiles];
    __shared__ int s_gitems[GM_MAXBEAMS_DEV / 32 + 2];  // prefix of the chunk counts of the groups of 32 neighbouring beams
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;

    long long tph = clock64();
#define PHASE(k) do { if (threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd(&A.counters->reg_phase[k], (unsigned long long)(now_ - tph)); tph = now_; } } while (0)
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) active[i] = 0u;
    if (threadIdx.x == 0) { s_free = 0; s_hit = 0; s_cow = 0; s_alloc = 0; }
    double lx, ly, lt;
    laser_pose(P, A.poses[3 * p], A.poses[3 * p + 1], A.poses[3 * p + 2], lx, ly, lt);
    const int p0x = w2m(lx, P.cx, P.delta, g.sx2), p0y = w2m(ly, P.cy, P.delta, g.sy2);
    __syncthreads();

    // ---- 0. per-beam endpoints
    for (unsigned b = threadIdx.x; b < P.nbeams; b += blockDim.x) {
        BeamRec R; R.p1x = R.p1y = 0; R.hx = R.hy = 0.f; R.flags = 0;
        double r = __ldg(A.ranges + b);
        bool ok = b >= P.beam_skip, hit;
        if (P.generate_map) {
            ok = ok && !(r > P.laser_max_range || r == 0.0 || isnan(r));
            hit = !(r >= P.usable_range);  // d < usableRange after clipping
            if (r > P.usable_range) r = P.usable_range;
        } else {
            ok = ok && !(r > P.laser_max_range || r > P.usable_range || r == 0.0 || isnan(r));
            hit = true;
        }
        if (ok) {
            double s, c;
            sincos(lt + __ldg(A.angles + b), &s, &c);
            const double phx = lx + r * c, phy = ly + r * s;
            R.p1x = w2m(phx, P.cx, P.delta, g.sx2); R.p1y = w2m(phy, P.cy, P.delta, g.sy2);
            R.hx = (float)phx; R.hy = (float)phy;
            R.flags = (P.generate_map ? 1 : 0) | (hit ? 2 : 0);
            if (hit && R.p1x >= 0 && R.p1y >= 0 && (R.p1x >> kPatchMag) < g.npx && (R.p1y >> kPatchMag) < g.npy) {
                const int e = (R.p1x >> kPatchMag) * g.npy + (R.p1y >> kPatchMag);
                atomicOr(&active[e >> 5], 1u << (e & 31));
                R.flags |= 4;  // the hit lands inside the map
            }
        }
        rec[b] = R;
        const int chunks = (R.flags & 1) ? (ray_cells(p0x, p0y, R.p1x, R.p1y) + kRayChunk - 1) / kRayChunk : 0;
        M.items[b + 1] = chunks;
        // a warp holds 32 neighbouring beams here: the longest of them sets the chunk count of the group (ray counting below)
        const int gmax = __reduce_max_sync(__activemask(), chunks);
        if (lane == 0) s_gitems[(b >> 5) + 1] = gmax;
    }
    if (threadIdx.x == 0) s_gitems[0] = 0;
    __syncthreads();
    // work items of the ray walks: items[b] .. items[b+1] are beam b's chunks (inclusive scan by one warp)
    if (warp == 0) {
        int run = 0;
        for (unsigned b0 = 0; b0 < P.nbeams; b0 += 32) {
            const unsigned b = b0 + lane;
            int v = b < P.nbeams ? M.items[b + 1] : 0;
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
            if (b < P.nbeams) M.items[b + 1] = run + v;
            run += __shfl_sync(0xffffffffu, v, 31);
        }
        if (lane == 0) M.items[0] = 0;
    } else if (warp == 1) {
        const int ng = ((int)P.nbeams + 31) >> 5;  // at most 64 groups: two per lane
        int run = 0;
        for (int g0 = 0; g0 < ng; g0 += 32) {
            const int gi = g0 + lane;
            int v = gi < ng ? s_gitems[gi + 1] : 0;
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
            if (gi < ng) s_gitems[gi + 1] = run + v;
            run += __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    const int nitems = M.items[P.nbeams];
    const int ngroups32 = ((int)P.nbeams + 31) >> 5, ngitems = s_gitems[ngroups32];
    PHASE(0);

    // ---- 1. active area: patches under the rays (from each chunk's end points, see item_patches)
    for (int it = threadIdx.x; it < nitems; it += blockDim.x) {
        int lo = 0, hi = (int)P.nbeams - 1;  // the beam whose chunk range holds item `it`
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (M.items[mid] <= it) lo = mid; else hi = mid - 1; }
        const BeamRec R = rec[lo];
        int qx[4], qy[4];
        const int nq = item_patches(p0x, p0y, R.p1x, R.p1y, it - M.items[lo], qx, qy);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < nq && (unsigned)qx[k] < (unsigned)g.npx && (unsigned)qy[k] < (unsigned)g.npy) {
                const int e = qx[k] * g.npy + qy[k];
                if (!((active[e >> 5] >> (e & 31)) & 1u)) atomicOr(&active[e >> 5], 1u << (e & 31));
            }
    }
    __syncthreads();
    PHASE(1);

    // ---- 1b. rank of every active patch: prefix[w] = number of active patches in words < w
    {
        const int per = (nwords + blockDim.x - 1) / blockDim.x;
        const int w0 = threadIdx.x * per, w1 = min(nwords, w0 + per);
        int c = 0;
        for (int w = w0; w < w1; w++) c += __popc(active[w]);
        int incl = c;  // inclusive scan over the block: warp scan, then warp totals
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_part[warp] = incl;
        __syncthreads();
        int base = 0;
        for (int w = 0; w < warp; w++) base += s_part[w];
        int run = base + incl - c;
        for (int w = w0; w < w1; w++) { M.prefix[w] = (unsigned short)run; run += __popc(active[w]); }
        if (threadIdx.x == blockDim.x - 1) { s_nactive = base + incl; atomicMax(&A.counters->max_active, base + incl); }
    }
    __syncthreads();
    const int nactive = s_nactive;
    PHASE(2);

    // directory entry of the rank-th active patch, from the prefix sums
    auto entry_of_rank = [&](int rank) -> int {
        int lo = 0, hi = nwords - 1;  // last word whose prefix <= rank and that holds the rank-th bit
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((int)M.prefix[mid] <= rank) lo = mid; else hi = mid - 1; }
        return (lo << 5) + (int)__fns(active[lo], 0, rank - (int)M.prefix[lo] + 1);
    };

    // ---- 2. copy-on-write.  (a) a thread per active patch reads its directory entry and reference count and lists the
    // entries that need a patch of their own (the tile area is free until phase 3); (b) a warp per listed entry.
    int* work = reinterpret_cast<int*>(M.tiles);
    const int work_cap = max(A.tile_cap * kTileWords, A.sort_n * (kHitBytes / 4));
    for (int r0 = 0; r0 < nactive; r0 += work_cap) {
        if (threadIdx.x == 0) s_nwork = 0;
        __syncthreads();
        for (int r = r0 + threadIdx.x; r < min(nactive, r0 + work_cap); r += blockDim.x) {
            const int e = entry_of_rank(r);
            const int id = dirp[e];
            if (id < 0 || __ldcg(A.pool.refcnt + id) > 1) work[atomicAdd(&s_nwork, 1)] = e;
        }
        __syncthreads();
        const int nwork = s_nwork;
        for (int k = warp; k < nwork; k += nw) {
            const int e = work[k];
            const int id = dirp[e], mode = id < 0 ? 1 : 2;  // 1 zero-fill a new patch, 2 copy
            // a copy keeps the kind of its source; a new patch starts as a free-space patch (4 KB) and moves to a hit-capable id
            // only if an end point lands in it (phase 4)
            const bool src_h = mode == 2 && id < A.pool.cap_h;
            int nid = -1;
            if (lane == 0) nid = pool_alloc(A.pool, A.counters, src_h);
            nid = __shfl_sync(0xffffffffu, nid, 0);
            if (nid < 0) continue;
            const bool dst_h = nid < A.pool.cap_h;  // (a free-space patch may land on a hit-capable id when the other stack is empty)
            int4* dst = A.pool.cells + (size_t)nid * kPatchCells;
            int4* vdst = reinterpret_cast<int4*>(A.pool.visits + (size_t)nid * kPatchCells);
            int4* mdst = reinterpret_cast<int4*>(A.pool.means + (size_t)nid * kPatchCells);
            if (dst_h && !src_h) {  // hit planes of a patch without hits: n = 0, nothing occupied; means are only read where n > 0
                for (int i = lane; i < kPatchCells; i += 32) dst[i] = make_int4(0, 0, 0, 0);
                A.pool.occ[(size_t)nid * kPatch + lane] = 0u;
            }
            if (mode == 1) {
                for (int i = lane; i < kPatchCells / 4; i += 32) vdst[i] = make_int4(0, 0, 0, 0);
            } else {
                // four independent 16-byte loads in flight per lane, then the stores (source and destination come from
                // the same array: the compiler keeps a plain copy loop strictly load -> store -> load)
                const int4* vsrc = reinterpret_cast<const int4*>(A.pool.visits + (size_t)id * kPatchCells);
                for (int i0 = lane; i0 < kPatchCells / 4; i0 += 128) {
                    int4 c[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) c[k] = __ldcg(vsrc + i0 + 32 * k);
#pragma unroll
                    for (int k = 0; k < 4; k++) vdst[i0 + 32 * k] = c[k];
                }
                if (src_h) {
                    const int4* src = A.pool.cells + (size_t)id * kPatchCells;
                    bool hits = false;
                    for (int i0 = lane; i0 < kPatchCells; i0 += 128) {
                        int4 c[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) c[k] = __ldcg(src + i0 + 32 * k);
#pragma unroll
                        for (int k = 0; k < 4; k++) { dst[i0 + 32 * k] = c[k]; hits = hits || c[k].z != 0; }
                    }
                    if (__any_sync(0xffffffffu, hits)) {  // (a hit-capable id may hold a patch without hits: no mean to copy)
                        const int4* msrc = reinterpret_cast<const int4*>(A.pool.means + (size_t)id * kPatchCells);
                        for (int i0 = lane; i0 < kPatchCells; i0 += 128) {
                            int4 c[4];
#pragma unroll
                            for (int k = 0; k < 4; k++) c[k] = __ldcg(msrc + i0 + 32 * k);
#pragma unroll
                            for (int k = 0; k < 4; k++) mdst[i0 + 32 * k] = c[k];
                        }
                    }
                    A.pool.occ[(size_t)nid * kPatch + lane] = __ldcg(A.pool.occ + (size_t)id * kPatch + lane);
                }
            }
            __syncwarp();
            if (lane == 0) {
                A.pool.refcnt[nid] = 1;
                dirp[e] = nid;
                if (mode == 2) {
                    // release our share only AFTER the copy: whoever then reads refcnt==1 is the sole owner
                    // and may write in place (see DESIGN.md "copy-on-write protocol")
                    __threadfence();
                    if (atomicSub(&A.pool.refcnt[id], 1) == 1) A.pool.pending[atomicAdd(A.pool.pending_n, 1)] = id;
                    atomicAdd(&s_cow, 1ull);
                } else atomicAdd(&s_alloc, 1ull);
            }
        }
        __syncthreads();
    }
    __threadfence();
    __syncthreads();
    PHASE(3);

    // ---- 3. free-space updates (scanmatcher.cpp:317-325).  visits++ is counted per cell in shared-memory tiles
    // (16-bit counters: a cell is crossed by at most nbeams < 2048 rays), `tile_cap` active patches per pass, then
    // added to the now private patches with plain read-modify-writes: no global atomics.
    unsigned long long nfree = 0;
    for (int first = 0; first < nactive && P.generate_map; first += A.tile_cap) {
        const int ntile = min(A.tile_cap, nactive - first);
        for (int i = threadIdx.x; i < ntile * kTileWords; i += blockDim.x) M.tiles[i] = 0u;
        if ((int)threadIdx.x < ntile) {  // patch of tile t: the directory entry of rank first + t, found from the prefix sums
            s_tile_id[threadIdx.x] = dirp[entry_of_rank(first + threadIdx.x)];
        }
        __syncthreads();
        // the counting below keeps to shared memory: meanwhile pull the visit planes of this pass (4 KB per patch) into
        // L2, where the flush will find them
        for (int i = threadIdx.x; i < ntile * 32; i += blockDim.x) {
            const int id = s_tile_id[i >> 5];
            if (id >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.pool.visits + (size_t)id * kPatchCells + ((i & 31) << 5)));
        }
        // A work item is a chunk of kRayChunk steps, counted from the sensor cell, of 32 NEIGHBOURING beams, one beam per lane.
        // Which cells a ray visits does not depend on the order: every lane walks the reference's Bresenham (gridlinetraversal.h:
        // 20-108: major axis upward from the lower end point) from the sensor outward -- downward along the major axis when the
        // sensor is the upper end -- with the walk's state kept as (inc, e): minor = b_lo + bstep * inc, inc = floor((2 db j + da) /
        // (2 da)), e = the remainder.  Neighbouring beams cross the same cells until they fan out (0.25 deg: about 11 m at 5 cm):
        // lanes in the same step whose cells coincide form runs, and one shared-memory atomic per run adds the run's length.
        int gb = 0;
        for (int it = warp; it < ngitems; it += nw) {
            while (s_gitems[gb + 1] <= it) gb++;  // items come in ascending order: the search resumes where it stopped
            const int i0 = (it - s_gitems[gb]) * kRayChunk;
            const int b = (gb << 5) + lane;
            BeamRec R; R.p1x = p0x; R.p1y = p0y; R.hx = R.hy = 0.f; R.flags = 0;
            if (b < (int)P.nbeams) R = rec[b];
            const Line L = make_line(p0x, p0y, R.p1x, R.p1y);
            const int m = (R.flags & 1) ? L.n - 1 : 0;  // cells of the ray: every point but the end point
            const int steps = max(0, min(kRayChunk, m - i0));
            // the walk's state as the cell (x, y) itself and a remainder f that overflows (f >= 2 da) exactly when the minor
            // coordinate moves: upward f = e, downward f = 2 da - 1 - e (then e - 2 db < 0  <=>  f + 2 db >= 2 da)
            const int j = L.from_end ? m - i0 : i0;  // the sensor is the UPPER end of a from_end line: steps j = m down to 1
            int inc = 0, e = L.da;
            if (steps > 0 && L.da) {
                const unsigned num = 2u * (unsigned)L.db * (unsigned)j + (unsigned)L.da, den = 2u * (unsigned)L.da;
                int q = (int)((double)num * (1.0 / (double)den));  // exact after the one-step correction
                long long rem = (long long)num - (long long)q * den;
                if (rem >= (long long)den) { q++; rem -= den; } else if (rem < 0) { q--; rem += den; }
                inc = q; e = (int)rem;
            }
            const int two_db = 2 * L.db, two_da = 2 * L.da;
            int f = L.from_end ? two_da - 1 - e : e;
            const int dmaj = L.from_end ? -1 : 1, dmin = L.from_end ? -L.bstep : L.bstep;
            const int x = L.steep ? L.b_lo + L.bstep * inc : L.a_lo + j, y = L.steep ? L.a_lo + j : L.b_lo + L.bstep * inc;
            const int sx0 = L.steep ? 0 : dmaj, sy0 = L.steep ? dmaj : 0, sx1 = L.steep ? dmin : 0, sy1 = L.steep ? 0 : dmin;
            // the cell as ONE word, x in the upper and y in the lower half (k_bounds put both ends of every ray inside the map, and
            // the map has fewer than 65536 cells a side -- checked at the top of the kernel): a step is one add, "another patch"
            // one masked compare
            unsigned xy = ((unsigned)x << 16) | ((unsigned)y & 0xffffu);
            const unsigned d0 = (unsigned)(sx0 * 65536 + sy0), d1 = d0 + (unsigned)(sx1 * 65536 + sy1);
            const int k0 = sx0 * kPatch + sy0, k1 = k0 + sx1 * kPatch + sy1;  // the same steps in a tile's cell index (x-major)
            unsigned lxy = ~xy;  // cell of the last patch lookup (none yet)
            int base = -1;       // tile of that patch (slot << 10), and the cell's counter in it while the walk stays in the patch
            int kk = 0;
            const int smax = __reduce_max_sync(0xffffffffu, steps);
            for (int sidx = 0; sidx < smax; sidx++) {
                int key = -1;
                if (sidx < steps) {
                    if ((xy ^ lxy) & 0xffe0ffe0u) {  // another patch
                        lxy = xy; base = -1;
                        const int px = (int)(xy >> (16 + kPatchMag)), py = (int)((xy & 0xffffu) >> kPatchMag);
                        if (px < g.npx && py < g.npy) {
                            const int en = px * g.npy + py;
                            const int sl = (int)M.prefix[en >> 5] + __popc(active[en >> 5] & ((1u << (en & 31)) - 1u)) - first;
                            if ((unsigned)sl < (unsigned)ntile) base = sl << 10;
                        }
                        kk = base | (int)((xy >> (16 - kPatchMag)) & 0x3e0u) | (int)(xy & 31u);
                    }
                    if (base >= 0) key = kk;
                    f += two_db;
                    const bool carry = f >= two_da;
                    f -= carry ? two_da : 0;
                    xy += carry ? d1 : d0;
                    kk += carry ? k1 : k0;  // (meaningless once the walk leaves the patch: the next trip looks the new one up)
                }
                const int prev = __shfl_up_sync(0xffffffffu, key, 1);
                const bool head = lane == 0 || key != prev;
                const unsigned heads = __ballot_sync(0xffffffffu, head);
                if (head && key >= 0) {
                    const unsigned after = (heads >> lane) >> 1;  // (lane 31: nothing after)
                    const unsigned len = after ? (unsigned)__ffs(after) : 32u - (unsigned)lane;
                    atomicAdd(M.tiles + (key >> 1), len << ((key & 1) << 4));
                }
            }
        }
        __syncthreads();
        PHASE(4);
        // flush in units of a quarter tile (8 patch rows) dealt round-robin to the warps: tiles differ a lot in how many
        // of their cells were crossed
        for (int u = warp; u < ntile * 4; u += nw) {
            const int t = u >> 2, g8 = (u & 3) << 8, row0 = (u & 3) << 3;
            const int id = s_tile_id[t];
            if (id < 0) continue;
            int* vis = A.pool.visits + (size_t)id * kPatchCells;
            const int4* cells = A.pool.cells + (size_t)id * kPatchCells;
            unsigned* occ = A.pool.occ + (size_t)id * kPatch + row0;
            const unsigned* tile = M.tiles + t * kTileWords;
            // eight independent loads in flight per lane: the patch is private, plain read-modify-write of the visits
            // plane (a warp covers one row of the patch per k: 128 contiguous bytes).  More visits can only clear an
            // occupancy bit: lane k keeps the bitmap word of row k, and n is fetched for the few crossed cells whose bit is set
            // (a free-space patch has no bitmap, and nothing in it is occupied)
            const unsigned occ_old = (lane < 8 && id < A.pool.cap_h) ? __ldcg(occ + lane) : 0u;
            unsigned occ_new = occ_old;
            int v[8]; unsigned c[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int ci = g8 + 32 * k + lane;
                c[k] = (tile[ci >> 1] >> ((ci & 1) << 4)) & 0xffffu;
                v[k] = c[k] ? __ldcg(vis + ci) : 0;  // untouched cells cost no traffic
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int ci = g8 + 32 * k + lane;
                const int nvis = v[k] + (int)c[k];
                if (c[k]) { vis[ci] = nvis; nfree += c[k]; }  // update(false) c times
                const unsigned roww = __shfl_sync(0xffffffffu, occ_old, k);
                bool clear = false;
                if (c[k] && ((roww >> lane) & 1u)) clear = !occ_default(__ldcg(cells + ci).z, nvis);
                const unsigned cleared = __ballot_sync(0xffffffffu, clear);
                if (lane == k) occ_new &= ~cleared;
            }
            if (occ_new != occ_old) occ[lane] = occ_new;
        }
        __syncthreads();
        PHASE(5);
    }
    if (threadIdx.x == 0) s_nwork = 0;  // (for the promotion list below)
    __threadfence();
    __syncthreads();

    // ---- 4. endpoint hits, per cell in beam order (acc is a float sum: the order is part of the result).  The beams
    // are grouped by cell in a shared-memory hash table (it reuses the tile area): slot = the cell's table entry, and per
    // slot the first and last beam that hit it.  The first beam of a cell then adds up the cell's beams in ascending
    // order by scanning slot_of[] from itself to the last one -- a handful of neighbouring beams in practice.
    // First, the free-space patches that receive their first end point move to hit-capable ids ("promotion"): the first beam
    // that claims the directory entry lists it, then a warp per listed entry takes a hit-capable patch, copies the visit
    // counters, zeroes the hit planes and gives the old patch back.  (The patches are private since phase 2.)
    {
        int* promote = reinterpret_cast<int*>(M.tiles);  // the tile area is free again
        // when the free-space updates took a single pass, s_tile_id still lists the patch of every active entry by rank: the
        // usual answer ("already hit-capable") then costs no global load
        const bool ranked = P.generate_map && nactive <= A.tile_cap;
        for (int b = threadIdx.x; b < (int)P.nbeams; b += blockDim.x) {
            if (!(rec[b].flags & 4)) continue;
            const int e = (rec[b].p1x >> kPatchMag) * g.npy + (rec[b].p1y >> kPatchMag);
            if (ranked && s_tile_id[(int)M.prefix[e >> 5] + __popc(active[e >> 5] & ((1u << (e & 31)) - 1u))] < A.pool.cap_h) continue;
            const int id = dirp[e];
            if (id >= A.pool.cap_h && atomicCAS(dirp + e, id, -2 - id) == id) promote[atomicAdd(&s_nwork, 1)] = e;  // (at most nbeams entries)
        }
        __syncthreads();
        const int npromote = s_nwork;
        for (int k = warp; k < npromote; k += nw) {
            const int e = promote[k];
            const int id = -2 - dirp[e];
            int nid = -1;
            if (lane == 0) nid = pool_alloc(A.pool, A.counters, true);
            nid = __shfl_sync(0xffffffffu, nid, 0);
            if (nid < 0) { if (lane == 0) dirp[e] = id; continue; }  // pool exhausted (flagged): the entry keeps its patch, the hit phase skips it
            const int4* vsrc = reinterpret_cast<const int4*>(A.pool.visits + (size_t)id * kPatchCells);
            int4* vdst = reinterpret_cast<int4*>(A.pool.visits + (size_t)nid * kPatchCells);
            int4* dst = A.pool.cells + (size_t)nid * kPatchCells;
            for (int i = lane; i < kPatchCells / 4; i += 32) vdst[i] = __ldcg(vsrc + i);
            for (int i = lane; i < kPatchCells; i += 32) dst[i] = make_int4(0, 0, 0, 0);
            A.pool.occ[(size_t)nid * kPatch + lane] = 0u;
            __syncwarp();
            if (lane == 0) {
                A.pool.refcnt[nid] = 1;
                dirp[e] = nid;
                A.pool.refcnt[id] = 0;  // private since phase 2: this was its only reference
                A.pool.pending[atomicAdd(A.pool.pending_n, 1)] = id;
                atomicAdd(&s_alloc, 1ull);
            }
        }
        if (npromote) {  // (uniform: the usual scan promotes nothing)
            __threadfence();
            __syncthreads();
        }
    }
    const int hs = A.sort_n << 1;  // table entries: a power of two >= 2 * nbeams
    unsigned long long* hkey = M.keys;
    int* hfirst = reinterpret_cast<int*>(hkey + hs);
    int* hlast = hfirst + hs;
    int* slot_of = hlast + hs;
    for (int i = threadIdx.x; i < hs; i += blockDim.x) { hkey[i] = ~0ull; hfirst[i] = INT_MAX; hlast[i] = -1; }
    __syncthreads();
    for (int b = threadIdx.x; b < (int)P.nbeams; b += blockDim.x) {
        int slot = -1;
        if (rec[b].flags & 4) {
            const unsigned long long key = ((unsigned long long)(unsigned)rec[b].p1x << 32) | (unsigned)rec[b].p1y;
            slot = (int)(mix64(key) & (unsigned long long)(hs - 1));
            for (;;) {
                const unsigned long long prev = atomicCAS(hkey + slot, ~0ull, key);
                if (prev == ~0ull || prev == key) break;
                slot = (slot + 1) & (hs - 1);
            }
            atomicMin(hfirst + slot, b);
            atomicMax(hlast + slot, b);
        }
        slot_of[b] = slot;
    }
    __syncthreads();
    PHASE(6);
    unsigned long long nhit = 0;
    for (int b = threadIdx.x; b < (int)P.nbeams; b += blockDim.x) {
        const int slot = slot_of[b];
        if (slot < 0 || hfirst[slot] != b) continue;  // not the first beam of its cell
        const int x = rec[b].p1x, y = rec[b].p1y;
        int id = dirp[(x >> kPatchMag) * g.npy + (y >> kPatchMag)];
        if ((unsigned)id >= (unsigned)A.pool.cap_h) continue;  // no patch (pool exhausted, flagged)
        int* cell = reinterpret_cast<int*>(A.pool.cells + (size_t)id * kPatchCells + ((x & 31) << kPatchMag) + (y & 31));
        float ax = __int_as_float(__ldcg(cell)), ay = __int_as_float(__ldcg(cell + 1));
        int* visp = A.pool.visits + (size_t)id * kPatchCells + ((x & 31) << kPatchMag) + (y & 31);
        const int n0 = __ldcg(cell + 2), v0 = __ldcg(visp);  // the cell belongs to this thread alone now
        int cnt = 0;
        const int last = hlast[slot];
        for (int q = b; q <= last; q++)
            if (slot_of[q] == slot) { ax += rec[q].hx; ay += rec[q].hy; cnt++; }  // acc += (float)phit, smmap.h:51-52
        cell[0] = __float_as_int(ax); cell[1] = __float_as_int(ay);
        cell[2] = n0 + cnt; *visp = v0 + cnt;
        A.pool.means[(size_t)id * kPatchCells + ((x & 31) << kPatchMag) + (y & 31)] = cell_mean(__float_as_int(ax), __float_as_int(ay), n0 + cnt);
        const bool was = occ_default(n0, v0), is = occ_default(n0 + cnt, v0 + cnt);
        if (was != is) {  // other cells of the row may flip at the same time: atomics on the bitmap word
            unsigned* w = A.pool.occ + (size_t)id * kPatch + (x & 31);
            if (is) atomicOr(w, 1u << (y & 31)); else atomicAnd(w, ~(1u << (y & 31)));
        }
        nhit += cnt;
    }
    // statistics: one shared atomic per warp (64-bit shared atomics are CAS loops; a thousand on one address crawl)
    for (int o = 16; o > 0; o >>= 1) { nfree += __shfl_xor_sync(0xffffffffu, nfree, o); nhit += __shfl_xor_sync(0xffffffffu, nhit, o); }
    if (lane == 0 && nfree) atomicAdd(&s_free, nfree);
    if (lane == 0 && nhit) atomicAdd(&s_hit, nhit);
    __syncthreads();
    PHASE(7);
#undef PHASE
    if (threadIdx.x == 0) {
        atomicAdd(&A.counters->free_updates, s_free); atomicAdd(&A.counters->hit_updates, s_hit);
        atomicAdd(&A.counters->cow_copies, s_cow); atomicAdd(&A.counters->patch_allocs, s_alloc);
    }
}

// ------------------------------------------------------------------------------------------------
// Particle copy by patch reference (gridslamprocessor.hxx:184-213: Particle copy -> HierarchicalArray2D copy-ctor -> autoptr
// share++).  dir_alt[j] <- dir[idx[j]] and one more reference per patch; then the old generation's slots give theirs back
// (k_release_slots).  References are counted incrementally -- never recounted from one context's directories -- because a
// pool may be shared with clones of the filter (gm_clone), whose directories hold references too.
__global__ void k_remap(RemapArgs A) {
    const unsigned j = blockIdx.x, src = A.idx[j];
    const Geom g = A.geom[src];
    const int* s = A.dir + (size_t)src * A.dir_cap;
    int* d = A.dir_alt + (size_t)j * A.dir_cap;
    for (int e = threadIdx.x; e < g.npx * g.npy; e += blockDim.x) {
        int id = s[e];
        d[e] = id;
        if (id >= 0) atomicAdd(&A.refcnt[id], 1);
    }
    if (threadIdx.x == 0) A.geom_alt[j] = g;
}
// the references held by the directories of slots first .. first + gridDim.x - 1 are dropped (autoptr's destructor,
// autoptr.h:42-53): a patch nobody references any more goes back to the free list (through `pending`)
__global__ void k_release_slots(const Geom* __restrict__ geom, const int* __restrict__ dir, int dir_cap, unsigned first, Pool pool) {
    const unsigned slot = first + blockIdx.x;
    const Geom g = geom[slot];
    const int* d = dir + (size_t)slot * dir_cap;
    for (int e = threadIdx.x; e < g.npx * g.npy; e += blockDim.x) {
        const int id = d[e];
        if (id >= 0 && atomicSub(&pool.refcnt[id], 1) == 1) pool.pending[atomicAdd(pool.pending_n, 1)] = id;
    }
}
// one more reference for every patch the directories of slots 0 .. gridDim.x - 1 point at (a clone takes its shares)
__global__ void k_addref_slots(const Geom* __restrict__ geom, const int* __restrict__ dir, int dir_cap, Pool pool) {
    const Geom g = geom[blockIdx.x];
    const int* d = dir + (size_t)blockIdx.x * dir_cap;
    for (int e = threadIdx.x; e < g.npx * g.npy; e += blockDim.x) {
        const int id = d[e];
        if (id >= 0) atomicAdd(&pool.refcnt[id], 1);
    }
}
// released ids go back to the stack of their kind (one CTA)
__global__ void k_merge_free(Pool pool) {
    __shared__ int s_add[2];
    const int n = *pool.pending_n;
    int top_h = pool.free_top[0], top_f = pool.free_top[1];
    if (top_h < 0) top_h = 0;  // exhausted during the stage (error already flagged)
    if (top_f < 0) top_f = 0;  // (ran dry: the hit-capable stack served the rest)
    if (threadIdx.x < 2) s_add[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int id = pool.pending[i];
        if (id < pool.cap_h) pool.free_list[top_h + atomicAdd(&s_add[0], 1)] = id;
        else pool.free_list[pool.cap_h + top_f + atomicAdd(&s_add[1], 1)] = id;
    }
    __syncthreads();
    if (threadIdx.x == 0) { pool.free_top[0] = top_h + s_add[0]; pool.free_top[1] = top_f + s_add[1]; *pool.pending_n = 0; }
}
// MAP_CONSISTENCY_CHECK of the reference (gridslamprocessor.cpp:82-143: every patch's share count against the number of maps that
// point at it): recount the references from the directories of the n live particles and compare with pool.refcnt; out[0] =
// patches whose count differs, out[1] = referenced patches, out[2] = references to a patch id outside the pool
__global__ void k_recount(const Geom* __restrict__ geom, const int* __restrict__ dir, int dir_cap, int cap, int* __restrict__ recount, unsigned long long* __restrict__ out) {
    const Geom g = geom[blockIdx.x];
    const int* d = dir + (size_t)blockIdx.x * dir_cap;
    for (int e = threadIdx.x; e < g.npx * g.npy; e += blockDim.x) {
        const int id = d[e];
        if (id < 0) continue;
        if (id >= cap) { atomicAdd(&out[2], 1ull); continue; }
        atomicAdd(&recount[id], 1);
    }
}
__global__ void k_recount_compare(Pool pool, int* __restrict__ recount, unsigned long long* __restrict__ out) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= pool.cap) return;
    const int have = pool.refcnt[id], want = recount[id];
    recount[id] = 0;
    if (have != want) atomicAdd(&out[0], 1ull);
    if (want > 0) atomicAdd(&out[1], 1ull);
}
void launch_recount(const Geom* geom, const int* dir, int dir_cap, const Pool& pool, int* recount, unsigned long long* out, unsigned n, cudaStream_t st) {
    k_recount<<<n, 256, 0, st>>>(geom, dir, dir_cap, pool.cap, recount, out);
}
void launch_recount_compare(const Pool& pool, int* recount, unsigned long long* out, cudaStream_t st) {
    k_recount_compare<<<(pool.cap + 255) / 256, 256, 0, st>>>(pool, recount, out);
}

__global__ void k_pool_init(Pool pool) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pool.cap) {  // both stacks hand out ascending ids
        pool.free_list[i] = i < pool.cap_h ? pool.cap_h - 1 - i : pool.cap - 1 - (i - pool.cap_h);
        pool.refcnt[i] = 0;
    }
    if (i == 0) { pool.free_top[0] = pool.cap_h; pool.free_top[1] = pool.cap - pool.cap_h; *pool.pending_n = 0; }
}
__global__ void k_fill_int(int* p, size_t n, int v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_fill_geom(Geom* g, unsigned n, Geom v) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) g[i] = v;
}

// ------------------------------------------------------------------------------------------------
// normalize (gridslamprocessor.hxx:82-121).  exp() in parallel; the two sums are sequential in the
// reference, so one thread walks them in order (N <= 64K doubles: tens of microseconds).
// With `rows` (gm_scanmatch_weights) the log-weights first take this scan's log-likelihood, weight += l
// (gridslamprocessor.hxx:50-52; rows[5 i + 4] is particle i's l): the host applies the same single addition to its copy.
// The sequential walks read a block-staged copy in shared memory: thread 0's chain of dependent additions then never waits
// for a global load (behind a registration running on the other stream an L2 round trip per line tripled this kernel).
constexpr unsigned kSeqChunk = 4096;
__global__ void __launch_bounds__(1024) k_normalize(double* __restrict__ logw, const double* __restrict__ rows, unsigned n, double gain_sigma,
                                                    double* __restrict__ w, double* __restrict__ neff_out) {
    __shared__ double red[32];
    __shared__ double s_val;
    __shared__ double stage[kSeqChunk];
    double lmax = -1.7976931348623157e308;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
        double v = logw[i];
        if (rows) { v = v + rows[5 * (size_t)i + 4]; logw[i] = v; }
        lmax = v > lmax ? v : lmax;
    }
    lmax = warp_max(lmax);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lmax;
    __syncthreads();
    if (threadIdx.x == 0) { for (unsigned k = 1; k < (blockDim.x >> 5); k++) lmax = red[k] > lmax ? red[k] : lmax; s_val = lmax; }
    __syncthreads();
    lmax = s_val;
    const double gain = 1. / (gain_sigma * n);
    double chain = 0;  // thread 0's running sum, carried over the chunks
    for (unsigned base = 0; base < n; base += kSeqChunk) {
        const unsigned m = min(kSeqChunk, n - base);
        __syncthreads();
        for (unsigned i = threadIdx.x; i < m; i += blockDim.x) { const double e = exp(gain * (logw[base + i] - lmax)); stage[i] = e; w[base + i] = e; }
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll 8
            for (unsigned i = 0; i < m; i++) chain += stage[i];
        }
    }
    if (threadIdx.x == 0) s_val = chain;
    __syncthreads();
    const double wcum = s_val;
    chain = 0;
    for (unsigned base = 0; base < n; base += kSeqChunk) {
        const unsigned m = min(kSeqChunk, n - base);
        __syncthreads();
        for (unsigned i = threadIdx.x; i < m; i += blockDim.x) { const double x = w[base + i] / wcum; stage[i] = x; w[base + i] = x; }
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll 8
            for (unsigned i = 0; i < m; i++) { const double x = stage[i]; chain += x * x; }
        }
    }
    if (threadIdx.x == 0) *neff_out = 1. / chain;
}

// resampleIndexes (particlefilter.h:180-229).  The cumulative weights c[i] and the targets t[k] are the
// reference's sequentially rounded sums (two serial chains); since both are monotone, the walk
// "while (c[i] > target) emit i" is exactly idx[k] = min{ i : c[i] > t[k] }, found by binary search.
// n_out = the `nparticles` argument (processScan's adaptParticles): the interval is the total weight over n_out and n_out
// indexes come out (particlefilter.h:196-203).
__global__ void __launch_bounds__(1024) k_resample(const double* __restrict__ w, unsigned n, unsigned n_out, double u01, double* __restrict__ cum,
                                                   double* __restrict__ tgt, unsigned* __restrict__ idx, unsigned* __restrict__ emitted) {
    __shared__ double s_interval;
    __shared__ double stage[kSeqChunk];
    double c = 0;
    for (unsigned base = 0; base < n; base += kSeqChunk) {
        const unsigned m = min(kSeqChunk, n - base);
        __syncthreads();
        for (unsigned i = threadIdx.x; i < m; i += blockDim.x) stage[i] = w[base + i];
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll 8
            for (unsigned i = 0; i < m; i++) { c += stage[i]; stage[i] = c; }
        }
        __syncthreads();
        for (unsigned i = threadIdx.x; i < m; i += blockDim.x) cum[base + i] = stage[i];
    }
    if (threadIdx.x == 0) s_interval = c / n_out;
    __syncthreads();
    const double interval = s_interval;
    if (threadIdx.x == 0) {
        double t = interval * u01;
        for (unsigned k = 0; k < n_out; k++) { tgt[k] = t; t += interval; }
    }
    __syncthreads();
    const double total = cum[n - 1];
    unsigned cnt = 0;
    for (unsigned k = threadIdx.x; k < n_out; k += blockDim.x) {
        double t = tgt[k];
        unsigned r = 0;
        if (total > t) {
            unsigned lo = 0, hi = n - 1;  // smallest i with cum[i] > t
            while (lo < hi) { unsigned mid = (lo + hi) >> 1; if (cum[mid] > t) hi = mid; else lo = mid + 1; }
            r = lo; cnt++;
        }
        idx[k] = r;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(emitted, cnt);
}

// ------------------------------------------------------------------------------------------------
// Sum over the visited cells of one map of PointAccumulator::entropy() - log 2 (smmap.h:68-79; a cell nobody has looked at
// has entropy log 2 and contributes nothing).  registerScan's return value -- the sum of the entropy change of every cell
// update, scanmatcher.cpp:317-341 -- telescopes per cell, so it is this quantity after the registration minus before.
// Fixed thread -> cell assignment and a fixed reduction shape: the same map gives the same bits.
__global__ void __launch_bounds__(256) k_entropy(DigestArgs A, double* out) {
    __shared__ double red[8];
    const unsigned p = A.first + blockIdx.x;
    const Geom g = A.geom[p];
    const int* dirp = A.dir + (size_t)p * A.dir_cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double acc = 0.;
    for (int e = warp; e < g.npx * g.npy; e += nw) {
        const int id = dirp[e];
        if (id < 0) continue;
        for (int i = lane; i < kPatchCells; i += 32) {
            const int n = id < A.cap_h ? A.pool[(size_t)id * kPatchCells + i].z : 0, v = A.visits[(size_t)id * kPatchCells + i];
            if (!v) continue;
            double h = 0.;
            if (n != v && n != 0) { const double x = (double)n / (double)v; h = -(x * log(x) + (1. - x) * log(1. - x)); }
            acc += h - 0.6931471805599453;
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.; for (int w = 0; w < nw; w++) t += red[w]; out[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(256) k_digest(DigestArgs A) {
    __shared__ unsigned long long acc[7];
    const unsigned p = A.first + blockIdx.x;
    const Geom g = A.geom[p];
    const int* dirp = A.dir + (size_t)p * A.dir_cap;
    if (threadIdx.x < 7) acc[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long patches = 0, cells = 0, sn = 0, sv = 0, hc = 0, ha = 0, bad = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int e = warp; e < g.npx * g.npy; e += nw) {
        int id = dirp[e];
        if (id < 0) continue;
        if (lane == 0) patches++;
        int px = e / g.npy, py = e % g.npy;
        const int4* pc = A.pool + (size_t)id * kPatchCells;
        const bool hit_capable = id < A.cap_h;  // a free-space patch: visit counters only, n = 0 and acc = 0 everywhere
        for (int i = lane; i < kPatchCells; i += 32) {
            int4 c = hit_capable ? pc[i] : make_int4(0, 0, 0, 0);
            c.w = A.visits[(size_t)id * kPatchCells + i];
            // the warp holds row i >> 5 of the patch: its occupancy bits must equal the bitmap word
            const unsigned bits = __ballot_sync(0xffffffffu, occ_default(c.z, c.w));
            if (lane == 0 && hit_capable && bits != A.occ[(size_t)id * kPatch + (i >> kPatchMag)]) bad++;
            if (c.z > 0) {  // and the stored mean must be the one recomputed from acc and n, bit for bit
                const double2 want = cell_mean(c.x, c.y, c.z), have = A.means[(size_t)id * kPatchCells + i];
                if (__double_as_longlong(want.x) != __double_as_longlong(have.x) || __double_as_longlong(want.y) != __double_as_longlong(have.y)) bad++;
            }
            if (!c.w && !c.z) continue;
            int wx = px * kPatch + (i >> 5) - g.sx2, wy = py * kPatch + (i & 31) - g.sy2;
            cells++; sn += (unsigned)c.z; sv += (unsigned)c.w;
            unsigned long long key = ((unsigned long long)(unsigned)wx << 32) | (unsigned)wy;
            hc += mix64(key ^ mix64(((unsigned long long)(unsigned)c.z << 32) | (unsigned)c.w));
            ha += mix64(key ^ mix64(((unsigned long long)(unsigned)c.x << 32) | (unsigned)c.y) ^ 0x5555ull);
        }
    }
    atomicAdd(&acc[0], patches); atomicAdd(&acc[1], cells); atomicAdd(&acc[2], sn);
    atomicAdd(&acc[3], sv); atomicAdd(&acc[4], hc); atomicAdd(&acc[5], ha);
    if (bad) atomicAdd(&acc[6], bad);
    __syncthreads();
    if (threadIdx.x < 7) A.out[(size_t)blockIdx.x * 7 + threadIdx.x] = acc[threadIdx.x];
}

__global__ void k_export_list(const int* __restrict__ dirp, Geom g, int* __restrict__ list, int* __restrict__ count) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < g.npx * g.npy; e += gridDim.x * blockDim.x) {
        int id = dirp[e];
        if (id >= 0) { int k = atomicAdd(count, 1); list[2 * k] = e; list[2 * k + 1] = id; }
    }
}
__global__ void k_export_gather(const int4* __restrict__ pool, const int* __restrict__ visits, int cap_h, const int* __restrict__ list, int4* __restrict__ out) {
    const int k = blockIdx.x, id = list[2 * k + 1];
    for (int i = threadIdx.x; i < kPatchCells; i += blockDim.x) {  // gm_cell: {acc.x, acc.y, n, visits}
        int4 c = id < cap_h ? pool[(size_t)id * kPatchCells + i] : make_int4(0, 0, 0, 0);
        c.w = visits[(size_t)id * kPatchCells + i];
        out[(size_t)k * kPatchCells + i] = c;
    }
}

// occupancy grid of one particle (the nav_msgs::OccupancyGrid fill of gmapping/src/slam_gmapping.cpp:956-993):
// out[y * width + x] = -1 unknown (never visited), 100 if n / visits > occ_thresh, else 0
__global__ void k_export_occupancy(const int* __restrict__ dirp, Geom g, const int4* __restrict__ pool, const int* __restrict__ visits, int cap_h, double occ_thresh,
                                   signed char* __restrict__ out, int width, int height) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= width || y >= height) return;
    long long ci;
    signed char v = -1;
    if (cell_addr(g, dirp, x, y, ci)) {
        int4 c = ci < ((long long)cap_h << 10) ? __ldg(pool + ci) : make_int4(0, 0, 0, 0);
        c.w = __ldg(visits + ci);
        if (c.w) v = ((double)c.z / (double)c.w > occ_thresh) ? 100 : 0;
    }
    out[(size_t)y * width + x] = v;
}

// drop every patch reference of one particle (before its map is replaced by gm_upload_map)
__global__ void k_release_particle(int* dirp, int nent, Pool pool) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nent; e += gridDim.x * blockDim.x) {
        const int id = dirp[e];
        if (id >= 0 && atomicSub(&pool.refcnt[id], 1) == 1) pool.pending[atomicAdd(pool.pending_n, 1)] = id;
        dirp[e] = -1;
    }
}
// install uploaded patches: block k takes a free patch, copies the staged cells, writes the directory entry
__global__ void k_install_patches(int* dirp, Geom g, const int* __restrict__ coords, const int4* __restrict__ staged, Pool pool, Counters* ctr) {
    __shared__ int s_id;
    const int k = blockIdx.x;
    int any = 0;  // a patch with an end point in it (or a non-zero accumulator) needs the hit planes
    for (int i = threadIdx.x; i < kPatchCells; i += blockDim.x) { const int4 c = staged[(size_t)k * kPatchCells + i]; any |= c.x | c.y | c.z; }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) s_id = pool_alloc(pool, ctr, any != 0);
    __syncthreads();
    const int id = s_id;
    if (id < 0) return;
    const bool hit_capable = id < pool.cap_h;
    for (int i = threadIdx.x; i < kPatchCells; i += blockDim.x) {  // blockDim is a multiple of 32: a warp covers one row per trip
        const int4 c = staged[(size_t)k * kPatchCells + i];
        pool.visits[(size_t)id * kPatchCells + i] = c.w;
        if (!hit_capable) continue;
        pool.cells[(size_t)id * kPatchCells + i] = make_int4(c.x, c.y, c.z, 0);
        pool.means[(size_t)id * kPatchCells + i] = cell_mean(c.x, c.y, c.z);
        const unsigned bits = __ballot_sync(0xffffffffu, occ_default(c.z, c.w));
        if ((i & 31) == 0) pool.occ[(size_t)id * kPatch + (i >> kPatchMag)] = bits;
    }
    if (threadIdx.x == 0) { pool.refcnt[id] = 1; dirp[coords[2 * k] * g.npy + coords[2 * k + 1]] = id; }
}

// ------------------------------------------------------------------------------------------------ launchers
void launch_export_occupancy(const int* dirp, Geom g, const int4* pool, const int* visits, int cap_h, double occ_thresh, signed char* out, int width, int height, cudaStream_t st) {
    dim3 grid((width + 255) / 256, height);
    k_export_occupancy<<<grid, 256, 0, st>>>(dirp, g, pool, visits, cap_h, occ_thresh, out, width, height);
}
void launch_release_particle(int* dirp, int nent, const Pool& pool, cudaStream_t st) { k_release_particle<<<64, 256, 0, st>>>(dirp, nent, pool); }
void launch_install_patches(int* dirp, Geom g, const int* coords, const int4* staged, const Pool& pool, Counters* ctr, int n, cudaStream_t st) {
    k_install_patches<<<n, 256, 0, st>>>(dirp, g, coords, staged, pool, ctr);
}
size_t scanmatch_smem_bytes(unsigned nbeams) { return (size_t)3 * nbeams * sizeof(double2); }
cudaError_t scanmatch_prepare(size_t smem) {
    cudaError_t e = cudaFuncSetAttribute(k_scanmatch<true, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_scanmatch<false, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_scanmatch<true, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_scanmatch<false, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_scanmatch<true, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_scanmatch<false, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return e;
}
void launch_scanmatch(const ScanMatchArgs& a, unsigned n, bool wide, cudaStream_t st) {
    const size_t smem = scanmatch_smem_bytes(a.P.nbeams);
    const bool k1 = a.P.kernel_size == 1 && a.P.thr_is_tenth;
    if (a.query < 0) {
        if (wide) {
            if (k1) k_scanmatch<true, false, 1><<<n, 2 * kSmThreads, smem, st>>>(a);
            else k_scanmatch<false, false, 1><<<n, 2 * kSmThreads, smem, st>>>(a);
        } else {
            if (k1) k_scanmatch<true, false, 0><<<n, kSmThreads, smem, st>>>(a);
            else k_scanmatch<false, false, 0><<<n, kSmThreads, smem, st>>>(a);
        }
    } else {
        if (k1) k_scanmatch<true, true, 0><<<n, kSmThreads, smem, st>>>(a);
        else k_scanmatch<false, true, 0><<<n, kSmThreads, smem, st>>>(a);
    }
}

void launch_bounds(const BoundsArgs& a, unsigned n, unsigned threads, cudaStream_t st) { k_bounds<<<n, threads, 0, st>>>(a); }
void launch_relayout(const RelayoutArgs& a, unsigned n, cudaStream_t st) { k_relayout<<<n, 256, 0, st>>>(a); }
void launch_relayout_commit(const RelayoutArgs& a, unsigned n, cudaStream_t st) { k_relayout_commit<<<n, 256, 0, st>>>(a); }
size_t register_smem_bytes(int bitmap_words, int sort_n, int nbeams, int tile_cap) {
    const size_t tiles = std::max((size_t)tile_cap * kTileWords * 4, (size_t)sort_n * kHitBytes);  // the hit table reuses the tile area
    return align16((size_t)bitmap_words * 4) + align16((size_t)bitmap_words * 2) + align16((size_t)nbeams * sizeof(BeamRec)) +
           align16((size_t)(nbeams + 1) * 4) + tiles;
}
cudaError_t register_prepare(size_t smem) {
    return cudaFuncSetAttribute(k_register, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
void launch_register(const RegisterArgs& a, unsigned n, size_t smem, cudaStream_t st) { k_register<<<n, kRegThreads, smem, st>>>(a); }
void launch_remap(const RemapArgs& a, unsigned n, cudaStream_t st) { k_remap<<<n, 256, 0, st>>>(a); }
void launch_release_slots(const Geom* geom, const int* dir, int dir_cap, unsigned first, unsigned count, const Pool& pool, cudaStream_t st) {
    if (count) k_release_slots<<<count, 256, 0, st>>>(geom, dir, dir_cap, first, pool);
}
void launch_addref_slots(const Geom* geom, const int* dir, int dir_cap, unsigned count, const Pool& pool, cudaStream_t st) {
    if (count) k_addref_slots<<<count, 256, 0, st>>>(geom, dir, dir_cap, pool);
}
void launch_merge_free(const Pool& p, cudaStream_t st) { k_merge_free<<<1, 1024, 0, st>>>(p); }
void launch_pool_init(const Pool& p, cudaStream_t st) { k_pool_init<<<(p.cap + 255) / 256, 256, 0, st>>>(p); }
void launch_fill_int(int* p, size_t n, int v, cudaStream_t st) { k_fill_int<<<1184, 256, 0, st>>>(p, n, v); }
void launch_fill_geom(Geom* g, unsigned n, Geom v, cudaStream_t st) { k_fill_geom<<<(n + 255) / 256, 256, 0, st>>>(g, n, v); }
void launch_normalize(const double* logw, const double* rows, unsigned n, double gain, double* w, double* neff, cudaStream_t st) {
    k_normalize<<<1, 1024, 0, st>>>(const_cast<double*>(logw), rows, n, gain, w, neff);
}
void launch_resample(const double* w, unsigned n, unsigned n_out, double u01, double* cum, double* tgt, unsigned* idx, unsigned* emitted, cudaStream_t st) {
    k_resample<<<1, 1024, 0, st>>>(w, n, n_out, u01, cum, tgt, idx, emitted);
}
void launch_digest(const DigestArgs& a, unsigned n, cudaStream_t st) { k_digest<<<n, 256, 0, st>>>(a); }
void launch_entropy(const DigestArgs& a, double* out, unsigned n, cudaStream_t st) { k_entropy<<<n, 256, 0, st>>>(a, out); }
void launch_export_list(const int* dirp, Geom g, int* list, int* count, cudaStream_t st) { k_export_list<<<64, 256, 0, st>>>(dirp, g, list, count); }
void launch_export_gather(const int4* pool, const int* visits, int cap_h, const int* list, int4* out, int n, cudaStream_t st) { k_export_gather<<<n, 256, 0, st>>>(pool, visits, cap_h, list, out); }

}  // namespace gm
