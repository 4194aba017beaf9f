// gm_device.cuh -- device-side data layout and helpers shared by the kernels (sm_100a).
//
// HBM layout:
//   pool   : int4[pool_cap * 1024]   patch p, cell (lx,ly) at pool[p*1024 + lx*32 + ly] = {acc.x, acc.y, n, unused (0)}
//            (the reference's PointAccumulator, smmap.h:35-36, without its visit counter; x-major like
//            Array2D::m_cells[x][y], array2d.h:39, so a window's yy-inner cells are contiguous)
//   visits : int[pool_cap * 1024]    PointAccumulator::visits of the same cells, in a plane of its own: the free-space
//            update of a scan touches ~40 visit counters for every cell it hits, and eight of them share a 32-byte sector
//            here (two would inside the 16-byte cells)
//   dir    : int[N * dir_cap]        particle i, patch (px,py) at dir[i*dir_cap + px*npy + py], -1 = unallocated
//   geom   : Geom[N]                 per-particle m_sizeX2/m_sizeY2 + directory dims (diverge after Map::resize)
//   means  : double2[pool_cap * 1024] PointAccumulator::mean() = (1./n) * acc of the same cells (smmap.h:24), valid where n > 0:
//            derived data, rewritten whenever a hit changes acc and n (at most one cell per beam and scan) and read by the
//            scan matcher ~10^5 times as often -- one 16-byte load instead of cell -> 1/n lookup -> two conversions and products
//   refcnt : int[pool_cap]           number of directory entries that reference a patch (autoptr::shares)
//   occ    : unsigned[pool_cap * 32] occupancy bitmap: bit ly of word p*32 + lx == (10 * n > visits) of that cell, i.e. the cell
//            is occupied at the default fullnessThreshold 0.1 (smmap.h:25).  Derived data, kept in step by every writer of
//            `pool`; the scan matcher reads a 3x3 window's occupancy from three words instead of nine cells
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gm {

constexpr int kPatchMag = 5;
constexpr int kPatch = 32;
constexpr int kPatchCells = 1024;
constexpr int kLineCap = 20000;  // scanmatcher.cpp:55
#define GM_MAXBEAMS_DEV 2048        // LASER_MAXBEAMS, scanmatcher.h:11 (== GM_MAXBEAMS of the C ABI)
constexpr int kInvTable = 4096;  // exactly rounded 1./n for n < kInvTable (PointAccumulator::mean, smmap.h:24)
constexpr int kMaxRanks = 16;   // GPUs of one node a filter can be sharded over
constexpr int kExpTable = 128;   // 2^(j/128), stored behind the reciprocals (exp_neg in gm_kernels.cu)

struct Geom { int sx2, sy2, npx, npy; };

struct DevParams {
    double cx, cy, delta;
    double laser_pose[3];
    double laser_max_range, usable_range, gaussian_sigma, likelihood_sigma;
    double opt_linear_delta, opt_angular_delta;
    double enlarge_step, fullness_threshold, ang_rel, lin_rel, free_cell_ratio;
    int kernel_size, generate_map;
    unsigned opt_iters, lskip, beam_skip, nbeams;
    int thr_is_tenth;  // fullness_threshold is exactly the default 0.1 -> exact integer occupancy tests
    // derived on the host with the reference's own expressions (IEEE double, no contraction)
    double inv_delta;       // 1. / delta
    double free_off_score;  // map.getDelta() * freeDelta, freeDelta = delta * freeCellRatio (scanmatcher.h:187,208)
    double free_off_lik;    // freeDelta (scanmatcher.h:290,320)
    double cexp;            // -1. / gaussianSigma (scanmatcher.h:257,368)
    double clik;            // -1. / likelihoodSigma (scanmatcher.h:375)
    double no_hit;          // nullLikelihood / likelihoodSigma, nullLikelihood = -.5 (scanmatcher.cpp:15, scanmatcher.h:288)
    // world2map(pfree - phit) minus sizeX2/Y2 when it is the same for every beam ([0] score, [1] likelihoodAndScore; [.][0] x,
    // [.][1] y), INT_MIN in [.][0] when it is not: |pfree - phit| <= free_off, so round((d - c) / delta) is constant as soon
    // as it agrees at d = -free_off and d = +free_off (the map from d to the index is monotone)
    int free_q[2][2];
};

struct Counters {  // zeroed per stage, read back by the host
    unsigned long long score_evals, beam_evals, cow_copies, patch_allocs, free_updates, hit_updates, resizes;
    int error;          // gm_status of the first device-side failure
    int need_dir_cap;   // largest npx*npy any particle needs (k_bounds)
    int any_resize;
    int max_active;     // largest active-patch count of a particle (k_register)
    // k_register: SM cycles summed over CTAs per phase (beam setup, ray marking, ranks, copy-on-write, ray counting,
    // tile flush, hit sort, hit accumulation) -- a profiling aid read through gm_stats
    unsigned long long reg_phase[8];
};

// Two kinds of patches share one id space.  Ids [0, cap_h) are HIT-CAPABLE: they own a slice of every plane.  Ids [cap_h, cap)
// are FREE-SPACE patches: no cell of theirs has ever been an end point (n == 0 everywhere, acc == 0, no mean, no occupied bit),
// so all they own is their slice of the visit counters -- 4 KB instead of 36 KB.  About half of the patches of an indoor map are
// of this kind (profiles/r02_summary.md).  A free-space patch that receives its first hit moves to a hit-capable id
// (k_register, "promotion"); a hit-capable id may also hold a patch without hits (the free-space stack ran dry, or the patch
// was allocated by an upload).  Readers treat id >= cap_h as "n = 0".
struct Pool {
    int4* cells;       // [(cap_h+1)*1024] {acc.x, acc.y, n, 0}; patch `cap_h` is all zero
    int* visits;       // [(cap+1)*1024]
    double2* means;    // [(cap_h+1)*1024] mean() where n > 0
    unsigned* occ;     // [(cap_h+1)*32] occupancy bitmap (see the layout note above); patch `cap_h` is all zero
    int* refcnt;       // [cap]
    int* free_list;    // [cap] two stacks of free patch ids: hit-capable ones in [0, cap_h), free-space ones in [cap_h, cap)
    int* free_top;     // [2] number of ids on each stack (0: hit-capable, 1: free-space)
    int* pending;      // [cap] ids released during a kernel (merged afterwards)
    int* pending_n;
    int cap, cap_h;
};

// take a patch off a free stack (no pushes happen while a kernel runs: released patches go to `pending`).  hit: the patch must be
// able to hold end points; otherwise a free-space patch is preferred and a hit-capable one taken when there is none left
__device__ __forceinline__ int pool_alloc(const Pool& pool, Counters* ctr, bool hit) {
    if (!hit && pool.cap > pool.cap_h) {
        const int k = atomicSub(pool.free_top + 1, 1) - 1;
        if (k >= 0) return pool.free_list[pool.cap_h + k];
    }
    const int k = atomicSub(pool.free_top, 1) - 1;
    if (k < 0) { atomicCAS(&ctr->error, 0, -3); return -1; }
    return pool.free_list[k];
}

// ---- world <-> map (map.h:283-298).  No FMA contraction: this file is compiled with --fmad=false.
__device__ __forceinline__ int w2m(double v, double c, double delta, int s2) {
    return (int)round((v - c) / delta) + s2;
}
__device__ __forceinline__ double m2w(int i, double c, double delta, int s2) { return (i - s2) * delta + c; }

// const Map::cell(): unknown (n=0, visits=0) when outside or unallocated (map.h:347-354)
__device__ __forceinline__ bool cell_addr(const Geom& g, const int* __restrict__ dirp, int x, int y, long long& idx) {
    if (x < 0 || y < 0) return false;
    int px = x >> kPatchMag, py = y >> kPatchMag;
    if (px >= g.npx || py >= g.npy) return false;
    int id = __ldg(dirp + px * g.npy + py);
    if (id < 0) return false;
    idx = (long long)id * kPatchCells + ((x & 31) << kPatchMag) + (y & 31);
    return true;
}

// PointAccumulator::operator double (smmap.h:25) compared against fullnessThreshold.
// For thr == 0.1 the FP64 quotient n/visits can never round across the threshold unless 10n == visits
// (|n/v - 1/10| >= 1/(10 v) >> ulp), so the tests are done in exact integer arithmetic.
static __device__ __noinline__ double occ_ratio(int n, int v) { return (double)n / (double)v; }  // cold: non-default thresholds only
__device__ __forceinline__ bool occ_gt(int n, int v, const DevParams& P) {
    if (!v) return -1.0 > P.fullness_threshold;
    return P.thr_is_tenth ? (10ll * n > (long long)v) : (occ_ratio(n, v) > P.fullness_threshold);
}
__device__ __forceinline__ bool occ_lt(int n, int v, const DevParams& P) {
    if (!v) return -1.0 < P.fullness_threshold;
    return P.thr_is_tenth ? (10ll * n < (long long)v) : (occ_ratio(n, v) < P.fullness_threshold);
}

// PointAccumulator::mean() (smmap.h:24): 1./n * Point(acc.x, acc.y), IEEE double division and products, no contraction
__device__ __forceinline__ double2 cell_mean(int acc_x_bits, int acc_y_bits, int n) {
    if (n == 0) return make_double2(0., 0.);
    const double inv = 1. / (double)n;
    return make_double2((double)__int_as_float(acc_x_bits) * inv, (double)__int_as_float(acc_y_bits) * inv);
}

// the occupancy-bitmap predicate: n / visits > 0.1 decided in integers (implies visits > 0 as n >= 0)
__device__ __forceinline__ bool occ_default(int n, int v) { return 10ll * n > (long long)v; }

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

// ---- closed-form Bresenham (gridlinetraversal.h:20-124).  The reference walks the major axis upward
// from the lower endpoint with d0 = 2db-da, stepping the minor axis when d >= 0; after j steps the
// minor coordinate has moved floor((2*db*j + da) / (2*da)) cells.  The point list is then put into
// start->end order.  Random access lets a warp own one ray with lanes on consecutive cells.
struct Line {
    int a_lo, b_lo, da, db, bstep, n;  // n = number of points
    bool steep, from_end;
};
__device__ __forceinline__ Line make_line(int x0, int y0, int x1, int y1) {
    Line L;
    int dx = abs(x1 - x0), dy = abs(y1 - y0);
    L.steep = !(dy <= dx);
    int a0 = L.steep ? y0 : x0, b0 = L.steep ? x0 : y0, a1 = L.steep ? y1 : x1, b1 = L.steep ? x1 : y1;
    L.da = L.steep ? dy : dx; L.db = L.steep ? dx : dy;
    L.from_end = a0 > a1;
    L.a_lo = L.from_end ? a1 : a0; L.b_lo = L.from_end ? b1 : b0;
    int dirflag = L.from_end ? -1 : 1;
    L.bstep = ((b1 - b0) * dirflag) > 0 ? 1 : -1;
    L.n = L.da + 1;
    return L;
}
// i-th point in start->end order
__device__ __forceinline__ void line_point(const Line& L, int i, int& x, int& y) {
    int j = L.from_end ? (L.n - 1 - i) : i;
    int inc = L.da ? (int)((2u * (unsigned)L.db * (unsigned)j + (unsigned)L.da) / (2u * (unsigned)L.da)) : 0;
    int a = L.a_lo + j, b = L.b_lo + L.bstep * inc;
    x = L.steep ? b : a; y = L.steep ? a : b;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace gm
