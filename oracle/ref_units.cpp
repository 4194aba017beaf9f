// oracle/ref_units.cpp -- TEST INFRASTRUCTURE.  Known-answer vectors for the small building blocks of the hot path, produced by
// calling the UNMODIFIED reference directly (compiled in place from /root/reference by `make -C oracle ref`):
//   GridLineTraversal::gridLine          include/gmapping/scanmatcher/gridlinetraversal.h:110-124
//   uniform_resampler::resampleIndexes   include/gmapping/particlefilter/particlefilter.h:180-229
//   sampleGaussian / pf_ran_gaussian     utils/stat.cpp:31-65
//   MotionModel::drawFromMotion          gridfastslam/motionmodel.cpp:40-60
//   Map::world2map / map2world           include/gmapping/grid/map.h:283-298
//   point / orientedpoint algebra        include/gmapping/utils/point.h:22-220;  evalLogGaussian  utils/stat.cpp:73-77
// Output: one binary file of tagged records (the format of ref_driver's traces), turned into tests/golden/units.npz by
// tools/make_golden.py units.  The inputs come from a fixed 64-bit LCG so that the test can regenerate them.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <gmapping/gridfastslam/motionmodel.h>
#include <gmapping/particlefilter/particlefilter.h>
#include <gmapping/scanmatcher/gridlinetraversal.h>
#include <gmapping/scanmatcher/smmap.h>
#include <gmapping/utils/stat.h>

using namespace GMapping;

static FILE* g_f;
static void rec(const char tag[4], const void* p, size_t bytes) {
    uint64_t n = bytes;
    fwrite(tag, 1, 4, g_f); fwrite(&n, 8, 1, g_f);
    if (bytes) fwrite(p, 1, bytes, g_f);
}
static uint64_t g_state = 0x243F6A8885A308D3ull;
static uint64_t lcg() { g_state = g_state * 6364136223846793005ull + 1442695040888963407ull; return g_state >> 11; }
static double uni() { return (double)lcg() / 9007199254740992.0; }  // [0, 1)
static uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: ref_units out.bin\n"); return 2; }
    g_f = fopen(argv[1], "wb");
    if (!g_f) return 1;

    // ---- grid lines: all octants, degenerate and long ones.  Per line: x0 y0 x1 y1, point count, order-sensitive hash of the points
    {
        std::vector<int64_t> rows;
        std::vector<IntPoint> buf(20000);
        GridLineTraversalLine line;
        line.points = buf.data();
        for (int k = 0; k < 4000; k++) {
            int x0 = (int)(lcg() % 4000), y0 = (int)(lcg() % 4000), x1, y1;
            const int kind = k % 8;
            if (kind == 0) { x1 = x0; y1 = y0; }                                   // a point
            else if (kind == 1) { x1 = x0 + (int)(lcg() % 600) - 300; y1 = y0; }   // horizontal
            else if (kind == 2) { x1 = x0; y1 = y0 + (int)(lcg() % 600) - 300; }   // vertical
            else if (kind == 3) { int d = (int)(lcg() % 600) - 300; x1 = x0 + d; y1 = y0 + ((lcg() & 1) ? d : -d); }  // diagonals
            else { x1 = x0 + (int)(lcg() % 1200) - 600; y1 = y0 + (int)(lcg() % 1200) - 600; }
            line.num_points = 0;
            GridLineTraversal::gridLine(IntPoint(x0, y0), IntPoint(x1, y1), &line);
            uint64_t h = 0;
            for (int i = 0; i < line.num_points; i++)
                h = mix64(h ^ (((uint64_t)(uint32_t)line.points[i].x << 32) | (uint32_t)line.points[i].y));
            const int64_t row[6] = {x0, y0, x1, y1, line.num_points, (int64_t)h};
            rows.insert(rows.end(), row, row + 6);
        }
        rec("LINE", rows.data(), rows.size() * sizeof(int64_t));
    }

    // ---- systematic resampling.  Per case: n, seed of drand48, then the weights and the indexes
    {
        uniform_resampler<double, double> resampler;
        const int sizes[10] = {1, 2, 3, 7, 30, 31, 64, 257, 1000, 4096};
        for (int c = 0; c < 40; c++) {
            const int n = sizes[c % 10];
            std::vector<double> w(n);
            double sum = 0;
            const int shape = c / 10;  // 0: uniform-ish, 1: peaked, 2: one dominant, 3: with exact zeros
            for (int i = 0; i < n; i++) {
                double v = uni();
                if (shape == 1) v = v * v * v * v;
                if (shape == 2) v = (i == n / 2) ? 1e6 : v * 1e-3;
                if (shape == 3 && (i % 3) == 0) v = 0;
                w[i] = v; sum += v;
            }
            if (sum > 0) for (int i = 0; i < n; i++) w[i] /= sum;
            else w[0] = 1;
            const long seed = 1000 + c;
            srand48(seed);
            std::vector<unsigned int> idx = resampler.resampleIndexes(w, 0);
            const double hdr[2] = {(double)n, (double)seed};
            rec("RSHD", hdr, sizeof hdr);
            rec("RSWT", w.data(), w.size() * sizeof(double));
            rec("RSIX", idx.data(), idx.size() * sizeof(unsigned int));
        }
    }

    // ---- the same with the `nparticles` argument (processScan's adaptParticles): fewer and more indexes than weights
    {
        uniform_resampler<double, double> resampler;
        const int sizes[6][2] = {{12, 8}, {12, 30}, {30, 7}, {257, 64}, {1000, 4096}, {4096, 512}};
        for (int c = 0; c < 18; c++) {
            const int n = sizes[c % 6][0], n_out = sizes[c % 6][1];
            std::vector<double> w(n);
            double sum = 0;
            const int shape = c / 6;  // 0: uniform-ish, 1: peaked, 2: with exact zeros
            for (int i = 0; i < n; i++) {
                double v = uni();
                if (shape == 1) v = v * v * v * v;
                if (shape == 2 && (i % 3) == 0) v = 0;
                w[i] = v; sum += v;
            }
            for (int i = 0; i < n; i++) w[i] /= sum;
            const long seed = 2000 + c;
            srand48(seed);
            std::vector<unsigned int> idx = resampler.resampleIndexes(w, n_out);
            const double hdr[3] = {(double)n, (double)n_out, (double)seed};
            rec("RAHD", hdr, sizeof hdr);
            rec("RAWT", w.data(), w.size() * sizeof(double));
            rec("RAIX", idx.data(), idx.size() * sizeof(unsigned int));
        }
    }

    // ---- Gaussian sampler: sampleGaussian(sigma, seed) seeds and draws; then a run of draws incl. sigma == 0 (no draw consumed)
    {
        std::vector<double> out;
        out.push_back(sampleGaussian(1.0, 4242));
        const double sigmas[6] = {1.0, 0.05, 0.0, 3.5, 1e-9, 0.0};
        for (int i = 0; i < 600; i++) out.push_back(sampleGaussian(sigmas[i % 6]));
        out.push_back(drand48());  // the stream position after all of it
        rec("GAUS", out.data(), out.size() * sizeof(double));
    }

    // ---- motion model: drawFromMotion(p, pnew, pold) for random poses; rows of 12: p, pnew, pold, result
    {
        MotionModel mm;
        mm.srr = 0.1; mm.srt = 0.2; mm.str = 0.1; mm.stt = 0.2;
        srand48(77);
        std::vector<double> rows;
        for (int k = 0; k < 300; k++) {
            OrientedPoint p(uni() * 20 - 10, uni() * 20 - 10, uni() * 6.2 - 3.1), pold(uni() * 20 - 10, uni() * 20 - 10, uni() * 6.2 - 3.1);
            OrientedPoint pnew(pold.x + (uni() - .5) * (k % 5 == 0 ? 0. : 1.), pold.y + (uni() - .5) * (k % 5 == 0 ? 0. : 1.), pold.theta + (uni() - .5) * (k % 7 == 0 ? 0. : .6));
            const OrientedPoint r = mm.drawFromMotion(p, pnew, pold);
            const double row[12] = {p.x, p.y, p.theta, pnew.x, pnew.y, pnew.theta, pold.x, pold.y, pold.theta, r.x, r.y, r.theta};
            rows.insert(rows.end(), row, row + 12);
        }
        rows.push_back(drand48());
        rec("MOTN", rows.data(), rows.size() * sizeof(double));
    }

    // ---- world2map / map2world incl. exact half-cell positions and negative coordinates, on a centred and an off-centre map
    {
        std::vector<double> rows;
        for (int m = 0; m < 2; m++) {
            const double cx = m ? 20.025 : 0., cy = m ? -7.5 : 0., delta = m ? 0.025 : 0.05;
            ScanMatcherMap map(Point(cx, cy), 100., 80., delta);
            for (int k = 0; k < 400; k++) {
                double x = uni() * 120 - 60, y = uni() * 100 - 50;
                if (k % 4 == 0) { x = cx + ((int)(lcg() % 400) - 200 + 0.5) * delta; y = cy + ((int)(lcg() % 400) - 200 - 0.5) * delta; }  // rounding boundaries
                const IntPoint ip = map.world2map(Point(x, y));
                const Point back = map.map2world(ip);
                const double row[8] = {(double)m, x, y, (double)ip.x, (double)ip.y, back.x, back.y, map.isInside(ip) ? 1. : 0.};
                rows.insert(rows.end(), row, row + 8);
            }
        }
        rec("W2MP", rows.data(), rows.size() * sizeof(double));
    }
    // ---- point algebra (utils/point.h) and evalLogGaussian (utils/stat.cpp): rows of 23 doubles per random pair of poses
    {
        std::vector<double> rows;
        for (int k = 0; k < 500; k++) {
            const OrientedPoint a(uni() * 40 - 20, uni() * 40 - 20, uni() * 12 - 6), b(uni() * 40 - 20, uni() * 40 - 20, uni() * 12 - 6);
            const Point q(uni() * 4 - 2, uni() * 4 - 2);
            const double v = uni() * 3 - 1.5;
            const OrientedPoint d = absoluteDifference(a, b), s = absoluteSum(a, b), plus = a + b, minus = a - b, scaled = a * v;
            const Point moved = absoluteSum(a, q), pq = Point(a.x, a.y) - q;
            const double row[23] = {d.x, d.y, d.theta, s.x, s.y, s.theta, plus.x, plus.y, plus.theta, minus.x, minus.y, minus.theta,
                                    scaled.x, scaled.y, scaled.theta, moved.x, moved.y, pq * q, euclidianDist(a, b), euclidianDist(a, q),
                                    euclidianDist(q, a), evalLogGaussian(k % 9 == 0 ? -1. : uni() * 2, uni() * 3 - 1.5), (double)k};
            rows.insert(rows.end(), row, row + 23);
        }
        rec("PTOP", rows.data(), rows.size() * sizeof(double));
    }
    // ---- ScanMatcherMap / HierarchicalArray2D / PointAccumulator on the host (grid/map.h, grid/harray2d.h, scanmatcher/smmap.h):
    // on-demand patch allocation through the non-const cell(), copies that SHARE patches until allocActiveArea() gives the
    // active ones a private copy, resize (crop + extend), toDoubleMap.  One summary row per stage.
    {
        std::vector<double> rows;
        struct Sum {
            static void of(const ScanMatcherMap& m, std::vector<double>& out) {
                double patches = 0, cells = 0, n = 0, visits = 0, occ = 0, ent = 0, mx = 0, my = 0;
                uint64_t h = 0;
                const HierarchicalArray2D<PointAccumulator>& st = m.storage();
                for (int px = 0; px < st.getXSize(); px++)
                    for (int py = 0; py < st.getYSize(); py++) {
                        if (!st.m_cells[px][py].m_reference) continue;
                        patches++;
                        const Array2D<PointAccumulator>& patch = *st.m_cells[px][py];
                        for (int lx = 0; lx < 32; lx++)
                            for (int ly = 0; ly < 32; ly++) {
                                const PointAccumulator& a = patch.m_cells[lx][ly];
                                if (!a.visits && !a.n) continue;
                                cells++; n += a.n; visits += a.visits; occ += (double)a; ent += a.entropy();
                                if (a.n) { const Point mu = a.mean(); mx += mu.x; my += mu.y; }
                                uint32_t ax, ay; memcpy(&ax, &a.acc.x, 4); memcpy(&ay, &a.acc.y, 4);
                                h += mix64(((uint64_t)(uint32_t)(px * 32 + lx) << 32 | (uint32_t)(py * 32 + ly)) ^ mix64(((uint64_t)ax << 32) | ay) ^ ((uint64_t)a.n << 20) ^ (uint64_t)a.visits);
                            }
                    }
                const IntPoint c = m.world2map(m.getCenter());
                const double row[16] = {patches, cells, n, visits, occ, ent, mx, my, (double)(h >> 32), (double)(h & 0xffffffffu), (double)c.x, (double)c.y,
                                        (double)m.getMapSizeX(), (double)m.getMapSizeY(), (double)st.getXSize(), (double)st.getYSize()};
                out.insert(out.end(), row, row + 16);
            }
        };
        ScanMatcherMap map(Point(1.0, -2.0), 20., 16., 0.05);
        for (int k = 0; k < 3000; k++) {
            const Point p(1.0 + (uni() - .5) * 14, -2.0 + (uni() - .5) * 10);
            const IntPoint ip = map.world2map(p);
            if (map.isInside(ip)) map.cell(ip).update(k % 3 != 0, p);
        }
        Sum::of(map, rows);
        ScanMatcherMap copy(map);  // shares every patch with `map`
        for (int k = 0; k < 500; k++) {
            const Point p(1.0 + (uni() - .5) * 6, -2.0 + (uni() - .5) * 6);
            const IntPoint ip = copy.world2map(p);
            if (copy.isInside(ip)) copy.cell(ip).update(true, p);  // lands in the shared patch (or in a new one owned by the copy)
        }
        Sum::of(map, rows); Sum::of(copy, rows);
        HierarchicalArray2D<PointAccumulator>::PointSet active;
        for (int k = 0; k < 40; k++) active.insert(copy.storage().patchIndexes(copy.world2map(Point(1.0 + (uni() - .5) * 8, -2.0 + (uni() - .5) * 8))));
        copy.storage().setActiveArea(active, true);
        copy.storage().allocActiveArea();  // private copies of the active patches
        for (int k = 0; k < 500; k++) {
            const Point p(1.0 + (uni() - .5) * 8, -2.0 + (uni() - .5) * 8);
            const IntPoint ip = copy.world2map(p);
            if (copy.isInside(ip)) copy.cell(ip).update(k % 2 == 0, p);
        }
        Sum::of(map, rows); Sum::of(copy, rows);
        copy.resize(-6.3, -9.1, 12.7, 3.2);  // crops on one side, extends on the other
        Sum::of(copy, rows);
        {
            Map<double, DoubleArray2D, false>* dm = copy.toDoubleMap();
            double s = 0, unknown = 0;
            for (int x = 0; x < dm->getMapSizeX(); x++)
                for (int y = 0; y < dm->getMapSizeY(); y++) { const double v = dm->cell(IntPoint(x, y)); if (v < 0) unknown++; else s += v * (1 + (x % 7) + 3 * (y % 5)); }
            const double row[16] = {s, unknown, (double)dm->getMapSizeX(), (double)dm->getMapSizeY()};
            rows.insert(rows.end(), row, row + 16);
            delete dm;
        }
        rec("HMAP", rows.data(), rows.size() * sizeof(double));
    }
    fclose(g_f);
    return 0;
}
