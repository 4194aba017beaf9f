// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// Log-driven driver that links the UNMODIFIED reference sources (compiled in place from
// /root/reference by oracle/Makefile, objects only under oracle/_ref/).  It restates the
// reference's offline loop (openslam_gmapping/gui/gsp_thread.cpp:305-337, 393-394, 417-453,
// which itself does not build with a modern g++ and needs Qt3) using only the reference's
// public/protected API, and
//   * `trace` mode: dumps a binary per-stage trace (golden vectors) through the virtual hooks
//     onOdometryUpdate / onScanmatchUpdate / onResampleUpdate (gridslamprocessor.h:214-216),
//   * `time`  mode: times processScan() with steady_clock (the CPU baseline of bench.py).
//
// Nothing here is copied from the reference; it only *calls* it.
#include <chrono>
#include <csetjmp>
#include <csignal>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <set>
#include <vector>

#include <gmapping/gridfastslam/gridslamprocessor.h>
#include <gmapping/log/carmenconfiguration.h>
#include <gmapping/log/sensorlog.h>
#include <gmapping/log/sensorstream.h>
#include <gmapping/scanmatcher/scanmatcherprocessor.h>
#include <gmapping/utils/stat.h>

using namespace GMapping;

namespace {

struct Opts {
    std::string mode = "time", log, out, gfs;
    int particles = 30, max_scans = 1 << 30, maps_every = 0, seed = 12345, probe_particles = 2;
    double xmin = -100, ymin = -100, xmax = 100, ymax = 100, delta = 0.05;
    double sigma = 0.05, maxrange = 80, maxurange = 80, lstep = 0.05, astep = 0.05, lsigma = 0.075, ogain = 3;
    int kernel = 1, iterations = 5, lskip = 0;
    double srr = 0.01, srt = 0.01, str = 0.01, stt = 0.01;
    double linear_update = 1.0, angular_update = 0.5, resample_thr = 0.5, min_score = 0.0, period = -1.0;
    bool generate_map = true, init_from_log = true, quiet = true;
};

struct Writer {
    FILE* f = nullptr;
    void rec(const char tag[4], const void* p, size_t bytes) {
        if (!f) return;
        uint64_t n = bytes;
        fwrite(tag, 1, 4, f);
        fwrite(&n, 8, 1, f);
        if (bytes) fwrite(p, 1, bytes, f);
    }
    template <class T> void vec(const char tag[4], const std::vector<T>& v) { rec(tag, v.data(), v.size() * sizeof(T)); }
};

uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

// order-independent digest of a particle map in WORLD cell coordinates (ix - sizeX2, iy - sizeY2)
struct MapDigest { int64_t patches, cells, sum_n, sum_visits; uint64_t hash_counts, hash_acc; double sum_accx, sum_accy; };

class TraceProcessor : public GridSlamProcessor {
  public:
    TraceProcessor(std::ostream& info) : GridSlamProcessor(info) {}
    Writer w;
    const double* cur_ranges = nullptr;
    bool resampled = false;
    int probe_particles = 2;
    std::vector<double> pre_pose;

    void onOdometryUpdate() override {
        pre_pose.clear();
        for (auto& p : m_particles) { pre_pose.push_back(p.pose.x); pre_pose.push_back(p.pose.y); pre_pose.push_back(p.pose.theta); }
    }
    void onScanmatchUpdate() override {
        if (!w.f) return;
        // per particle: re-run the (const) matcher calls on the pre-update maps to expose score/s/l/c
        std::vector<double> rows;
        for (size_t i = 0; i < m_particles.size(); i++) {
            const Particle& p = m_particles[i];
            OrientedPoint init(pre_pose[3 * i], pre_pose[3 * i + 1], pre_pose[3 * i + 2]), corr;
            double sc = m_matcher.optimize(corr, p.map, init, cur_ranges);
            double s, l; unsigned c = m_matcher.likelihoodAndScore(s, l, p.map, p.pose, cur_ranges);
            double row[12] = {init.x, init.y, init.theta, sc, corr.x, corr.y, corr.theta, s, l, (double)c, p.pose.x, p.pose.y};
            rows.insert(rows.end(), row, row + 12); rows.push_back(p.pose.theta); rows.push_back(p.weight); rows.push_back(p.weightSum);
        }
        w.vec("SMAT", rows);
        // stage-level probes: score / likelihoodAndScore at perturbed poses for the first few particles
        std::vector<double> probes;
        static const double dxs[5][3] = {{0, 0, 0}, {0.05, 0, 0}, {0, -0.05, 0.01}, {-0.025, 0.0125, -0.05}, {0.3, 0.2, 0.1}};
        for (int i = 0; i < probe_particles && i < (int)m_particles.size(); i++)
            for (auto& d : dxs) {
                const Particle& p = m_particles[i];
                OrientedPoint q(p.pose.x + d[0], p.pose.y + d[1], p.pose.theta + d[2]);
                double s2, l2; unsigned c2 = m_matcher.likelihoodAndScore(s2, l2, p.map, q, cur_ranges);
                double s1 = m_matcher.score(p.map, q, cur_ranges);
                double row[8] = {(double)i, q.x, q.y, q.theta, s1, s2, l2, (double)c2};
                probes.insert(probes.end(), row, row + 8);
            }
        w.vec("PROB", probes);
        // sampled variants on particle 0's map: likelihood() over the pose grid and optimize(mean, cov, ...)
        {
            const Particle& p = m_particles[0];
            OrientedPoint q(pre_pose[0], pre_pose[1], pre_pose[2]);
            double lmax; OrientedPoint mean; ScanMatcher::CovarianceMatrix cov;
            const double lsum = m_matcher.likelihood(lmax, mean, cov, p.map, q, cur_ranges);
            OrientedPoint omean; ScanMatcher::CovarianceMatrix ocov;
            const double osc = m_matcher.optimize(omean, ocov, p.map, q, cur_ranges);
            const double row[22] = {lsum, lmax, mean.x, mean.y, mean.theta, cov.xx, cov.yy, cov.tt, cov.xy, cov.xt, cov.yt,
                                    osc, omean.x, omean.y, omean.theta, ocov.xx, ocov.yy, ocov.tt, ocov.xy, ocov.xt, ocov.yt, 0};
            w.rec("LIKE", row, sizeof row);
        }
    }
    void onResampleUpdate() override {
        resampled = true;
        if (!w.f) return;
        double ne = m_neff; w.rec("RNEF", &ne, 8);
        w.vec("RWGT", m_weights);
        std::vector<uint32_t> idx(m_indexes.begin(), m_indexes.end());
        w.vec("RIDX", idx);
    }
    void dumpState() {
        if (!w.f) return;
        std::vector<double> st;
        for (auto& p : m_particles) {
            double row[6] = {p.pose.x, p.pose.y, p.pose.theta, p.weight, p.weightSum, (double)p.previousIndex};
            st.insert(st.end(), row, row + 6);
        }
        w.vec("STAT", st);
        double ne = m_neff; w.rec("NEFF", &ne, 8);
        w.vec("WGTS", m_weights);
        // trajectory tree as updateTreeWeights left it: per particle, walking node -> root through the public TNode fields:
        // chain length, the leaf's accWeight, sum of accWeight, position-weighted sum of accWeight, sum of visitCounter, sum of childs
        std::vector<double> tree;
        for (auto& p : m_particles) {
            double depth = 0, sum = 0, wsum = 0, visits = 0, childs = 0;
            for (const TNode* n = p.node; n; n = n->parent) {
                depth += 1; sum += n->accWeight; wsum += depth * n->accWeight; visits += n->visitCounter; childs += n->childs;
            }
            const double row[6] = {depth, p.node ? p.node->accWeight : 0., sum, wsum, visits, childs};
            tree.insert(tree.end(), row, row + 6);
        }
        w.vec("TREE", tree);
    }
    static MapDigest digest(const ScanMatcherMap& m, std::vector<int32_t>* full) {
        MapDigest d; memset(&d, 0, sizeof d);
        IntPoint c = m.world2map(m.getCenter());  // == (sizeX2, sizeY2): round(0) == 0
        const auto& st = m.storage();
        for (int px = 0; px < st.getXSize(); px++)
            for (int py = 0; py < st.getYSize(); py++) {
                const auto& ptr = st.m_cells[px][py];
                if (!ptr.m_reference) continue;
                d.patches++;
                const Array2D<PointAccumulator>& patch = *ptr;
                for (int lx = 0; lx < 32; lx++)
                    for (int ly = 0; ly < 32; ly++) {
                        const PointAccumulator& a = patch.m_cells[lx][ly];
                        if (!a.visits && !a.n) continue;
                        int32_t wx = px * 32 + lx - c.x, wy = py * 32 + ly - c.y;
                        d.cells++; d.sum_n += a.n; d.sum_visits += a.visits; d.sum_accx += a.acc.x; d.sum_accy += a.acc.y;
                        uint64_t key = ((uint64_t)(uint32_t)wx << 32) | (uint32_t)wy;
                        d.hash_counts += mix64(key ^ mix64(((uint64_t)(uint32_t)a.n << 32) | (uint32_t)a.visits));
                        uint32_t ax, ay; memcpy(&ax, &a.acc.x, 4); memcpy(&ay, &a.acc.y, 4);
                        d.hash_acc += mix64(key ^ mix64(((uint64_t)ax << 32) | ay) ^ 0x5555);
                        if (full) { int32_t row[6] = {wx, wy, a.n, a.visits, (int32_t)ax, (int32_t)ay}; full->insert(full->end(), row, row + 6); }
                    }
            }
        return d;
    }
    void dumpMaps(bool full_first) {
        if (!w.f) return;
        std::vector<MapDigest> ds;
        for (size_t i = 0; i < m_particles.size(); i++) ds.push_back(digest(m_particles[i].map, nullptr));
        w.vec("MDIG", ds);
        std::vector<int32_t> geo;
        for (auto& p : m_particles) {
            IntPoint c = p.map.world2map(p.map.getCenter());
            int32_t row[6] = {c.x, c.y, p.map.getMapSizeX(), p.map.getMapSizeY(), p.map.storage().getXSize(), p.map.storage().getYSize()};
            geo.insert(geo.end(), row, row + 6);
        }
        w.vec("MGEO", geo);
        if (full_first) {
            int bi = getBestParticleIndex();
            std::vector<int32_t> cells; cells.push_back(bi);
            digest(m_particles[bi].map, &cells);
            w.vec("MCEL", cells);
        }
    }
    // What the ROS node does every map_update_interval (gmapping/src/slam_gmapping.cpp:866-993, SlamGMapping::updateMap),
    // restated against the same public API: a standalone ScanMatcher replays the best particle's trajectory (leaf -> root)
    // into a fresh ScanMatcherMap, then every cell is thresholded into the occupancy-grid message.  Record: digest +
    // geometry of that map, the counts of unknown / free / occupied grid cells and a hash of the grid.
    void dumpRosMap(double xmin, double ymin, double xmax, double ymax, double delta, double maxRange, double maxUrange, double occ_thresh) {
        if (!w.f) return;
        ScanMatcher matcher;
        matcher.setLaserParameters(m_matcher.laserBeams(), const_cast<double*>(m_matcher.laserAngles()), m_matcher.getlaserPose());
        matcher.setlaserMaxRange(maxRange);
        matcher.setusableRange(maxUrange);
        matcher.setgenerateMap(true);
        Particle best = getParticles()[getBestParticleIndex()];  // a copy, like slam_gmapping.cpp:879
        Point center;
        center.x = (xmin + xmax) / 2.0;
        center.y = (ymin + ymax) / 2.0;
        ScanMatcherMap smap(center, xmin, ymin, xmax, ymax, delta);
        int nodes = 0;
        for (TNode* n = best.node; n; n = n->parent) {
            if (!n->reading) continue;
            matcher.invalidateActiveArea();
            matcher.computeActiveArea(smap, n->pose, &((*n->reading)[0]));
            matcher.registerScan(smap, n->pose, &((*n->reading)[0]));
            nodes++;
        }
        MapDigest d = digest(smap, nullptr);
        int64_t unknown = 0, freec = 0, occupied = 0;
        uint64_t h = 0;
        for (int x = 0; x < smap.getMapSizeX(); x++)
            for (int y = 0; y < smap.getMapSizeY(); y++) {
                IntPoint p(x, y);
                const double occ = smap.cell(p);
                const int v = occ < 0 ? -1 : (occ > occ_thresh ? 100 : 0);
                if (v < 0) { unknown++; continue; }
                (v ? occupied : freec)++;
                h += mix64((((uint64_t)(uint32_t)x << 32) | (uint32_t)y) ^ mix64((uint64_t)v + 7));
            }
        const Point wmin = smap.map2world(IntPoint(0, 0)), wmax = smap.map2world(IntPoint(smap.getMapSizeX(), smap.getMapSizeY()));
        w.rec("UMAP", &d, sizeof d);
        const double info[12] = {(double)nodes, (double)smap.getMapSizeX(), (double)smap.getMapSizeY(), wmin.x, wmin.y, wmax.x, wmax.y,
                                 (double)unknown, (double)freec, (double)occupied, (double)(h >> 32), (double)(h & 0xffffffffu)};
        w.rec("UINF", info, sizeof info);
    }
    // getTrajectories() (gridslamprocessor_tree.cpp:57-147): a deep copy of the trajectory tree, one leaf per particle, shared
    // ancestors copied once.  Per leaf, walking the COPY: chain length, sums of pose x / y / theta, sum of accWeight, sum of
    // childs, nodes that carry a reading, nodes shared with the previous particle's chain.  Then the copy is freed leaf by leaf.
    void dumpTrajectories() {
        if (!w.f) return;
        TNodeVector leaves = getTrajectories();
        std::vector<double> rows;
        std::set<const TNode*> prev;
        for (size_t i = 0; i < leaves.size(); i++) {
            double depth = 0, sx = 0, sy = 0, st = 0, acc = 0, childs = 0, readings = 0, shared = 0;
            std::set<const TNode*> mine;
            for (const TNode* n = leaves[i]; n; n = n->parent) {
                depth += 1; sx += n->pose.x; sy += n->pose.y; st += n->pose.theta; acc += n->accWeight; childs += n->childs;
                readings += n->reading ? 1 : 0; shared += prev.count(n) ? 1 : 0;
                mine.insert(n);
            }
            bool isOriginal = false;  // the copy must not alias the filter's own nodes
            for (auto& p : m_particles) for (const TNode* n = p.node; n; n = n->parent) if (mine.count(n)) isOriginal = true;
            const double row[9] = {depth, sx, sy, st, acc, childs, readings, shared, isOriginal ? 1. : 0.};
            rows.insert(rows.end(), row, row + 9);
            prev.swap(mine);
        }
        w.vec("TRAJ", rows);
        for (size_t i = 0; i < leaves.size(); i++) delete leaves[i];  // ~TNode releases ancestors when their last child goes
    }
    size_t size() const { return m_particles.size(); }
    int count() const { return m_count; }
};

// integrateScanSequence (gridslamprocessor_tree.cpp:149-224): a second, small filter replays the newest nodes of the first
// filter's best trajectory.  The reference version moves every particle along the path (pose composed relative to the path, the
// node's weight difference added), calls computeActiveArea -- which may resize the particle maps -- and grows one TNode per
// particle and step; it registers nothing.  With assertions enabled the reference then ABORTS while freeing its private copy
// of the path (the copied child counts are never released, assert(!childs) in ~TNode): all of its work is done by then, so
// the abort is caught and the state dumped.  Record per particle: pose, weight, weightSum, trajectory depth, leaf pose, map geometry.
sigjmp_buf g_abort_jmp;
void onAbort(int) { siglongjmp(g_abort_jmp, 1); }

void dumpIntegrate(Writer& w, const SensorMap& sensorMap, const TraceProcessor& src, const Opts& o, std::ostream& info) {
    if (!w.f) return;
    typedef GridSlamProcessor::TNode TNode;
    const GridSlamProcessor::Particle& best = src.getParticles()[src.getBestParticleIndex()];
    std::vector<const TNode*> path;  // newest first, nodes that carry a reading only (the root has none)
    for (const TNode* n = best.node; n && path.size() < 6; n = n->parent)
        if (n->reading) path.push_back(n);
    TNode* leaf = 0;
    for (size_t k = path.size(); k-- > 0;) {
        TNode* c = new TNode(path[k]->pose, 0.25 * (double)(path.size() - k), leaf, 0);
        c->reading = path[k]->reading;
        leaf = c;
    }
    TraceProcessor g2(info);
    g2.setSensorMap(sensorMap);
    g2.setMatchingParameters(o.maxurange, o.maxrange, o.sigma, o.kernel, o.lstep, o.astep, o.iterations, o.lsigma, o.ogain, o.lskip);
    g2.setMotionModelParameters(o.srr, o.srt, o.str, o.stt);
    g2.setUpdateDistances(o.linear_update, o.angular_update, o.resample_thr);
    g2.setgenerateMap(o.generate_map);
    // a map far too small for the scans: computeActiveArea has to resize it
    const OrientedPoint start(path.empty() ? 0. : path.back()->pose.x + 0.3, path.empty() ? 0. : path.back()->pose.y - 0.2, 0.4);
    g2.GridSlamProcessor::init(3, start.x - 1.6, start.y - 1.6, start.x + 1.6, start.y + 1.6, o.delta, start);
    if (leaf) {
        struct sigaction sa, old;
        memset(&sa, 0, sizeof sa);
        sa.sa_handler = onAbort;
        sigemptyset(&sa.sa_mask);
        sigaction(SIGABRT, &sa, &old);
        if (sigsetjmp(g_abort_jmp, 1) == 0) g2.integrateScanSequence(leaf);
        sigaction(SIGABRT, &old, 0);
    }
    std::vector<double> rows;
    for (size_t i = 0; i < g2.getParticles().size(); i++) {
        const GridSlamProcessor::Particle& p = g2.getParticles()[i];
        double depth = 0;
        for (const TNode* n = p.node; n; n = n->parent) depth += 1;
        const IntPoint c = p.map.world2map(p.map.getCenter());
        const double row[16] = {p.pose.x, p.pose.y, p.pose.theta, p.weight, p.weightSum, depth, p.node ? p.node->pose.x : 0., p.node ? p.node->pose.y : 0.,
                                p.node ? p.node->pose.theta : 0., (double)c.x, (double)c.y, (double)p.map.getMapSizeX(), (double)p.map.getMapSizeY(),
                                (double)p.map.storage().getXSize(), (double)p.map.storage().getYSize(), (double)path.size()};
        rows.insert(rows.end(), row, row + 16);
    }
    w.vec("ISEQ", rows);
    delete leaf;  // our own path: ~TNode frees the ancestors
}

// ScanMatcherProcessor (scanmatcher/scanmatcherprocessor.cpp): incremental scan matching of the log's first readings against one
// map, no particle filter.  Record: pose after every reading, then digest + geometry of its map.
void dumpScanMatcherProcessor(Writer& w, const SensorMap& sensorMap, const std::string& log, const Opts& o, int max_readings) {
    if (!w.f) return;
    ScanMatcherProcessor smp(o.xmin, o.ymin, o.xmax, o.ymax, o.delta, o.delta);
    smp.setSensorMap(sensorMap);
    smp.setMatchingParameters(o.maxurange, o.maxrange, o.sigma, o.kernel, o.lstep, o.astep, o.iterations, false);
    smp.setRegistrationParameters(300., 150.);
    smp.matcher().setgenerateMap(o.generate_map);  // the reference's ScanMatcher leaves m_generateMap uninitialised
    smp.init();
    std::ifstream is(log.c_str());
    InputSensorStream input(sensorMap, is);
    std::vector<double> poses, last;
    int n = 0;
    while (input && n < max_readings) {
        const SensorReading* r = nullptr;
        input >> r;
        if (!r) continue;
        const RangeReading* rr = dynamic_cast<const RangeReading*>(r);
        if (rr) {
            smp.processScan(*rr);
            const OrientedPoint p = smp.getPose();
            poses.push_back(p.x); poses.push_back(p.y); poses.push_back(p.theta);
            last.resize(rr->size());
            rr->rawView(last.data(), smp.getMap().getDelta());
            n++;
        }
        delete r;
    }
    w.vec("SMPP", poses);
    // ScanMatcher::icpOptimize / icpStep (scanmatcher.cpp:356-375, scanmatcher.h:88-165) on the map built above, from two poses
    // near the last one (the second with a heading beyond pi: the returned heading is wrapped).  Record: pose + score, twice each
    if (!last.empty()) {
        const OrientedPoint at = smp.getPose();
        const OrientedPoint inits[2] = {OrientedPoint(at.x + 0.03, at.y - 0.02, at.theta + 0.01), OrientedPoint(at.x - 0.05, at.y + 0.04, at.theta + 2 * M_PI + 0.02)};
        std::vector<double> icp;
        for (int k = 0; k < 2; k++) {
            OrientedPoint a(0, 0, 0), b(0, 0, 0);
            const double so = smp.matcher().icpOptimize(a, smp.getMap(), inits[k], last.data());
            const double ss = smp.matcher().icpStep(b, smp.getMap(), inits[k], last.data());
            const double row[8] = {a.x, a.y, a.theta, so, b.x, b.y, b.theta, ss};
            icp.insert(icp.end(), row, row + 8);
        }
        w.vec("ICPO", icp);
        // ScanMatcher::registerScan's return value (scanmatcher.cpp:317-341: the summed entropy change of the cell updates) for
        // two registrations of that scan into a copy of the map, the second from a slightly different pose
        ScanMatcherMap copy(smp.getMap());
        std::vector<double> gains;
        for (int k = 0; k < 2; k++) {
            smp.matcher().invalidateActiveArea();
            gains.push_back(smp.matcher().registerScan(copy, OrientedPoint(at.x + 0.11 * k, at.y - 0.07 * k, at.theta + 0.05 * k), last.data()));
        }
        w.vec("RSEN", gains);
    }
    const ScanMatcherMap& m = smp.getMap();
    MapDigest d = TraceProcessor::digest(m, nullptr);
    w.rec("SMPD", &d, sizeof d);
    IntPoint c = m.world2map(m.getCenter());
    const int32_t geo[6] = {c.x, c.y, m.getMapSizeX(), m.getMapSizeY(), m.storage().getXSize(), m.storage().getYSize()};
    w.rec("SMPG", geo, sizeof geo);
}

bool arg(int& i, int argc, char** argv, const char* name, std::string& v) {
    if (strcmp(argv[i], name)) return false;
    if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", name); exit(2); }
    v = argv[++i]; return true;
}

}  // namespace

int main(int argc, char** argv) {
    Opts o;
    for (int i = 1; i < argc; i++) {
        std::string v;
#define S(flag, field) if (arg(i, argc, argv, flag, v)) { o.field = v; continue; }
#define I(flag, field) if (arg(i, argc, argv, flag, v)) { o.field = atoi(v.c_str()); continue; }
#define D(flag, field) if (arg(i, argc, argv, flag, v)) { o.field = atof(v.c_str()); continue; }
        S("-mode", mode) S("-log", log) S("-out", out) S("-gfs", gfs)
        I("-particles", particles) I("-scans", max_scans) I("-maps-every", maps_every) I("-seed", seed) I("-probe", probe_particles)
        D("-xmin", xmin) D("-ymin", ymin) D("-xmax", xmax) D("-ymax", ymax) D("-delta", delta)
        D("-sigma", sigma) D("-maxrange", maxrange) D("-maxUrange", maxurange) D("-lstep", lstep) D("-astep", astep)
        D("-lsigma", lsigma) D("-ogain", ogain) I("-kernelSize", kernel) I("-iterations", iterations) I("-lskip", lskip)
        D("-srr", srr) D("-srt", srt) D("-str", str) D("-stt", stt)
        D("-linearUpdate", linear_update) D("-angularUpdate", angular_update) D("-resampleThreshold", resample_thr)
        D("-minimumScore", min_score) D("-period", period)
        if (!strcmp(argv[i], "-no-generateMap")) { o.generate_map = false; continue; }
        if (!strcmp(argv[i], "-init-zero")) { o.init_from_log = false; continue; }
        if (!strcmp(argv[i], "-verbose")) { o.quiet = false; continue; }
        fprintf(stderr, "unknown flag %s\n", argv[i]); return 2;
    }
    if (o.log.empty()) { fprintf(stderr, "usage: ref_driver -mode trace|time -log file.log [-out trace.bin] ...\n"); return 2; }

    // silence the reference's unconditional cerr/cout chatter (gridslamprocessor.hxx:224-297)
    std::ofstream devnull("/dev/null");
    std::streambuf* old_cerr = std::cerr.rdbuf();
    std::streambuf* old_cout = std::cout.rdbuf();
    if (o.quiet) { std::cerr.rdbuf(devnull.rdbuf()); std::cout.rdbuf(devnull.rdbuf()); }

    std::ifstream cfg_is(o.log.c_str());
    if (!cfg_is) { fprintf(stderr, "cannot open %s\n", o.log.c_str()); return 1; }
    CarmenConfiguration conf;
    conf.load(cfg_is);
    cfg_is.close();
    SensorMap sensorMap = conf.computeSensorMap();

    std::ifstream is(o.log.c_str());
    InputSensorStream input(sensorMap, is);

    std::ostream* info = o.quiet ? static_cast<std::ostream*>(&devnull) : &std::cout;
    TraceProcessor gsp(*info);
    gsp.probe_particles = o.probe_particles;
    gsp.setSensorMap(sensorMap);
    gsp.setMatchingParameters(o.maxurange, o.maxrange, o.sigma, o.kernel, o.lstep, o.astep, o.iterations, o.lsigma, o.ogain, o.lskip);
    gsp.setMotionModelParameters(o.srr, o.srt, o.str, o.stt);
    gsp.setUpdateDistances(o.linear_update, o.angular_update, o.resample_thr);
    gsp.setUpdatePeriod(o.period);
    gsp.setgenerateMap(o.generate_map);
    gsp.setminimumScore(o.min_score);

    // the .gfs trace the reference's log drivers write (gsp_thread.cpp:330-337 opens it the same way, before init)
    if (!o.gfs.empty()) gsp.outputStream().open(o.gfs.c_str());

    bool inited = false;
    if (o.mode == "trace" && !o.out.empty()) {
        gsp.w.f = fopen(o.out.c_str(), "wb");
        if (!gsp.w.f) { fprintf(stderr, "cannot write %s\n", o.out.c_str()); return 1; }
    }

    std::vector<double> ms;
    int processed_n = 0, read_n = 0;
    while (input && processed_n < o.max_scans) {
        const SensorReading* r = nullptr;
        input >> r;
        if (!r) continue;
        const RangeReading* rr = dynamic_cast<const RangeReading*>(r);
        if (!rr) { delete r; continue; }
        if (!inited) {
            OrientedPoint init = o.init_from_log ? rr->getPose() : OrientedPoint(0, 0, 0);
            gsp.GridSlamProcessor::init(o.particles, o.xmin, o.ymin, o.xmax, o.ymax, o.delta, init);
            sampleGaussian(1, o.seed);  // srand48(seed), as gsp_thread.cpp:393-394 / slam_gmapping.cpp:676
            if (gsp.w.f) {
                double hdr[8] = {(double)o.particles, (double)rr->size(), init.x, init.y, init.theta, o.delta, (double)o.seed, 0};
                gsp.w.rec("HEAD", hdr, sizeof hdr);
                std::vector<double> ang(gsp.m_matcher.laserAngles(), gsp.m_matcher.laserAngles() + gsp.m_matcher.laserBeams());
                gsp.w.vec("ANGL", ang);
            }
            inited = true;
        }
        std::vector<double> plain(rr->begin(), rr->end());
        gsp.cur_ranges = plain.data();
        gsp.resampled = false;
        if (gsp.w.f) {
            double rd[4] = {rr->getPose().x, rr->getPose().y, rr->getPose().theta, rr->getTime()};
            gsp.w.rec("READ", rd, sizeof rd);
            gsp.w.vec("RNGS", plain);
        }
        auto t0 = std::chrono::steady_clock::now();
        bool processed = gsp.processScan(*rr);
        auto t1 = std::chrono::steady_clock::now();
        read_n++;
        if (gsp.w.f) {
            gsp.w.vec("ODOM", gsp.pre_pose);
            int32_t flags[2] = {processed ? 1 : 0, gsp.resampled ? 1 : 0};
            gsp.w.rec("PROC", flags, sizeof flags);
        }
        if (processed) {
            ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
            processed_n++;
            gsp.dumpState();
            if (o.maps_every > 0 && (processed_n % o.maps_every == 0 || processed_n == 1)) gsp.dumpMaps(false);
        }
        // the filter keeps a pointer to its own copy of the reading (gridslamprocessor.cpp:556-561)
        delete r;
    }
    if (gsp.w.f) {
        gsp.dumpMaps(true);
        gsp.dumpTrajectories();
        gsp.dumpRosMap(o.xmin, o.ymin, o.xmax, o.ymax, o.delta, o.maxrange, o.maxurange, 0.25);
        dumpScanMatcherProcessor(gsp.w, sensorMap, o.log, o, 12);
        {   // SensorLog (log/sensorlog.cpp): the whole log in memory and the bounding box gfs_nogui's -autosize uses
            SensorLog slog(sensorMap);
            std::ifstream ls(o.log.c_str());
            slog.load(ls);
            double bx0, by0, bx1, by1;
            const OrientedPoint start = slog.boundingBox(bx0, by0, bx1, by1);
            const double row[8] = {(double)slog.size(), bx0, by0, bx1, by1, start.x, start.y, start.theta};
            gsp.w.rec("BBOX", row, sizeof row);
        }
        dumpIntegrate(gsp.w, sensorMap, gsp, o, *info);
        gsp.w.rec("DONE", nullptr, 0);
        fclose(gsp.w.f);
    }
    if (gsp.outputStream().is_open()) gsp.outputStream().close();
    std::cerr.rdbuf(old_cerr);
    std::cout.rdbuf(old_cout);
    double sum = 0, sum_sm = 0;
    for (size_t i = 0; i < ms.size(); i++) { sum += ms[i]; if (i > 0) sum_sm += ms[i]; }
    // JSON on stdout: mean over processed scans; `ms_matched` excludes the first (registration-only) scan
    printf("{\"impl\": \"reference\", \"particles\": %d, \"read\": %d, \"processed\": %d, \"ms_total\": %.3f, "
           "\"ms_per_scan\": %.4f, \"ms_per_matched_scan\": %.4f}\n",
           o.particles, read_n, processed_n, sum, ms.empty() ? 0.0 : sum / ms.size(),
           ms.size() > 1 ? sum_sm / (ms.size() - 1) : 0.0);
    return 0;
}
