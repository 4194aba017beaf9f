// capi_shim.cpp -- a flat C view of GMapping::GridSlamProcessor for bench.py and the tests (ctypes): every call
// goes through the same C++ API a C++ consumer would use.
#include <set>
#include <gmapping/gridfastslam/gridslamprocessor.h>
#include <gmapping/utils/stat.h>

#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>

#include "gm_b200.h"

using namespace GMapping;

namespace {
struct Counting : public GridSlamProcessor {
    explicit Counting(std::ostream& os) : GridSlamProcessor(os), resamples(0), lastResampled(false) {}
    Counting(const Counting& o) : GridSlamProcessor(o), resamples(o.resamples), lastResampled(false) {}
    void onResampleUpdate() override { resamples++; lastResampled = true; }
    int resamples;
    bool lastResampled;
};
struct Handle {
    std::ofstream devnull;
    Counting* gsp;
    RangeSensor* laser;
    SensorMap smap;
    Handle(bool verbose) : devnull("/dev/null"), gsp(new Counting(verbose ? std::cout : devnull)), laser(0) {}
    // clone: the filter is deep-copied (GridSlamProcessor copy constructor); the sensor is re-created
    Handle(const Handle& o) : devnull("/dev/null"), gsp(new Counting(*o.gsp)), laser(new RangeSensor(*o.laser)) { smap.insert(std::make_pair(laser->getName(), laser)); }
    ~Handle() { delete gsp; delete laser; }
};
}  // namespace

extern "C" {

// (the counting subclass reads no map in its hook: it takes the default of a plain GridSlamProcessor, see getregisterWithScanMatch)
void* gmh_create(int verbose) { Handle* h = new Handle(verbose != 0); h->gsp->setregisterWithScanMatch(true); return h; }
void gmh_destroy(void* h) { delete static_cast<Handle*>(h); }
void* gmh_clone(void* h) { return new Handle(*static_cast<Handle*>(h)); }
int gmh_dist_unique_id(void* id128) { return gm_dist_unique_id(id128); }
void gmh_set_distributed(void* h, int rank, int world, const void* id) { static_cast<Handle*>(h)->gsp->setDistributed(rank, world, id); }
void gmh_set_lazy_tree(void* h, int on) { static_cast<Handle*>(h)->gsp->setlazyTreeWeights(on != 0); }
void gmh_refresh_tree(void* h) { static_cast<Handle*>(h)->gsp->refreshTreeWeights(); }
// the trajectory tree seen from each particle, walking node -> root: chain length, the leaf's accWeight, sum of accWeight,
// position-weighted sum of accWeight, sum of visitCounter, sum of childs (6 doubles per particle; the rows the tests' golden
// traces hold for the reference)
void gmh_tree_rows(void* hv, double* rows) {
    const GridSlamProcessor::ParticleVector& ps = static_cast<Handle*>(hv)->gsp->getParticles();
    for (size_t i = 0; i < ps.size(); i++) {
        double depth = 0, sum = 0, wsum = 0, visits = 0, childs = 0;
        for (const GridSlamProcessor::TNode* n = ps[i].node; n; n = n->parent) {
            depth += 1; sum += n->accWeight; wsum += depth * n->accWeight; visits += n->visitCounter; childs += n->childs;
        }
        double* r = rows + 6 * i;
        r[0] = depth; r[1] = ps[i].node ? ps[i].node->accWeight : 0.; r[2] = sum; r[3] = wsum; r[4] = visits; r[5] = childs;
    }
}
void gmh_set_register_with_scanmatch(void* h, int on) { static_cast<Handle*>(h)->gsp->setregisterWithScanMatch(on != 0); }
void gmh_set_early_registration(void* h, int on) { static_cast<Handle*>(h)->gsp->setearlyRegistration(on != 0); }

// params: maxUrange maxrange sigma kernelSize lstep astep iterations lsigma ogain lskip srr srt str stt linearUpdate
//         angularUpdate resampleThreshold period generateMap minimumScore            (20 doubles)
void gmh_setup(void* hv, unsigned beams, const double* angles, const double* laser_pose, const double* p) {
    Handle* h = static_cast<Handle*>(hv);
    delete h->laser;
    h->laser = new RangeSensor("FLASER", beams, 0., OrientedPoint(laser_pose[0], laser_pose[1], laser_pose[2]), 0., p[1]);
    for (unsigned i = 0; i < beams; i++) h->laser->beams()[i].pose.theta = angles[i];  // the caller's table, verbatim
    h->laser->updateBeamsLookup();
    h->smap.clear();
    h->smap.insert(std::make_pair(h->laser->getName(), h->laser));
    GridSlamProcessor& g = *h->gsp;
    g.setSensorMap(h->smap);
    g.setMatchingParameters(p[0], p[1], p[2], (int)p[3], p[4], p[5], (int)p[6], p[7], p[8], (unsigned int)p[9]);
    g.setMotionModelParameters(p[10], p[11], p[12], p[13]);
    g.setUpdateDistances(p[14], p[15], p[16]);
    g.setUpdatePeriod(p[17]);
    g.setgenerateMap(p[18] != 0);
    g.setminimumScore(p[19]);
}

void gmh_init(void* hv, unsigned particles, double xmin, double ymin, double xmax, double ymax, double delta, const double* pose, unsigned long seed) {
    Handle* h = static_cast<Handle*>(hv);
    h->gsp->init(particles, xmin, ymin, xmax, ymax, delta, OrientedPoint(pose[0], pose[1], pose[2]));
    if (seed) sampleGaussian(1, seed);  // srand48(seed) + one draw, like the reference's drivers
}

int gmh_process_scan(void* hv, const double* ranges, unsigned n, const double* odom, double stamp) {
    Handle* h = static_cast<Handle*>(hv);
    RangeReading rr(n, ranges, h->laser, stamp);
    rr.setPose(OrientedPoint(odom[0], odom[1], odom[2]));
    h->gsp->lastResampled = false;
    return h->gsp->processScan(rr) ? 1 : 0;
}

// GridSlamProcessor::processScan(reading, adaptParticles): a resampling at this reading continues with `adapt` particles
int gmh_process_scan_adapt(void* hv, const double* ranges, unsigned n, const double* odom, double stamp, int adapt) {
    Handle* h = static_cast<Handle*>(hv);
    RangeReading rr(n, ranges, h->laser, stamp);
    rr.setPose(OrientedPoint(odom[0], odom[1], odom[2]));
    h->gsp->lastResampled = false;
    return h->gsp->processScan(rr, adapt) ? 1 : 0;
}
void gmh_set_max_particles(void* h, unsigned n) { static_cast<Handle*>(h)->gsp->setmaxParticles(n); }

unsigned gmh_particles(void* hv) { return (unsigned)static_cast<Handle*>(hv)->gsp->getParticles().size(); }
unsigned gmh_local_particles(void* hv) { return static_cast<Handle*>(hv)->gsp->localParticles(); }
unsigned gmh_first_local(void* hv) { return static_cast<Handle*>(hv)->gsp->firstLocalParticle(); }
double gmh_neff(void* hv) { return static_cast<Handle*>(hv)->gsp->getneff(); }
int gmh_best(void* hv) { return static_cast<Handle*>(hv)->gsp->getBestParticleIndex(); }
int gmh_resamples(void* hv) { return static_cast<Handle*>(hv)->gsp->resamples; }
int gmh_last_resampled(void* hv) { return static_cast<Handle*>(hv)->gsp->lastResampled ? 1 : 0; }

// rows of 6: x y theta weight weightSum previousIndex
void gmh_state(void* hv, double* rows) {
    const GridSlamProcessor::ParticleVector& ps = static_cast<Handle*>(hv)->gsp->getParticles();
    for (size_t i = 0; i < ps.size(); i++) {
        double* r = rows + 6 * i;
        r[0] = ps[i].pose.x; r[1] = ps[i].pose.y; r[2] = ps[i].pose.theta; r[3] = ps[i].weight; r[4] = ps[i].weightSum; r[5] = ps[i].previousIndex;
    }
}

void gmh_indexes(void* hv, unsigned* out) {
    const std::vector<unsigned int>& idx = static_cast<Handle*>(hv)->gsp->getIndexes();
    if (!idx.empty()) memcpy(out, idx.data(), sizeof(unsigned) * idx.size());
}

// out[36]: ... then msHostMotion msHostScanMatch msHostTreeWeights msHostResample, 6 device totals, 4 host totals, 3 work totals
// out[16]: msScanMatch msWeights msRemap msRegister msMigrate msAllGather beamEvals scoreEvals cowCopies patchAllocs freeUpdates
//          hitUpdates poolInUse launches migratedParticles migratedBytes
void gmh_device_stats(void* hv, double* out) {
    const GridSlamProcessor::DeviceStats d = static_cast<Handle*>(hv)->gsp->deviceStats();
    const double v[36] = {d.msScanMatch, d.msWeights, d.msRemap, d.msRegister, d.msMigrate, d.msAllGather, (double)d.beamEvals, (double)d.scoreEvals,
                          (double)d.cowCopies, (double)d.patchAllocs, (double)d.freeUpdates, (double)d.hitUpdates, (double)d.poolInUse,
                          (double)d.launches, (double)d.migratedParticles, (double)d.migratedBytes,
                          d.msHostMotion, d.msHostScanMatch, d.msHostTreeWeights, d.msHostResample,
                          d.totalMs[0], d.totalMs[1], d.totalMs[2], d.totalMs[3], d.totalMs[4], d.totalMs[5],
                          d.totalHostMs[0], d.totalHostMs[1], d.totalHostMs[2], d.totalHostMs[3],
                          (double)d.totalBeamEvals, (double)d.totalScoreEvals, (double)d.totalScanMatches,
                          (double)d.totalMigratedParticles, (double)d.totalMigratedBytes, (double)d.poolHitInUse};
    memcpy(out, v, sizeof v);
}

// 6 words per LOCAL particle: patches, cells, sum n, sum visits, hash(cell, n, visits), hash(acc)
void gmh_map_digests(void* hv, unsigned long long* out) {
    const std::vector<unsigned long long> d = static_cast<Handle*>(hv)->gsp->mapDigests();
    if (!d.empty()) memcpy(out, d.data(), sizeof(unsigned long long) * d.size());
}

// getTrajectories(): deep copy of the trajectory tree; returns the number of nodes in the copy and frees it again
// (-> tests: shape; tools: what a clone() costs on the host)
unsigned long long gmh_copy_trajectories(void* hv, int count_nodes) {
    GridSlamProcessor::TNodeVector t = static_cast<Handle*>(hv)->gsp->getTrajectories();
    unsigned long long nodes = t.size();
    std::set<const GridSlamProcessor::TNode*> seen;
    if (count_nodes) nodes = 0;
    for (size_t i = 0; count_nodes && i < t.size(); i++)
        for (const GridSlamProcessor::TNode* n = t[i]; n && seen.insert(n).second; n = n->parent) nodes++;
    for (size_t i = 0; i < t.size(); i++) delete t[i];
    return nodes;
}

unsigned long long gmh_check_maps(void* hv) { return static_cast<Handle*>(hv)->gsp->checkMapConsistency(); }

// occupancy grid of one particle on the GPU; data has width*height bytes (query the size first with data == null)
void gmh_occupancy_grid(void* hv, unsigned particle, double thresh, signed char* data, int* width, int* height, double* origin) {
    std::vector<signed char> buf;
    static_cast<Handle*>(hv)->gsp->exportOccupancyGrid(particle, thresh, buf, *width, *height, origin[0], origin[1]);
    if (data) memcpy(data, buf.data(), buf.size());
}

// occupancy of the best particle's map around a world point, through the host mirror (Particle::map.cell): a smoke check of
// the lazy download path.  Returns n / visits or -1.
double gmh_occupancy(void* hv, unsigned particle, double wx, double wy) {
    const GridSlamProcessor::ParticleVector& ps = static_cast<Handle*>(hv)->gsp->getParticles();
    const ScanMatcherMap& m = ps[particle].map;
    return (double)m.cell(m.world2map(Point(wx, wy)));
}

}  // extern "C"
