// gridslamprocessor.cpp -- the per-scan particle update with the heavy stages on the B200
// (reference: gridfastslam/gridslamprocessor.cpp, gridslamprocessor_tree.cpp, gridslamprocessor.hxx).
//
// processScan keeps the reference's order of side effects exactly -- motion-model draws for every particle on
// every reading, the gate, scanMatch, onScanmatchUpdate, tree weights, resample (one drand48 draw only when it
// fires), tree weights again -- because the drand48 stream, the .gfs trace and the virtual hooks are all visible
// to users.  The four stages that are O(particles x beams) are single calls into the C ABI.
#include <gmapping/gridfastslam/gridslamprocessor.h>

#include <typeinfo>
#include <gmapping/utils/stat.h>
#include <stdlib.h>

#include <cassert>
#include <cmath>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <set>
#include <string>

#include <chrono>

#include "devicemap.h"

namespace GMapping {

using std::endl;

namespace {
struct Stopwatch {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double lap() {
        const std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
        const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        t0 = t1;
        return ms;
    }
};
}  // namespace

static const double kOdometryJumpWarning = 20;  // metres between two readings (reference: gridslamprocessor.cpp:17)

// ------------------------------------------------------------------------------------------ trajectory tree
GridSlamProcessor::TNode::TNode(const OrientedPoint& p, double w, TNode* n, unsigned int c)
    : pose(p), weight(w), accWeight(0), gweight(0), parent(n), reading(0), childs(c), visitCounter(0), flag(false), epoch(0), copy(0), copyEpoch(0) {
    if (parent) parent->childs++;
}

GridSlamProcessor::TNode::~TNode() {
    if (parent && --parent->childs == 0) delete parent;
    assert(!childs);
}

// slab allocation of trajectory nodes: a free list per thread over slabs that are never returned (a node may be deleted
// by another thread than the one that made it; it then joins that thread's list)
namespace {
struct NodeArena {
    struct Free { Free* next; };
    Free* head = nullptr;
};
NodeArena& nodeArena() {
    static thread_local NodeArena* a = new NodeArena;  // leaked on purpose: nodes may be deleted during thread shutdown
    return *a;
}
}  // namespace

void* GridSlamProcessor::TNode::operator new(std::size_t size) {
    if (size != sizeof(TNode)) return ::operator new(size);
    NodeArena& a = nodeArena();
    if (!a.head) {
        const std::size_t count = 4096, stride = (sizeof(TNode) + 15) & ~(std::size_t)15;
        char* slab = static_cast<char*>(::operator new(count * stride));
        for (std::size_t i = 0; i < count; i++) {
            NodeArena::Free* f = reinterpret_cast<NodeArena::Free*>(slab + (count - 1 - i) * stride);
            f->next = a.head; a.head = f;
        }
    }
    NodeArena::Free* f = a.head;
    a.head = f->next;
    return f;
}

void GridSlamProcessor::TNode::operator delete(void* p, std::size_t size) noexcept {
    if (!p) return;
    if (size != sizeof(TNode)) { ::operator delete(p); return; }
    NodeArena& a = nodeArena();
    NodeArena::Free* f = static_cast<NodeArena::Free*>(p);
    f->next = a.head; a.head = f;
}

GridSlamProcessor::Particle::Particle(const ScanMatcherMap& m)
    : map(m), pose(0, 0, 0), previousPose(0, 0, 0), weight(0), weightSum(0), gweight(0), previousIndex(0), node(0) {}

// ------------------------------------------------------------------------------------------ construction
static void defaults(GridSlamProcessor& g) {
    g.setobsSigmaGain(1);
    g.setresampleThreshold(0.5);
    g.setminimumScore(0.);
}

GridSlamProcessor::GridSlamProcessor()
    : m_beams(0), last_update_time_(0.0), period_(5.0), m_count(0), m_readingCount(0), m_linearDistance(0), m_angularDistance(0),
      m_neff(0), m_xmin(0), m_ymin(0), m_xmax(0), m_ymax(0), m_delta(0), m_regScore(0), m_critScore(0), m_maxMove(0),
      m_linearThresholdDistance(0), m_angularThresholdDistance(0), m_infoStream(std::cout), m_rank(0), m_world(1), m_firstLocal(0),
      m_localCount(0), m_treeEpoch(0) {
    m_motionModel.srr = m_motionModel.srt = m_motionModel.str = m_motionModel.stt = 0;
    m_hostMs[0] = m_hostMs[1] = m_hostMs[2] = m_hostMs[3] = 0;
    defaults(*this);
}

GridSlamProcessor::GridSlamProcessor(std::ostream& infoS)
    : m_beams(0), last_update_time_(0.0), period_(5.0), m_count(0), m_readingCount(0), m_linearDistance(0), m_angularDistance(0),
      m_neff(0), m_xmin(0), m_ymin(0), m_xmax(0), m_ymax(0), m_delta(0), m_regScore(0), m_critScore(0), m_maxMove(0),
      m_linearThresholdDistance(0), m_angularThresholdDistance(0), m_infoStream(infoS), m_rank(0), m_world(1), m_firstLocal(0),
      m_localCount(0), m_treeEpoch(0) {
    m_motionModel.srr = m_motionModel.srt = m_motionModel.str = m_motionModel.stt = 0;
    m_hostMs[0] = m_hostMs[1] = m_hostMs[2] = m_hostMs[3] = 0;
    defaults(*this);
}

// deep copy of the filter (reference: gridslamprocessor.cpp:38-97): parameters and bookkeeping verbatim, the trajectory
// tree through getTrajectories(), the particle maps through a device-side clone of the context.  Like the reference's,
// the copy reports to cout, has no .gfs stream and restarts its update period.
GridSlamProcessor::GridSlamProcessor(const GridSlamProcessor& gsp)
    : m_matcher(gsp.m_matcher), m_minimumScore(gsp.m_minimumScore), m_beams(gsp.m_beams), last_update_time_(0.0), period_(5.0),
      m_indexes(gsp.m_indexes), m_weights(gsp.m_weights), m_motionModel(gsp.m_motionModel), m_resampleThreshold(gsp.m_resampleThreshold),
      m_count(gsp.m_count), m_readingCount(gsp.m_readingCount), m_lastPartPose(gsp.m_lastPartPose), m_odoPose(gsp.m_odoPose), m_pose(gsp.m_pose),
      m_linearDistance(gsp.m_linearDistance), m_angularDistance(gsp.m_angularDistance), m_neff(gsp.m_neff), m_xmin(gsp.m_xmin), m_ymin(gsp.m_ymin),
      m_xmax(gsp.m_xmax), m_ymax(gsp.m_ymax), m_delta(gsp.m_delta), m_regScore(gsp.m_regScore), m_critScore(gsp.m_critScore), m_maxMove(gsp.m_maxMove),
      m_linearThresholdDistance(gsp.m_linearThresholdDistance), m_angularThresholdDistance(gsp.m_angularThresholdDistance),
      m_obsSigmaGain(gsp.m_obsSigmaGain), m_infoStream(std::cout), m_rank(0), m_world(1), m_firstLocal(0), m_localCount(gsp.m_localCount),
      m_treeEpoch(0), m_lazyTree(gsp.m_lazyTree), m_maxParticles(gsp.m_maxParticles), m_particleSlots(gsp.m_particleSlots), m_earlyRegistration(gsp.m_earlyRegistration), m_fusedRegister(gsp.m_fusedRegister) {
    if (gsp.m_world > 1) b200::fatal("GridSlamProcessor copy / clone(): a sharded filter cannot be cloned");
    m_hostMs[0] = m_hostMs[1] = m_hostMs[2] = m_hostMs[3] = 0;
    m_hostMsTotal[0] = m_hostMsTotal[1] = m_hostMsTotal[2] = m_hostMsTotal[3] = 0;
    if (!gsp.m_device) return;
    m_device.reset(new b200::DeviceContext(*gsp.m_device, 0));
    const TNodeVector leaves = gsp.getTrajectories();
    ScanMatcherMap mirror(Point((m_xmin + m_xmax) * .5, (m_ymin + m_ymax) * .5), 0., 0., m_delta);
    m_particles.reserve(std::max(gsp.m_particles.size(), (size_t)m_particleSlots));
    for (size_t i = 0; i < gsp.m_particles.size(); i++) {
        const Particle& src = gsp.m_particles[i];
        m_particles.push_back(Particle(mirror));
        Particle& p = m_particles.back();
        p.pose = src.pose; p.previousPose = src.previousPose; p.weight = src.weight; p.weightSum = src.weightSum; p.gweight = src.gweight;
        p.previousIndex = src.previousIndex;
        p.node = leaves[i];
        std::shared_ptr<b200::DeviceMapResidency> r(new b200::DeviceMapResidency);
        r->ctx = m_device;
        r->slot = (unsigned int)i;
        r->stale = true;
        p.map.setResidency(r);
    }
    if (m_count > 0) updateTreeWeights(false);
}

GridSlamProcessor* GridSlamProcessor::clone() const { return new GridSlamProcessor(*this); }

// free the trajectory tree: deleting a leaf walks up and frees every ancestor that loses its last child.
// Before the first processed scan all particles share the root, so each distinct node is deleted once.
static void releaseLeaves(GridSlamProcessor::ParticleVector& particles) {
    std::set<GridSlamProcessor::TNode*> leaves;
    for (GridSlamProcessor::ParticleVector::iterator it = particles.begin(); it != particles.end(); ++it) {
        if (it->node) leaves.insert(it->node);
        it->node = 0;
    }
    for (std::set<GridSlamProcessor::TNode*>::iterator it = leaves.begin(); it != leaves.end(); ++it) delete *it;
}

GridSlamProcessor::~GridSlamProcessor() { releaseLeaves(m_particles); }

// the beam table comes from the front laser ("FLASER", or "ROBOTLASER1" for the newer CARMEN format)
void GridSlamProcessor::setSensorMap(const SensorMap& smap) {
    SensorMap::const_iterator it = smap.find(std::string("FLASER"));
    if (it == smap.end()) it = smap.find(std::string("ROBOTLASER1"));
    assert(it != smap.end());
    const RangeSensor* rs = dynamic_cast<const RangeSensor*>(it->second);
    assert(rs && rs->beams().size());
    m_beams = static_cast<unsigned int>(rs->beams().size());
    std::vector<double> angles(m_beams);
    for (unsigned int i = 0; i < m_beams; i++) angles[i] = rs->beams()[i].pose.theta;
    m_matcher.setLaserParameters(m_beams, angles.data(), rs->getPose());
}

void GridSlamProcessor::setMatchingParameters(double urange, double range, double sigma, int kernsize, double lopt, double aopt, int iterations,
                                              double likelihoodSigma, double likelihoodGain, unsigned int likelihoodSkip) {
    m_obsSigmaGain = likelihoodGain;
    m_matcher.setMatchingParameters(urange, range, sigma, kernsize, lopt, aopt, iterations, likelihoodSigma, likelihoodSkip);
    if (m_infoStream)
        m_infoStream << " -maxUrange " << urange << " -maxUrange " << range << " -sigma     " << sigma << " -kernelSize " << kernsize
                     << " -lstep " << lopt << " -lobsGain " << m_obsSigmaGain << " -astep " << aopt << endl;
}

void GridSlamProcessor::setMotionModelParameters(double srr, double srt, double str, double stt) {
    m_motionModel.srr = srr; m_motionModel.srt = srt; m_motionModel.str = str; m_motionModel.stt = stt;
    if (m_infoStream) m_infoStream << " -srr " << srr << " -srt " << srt << " -str " << str << " -stt " << stt << endl;
}

void GridSlamProcessor::setUpdateDistances(double linear, double angular, double resampleThreshold) {
    m_linearThresholdDistance = linear;
    m_angularThresholdDistance = angular;
    m_resampleThreshold = resampleThreshold;
    if (m_infoStream) m_infoStream << " -linearUpdate " << linear << " -angularUpdate " << angular << " -resampleThreshold " << m_resampleThreshold << endl;
}

// `size` particles at initialPose, all on empty maps covering [xmin,xmax] x [ymin,ymax] at resolution delta.
// The maps are created on the device; the host Particle::map objects start as empty mirrors.
void GridSlamProcessor::init(unsigned int size, double xmin, double ymin, double xmax, double ymax, double delta, OrientedPoint initialPose) {
    m_xmin = xmin; m_ymin = ymin; m_xmax = xmax; m_ymax = ymax; m_delta = delta;
    if (m_infoStream)
        m_infoStream << " -xmin " << m_xmin << " -xmax " << m_xmax << " -ymin " << m_ymin << " -ymax " << m_ymax << " -delta " << m_delta
                     << " -particles " << size << endl;
    releaseLeaves(m_particles);
    m_particles.clear();

    if (size % m_world) b200::fatal("GridSlamProcessor::init: the particle count must be a multiple of the number of ranks");
    m_localCount = size / m_world;
    m_firstLocal = m_rank * m_localCount;
    unsigned long long pool = (unsigned long long)m_localCount * 512;  // patches; GM_B200_POOL_PATCHES overrides
    if (pool < 65536) pool = 65536;
    // sharded: as many spare slots as local particles, for resampled parents imported from other ranks
    const unsigned int slots = m_world > 1 ? 2 * m_localCount : std::max(m_localCount, m_maxParticles);  // (adaptParticles: one GPU only)
    m_particleSlots = m_world > 1 ? m_localCount : slots;
    m_device.reset(new b200::DeviceContext(slots, pool, std::max(size, m_maxParticles)));
    m_device->configure(m_matcher);
    const double center[2] = {(xmin + xmax) * .5, (ymin + ymax) * .5};
    m_device->check(gm_init_particles(m_device->h, m_localCount, center, xmax - xmin, ymax - ymin, delta), "gm_init_particles");
    m_device->particles = m_localCount;
    if (m_world > 1) m_device->check(gm_dist_init(m_device->h, m_rank, m_world, m_ncclId), "gm_dist_init");
    // the registration kernels keep running while processScan finishes its host-side bookkeeping (tree weights) and the
    // caller prepares the next reading; the next device call collects them
    m_device->check(gm_set_async_register(m_device->h, 1), "gm_set_async_register");
    m_device->check(gm_set_fused_register(m_device->h, getregisterWithScanMatch() && m_earlyRegistration ? 1 : 0), "gm_set_fused_register");

    TNode* root = new TNode(initialPose, 0, 0, 0);
    ScanMatcherMap mirror(Point(center[0], center[1]), 0., 0., delta);  // geometry + cells arrive with the first pull
    // room for every particle the filter may ever hold (adaptParticles): growing must not reallocate -- copying a Particle turns
    // its map into a plain host copy
    m_particles.reserve(std::max(size, m_particleSlots));
    for (unsigned int i = 0; i < size; i++) {
        m_particles.push_back(Particle(mirror));
        Particle& p = m_particles.back();
        p.pose = initialPose;
        p.previousPose = initialPose;
        p.setWeight(0);
        p.previousIndex = 0;
        p.node = root;  // all particles share the root; it is not counted as their child (reference: gridslamprocessor.cpp:381-396)
        if (i >= m_firstLocal && i < m_firstLocal + m_localCount) {  // this rank holds the particle's map
            std::shared_ptr<b200::DeviceMapResidency> r(new b200::DeviceMapResidency);
            r->ctx = m_device;
            r->slot = i - m_firstLocal;
            r->stale = true;
            p.map.setResidency(r);
        }
    }
    m_neff = (double)size;
    m_count = 0;
    m_readingCount = 0;
    m_linearDistance = m_angularDistance = 0;
}

void GridSlamProcessor::processTruePos(const OdometryReading& o) {
    const OdometrySensor* os = dynamic_cast<const OdometrySensor*>(o.getSensor());
    if (os && os->isIdeal() && m_outputStream) {
        m_outputStream << std::setiosflags(std::ios::fixed) << std::setprecision(3);
        m_outputStream << "SIMULATOR_POS " << o.getPose().x << " " << o.getPose().y << " ";
        m_outputStream << std::setiosflags(std::ios::fixed) << std::setprecision(6) << o.getPose().theta << " " << o.getTime() << endl;
    }
}

// a subclass may have edited a particle's map through the host mirror (m_particles is protected; non-const cell() / storage()
// / resize()): such maps go back to the device before the next stage reads them
void GridSlamProcessor::flushHostEdits() {
    for (size_t j = 0; j < m_localCount; j++) {
        Particle& p = m_particles[m_firstLocal + j];
        if (p.map.residency() && p.map.residency()->hostDirty) p.map.residency()->push(&p.map);
    }
}

void GridSlamProcessor::markMapsStale() {
    for (ParticleVector::iterator it = m_particles.begin(); it != m_particles.end(); ++it)
        if (it->map.residency()) it->map.residency()->stale = true;
}

// ------------------------------------------------------------------------------------------ the GPU stages
// scanMatch (reference: gridslamprocessor.hxx:15-75): optimize + accept/reject + likelihoodAndScore for all particles, and -- in
// the same stream-ordered device sequence, one synchronisation -- the normalize() of the updateTreeWeights that follows
// (gridslamprocessor.cpp:586): the device adds each particle's l to the log-weight it was given, exactly like the loop below.
// Sharded: the rows of every rank's particles come back in global order (they reach every GPU from the kernel's epilogue).
void GridSlamProcessor::scanMatch(const double* plainReading, const RangeReading* reading) {
    const size_t n = m_particles.size(), nl = m_localCount;
    m_poseBuf.resize(3 * nl); m_rowBuf.resize(5 * n); m_logwBuf.resize(n);
    for (size_t j = 0; j < nl; j++) {
        const OrientedPoint& q = m_particles[m_firstLocal + j].pose;
        m_poseBuf[3 * j] = q.x; m_poseBuf[3 * j + 1] = q.y; m_poseBuf[3 * j + 2] = q.theta;
    }
    for (size_t i = 0; i < n; i++) m_logwBuf[i] = m_particles[i].weight;
    m_weights.resize(n);
    m_device->configure(m_matcher);
    flushHostEdits();
    m_device->check(gm_scanmatch_weights_begin(m_device->h, plainReading, m_poseBuf.data(), m_minimumScore, m_logwBuf.data(), m_obsSigmaGain),
                    "gm_scanmatch_weights_begin");
    // While the GPU matches: the trajectory nodes this scan will add, one per particle, hung below the particles' present leaves
    // (what resample() does for a filter that does not resample -- nearly every scan; a resampling re-hangs them, see there).
    // Their poses are filled in once the scan match has delivered them.
    m_newNodes.resize(n);
    for (size_t i = 0; i < n; i++) {
        TNode* node = new TNode(m_particles[i].pose, 0.0, m_particles[i].node, 0);
        node->reading = reading;
        m_newNodes[i] = node;
    }
    m_device->check(gm_scanmatch_weights_end(m_device->h, m_rowBuf.data(), m_weights.data(), &m_neff), "gm_scanmatch_weights_end");
    double sumScore = 0;
    for (size_t i = 0; i < n; i++) {
        Particle& p = m_particles[i];
        const double* row = &m_rowBuf[5 * i];
        const double score = row[3], lik = row[4];
        sumScore += score;
        if (!(score > m_minimumScore) && m_infoStream) {
            m_infoStream << "Scan Matching Failed, using odometry. Likelihood=" << lik << endl;
            m_infoStream << "lp:" << m_lastPartPose.x << " " << m_lastPartPose.y << " " << m_lastPartPose.theta << endl;
            m_infoStream << "op:" << m_odoPose.x << " " << m_odoPose.y << " " << m_odoPose.theta << endl;
        }
        p.pose = OrientedPoint(row[0], row[1], row[2]);
        p.weight += lik;
        p.weightSum += lik;
        m_logwBuf[i] = p.weight;  // what m_weights / m_neff were computed from
    }
    m_weightsChanged = false;
    if (m_infoStream) m_infoStream << "Average Scan Matching Score=" << sumScore / n << endl;
}

// normalize (reference: gridslamprocessor.hxx:82-121): weights = softmax(gain * log-weights), Neff = 1 / sum w^2
void GridSlamProcessor::normalize() {
    const size_t n = m_particles.size();
    m_logwBuf.resize(n);
    for (size_t i = 0; i < n; i++) m_logwBuf[i] = m_particles[i].weight;
    m_weights.resize(n);
    m_device->check(gm_normalize_neff(m_device->h, m_logwBuf.data(), (uint32_t)n, m_obsSigmaGain, m_weights.data(), &m_neff), "gm_normalize_neff");
}

// particle copy by patch reference (when `parents` is given) + computeActiveArea + registerScan for every particle
void GridSlamProcessor::registerAll(const double* plainReading, const unsigned int* parents) {
    const size_t nl = m_localCount;
    m_poseBuf.resize(3 * nl);
    for (size_t j = 0; j < nl; j++) {
        const OrientedPoint& q = m_particles[m_firstLocal + j].pose;
        m_poseBuf[3 * j] = q.x; m_poseBuf[3 * j + 1] = q.y; m_poseBuf[3 * j + 2] = q.theta;
    }
    m_device->configure(m_matcher);
    flushHostEdits();
    if (m_world > 1)  // parents are global indexes; maps of parents on other ranks migrate here first
        m_device->check(gm_dist_register(m_device->h, plainReading, m_poseBuf.data(), parents), "gm_dist_register");
    else
        m_device->check(gm_register(m_device->h, plainReading, m_poseBuf.data(), parents), "gm_register");
    markMapsStale();
}

// particle copy by patch reference alone: the scan is already in the parents' maps (early registration)
void GridSlamProcessor::copyAll(const unsigned int* parents) {
    flushHostEdits();
    if (m_world > 1) m_device->check(gm_dist_copy_particles(m_device->h, parents), "gm_dist_copy_particles");
    else m_device->check(gm_copy_particles(m_device->h, parents), "gm_copy_particles");
    markMapsStale();
}

// resample (reference: gridslamprocessor.hxx:131-302)
bool GridSlamProcessor::resample(const double* plainReading, int adaptSize, const RangeReading* reading) {
    const size_t n = m_particles.size();
    const bool prepared = m_newNodes.size() == n;  // scanMatch() made this scan's nodes while the GPU was busy
    if (!(m_neff < m_resampleThreshold * n)) {
        // no resampling: every particle grows its own trajectory by one node and registers the scan
        for (size_t i = 0; i < n; i++) {
            Particle& p = m_particles[i];
            TNode* node = prepared ? m_newNodes[i] : new TNode(p.pose, 0.0, p.node, 0);
            node->pose = p.pose;
            node->reading = reading;
            p.node = node;
            p.previousIndex = (int)i;
        }
        m_newNodes.clear();
        if (!m_registeredEarly) registerAll(plainReading, 0);
        m_device->mapsAhead = false;
        return false;
    }
    if (m_infoStream) m_infoStream << "*************RESAMPLE***************" << endl;
    // adaptParticles (gridslamprocessor.hxx:153-156, particlefilter.h:196-203): the new generation has that many particles
    const size_t nn = adaptSize > 0 ? (size_t)adaptSize : n;
    if (nn != n) {
        if (m_world > 1) b200::fatal("processScan(reading, adaptParticles): a sharded filter keeps its particle count");
        if (nn > m_particleSlots)
            b200::fatal("processScan(reading, adaptParticles): more particles than the filter was set up for; call setmaxParticles() before init()");
    }
    // systematic resampling; the single uniform draw is taken here, on the host stream (particlefilter.h:199)
    const double u01 = drand48();
    m_indexes.resize(nn);
    uint32_t emitted = 0;
    static_assert(sizeof(unsigned int) == sizeof(uint32_t), "index type");
    m_device->check(gm_resample_indexes_n(m_device->h, m_weights.data(), (uint32_t)n, (uint32_t)nn, u01, m_indexes.data(), &emitted), "gm_resample_indexes");
    if (m_outputStream.is_open()) {
        m_outputStream << "RESAMPLE " << m_indexes.size() << " ";
        for (size_t i = 0; i < m_indexes.size(); i++) m_outputStream << m_indexes[i] << " ";
        m_outputStream << endl;
    }
    onResampleUpdate();

    // new generation: slot j continues parent m_indexes[j].  The maps stay where they are (slot j's host mirror is
    // bound to device slot j); gm_register re-points slot j's patch directory at its parent's patches.
    struct Carry { OrientedPoint pose, previousPose; double weightSum, gweight; TNode* node; };
    std::vector<Carry> next(nn);
    std::vector<char> kept(n, 0);
    for (size_t j = 0; j < nn; j++) {
        const Particle& src = m_particles[m_indexes[j]];
        // a parent's first child takes the node prepared below the parent's leaf, further children get nodes of their own
        TNode* node = (prepared && !kept[m_indexes[j]]) ? m_newNodes[m_indexes[j]] : new TNode(src.pose, 0, src.node, 0);
        kept[m_indexes[j]] = 1;
        node->pose = src.pose;
        node->reading = reading;
        next[j].pose = src.pose; next[j].previousPose = src.previousPose;
        next[j].weightSum = src.weightSum; next[j].gweight = src.gweight; next[j].node = node;
    }
    for (size_t i = 0; i < n; i++)
        if (!kept[i]) {  // leaves of trajectories that died out (deleting the prepared node takes its parent, the leaf, along)
            if (prepared) delete m_newNodes[i]; else delete m_particles[i].node;
            m_particles[i].node = 0;
        }
    m_newNodes.clear();
    if (nn != n) {
        // (the reference's bookkeeping of the leaves that died walks `j < m_indexes.size()`: with fewer particles it leaks the
        // unselected ones beyond the new count, with more it reads past the old vector; here every unselected leaf is freed)
        flushHostEdits();  // a parent may sit in a slot that is about to go
        if (nn < n) m_particles.erase(m_particles.begin() + nn, m_particles.end());
        else {
            ScanMatcherMap mirror(Point((m_xmin + m_xmax) * .5, (m_ymin + m_ymax) * .5), 0., 0., m_delta);
            if (nn > m_particles.capacity()) b200::fatal("processScan(reading, adaptParticles): internal: the particle vector was not reserved for growth");
            for (size_t j = n; j < nn; j++) {
                m_particles.push_back(Particle(mirror));
                std::shared_ptr<b200::DeviceMapResidency> r(new b200::DeviceMapResidency);
                r->ctx = m_device;
                r->slot = (unsigned int)j;
                r->stale = true;
                m_particles.back().map.setResidency(r);
            }
        }
        m_localCount = (unsigned int)nn;
        m_device->particles = (unsigned int)nn;
    }
    for (size_t j = 0; j < nn; j++) {
        Particle& p = m_particles[j];
        p.pose = next[j].pose; p.previousPose = next[j].previousPose;
        p.weightSum = next[j].weightSum; p.gweight = next[j].gweight; p.node = next[j].node;
        p.previousIndex = (int)m_indexes[j];
        p.setWeight(0);
    }
    m_weightsChanged = true;
    // a child starts from its parent's pose and map: registering the scan in the parent before the copy (early
    // registration) or in the child after it (the reference's order) writes the same cells
    if (nn != n) {  // the copies change the number of device maps; then, in the reference's order, the scan goes into every child
        m_device->check(gm_copy_particles_n(m_device->h, m_indexes.data(), (uint32_t)nn), "gm_copy_particles_n");
        markMapsStale();
        if (!m_registeredEarly) registerAll(plainReading, 0);
    } else if (m_registeredEarly) copyAll(m_indexes.data());
    else registerAll(plainReading, m_indexes.data());
    m_device->mapsAhead = false;
    return true;
}

// ------------------------------------------------------------------------------------------ tree weights
// normalize() unless no log-weight changed since the last time (the second updateTreeWeights of a processScan that did
// not resample sees the log-weights of the first: same weights, same Neff)
void GridSlamProcessor::normalizeIfNeeded() {
    const size_t n = m_particles.size();
    bool same = !m_weightsChanged && m_weights.size() == n && m_logwBuf.size() == n;
    // a subclass may have touched a log-weight in one of the hooks (m_particles is protected): compare with what was normalised
    for (size_t i = 0; same && i < n; i++) same = m_particles[i].weight == m_logwBuf[i];
    if (!same) { normalize(); m_weightsChanged = false; }
}

void GridSlamProcessor::updateTreeWeights(bool weightsAlreadyNormalized) {
    if (!weightsAlreadyNormalized) normalizeIfNeeded();
    if (m_lazyTree) { m_treeDirty = true; return; }
    propagateWeights();
}

void GridSlamProcessor::refreshTreeWeights() {
    if (!m_treeDirty || m_weights.size() != m_particles.size()) return;
    propagateWeights();
    m_treeDirty = false;
}

// zero accWeight / visitCounter on every node reachable from a particle (reference: gridslamprocessor_tree.cpp:255-268);
// the epoch stamp stops each walk at the first ancestor another particle already reset in this pass.  propagateWeights()
// does not need it: it resets each node on its first visit of a pass.
void GridSlamProcessor::resetTree() {
    m_treeEpoch++;
    for (ParticleVector::iterator it = m_particles.begin(); it != m_particles.end(); ++it)
        for (TNode* n = it->node; n && n->epoch != m_treeEpoch; n = n->parent) {
            n->accWeight = 0;
            n->visitCounter = 0;
            n->epoch = m_treeEpoch;
        }
}

// push a leaf's weight towards the root: a node forwards its sum once all of its children have reported.  A node seen for
// the first time in this pass (stale epoch stamp) starts from zero -- the reference's resetTree() folded into the walk.
static double pushUp(GridSlamProcessor::TNode* n, double weight, unsigned int epoch) {
    while (n) {
        if (n->epoch != epoch) { n->epoch = epoch; n->visitCounter = 0; n->accWeight = 0; }
        n->visitCounter++;
        n->accWeight += weight;
        assert(n->visitCounter <= n->childs);
        if (n->visitCounter != n->childs) return 0;
        weight = n->accWeight;
        n = n->parent;
    }
    return weight;
}

// resetTree + propagateWeights of the reference (gridslamprocessor_tree.cpp:255-346) in ONE walk: every node below which a
// particle lives is reached by this walk (a node waits for all of its children, and every child subtree ends in leaves that
// report), in the same order and with the same additions as the reference's second pass, so accWeight / visitCounter end
// up bit-identical; the separate zeroing pass over every ancestry chain (half of the pointer chasing) is gone.
double GridSlamProcessor::propagateWeights() {
    const unsigned int epoch = ++m_treeEpoch;
    double rootWeight = 0, leafSum = 0;
    for (size_t i = 0; i < m_particles.size(); i++) {
        const double w = m_weights[i];
        leafSum += w;
        TNode* n = m_particles[i].node;
        n->epoch = epoch;
        n->visitCounter = 0;
        n->accWeight = w;
        rootWeight += pushUp(n->parent, n->accWeight, epoch);
    }
    if (fabs(leafSum - 1.0) > 0.0001 || fabs(rootWeight - 1.0) > 0.0001) {
        std::cerr << "ERROR: root->accWeight=" << rootWeight << "    sum_leaf_weights=" << leafSum << endl;
        assert(0);
    }
    return rootWeight;
}

// ------------------------------------------------------------------------------------------ processScan
bool GridSlamProcessor::processScan(const RangeReading& reading, int adaptParticles) {
    const OrientedPoint relPose = reading.getPose();
    if (!m_count) m_lastPartPose = m_odoPose = relPose;

    Stopwatch sw;
    m_hostMs[0] = m_hostMs[1] = m_hostMs[2] = m_hostMs[3] = 0;
    // motion model: every reading, every particle, in order -- this is where the drand48 stream advances
    // (MotionModel::drawFromMotion(p, pnew, pold) with its particle-independent part -- the odometry step in the robot
    // frame and the three noise scales -- evaluated once per reading; same expressions, same draws, same results)
    {
        const MotionModel& mm = m_motionModel;
        const double sxy = 0.3 * mm.srr;
        const OrientedPoint delta = absoluteDifference(relPose, m_odoPose);
        const double sx = mm.srr * fabs(delta.x) + mm.str * fabs(delta.theta) + sxy * fabs(delta.y);
        const double sy = mm.srr * fabs(delta.y) + mm.str * fabs(delta.theta) + sxy * fabs(delta.x);
        const double st = mm.stt * fabs(delta.theta) + mm.srt * sqrt(delta.x * delta.x + delta.y * delta.y);
        MotionDraws& md = m_motionDraws;
        const size_t n = m_particles.size();
        if (md.valid && md.theta.size() == n && sx != 0 && sy != 0 && st != 0 && Drand48Stream::peek() == md.before) {
            // the samples were drawn ahead of time (prepareMotionDraws): same draws, same expressions, the stream moves to where they end
            Drand48Stream::put(md.after);
            for (size_t i = 0; i < n; i++) {
                OrientedPoint& pose = m_particles[i].pose;
                OrientedPoint noisy(delta);
                noisy.x += sx * md.b[3 * i] * md.r[3 * i];
                noisy.y += sy * md.b[3 * i + 1] * md.r[3 * i + 1];
                noisy.theta += st * md.b[3 * i + 2] * md.r[3 * i + 2];
                noisy.theta = fmod(noisy.theta, 2 * M_PI);
                if (noisy.theta > M_PI) noisy.theta -= 2 * M_PI;
                if (pose.theta == md.theta[i]) {  // absoluteSum(pose, noisy) with the prepared sin / cos of pose.theta
                    const double s = md.sn[i], c = md.cs[i];
                    pose = OrientedPoint(c * noisy.x - s * noisy.y, s * noisy.x + c * noisy.y, noisy.theta) + pose;
                } else {
                    pose = absoluteSum(pose, noisy);
                }
            }
        } else {
            Drand48Stream draws;  // the drand48() stream, stepped inline for these 3 x N samples
            for (ParticleVector::iterator it = m_particles.begin(); it != m_particles.end(); ++it) {
                OrientedPoint noisy(delta);
                noisy.x += draws.gaussian(sx);
                noisy.y += draws.gaussian(sy);
                noisy.theta += draws.gaussian(st);
                noisy.theta = fmod(noisy.theta, 2 * M_PI);
                if (noisy.theta > M_PI) noisy.theta -= 2 * M_PI;
                it->pose = absoluteSum(it->pose, noisy);
            }
        }
        md.valid = false;
    }
    m_hostMs[0] = sw.lap();

    if (m_outputStream.is_open()) {
        m_outputStream << std::setiosflags(std::ios::fixed) << std::setprecision(6);
        m_outputStream << "ODOM ";
        m_outputStream << std::setiosflags(std::ios::fixed) << std::setprecision(3) << m_odoPose.x << " " << m_odoPose.y << " ";
        m_outputStream << std::setiosflags(std::ios::fixed) << std::setprecision(6) << m_odoPose.theta << " ";
        m_outputStream << reading.getTime() << endl;
        m_outputStream << std::setiosflags(std::ios::fixed) << std::setprecision(6);
        m_outputStream << "ODO_UPDATE " << m_particles.size() << " ";
        for (ParticleVector::iterator it = m_particles.begin(); it != m_particles.end(); ++it) {
            m_outputStream << std::setiosflags(std::ios::fixed) << std::setprecision(3) << it->pose.x << " " << it->pose.y << " ";
            m_outputStream << std::setiosflags(std::ios::fixed) << std::setprecision(6) << it->pose.theta << " " << it->weight << " ";
        }
        m_outputStream << reading.getTime() << endl;
    }
    onOdometryUpdate();

    OrientedPoint move = relPose - m_odoPose;
    move.theta = atan2(sin(move.theta), cos(move.theta));
    m_linearDistance += sqrt(move * move);
    m_angularDistance += fabs(move.theta);
    if (m_linearDistance > kOdometryJumpWarning)
        std::cerr << "gmapping-b200: odometry jumped " << m_linearDistance << " m between two readings (from " << m_odoPose.x << " " << m_odoPose.y
                  << " to " << relPose.x << " " << relPose.y << "); continuing, the map may not fit" << endl;
    m_odoPose = relPose;

    bool processed = false;
    if (!m_count || m_linearDistance >= m_linearThresholdDistance || m_angularDistance >= m_angularThresholdDistance ||
        (period_ >= 0.0 && (reading.getTime() - last_update_time_) > period_)) {
        last_update_time_ = reading.getTime();
        if (m_outputStream.is_open()) {
            m_outputStream << std::setiosflags(std::ios::fixed) << std::setprecision(6);
            m_outputStream << "FRAME " << m_readingCount << " " << m_linearDistance << " " << m_angularDistance << endl;
        }
        if (m_infoStream) m_infoStream << "update frame " << m_readingCount << endl << "update ld=" << m_linearDistance << " ad=" << m_angularDistance << endl;
        assert(reading.size() == m_beams);
        const std::vector<double> plain(reading.begin(), reading.end());
        // the trajectory nodes keep their own copy of the scan (the caller may free `reading`)
        RangeReading* readingCopy = new RangeReading(reading.size(), &(reading[0]), static_cast<const RangeSensor*>(reading.getSensor()), reading.getTime());

        if (m_count > 0) {
            sw.lap();
            m_registeredEarly = false;
            const bool fused = getregisterWithScanMatch() && m_earlyRegistration;
            scanMatch(plain.data(), readingCopy);
            if (fused) { m_registeredEarly = true; m_device->mapsAhead = true; markMapsStale(); }  // the registration is already queued
            m_hostMs[1] = sw.lap();
            if (m_outputStream.is_open()) {
                m_outputStream << "LASER_READING " << reading.size() << " ";
                m_outputStream << std::setiosflags(std::ios::fixed) << std::setprecision(2);
                for (RangeReading::const_iterator b = reading.begin(); b != reading.end(); ++b) m_outputStream << *b << " ";
                const OrientedPoint p = reading.getPose();
                m_outputStream << std::setiosflags(std::ios::fixed) << std::setprecision(6);
                m_outputStream << p.x << " " << p.y << " " << p.theta << " " << reading.getTime() << endl;
                m_outputStream << "SM_UPDATE " << m_particles.size() << " ";
                for (ParticleVector::const_iterator it = m_particles.begin(); it != m_particles.end(); ++it) {
                    m_outputStream << std::setiosflags(std::ios::fixed) << std::setprecision(3) << it->pose.x << " " << it->pose.y << " ";
                    m_outputStream << std::setiosflags(std::ios::fixed) << std::setprecision(6) << it->pose.theta << " " << it->weight << " ";
                }
                m_outputStream << endl;
            }
            onScanmatchUpdate();
            // early registration: the scan goes into every particle's map at its scan-matched pose now, and the kernels run
            // while the host normalises the weights, decides about resampling and grows the trajectory tree.  Until the
            // point where the reference registers (inside resample()) the device maps are one scan ahead of the
            // reference's: reading Particle::map in that window (from onResampleUpdate) is refused, see setearlyRegistration
            sw.lap();
            normalizeIfNeeded();  // first: a one-CTA kernel that would otherwise queue behind the registration's grid
            m_hostMs[2] += sw.lap();
            if (m_earlyRegistration && !m_registeredEarly) { registerAll(plain.data(), 0); m_registeredEarly = true; m_device->mapsAhead = true; }
            sw.lap();
            // updateTreeWeights(false): the weights and Neff are needed now; the propagation into the trajectory tree can only be
            // observed before the second updateTreeWeights below through onResampleUpdate(), i.e. when the filter resamples --
            // otherwise it is skipped here and the second call, which recomputes every node, does it once
            normalizeIfNeeded();
            if (m_neff < m_resampleThreshold * m_particles.size() && !m_lazyTree) propagateWeights();
            m_hostMs[2] += sw.lap();
            if (m_infoStream) m_infoStream << "neff= " << m_neff << endl;
            if (m_outputStream.is_open()) {
                m_outputStream << std::setiosflags(std::ios::fixed) << std::setprecision(6);
                m_outputStream << "NEFF " << m_neff << endl;
            }
            sw.lap();
            resample(plain.data(), adaptParticles, readingCopy);
            m_hostMs[3] = sw.lap();
        } else {
            if (m_infoStream) m_infoStream << "Registering First Scan" << endl;
            for (ParticleVector::iterator it = m_particles.begin(); it != m_particles.end(); ++it) {
                TNode* node = new TNode(it->pose, 0., it->node, 0);
                node->reading = readingCopy;
                it->node = node;
            }
            registerAll(plain.data(), 0);
        }
        sw.lap();
        updateTreeWeights(false);
        m_hostMs[2] += sw.lap();
        m_lastPartPose = m_odoPose;
        m_linearDistance = 0;
        m_angularDistance = 0;
        m_count++;
        processed = true;
        for (ParticleVector::iterator it = m_particles.begin(); it != m_particles.end(); ++it) it->previousPose = it->pose;
        // the GPU is registering the scan: the host uses the time for the part of the next reading's motion update that does
        // not depend on the odometry (outside the stage times: it overlaps device work)
        if (m_count > 1 && m_earlyRegistration) prepareMotionDraws();
    }
    for (int k = 0; k < 4; k++) m_hostMsTotal[k] += m_hostMs[k];
    if (m_outputStream.is_open()) m_outputStream << std::flush;
    m_readingCount++;
    return processed;
}

// The registration queued on the device right behind the scan match: what it takes away is the window in which an overridden
// onScanmatchUpdate() could still read the maps without this scan in them.  A filter object that is a plain GridSlamProcessor
// has no such hook (the ROS wrapper and gfs-carmen use it that way), neither has a sharded one a use for it: on by default
// there, off for subclasses unless they ask (setregisterWithScanMatch).
bool GridSlamProcessor::getregisterWithScanMatch() const {
    if (m_fusedRegister >= 0) return m_fusedRegister > 0;
    return m_world > 1 || typeid(*this) == typeid(GridSlamProcessor);
}

void GridSlamProcessor::prepareMotionDraws() {
    MotionDraws& md = m_motionDraws;
    const size_t n = m_particles.size();
    md.b.resize(3 * n); md.r.resize(3 * n); md.theta.resize(n); md.sn.resize(n); md.cs.resize(n);
    md.before = Drand48Stream::peek();
    Drand48Stream ahead(md.before);
    for (size_t i = 0; i < n; i++) {
        ahead.polar(md.b[3 * i], md.r[3 * i]);
        ahead.polar(md.b[3 * i + 1], md.r[3 * i + 1]);
        ahead.polar(md.b[3 * i + 2], md.r[3 * i + 2]);
        const double th = m_particles[i].pose.theta;
        md.theta[i] = th; md.sn[i] = sin(th); md.cs[i] = cos(th);
    }
    md.after = ahead.state();
    md.valid = true;
}

std::ofstream& GridSlamProcessor::outputStream() { return m_outputStream; }
std::ostream& GridSlamProcessor::infoStream() { return m_infoStream; }

// argmax of weightSum, first wins (reference: gridslamprocessor.cpp:677-688)
int GridSlamProcessor::getBestParticleIndex() const {
    unsigned int best = 0;
    double bw = -std::numeric_limits<double>::max();
    for (unsigned int i = 0; i < m_particles.size(); i++)
        if (bw < m_particles[i].weightSum) { bw = m_particles[i].weightSum; best = i; }
    return (int)best;
}

void GridSlamProcessor::onScanmatchUpdate() {}
void GridSlamProcessor::onResampleUpdate() {}
void GridSlamProcessor::onOdometryUpdate() {}

// ------------------------------------------------------------------------------------------ trajectories
// deep copy of the trajectory tree: one new leaf per particle, shared ancestors copied once
// (reference: gridslamprocessor_tree.cpp:57-147)
GridSlamProcessor::TNodeVector GridSlamProcessor::getTrajectories() const {
    const_cast<GridSlamProcessor*>(this)->refreshTreeWeights();
    // every node remembers its copy for the duration of this call (TNode::copy, stamped with a fresh epoch): one pass, no map
    if (++m_copyEpoch == 0) {  // the stamp wrapped: no stale stamp may look current
        for (ParticleVector::const_iterator it = m_particles.begin(); it != m_particles.end(); ++it)
            for (const TNode* n = it->node; n; n = n->parent) n->copyEpoch = 0;
        m_copyEpoch = 1;
    }
    const unsigned int stamp = m_copyEpoch;
    TNodeVector leaves;
    leaves.reserve(m_particles.size());
    std::vector<const TNode*> chain;
    for (ParticleVector::const_iterator it = m_particles.begin(); it != m_particles.end(); ++it) {
        // collect the not-yet-copied part of this particle's chain, then build it top-down
        chain.clear();
        const TNode* n = it->node;
        while (n && n->copyEpoch != stamp) { chain.push_back(n); n = n->parent; }
        TNode* parentCopy = n ? n->copy : 0;
        for (size_t k = chain.size(); k-- > 0;) {
            const TNode* src = chain[k];
            TNode* c = new TNode(src->pose, src->weight, parentCopy, 0);
            c->reading = src->reading;
            c->accWeight = src->accWeight;
            c->gweight = src->gweight;
            src->copy = c; src->copyEpoch = stamp;
            parentCopy = c;
        }
        leaves.push_back(it->node ? it->node->copy : 0);
    }
    return leaves;
}

// move every particle along a stored trajectory (reference: gridslamprocessor_tree.cpp:149-224), oldest node first: the pose
// is composed relative to the path, the node's weight difference is added to weight and weightSum, computeActiveArea runs for
// the node's scan (it may resize the maps; nothing is registered) and every particle's own trajectory grows by one node.
// The reference's quirks are kept: the rotation uses the NEW path heading (oldPose is overwritten before it is used), and
// the weight difference is always taken against the FIRST node (oldWeight is never advanced).  Every node needs a reading,
// like in the reference (it dereferences aux->reading); m_count and the tree weights are left alone.
// (The reference then frees its private copy of the path with `delete` in a loop, which trips assert(!childs) for any path
// longer than one node; there is no private copy here.)
void GridSlamProcessor::integrateScanSequence(TNode* node) {
    std::vector<const TNode*> order;
    for (const TNode* n = node; n; n = n->parent) order.push_back(n);
    if (m_infoStream) m_infoStream << "Restoring State Nodes=" << (double)order.size() << endl;
    bool first = true;
    double oldWeight = 0;
    OrientedPoint oldPose;
    const size_t nl = m_localCount;
    for (size_t k = order.size(); k-- > 0;) {
        const TNode* aux = order[k];
        if (first) { oldPose = aux->pose; first = false; oldWeight = aux->weight; }
        const OrientedPoint dp = aux->pose - oldPose;
        const double dw = aux->weight - oldWeight;
        oldPose = aux->pose;
        // (the reference dereferences aux->reading unconditionally, i.e. crashes on a node without one -- the root of every
        // trajectory getTrajectories() returns; here such a node moves the particles but has no scan to bound)
        std::vector<double> plain;
        if (aux->reading) { assert(aux->reading->size() == m_beams); plain.assign(aux->reading->begin(), aux->reading->end()); }
        for (ParticleVector::iterator it = m_particles.begin(); it != m_particles.end(); ++it) {
            const double s = sin(oldPose.theta - it->pose.theta), c = cos(oldPose.theta - it->pose.theta);
            it->pose.x += c * dp.x - s * dp.y;
            it->pose.y += s * dp.x + c * dp.y;
            it->pose.theta += dp.theta;
            it->pose.theta = atan2(sin(it->pose.theta), cos(it->pose.theta));
            it->weight += dw;
            it->weightSum += dw;
            it->node = new TNode(it->pose, 0.0, it->node);
        }
        if (plain.empty()) continue;
        // computeActiveArea for every (local) particle in one call: the bounding box of the scan and the Map::resize it triggers
        m_poseBuf.resize(3 * nl);
        for (size_t j = 0; j < nl; j++) {
            const OrientedPoint& q = m_particles[m_firstLocal + j].pose;
            m_poseBuf[3 * j] = q.x; m_poseBuf[3 * j + 1] = q.y; m_poseBuf[3 * j + 2] = q.theta;
        }
        m_device->configure(m_matcher);
        flushHostEdits();
        m_device->check(gm_compute_active_area(m_device->h, plain.data(), m_poseBuf.data()), "gm_compute_active_area");
        markMapsStale();
    }
    m_weightsChanged = true;
}

// ------------------------------------------------------------------------------------------ B200 additions
GridSlamProcessor::DeviceStats GridSlamProcessor::deviceStats() const {
    DeviceStats d = DeviceStats();
    if (!m_device) return d;
    gm_stats s;
    m_device->check(gm_get_stats(m_device->h, &s), "gm_get_stats");
    d.msScanMatch = s.ms_scanmatch; d.msWeights = s.ms_weights; d.msRemap = s.ms_remap; d.msRegister = s.ms_register;
    d.beamEvals = s.beam_evals; d.scoreEvals = s.score_evals; d.cowCopies = s.cow_copies; d.patchAllocs = s.patch_allocs;
    d.freeUpdates = s.free_updates; d.hitUpdates = s.hit_updates; d.poolInUse = s.pool_in_use; d.poolHitInUse = s.pool_hit_in_use; d.launches = s.launches;
    for (int k = 0; k < 6; k++) d.totalMs[k] = s.total_ms[k];
    for (int k = 0; k < 4; k++) d.totalHostMs[k] = m_hostMsTotal[k];
    d.totalBeamEvals = s.total_beam_evals; d.totalScoreEvals = s.total_score_evals; d.totalScanMatches = s.total_scanmatches;
    d.totalMigratedParticles = s.total_migrated_particles; d.totalMigratedBytes = s.total_migrated_bytes;
    d.msHostMotion = m_hostMs[0]; d.msHostScanMatch = m_hostMs[1]; d.msHostTreeWeights = m_hostMs[2]; d.msHostResample = m_hostMs[3];
    d.msMigrate = s.ms_migrate; d.msAllGather = s.ms_allgather; d.migratedParticles = s.migrated_particles; d.migratedBytes = s.migrated_bytes;
    return d;
}

void GridSlamProcessor::setDistributed(int rank, int world, const void* ncclId) {
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !ncclId)) b200::fatal("GridSlamProcessor::setDistributed: bad rank / world / id");
    m_rank = rank;
    m_world = world;
    // every rank carries the whole trajectory tree: the weight propagation (O(particles) pointer chasing per scan, replicated)
    // is deferred to the readers of TNode::accWeight -- getTrajectories() / refreshTreeWeights() -- unless asked otherwise
    if (world > 1) m_lazyTree = true;
    if (ncclId) memcpy(m_ncclId, ncclId, sizeof m_ncclId);
}

void GridSlamProcessor::exportOccupancyGrid(unsigned int particle, double occThresh, std::vector<signed char>& data, int& width, int& height,
                                            double& originX, double& originY) const {
    if (particle < m_firstLocal || particle >= m_firstLocal + m_localCount) b200::fatal("exportOccupancyGrid: the particle's map lives on another rank");
    const unsigned int slot = particle - m_firstLocal;
    gm_map_info info;
    m_device->check(gm_map_info_get(m_device->h, slot, &info), "gm_map_info_get");
    width = info.patches_x * 32;
    height = info.patches_y * 32;
    data.resize((size_t)width * height);
    m_device->check(gm_export_occupancy(m_device->h, slot, occThresh, reinterpret_cast<int8_t*>(data.data()), (uint32_t)width, (uint32_t)height),
                    "gm_export_occupancy");
    originX = (0 - info.size_x2) * info.delta + info.center_x;  // Map::map2world(0, 0)
    originY = (0 - info.size_y2) * info.delta + info.center_y;
}

unsigned long long GridSlamProcessor::checkMapConsistency() const {
    uint64_t bad = 0, used = 0;
    m_device->check(gm_check_refcounts(m_device->h, &bad, &used), "gm_check_refcounts");
    return used;
}

std::vector<unsigned long long> GridSlamProcessor::mapDigests() const {
    const size_t n = m_localCount;
    std::vector<gm_digest> d(n);
    if (n) m_device->check(gm_map_digest(m_device->h, 0, (uint32_t)n, d.data()), "gm_map_digest");
    std::vector<unsigned long long> out(6 * n);
    for (size_t i = 0; i < n; i++) {
        out[6 * i] = d[i].patches; out[6 * i + 1] = d[i].cells; out[6 * i + 2] = d[i].sum_n; out[6 * i + 3] = d[i].sum_visits;
        out[6 * i + 4] = d[i].hash_counts; out[6 * i + 5] = d[i].hash_acc;
    }
    return out;
}

}  // namespace GMapping
