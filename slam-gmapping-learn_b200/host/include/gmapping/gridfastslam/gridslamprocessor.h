// gridslamprocessor.h -- GMapping::GridSlamProcessor with the reference's API
// (include/gmapping/gridfastslam/gridslamprocessor.h:35-367).
//
// What stays on the host: the drand48-driven motion model, the update gate, the trajectory tree,
// the .gfs trace and the weight bookkeeping.  What runs on the B200 through the C ABI (include/gm_b200.h):
// scanMatch (gm_scanmatch), normalize/Neff (gm_normalize_neff), resampleIndexes (gm_resample_indexes) and the
// particle copy + registerScan loop (gm_register).  Particle::map is a device-resident ScanMatcherMap: reading it
// from the host downloads that particle's patches on first use after every update (grid/map.h).
#ifndef GMB200_GRIDSLAMPROCESSOR_H
#define GMB200_GRIDSLAMPROCESSOR_H
#include <stdint.h>
#include <gmapping/gridfastslam/gridfastslam_export.h>
#include <gmapping/gridfastslam/motionmodel.h>
#include <gmapping/particlefilter/particlefilter.h>
#include <gmapping/scanmatcher/scanmatcher.h>
#include <gmapping/sensor/sensor_odometry/odometryreading.h>
#include <gmapping/sensor/sensor_range/rangereading.h>
#include <gmapping/sensor/sensor_range/rangesensor.h>
#include <gmapping/utils/macro_params.h>
#include <gmapping/utils/point.h>

#include <climits>
#include <cstddef>
#include <deque>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <vector>

namespace GMapping {

namespace b200 { class DeviceContext; }

class GRIDFASTSLAM_EXPORT GridSlamProcessor {
  public:
    // node of the reversed trajectory tree: every particle points at its newest node, nodes point at their parent,
    // `childs` counts the references; deleting a node whose parent has no other child deletes the parent too
    struct TNode {
        TNode(const OrientedPoint& pose, double weight, TNode* parent = 0, unsigned int childs = 0);
        ~TNode();
        // B200 build: nodes come from per-thread slabs (one is created per particle and scan; `new` / `delete` work as ever)
        static void* operator new(std::size_t size);
        static void operator delete(void* p, std::size_t size) noexcept;
        OrientedPoint pose;
        double weight;
        double accWeight;  // sum of the (normalised) weights of the leaves below this node
        double gweight;
        TNode* parent;
        const RangeReading* reading;
        unsigned int childs;
        mutable unsigned int visitCounter;
        mutable bool flag;
        mutable unsigned int epoch;  // B200 build: pass stamp so that shared ancestors are reset once per pass
        // B200 build: getTrajectories() notes each node's copy here (valid while copyEpoch is the filter's current one) instead of
        // looking nodes up in a map
        mutable TNode* copy;
        mutable unsigned int copyEpoch;
    };
    typedef std::vector<GridSlamProcessor::TNode*> TNodeVector;
    typedef std::deque<GridSlamProcessor::TNode*> TNodeDeque;

    struct Particle {
        Particle(const ScanMatcherMap& map);
        inline operator double() const { return weight; }
        inline operator OrientedPoint() const { return pose; }
        inline void setWeight(double w) { weight = w; }
        ScanMatcherMap map;  // host mirror of the particle's device-resident map
        OrientedPoint pose;
        OrientedPoint previousPose;
        double weight;     // log-weight accumulated since the last resampling
        double weightSum;  // log-weight accumulated over the whole run
        double gweight;
        int previousIndex;
        TNode* node;
    };
    typedef std::vector<Particle> ParticleVector;

    GridSlamProcessor();
    GridSlamProcessor(std::ostream& infoStr);
    // deep copy of the filter including the trajectory tree and (device-side) the particle maps
    GridSlamProcessor* clone() const;
    virtual ~GridSlamProcessor();

    void setSensorMap(const SensorMap& smap);
    void init(unsigned int size, double xmin, double ymin, double xmax, double ymax, double delta,
              OrientedPoint initialPose = OrientedPoint(0, 0, 0));
    void setMatchingParameters(double urange, double range, double sigma, int kernsize, double lopt, double aopt, int iterations,
                               double likelihoodSigma = 1, double likelihoodGain = 1, unsigned int likelihoodSkip = 0);
    void setMotionModelParameters(double srr, double srt, double str, double stt);
    void setUpdateDistances(double linear, double angular, double resampleThreshold);
    void setUpdatePeriod(double p) { period_ = p; }

    void processTruePos(const OdometryReading& odometry);
    // one laser reading: motion-model sampling always; when the robot moved enough (or the period elapsed) scan
    // matching, weighting, resampling and map update.  Returns whether the filter was updated.
    bool processScan(const RangeReading& reading, int adaptParticles = 0);

    TNodeVector getTrajectories() const;
    void integrateScanSequence(TNode* node);

    ScanMatcher m_matcher;
    std::ofstream& outputStream();
    std::ostream& infoStream();
    inline const ParticleVector& getParticles() const { return m_particles; }
    inline const std::vector<unsigned int>& getIndexes() const { return m_indexes; }
    int getBestParticleIndex() const;

    virtual void onOdometryUpdate();
    virtual void onResampleUpdate();
    virtual void onScanmatchUpdate();

    // parameter accessors forwarded to the matcher / the motion model: get<name>() / set<name>(value), the names the
    // reference generates with MEMBER_PARAM_SET_GET / STRUCT_PARAM_SET_GET (gridslamprocessor.h:220-245)
#define GMB200_MATCHER_PARAM(type, name)                               \
    inline type get##name() const { return m_matcher.get##name(); }   \
    inline void set##name(type value) { m_matcher.set##name(value); }
#define GMB200_MOTION_PARAM(name)                                      \
    inline double get##name() const { return m_motionModel.name; }    \
    inline void set##name(double value) { m_motionModel.name = value; }
    GMB200_MATCHER_PARAM(double, laserMaxRange) GMB200_MATCHER_PARAM(double, usableRange)
    GMB200_MATCHER_PARAM(double, gaussianSigma) GMB200_MATCHER_PARAM(double, likelihoodSigma)
    GMB200_MATCHER_PARAM(int, kernelSize)
    GMB200_MATCHER_PARAM(double, optAngularDelta) GMB200_MATCHER_PARAM(double, optLinearDelta)
    GMB200_MATCHER_PARAM(unsigned int, optRecursiveIterations) GMB200_MATCHER_PARAM(unsigned int, likelihoodSkip)
    GMB200_MATCHER_PARAM(double, llsamplerange) GMB200_MATCHER_PARAM(double, lasamplerange)
    GMB200_MATCHER_PARAM(double, llsamplestep) GMB200_MATCHER_PARAM(double, lasamplestep)
    GMB200_MATCHER_PARAM(bool, generateMap)
    GMB200_MATCHER_PARAM(bool, enlargeStep)  // sic: the reference forwards this double as a bool
    GMB200_MATCHER_PARAM(OrientedPoint, laserPose)
    GMB200_MOTION_PARAM(srr) GMB200_MOTION_PARAM(srt) GMB200_MOTION_PARAM(str) GMB200_MOTION_PARAM(stt)
#undef GMB200_MATCHER_PARAM
#undef GMB200_MOTION_PARAM

    PARAM_SET_GET(double, minimumScore, protected, public, public);

    // ---- B200 additions (not in the reference API)
    // device time of the last processed scan's stages and the work they did (C ABI gm_stats)
    struct DeviceStats {
        double msScanMatch, msWeights, msRemap, msRegister, msMigrate, msAllGather;
        unsigned long long beamEvals, scoreEvals, cowCopies, patchAllocs, freeUpdates, hitUpdates, poolInUse, launches;
        unsigned long long poolHitInUse;  // of poolInUse, the patches that hold end points (36 KB each; the others 4 KB)
        unsigned long long migratedParticles, migratedBytes;
        // wall clock of the host-side parts of the last processScan (ms): motion model, scanMatch call incl. transfers,
        // both updateTreeWeights (normalize on the GPU + tree propagation), resample call incl. registration
        double msHostMotion, msHostScanMatch, msHostTreeWeights, msHostResample;
        // running totals since init(): device ms per stage (scanmatch, weights, remap, register, migrate, allgather), host ms
        // (motion, scanMatch call, tree weights, resample call), work done
        double totalMs[6], totalHostMs[4];
        unsigned long long totalBeamEvals, totalScoreEvals, totalScanMatches, totalMigratedParticles, totalMigratedBytes;
    };
    DeviceStats deviceStats() const;
    // Multi-GPU sharding (call before init): this process is rank `rank` of `world`, one GPU each; ncclId is the 128-byte
    // id from gm_dist_unique_id(), the same on every rank.  Every rank keeps the full host state (poses, weights,
    // trajectory tree, drand48 stream -- seed all ranks identically); rank r owns the device maps of particles
    // [r*N/world, (r+1)*N/world), the other particles' Particle::map stay empty on this rank.
    void setDistributed(int rank, int world, const void* ncclId);
    inline unsigned int firstLocalParticle() const { return m_firstLocal; }
    inline unsigned int localParticles() const { return m_localCount; }
    // updateTreeWeights() walks every node of the trajectory tree (O(particles x scans) while nothing is resampled away).
    // With lazy = true processScan only normalises the weights (Neff, resampling are unaffected); TNode::accWeight is
    // brought up to date by getTrajectories() or refreshTreeWeights().  Default: false, the reference's behaviour.
    // processScan(reading, adaptParticles) may change the particle count while resampling (gridslamprocessor.hxx:153-156): any
    // count up to the larger of init()'s and this one (call before init(); the device slots are sized there).  One GPU only.
    void setmaxParticles(unsigned int n) { m_maxParticles = n; }
    unsigned int getmaxParticles() const { return m_maxParticles; }
    void setlazyTreeWeights(bool lazy) { m_lazyTree = lazy; }
    bool getlazyTreeWeights() const { return m_lazyTree; }
    // B200 build: register the scan right after scan matching (before Neff and the resampling decision are known) so the
    // kernels overlap the host-side bookkeeping; a resampling then only copies particles (same maps as the reference's
    // copy-then-register order).  On by default.  Off: the reference's order, and Particle::map may be read
    // inside onResampleUpdate().
    void setearlyRegistration(bool early) { m_earlyRegistration = early; }
    bool getearlyRegistration() const { return m_earlyRegistration; }
    // B200 build: queue the registration on the device right behind the scan match (gm_set_fused_register), so that it runs while
    // the weights come back to the host -- on a sharded filter, while this rank waits for the other ranks' results.  The device
    // maps are then one scan ahead already inside onScanmatchUpdate() (reading Particle::map there aborts with a message, as in
    // onResampleUpdate()).  Needs early registration.  Default: on for a sharded filter, off on a single GPU.  Call before init().
    void setregisterWithScanMatch(bool on) { m_fusedRegister = on ? 1 : 0; }
    bool getregisterWithScanMatch() const;  // default: on for a sharded filter and for a plain GridSlamProcessor (no subclass, no hook)
    void refreshTreeWeights();
    // occupancy grid of one local particle's map computed on the GPU (what gmapping/src/slam_gmapping.cpp:956-993 fills by
    // walking smap.cell(p) on the host): data[y * width + x] = -1 unknown, 100 if n/visits > occThresh, else 0;
    // (originX, originY) = world position of cell (0, 0)
    void exportOccupancyGrid(unsigned int particle, double occThresh, std::vector<signed char>& data, int& width, int& height,
                             double& originX, double& originY) const;
    // the reference's MAP_CONSISTENCY_CHECK (gridfastslam/gridslamprocessor.cpp:82-143: every patch's share count against the
    // number of particle maps that point at it) on the device: aborts with a message on any difference; returns the number of
    // distinct patches the maps of this filter (and of its clones, which share the pool) reference
    unsigned long long checkMapConsistency() const;
    // order-independent digests of the LOCAL particle maps (patches, cells, sum n, sum visits, hash of (cell, n, visits), hash of acc)
    std::vector<unsigned long long> mapDigests() const;

  protected:
    GridSlamProcessor(const GridSlamProcessor& gsp);

    unsigned int m_beams;
    double last_update_time_;
    double period_;
    ParticleVector m_particles;
    std::vector<unsigned int> m_indexes;
    std::vector<double> m_weights;
    MotionModel m_motionModel;
    PARAM_SET_GET(double, resampleThreshold, protected, public, public);
    int m_count, m_readingCount;
    OrientedPoint m_lastPartPose;
    OrientedPoint m_odoPose;
    OrientedPoint m_pose;
    double m_linearDistance, m_angularDistance;
    PARAM_GET(double, neff, protected, public);
    PARAM_GET(double, xmin, protected, public);
    PARAM_GET(double, ymin, protected, public);
    PARAM_GET(double, xmax, protected, public);
    PARAM_GET(double, ymax, protected, public);
    PARAM_GET(double, delta, protected, public);
    PARAM_SET_GET(double, regScore, protected, public, public);
    PARAM_SET_GET(double, critScore, protected, public, public);
    PARAM_SET_GET(double, maxMove, protected, public, public);
    PARAM_SET_GET(double, linearThresholdDistance, protected, public, public);
    PARAM_SET_GET(double, angularThresholdDistance, protected, public, public);
    PARAM_SET_GET(double, obsSigmaGain, protected, public, public);
    std::ofstream m_outputStream;
    std::ostream& m_infoStream;

  private:
    void scanMatch(const double* plainReading, const RangeReading* reading);
    void normalize();
    void normalizeIfNeeded();
    bool resample(const double* plainReading, int adaptParticles, const RangeReading* rr = 0);
    void updateTreeWeights(bool weightsAlreadyNormalized = false);
    void resetTree();
    double propagateWeights();
    void registerAll(const double* plainReading, const unsigned int* parents);
    void copyAll(const unsigned int* parents);
    void markMapsStale();
    void flushHostEdits();

    std::shared_ptr<b200::DeviceContext> m_device;
    int m_rank, m_world;
    unsigned char m_ncclId[128];
    unsigned int m_firstLocal, m_localCount;
    unsigned int m_treeEpoch;
    mutable unsigned int m_copyEpoch = 0;
    bool m_lazyTree = false;
    unsigned int m_maxParticles = 0, m_particleSlots = 0;  // setmaxParticles(); device slots for particle maps (set by init)
    bool m_earlyRegistration = true, m_registeredEarly = false;
    int m_fusedRegister = -1;  // -1: by default (see getregisterWithScanMatch)
    bool m_weightsChanged = true;  // a particle's log-weight changed since the last normalize()
    mutable bool m_treeDirty = false;
    double m_hostMs[4];
    double m_hostMsTotal[4] = {0, 0, 0, 0};
    std::vector<double> m_poseBuf, m_rowBuf, m_logwBuf;
    // the next reading's motion samples, drawn ahead (while the GPU registers the scan) on a look-ahead copy of the drand48
    // stream: b, r of the 3 N polar samples, sin / cos of every particle's heading.  Adopted by the next processScan only if
    // the C library's generator is still where the look-ahead started and no noise scale is zero (sampleGaussian(0) draws nothing)
    struct MotionDraws {
        bool valid = false;
        uint64_t before = 0, after = 0;
        std::vector<double> b, r, theta, sn, cs;
    } m_motionDraws;
    void prepareMotionDraws();
    TNodeVector m_newNodes;  // this scan's trajectory nodes, made while the GPU matches the scan (scanMatch -> resample)
};

typedef std::multimap<const GridSlamProcessor::TNode*, GridSlamProcessor::TNode*> TNodeMultimap;

}  // namespace GMapping
#endif
