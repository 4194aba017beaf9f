// scanmatcherprocessor.h -- GMapping::ScanMatcherProcessor with the reference's API
// (include/gmapping/scanmatcher/scanmatcherprocessor.h): incremental scan matching against ONE map, no particle filter
// (SURVEY.md §8 row f4).  The map is an ordinary ScanMatcherMap; the first matcher call makes it device-resident
// (single-map context), so optimize() and registerScan() run in the same sm_100a kernels as the particle filter's.
#ifndef GMB200_SCANMATCHERPROCESSOR_H
#define GMB200_SCANMATCHERPROCESSOR_H
#include <gmapping/log/sensorlog.h>
#include <gmapping/scanmatcher/scanmatcher.h>
#include <gmapping/scanmatcher/scanmatcher_export.h>
#include <gmapping/sensor/sensor_range/rangereading.h>
#include <gmapping/sensor/sensor_range/rangesensor.h>

#include <string>

namespace GMapping {

class SCANMATCHER_EXPORT ScanMatcherProcessor {
  public:
    // ---- construction: an empty map with the frame (centre, world size, resolution) of `m`, or from explicit bounds
    ScanMatcherProcessor(const ScanMatcherMap& m);
    ScanMatcherProcessor(double xmin, double ymin, double xmax, double ymax, double delta, double patchdelta);
    virtual ~ScanMatcherProcessor();

    // ---- configuration
    void setSensorMap(const SensorMap& smap, std::string sensorName = "FLASER");
    void setMatchingParameters(double urange, double range, double sigma, int kernsize, double lopt, double aopt, int iterations,
                               bool computeCovariance = false);
    void setRegistrationParameters(double regScore, double critScore);
    void setmaxMove(double mmove) { m_maxMove = mmove; }
    ScanMatcher& matcher() { return m_matcher; }
    bool useICP;  // icpOptimize instead of the hill-climb

    // ---- operation
    void init();
    // odometry step since the last reading rotated into the map frame -> hill-climb from there -> register the scan when the
    // match is poor enough to need it (score < regScore), at the odometry pose when it is worse than critScore
    virtual void processScan(const RangeReading& reading);
    OrientedPoint getPose() const;
    const ScanMatcherMap& getMap() const { return m_map; }

  protected:
    // what is being estimated
    ScanMatcherMap m_map;
    OrientedPoint m_pose, m_odoPose;  // pose estimate; odometry pose of the previous reading
    int m_count;                      // readings integrated so far
    bool m_first;
    // how
    ScanMatcher m_matcher;
    SensorMap m_sensorMap;
    unsigned int m_beams;
    double m_regScore, m_critScore, m_maxMove;
    bool m_computeCovariance;
};

}  // namespace GMapping
#endif
