// scanmatcherprocessor.cpp -- incremental scan matching against one device-resident map
// (reference behaviour: scanmatcher/scanmatcherprocessor.cpp:9-228)
#include <gmapping/scanmatcher/scanmatcherprocessor.h>

#include <cassert>
#include <cmath>
#include <iostream>
#include <vector>

namespace GMapping {

namespace {
const double kDefaultRegScore = 300.;  // register a scan when the match scores below this ...
const double kDefaultMaxMove = 1.;     // squared odometry jump (m^2) above which a reading is ignored
}  // namespace

ScanMatcherProcessor::ScanMatcherProcessor(const ScanMatcherMap& m)
    : useICP(false), m_map(m.getCenter(), m.getWorldSizeX(), m.getWorldSizeY(), m.getResolution()), m_pose(0, 0, 0), m_count(0), m_first(true),
      m_beams(0), m_regScore(kDefaultRegScore), m_critScore(.5 * kDefaultRegScore), m_maxMove(kDefaultMaxMove), m_computeCovariance(false) {}

// the reference passes `patchdelta`, not `delta`, as the map resolution (scanmatcherprocessor.cpp:25-27); kept
ScanMatcherProcessor::ScanMatcherProcessor(double xmin, double ymin, double xmax, double ymax, double delta, double patchdelta)
    : useICP(false), m_map(Point((xmax + xmin) * .5, (ymax + ymin) * .5), xmax - xmin, ymax - ymin, patchdelta), m_pose(0, 0, 0), m_count(0),
      m_first(true), m_beams(0), m_regScore(kDefaultRegScore), m_critScore(.5 * kDefaultRegScore), m_maxMove(kDefaultMaxMove),
      m_computeCovariance(false) {
    (void)delta;
}

ScanMatcherProcessor::~ScanMatcherProcessor() {}

void ScanMatcherProcessor::setSensorMap(const SensorMap& smap, std::string sensorName) {
    m_sensorMap = smap;
    SensorMap::const_iterator it = m_sensorMap.find(sensorName);
    assert(it != m_sensorMap.end());
    const RangeSensor* laser = dynamic_cast<const RangeSensor*>(it->second);
    assert(laser && laser->beams().size());
    m_beams = static_cast<unsigned int>(laser->beams().size());
    std::vector<double> angles(m_beams);
    for (unsigned int i = 0; i < m_beams; i++) angles[i] = laser->beams()[i].pose.theta;
    m_matcher.setLaserParameters(m_beams, angles.data(), laser->getPose());
}

void ScanMatcherProcessor::init() {
    m_first = true;
    m_pose = OrientedPoint(0, 0, 0);
    m_count = 0;
}

void ScanMatcherProcessor::processScan(const RangeReading& reading) {
    const OrientedPoint relPose = reading.getPose();
    if (!m_count) m_odoPose = relPose;

    // odometry increment, rotated from the odometry frame into the frame of the pose estimate
    const OrientedPoint move = relPose - m_odoPose;
    const double jump = move * move;
    if (jump > m_maxMove) {
        std::cerr << "Too big jump in the log file: " << jump << std::endl;
        std::cerr << "relPose=" << relPose.x << " " << relPose.y << std::endl;
        std::cerr << "ignoring" << std::endl;
        return;
    }
    const double dth = m_odoPose.theta - m_pose.theta;
    const double s = sin(dth), c = cos(dth);
    OrientedPoint step;
    step.x = c * move.x - s * move.y;
    step.y = s * move.x + c * move.y;
    step.theta = move.theta;
    m_pose = m_pose + step;
    m_pose.theta = atan2(sin(m_pose.theta), cos(m_pose.theta));
    m_odoPose = relPose;

    assert(reading.size() == m_beams);
    // beams whose end point is closer than one cell to the previous kept one are suppressed (set to DBL_MAX)
    std::vector<double> ranges(m_beams);
    reading.rawView(ranges.data(), m_map.getDelta());

    double score = 0;
    OrientedPoint matched = m_pose;
    if (m_count) {
        if (m_computeCovariance) {
            ScanMatcher::CovarianceMatrix cov;  // the reference decomposes it for a debug print only
            score = m_matcher.optimize(matched, cov, m_map, m_pose, ranges.data());
        } else if (useICP) {
            std::cerr << "USING ICP" << std::endl;
            score = m_matcher.icpOptimize(matched, m_map, m_pose, ranges.data());
        } else {
            score = m_matcher.optimize(matched, m_map, m_pose, ranges.data());
        }
    }
    if (!m_count || score < m_regScore) {
        m_matcher.invalidateActiveArea();
        m_matcher.registerScan(m_map, score < m_critScore ? m_pose : matched, ranges.data());
    }
    m_pose = matched;
    m_count++;
}

void ScanMatcherProcessor::setMatchingParameters(double urange, double range, double sigma, int kernsize, double lopt, double aopt, int iterations,
                                                 bool computeCovariance) {
    m_matcher.setMatchingParameters(urange, range, sigma, kernsize, lopt, aopt, iterations);
    m_computeCovariance = computeCovariance;
}

void ScanMatcherProcessor::setRegistrationParameters(double regScore, double critScore) {
    m_regScore = regScore;
    m_critScore = critScore;
}

OrientedPoint ScanMatcherProcessor::getPose() const { return m_pose; }

}  // namespace GMapping
