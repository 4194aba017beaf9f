// scanmatcher.h -- GMapping::ScanMatcher with the reference's API (include/gmapping/scanmatcher/scanmatcher.h:16-86).
//
// The object is the parameter block (laser table, ranges, sigmas, kernel size, hill-climb steps).  Its methods do
// not compute on the host: the map argument is made device-resident (grid/map.h MapResidency) and the work runs in
// the sm_100a kernels behind the C ABI (include/gm_b200.h).  There is no CPU fallback; without a B200 the first
// call aborts with the library's error message, the way the reference aborts on its asserts.
#ifndef GMB200_SCANMATCHER_H
#define GMB200_SCANMATCHER_H
#include <gmapping/scanmatcher/scanmatcher_export.h>
#include <gmapping/scanmatcher/smmap.h>
#include <gmapping/utils/gvalues.h>
#include <gmapping/utils/macro_params.h>
#include <gmapping/utils/stat.h>

#include <iostream>
#define LASER_MAXBEAMS 2048

struct gm_match_params;

namespace GMapping {

class SCANMATCHER_EXPORT ScanMatcher {
  public:
    typedef Covariance3 CovarianceMatrix;

    ScanMatcher();
    ~ScanMatcher();

    // greedy hill-climb around p on `map`; returns the best score, pnew = the climbed pose (scanmatcher.cpp:386-529)
    double optimize(OrientedPoint& pnew, const ScanMatcherMap& map, const OrientedPoint& p, const double* readings) const;
    // sum over beams of exp(-|mu|^2 / gaussianSigma) for the nearest matching cell in the kernel window (scanmatcher.h:171-259)
    double score(const ScanMatcherMap& map, const OrientedPoint& p, const double* readings) const;
    // score s, log-likelihood l and the number of matched beams (scanmatcher.h:271-380)
    unsigned int likelihoodAndScore(double& s, double& l, const ScanMatcherMap& map, const OrientedPoint& p, const double* readings) const;
    // bounding box of the scan at p, Map::resize when it leaves the map (scanmatcher.cpp:83-254)
    void computeActiveArea(ScanMatcherMap& map, const OrientedPoint& p, const double* readings);
    void invalidateActiveArea();
    // ray-trace the scan into the map: visits++ along each beam, hit accumulation at its end (scanmatcher.cpp:266-354).
    // The reference returns an entropy delta that no caller on the particle-update path reads; this returns 0.
    double registerScan(ScanMatcherMap& map, const OrientedPoint& p, const double* readings);

    void setLaserParameters(unsigned int beams, double* angles, const OrientedPoint& lpose);
    void setMatchingParameters(double urange, double range, double sigma, int kernsize, double lopt, double aopt, int iterations,
                               double likelihoodSigma = 1, unsigned int likelihoodSkip = 0);

    // sampled variants (scanmatcher.cpp:541-690, :711-780): the pose samples are scored on the GPU in
    // batched likelihoodAndScore queries, the moment sums over the few dozen samples are formed on the host
    double optimize(OrientedPoint& mean, CovarianceMatrix& cov, const ScanMatcherMap& map, const OrientedPoint& p, const double* readings) const;
    double likelihood(double& lmax, OrientedPoint& mean, CovarianceMatrix& cov, const ScanMatcherMap& map, const OrientedPoint& p,
                      const double* readings);
    // ICP variants (scanmatcher.cpp:356-375, scanmatcher.h:88-165): host functions; the reference's pose correction is commented out, see scanmatcher.cpp
    double icpOptimize(OrientedPoint& pnew, const ScanMatcherMap& map, const OrientedPoint& p, const double* readings) const;
    double icpStep(OrientedPoint& pret, const ScanMatcherMap& map, const OrientedPoint& p, const double* readings) const;

    inline const double* laserAngles() const { return m_laserAngles; }
    inline unsigned int laserBeams() const { return m_laserBeams; }
    static const double nullLikelihood;

    // B200: the parameter block in the C ABI's layout
    void exportParams(gm_match_params* out) const;

  protected:
    bool m_activeAreaComputed;
    unsigned int m_laserBeams;
    double m_laserAngles[LASER_MAXBEAMS];
    PARAM_SET_GET(OrientedPoint, laserPose, protected, public, public)
    PARAM_SET_GET(double, laserMaxRange, protected, public, public)
    PARAM_SET_GET(double, usableRange, protected, public, public)
    PARAM_SET_GET(double, gaussianSigma, protected, public, public)
    PARAM_SET_GET(double, likelihoodSigma, protected, public, public)
    PARAM_SET_GET(int, kernelSize, protected, public, public)
    PARAM_SET_GET(double, optAngularDelta, protected, public, public)
    PARAM_SET_GET(double, optLinearDelta, protected, public, public)
    PARAM_SET_GET(unsigned int, optRecursiveIterations, protected, public, public)
    PARAM_SET_GET(unsigned int, likelihoodSkip, protected, public, public)
    PARAM_SET_GET(double, llsamplerange, protected, public, public)
    PARAM_SET_GET(double, llsamplestep, protected, public, public)
    PARAM_SET_GET(double, lasamplerange, protected, public, public)
    PARAM_SET_GET(double, lasamplestep, protected, public, public)
    PARAM_SET_GET(bool, generateMap, protected, public, public)
    PARAM_SET_GET(double, enlargeStep, protected, public, public)
    PARAM_SET_GET(double, fullnessThreshold, protected, public, public)
    PARAM_SET_GET(double, angularOdometryReliability, protected, public, public)
    PARAM_SET_GET(double, linearOdometryReliability, protected, public, public)
    PARAM_SET_GET(double, freeCellRatio, protected, public, public)
    PARAM_SET_GET(unsigned int, initialBeamsSkip, protected, public, public)
};

}  // namespace GMapping
#endif
