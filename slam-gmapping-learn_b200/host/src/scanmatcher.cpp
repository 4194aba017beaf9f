// scanmatcher.cpp -- GMapping::ScanMatcher over the C ABI.  Every method that touches a map makes it device-resident
// and launches the CUDA path; nothing here walks cells on the host.
#include <gmapping/scanmatcher/scanmatcher.h>

#include <cassert>
#include <cmath>
#include <cstring>

#include "devicemap.h"

namespace GMapping {

const double ScanMatcher::nullLikelihood = -.5;  // reference: scanmatcher/scanmatcher.cpp:15

PointAccumulator* PointAccumulator::unknown_ptr = 0;
const PointAccumulator& PointAccumulator::Unknown() {
    if (!unknown_ptr) unknown_ptr = new PointAccumulator;
    return *unknown_ptr;
}

// defaults of the reference constructor (scanmatcher/scanmatcher.cpp:17-56); it leaves m_generateMap unset,
// every driver sets it -- here it starts false like a zeroed object would
ScanMatcher::ScanMatcher() : m_activeAreaComputed(false), m_laserBeams(0), m_laserPose(0, 0, 0) {
    std::memset(m_laserAngles, 0, sizeof m_laserAngles);
    m_laserMaxRange = 80;
    m_usableRange = 80;
    m_gaussianSigma = 1;
    m_likelihoodSigma = 1;
    m_kernelSize = 1;
    m_optAngularDelta = 0.1;
    m_optLinearDelta = 0.05;
    m_optRecursiveIterations = 3;
    m_likelihoodSkip = 0;
    m_llsamplerange = 0.01;
    m_llsamplestep = 0.01;
    m_lasamplerange = 0.005;
    m_lasamplestep = 0.005;
    m_generateMap = false;
    m_enlargeStep = 10.;
    m_fullnessThreshold = 0.1;
    m_angularOdometryReliability = 0.;
    m_linearOdometryReliability = 0.;
    m_freeCellRatio = sqrt(2.);
    m_initialBeamsSkip = 0;
}

ScanMatcher::~ScanMatcher() {}

void ScanMatcher::setLaserParameters(unsigned int beams, double* angles, const OrientedPoint& lpose) {
    assert(beams < LASER_MAXBEAMS);  // scanmatcher.cpp:704
    m_laserPose = lpose;
    m_laserBeams = beams;
    std::memcpy(m_laserAngles, angles, sizeof(double) * beams);
}

void ScanMatcher::setMatchingParameters(double urange, double range, double sigma, int kernsize, double lopt, double aopt, int iterations,
                                        double likelihoodSigma, unsigned int likelihoodSkip) {
    m_usableRange = urange;
    m_laserMaxRange = range;
    m_kernelSize = kernsize;
    m_optLinearDelta = lopt;
    m_optAngularDelta = aopt;
    m_optRecursiveIterations = iterations;
    m_gaussianSigma = sigma;
    m_likelihoodSigma = likelihoodSigma;
    m_likelihoodSkip = likelihoodSkip;
}

void ScanMatcher::exportParams(gm_match_params* p) const {
    std::memset(p, 0, sizeof *p);
    p->laser_max_range = m_laserMaxRange; p->usable_range = m_usableRange;
    p->gaussian_sigma = m_gaussianSigma; p->likelihood_sigma = m_likelihoodSigma;
    p->kernel_size = m_kernelSize; p->generate_map = m_generateMap ? 1 : 0;
    p->opt_linear_delta = m_optLinearDelta; p->opt_angular_delta = m_optAngularDelta;
    p->opt_recursive_iterations = m_optRecursiveIterations; p->likelihood_skip = m_likelihoodSkip;
    p->enlarge_step = m_enlargeStep; p->fullness_threshold = m_fullnessThreshold;
    p->angular_odometry_reliability = m_angularOdometryReliability; p->linear_odometry_reliability = m_linearOdometryReliability;
    p->free_cell_ratio = m_freeCellRatio; p->initial_beams_skip = m_initialBeamsSkip;
}

void ScanMatcher::invalidateActiveArea() { m_activeAreaComputed = false; }

double ScanMatcher::score(const ScanMatcherMap& map, const OrientedPoint& p, const double* readings) const {
    b200::DeviceMapResidency* r = b200::makeResident(map, *this);
    const uint32_t idx = r->slot;
    const double pose[3] = {p.x, p.y, p.theta};
    double s = 0;
    r->ctx->check(gm_score(r->ctx->h, 1, &idx, pose, readings, &s), "gm_score");
    return s;
}

unsigned int ScanMatcher::likelihoodAndScore(double& s, double& l, const ScanMatcherMap& map, const OrientedPoint& p, const double* readings) const {
    b200::DeviceMapResidency* r = b200::makeResident(map, *this);
    const uint32_t idx = r->slot;
    const double pose[3] = {p.x, p.y, p.theta};
    uint32_t c = 0;
    r->ctx->check(gm_likelihood_and_score(r->ctx->h, 1, &idx, pose, readings, &s, &l, &c), "gm_likelihood_and_score");
    return c;
}

double ScanMatcher::optimize(OrientedPoint& pnew, const ScanMatcherMap& map, const OrientedPoint& p, const double* readings) const {
    b200::DeviceMapResidency* r = b200::makeResident(map, *this);
    const uint32_t idx = r->slot;
    const double pose[3] = {p.x, p.y, p.theta};
    double out[3], sc = 0;
    r->ctx->check(gm_optimize(r->ctx->h, 1, &idx, pose, readings, out, &sc), "gm_optimize");
    pnew = OrientedPoint(out[0], out[1], out[2]);
    return sc;
}

void ScanMatcher::computeActiveArea(ScanMatcherMap& map, const OrientedPoint& p, const double* readings) {
    if (m_activeAreaComputed) return;
    b200::DeviceMapResidency* r = b200::makeResident(map, *this);
    if (r->ctx->particles == 1) {
        const double pose[3] = {p.x, p.y, p.theta};
        r->ctx->check(gm_compute_active_area(r->ctx->h, readings, pose), "gm_compute_active_area");
        r->stale = true;
    }
    // for a map inside a particle set the resize is carried out by the filter's own registration step
    m_activeAreaComputed = true;
}

double ScanMatcher::registerScan(ScanMatcherMap& map, const OrientedPoint& p, const double* readings) {
    b200::DeviceMapResidency* r = b200::makeResident(map, *this);
    if (r->ctx->particles != 1)
        b200::fatal("ScanMatcher::registerScan on one map of a device-resident particle set: register through GridSlamProcessor "
                    "(gm_register updates all particles in one launch)");
    const double pose[3] = {p.x, p.y, p.theta};
    r->ctx->check(gm_register(r->ctx->h, readings, pose, nullptr), "gm_register");
    r->stale = true;
    m_activeAreaComputed = true;
    return 0;
}

static double notOnPath(const char* what) {
    b200::fatal(std::string(what) + " is not part of the per-scan particle update and is not implemented in this build (SURVEY.md §8 f4)");
}
double ScanMatcher::icpOptimize(OrientedPoint&, const ScanMatcherMap&, const OrientedPoint&, const double*) const { return notOnPath("ScanMatcher::icpOptimize"); }
double ScanMatcher::optimize(OrientedPoint&, CovarianceMatrix&, const ScanMatcherMap&, const OrientedPoint&, const double*) const {
    return notOnPath("ScanMatcher::optimize(mean, cov, ...)");
}
double ScanMatcher::icpStep(OrientedPoint&, const ScanMatcherMap&, const OrientedPoint&, const double*) const { return notOnPath("ScanMatcher::icpStep"); }
double ScanMatcher::likelihood(double&, OrientedPoint&, CovarianceMatrix&, const ScanMatcherMap&, const OrientedPoint&, const double*) {
    return notOnPath("ScanMatcher::likelihood");
}

}  // namespace GMapping
