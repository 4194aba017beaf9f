// scanmatcher.cpp -- GMapping::ScanMatcher over the C ABI.  Every method that touches a map makes it device-resident
// and launches the CUDA path; nothing here walks cells on the host.
#include <gmapping/scanmatcher/scanmatcher.h>

#include <cassert>
#include <cmath>
#include <cstring>
#include <iostream>
#include <vector>

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
    // the return value (scanmatcher.cpp:317-341): the summed entropy change of every cell update, only formed when free space
    // is traced; it telescopes per cell, so it is the map's entropy figure after the registration minus before
    double before = 0, after = 0;
    if (m_generateMap) r->ctx->check(gm_map_entropy(r->ctx->h, 0, 1, &before), "gm_map_entropy");
    r->ctx->check(gm_register(r->ctx->h, readings, pose, nullptr), "gm_register");
    if (m_generateMap) r->ctx->check(gm_map_entropy(r->ctx->h, 0, 1, &after), "gm_map_entropy");
    r->stale = true;
    m_activeAreaComputed = true;
    return after - before;
}

// ------------------------------------------------------------------------------------------ sampled variants
// (reference: scanmatcher/scanmatcher.cpp:541-690 and :711-780).  The pose samples are evaluated on the GPU in one
// batched likelihoodAndScore query per step; only the few-dozen-term moment sums are formed here.
namespace {

struct Sample { OrientedPoint pose; double score, likelihood; };

// likelihoodAndScore of every pose on `map`, one launch
void evaluate(const ScanMatcher& m, const ScanMatcherMap& map, const double* readings, std::vector<Sample>& samples, size_t first) {
    b200::DeviceMapResidency* r = b200::makeResident(map, m);
    const size_t n = samples.size() - first;
    if (!n) return;
    std::vector<uint32_t> idx(n, r->slot), c(n);
    std::vector<double> poses(3 * n), s(n), l(n);
    for (size_t i = 0; i < n; i++) {
        const OrientedPoint& q = samples[first + i].pose;
        poses[3 * i] = q.x; poses[3 * i + 1] = q.y; poses[3 * i + 2] = q.theta;
    }
    r->ctx->check(gm_likelihood_and_score(r->ctx->h, (uint32_t)n, idx.data(), poses.data(), readings, s.data(), l.data(), c.data()),
                  "gm_likelihood_and_score");
    for (size_t i = 0; i < n; i++) { samples[first + i].score = s[i]; samples[first + i].likelihood = l[i]; }
}

// second moments of the samples about `mean`, weighted by Sample::likelihood (already exponentiated), divided by `norm`
ScanMatcher::CovarianceMatrix spread(const std::vector<Sample>& samples, const OrientedPoint& mean, double norm) {
    ScanMatcher::CovarianceMatrix cov = {0., 0., 0., 0., 0., 0.};
    for (size_t i = 0; i < samples.size(); i++) {
        OrientedPoint d = samples[i].pose - mean;
        d.theta = atan2(sin(d.theta), cos(d.theta));
        const double w = samples[i].likelihood;
        cov.xx += d.x * d.x * w; cov.yy += d.y * d.y * w; cov.tt += d.theta * d.theta * w;
        cov.xy += d.x * d.y * w; cov.xt += d.x * d.theta * w; cov.yt += d.y * d.theta * w;
    }
    cov.xx /= norm, cov.xy /= norm, cov.xt /= norm, cov.yy /= norm, cov.yt /= norm, cov.tt /= norm;
    return cov;
}

}  // namespace

// hill-climb like optimize(pnew, ...), but driven by likelihoodAndScore's score, remembering every evaluated pose;
// returns the best score, _mean = the climbed pose, _cov = likelihood-weighted spread of the evaluated poses
double ScanMatcher::optimize(OrientedPoint& _mean, CovarianceMatrix& _cov, const ScanMatcherMap& map, const OrientedPoint& init,
                             const double* readings) const {
    std::vector<Sample> moves(1);
    moves[0].pose = init;
    evaluate(*this, map, readings, moves, 0);
    OrientedPoint current = init;
    double bestScore = -1, currentScore = moves[0].score;
    double adelta = m_optAngularDelta, ldelta = m_optLinearDelta;
    unsigned int refinement = 0;
    do {
        if (bestScore >= currentScore) { refinement++; adelta *= .5; ldelta *= .5; }
        bestScore = currentScore;
        const size_t first = moves.size();
        for (int k = 0; k < 6; k++) {  // Front, Back, Left, Right, TurnLeft, TurnRight
            Sample c;
            c.pose = current; c.score = c.likelihood = 0;
            switch (k) {
                case 0: c.pose.x += ldelta; break;
                case 1: c.pose.x -= ldelta; break;
                case 2: c.pose.y -= ldelta; break;
                case 3: c.pose.y += ldelta; break;
                case 4: c.pose.theta += adelta; break;
                default: c.pose.theta -= adelta; break;
            }
            moves.push_back(c);
        }
        evaluate(*this, map, readings, moves, first);  // the six candidates in one launch
        OrientedPoint bestLocal = current;
        for (size_t i = first; i < moves.size(); i++)  // the score that drives this variant is likelihoodAndScore's, undamped
            if (moves[i].score > currentScore) { currentScore = moves[i].score; bestLocal = moves[i].pose; }
        current = bestLocal;
    } while (currentScore > bestScore || refinement < m_optRecursiveIterations);

    double lmax = -1e9;
    for (size_t i = 0; i < moves.size(); i++) lmax = moves[i].likelihood > lmax ? moves[i].likelihood : lmax;
    double lacc = 0;
    OrientedPoint mean(0, 0, 0);
    for (size_t i = 0; i < moves.size(); i++) {
        moves[i].likelihood = exp(moves[i].likelihood - lmax);
        mean = mean + moves[i].pose * moves[i].likelihood;
        lacc += moves[i].likelihood;
    }
    mean = mean * (1. / lacc);
    _cov = spread(moves, mean, lacc);
    _mean = current;
    return bestScore;
}

// likelihood of the scan integrated over a regular grid of poses around p (+-llsamplerange in x / y, +-lasamplerange in
// theta); returns log(sum exp(l)), _lmax = the largest log-likelihood, _mean / _cov = moments of the sampled poses
double ScanMatcher::likelihood(double& _lmax, OrientedPoint& _mean, CovarianceMatrix& _cov, const ScanMatcherMap& map, const OrientedPoint& p,
                               const double* readings) {
    std::vector<Sample> samples;
    for (double xx = -m_llsamplerange; xx <= m_llsamplerange; xx += m_llsamplestep)
        for (double yy = -m_llsamplerange; yy <= m_llsamplerange; yy += m_llsamplestep)
            for (double tt = -m_lasamplerange; tt <= m_lasamplerange; tt += m_lasamplestep) {
                Sample q;
                q.pose = OrientedPoint(p.x + xx, p.y + yy, p.theta + tt);
                q.score = q.likelihood = 0;
                samples.push_back(q);
            }
    evaluate(*this, map, readings, samples, 0);
    double lmax = -1e9;
    for (size_t i = 0; i < samples.size(); i++) lmax = samples[i].likelihood > lmax ? samples[i].likelihood : lmax;
    double lcum = 0, s = 0, c = 0;
    OrientedPoint mean(0, 0, 0);
    for (size_t i = 0; i < samples.size(); i++) {
        samples[i].likelihood = exp(samples[i].likelihood - lmax);
        lcum += samples[i].likelihood;
    }
    for (size_t i = 0; i < samples.size(); i++) {
        mean = mean + samples[i].pose * samples[i].likelihood;
        s += samples[i].likelihood * sin(samples[i].pose.theta);
        c += samples[i].likelihood * cos(samples[i].pose.theta);
    }
    mean = mean * (1. / lcum);
    s /= lcum;
    c /= lcum;
    mean.theta = atan2(s, c);  // circular mean for the heading
    _cov = spread(samples, mean, lcum);
    _mean = mean;
    _lmax = lmax;
    return log(lcum) + lmax;
}

// ICP variants (include/gmapping/scanmatcher/scanmatcher.h:88-165, scanmatcher/scanmatcher.cpp:356-375): outside the per-scan
// particle update, host functions over the map's mirror.  In the reference the call that would turn the
// point pairs into a pose correction (icpNonlinearStep) is commented out: icpStep pairs every selected beam's end point with
// the nearest occupied cell mean of its window, reports how many pairs it found on stderr, returns the pose it was given
// (heading wrapped) and score(map, p, readings).  icpOptimize therefore stops after one step with score(map, init, readings).
double ScanMatcher::icpStep(OrientedPoint& pret, const ScanMatcherMap& map, const OrientedPoint& p, const double* readings) const {
    OrientedPoint lp = p;
    lp.x += cos(p.theta) * m_laserPose.x - sin(p.theta) * m_laserPose.y;
    lp.y += sin(p.theta) * m_laserPose.x + cos(p.theta) * m_laserPose.y;
    lp.theta += m_laserPose.theta;
    // (the reference scales the free-cell offset by the resolution twice here: delta * (delta * freeCellRatio))
    const double freeOffset = map.getDelta() * (map.getDelta() * m_freeCellRatio);
    unsigned int skip = 0, npairs = 0;
    for (unsigned int b = m_initialBeamsSkip; b < m_laserBeams; b++) {
        const double r = readings[b];
        skip++;
        skip = skip > m_likelihoodSkip ? 0 : skip;
        if (r > m_usableRange || r == 0.0) continue;
        if (skip) continue;
        const double dirx = cos(lp.theta + m_laserAngles[b]), diry = sin(lp.theta + m_laserAngles[b]);
        const Point phit(lp.x + r * dirx, lp.y + r * diry);
        const IntPoint iphit = map.world2map(phit);
        const Point pfree = Point(lp.x + (r - freeOffset) * dirx, lp.y + (r - freeOffset) * diry) - phit;
        const IntPoint ipfree = map.world2map(pfree);
        bool found = false;
        for (int xx = -m_kernelSize; xx <= m_kernelSize && !found; xx++)
            for (int yy = -m_kernelSize; yy <= m_kernelSize; yy++) {
                const IntPoint pr = iphit + IntPoint(xx, yy), pf = pr + ipfree;
                const PointAccumulator& cell = map.cell(pr);
                const PointAccumulator& fcell = map.cell(pf);
                if (((double)cell) > m_fullnessThreshold && ((double)fcell) < m_fullnessThreshold) { found = true; break; }  // which cell wins does not matter: only the count is used
            }
        if (found) npairs++;
    }
    const OrientedPoint result(0, 0, 0);
    std::cerr << "result(" << npairs << ")=" << result.x << " " << result.y << " " << result.theta << std::endl;
    pret.x = p.x + result.x;
    pret.y = p.y + result.y;
    pret.theta = p.theta + result.theta;
    pret.theta = atan2(sin(pret.theta), cos(pret.theta));
    return score(map, p, readings);
}

double ScanMatcher::icpOptimize(OrientedPoint& pnew, const ScanMatcherMap& map, const OrientedPoint& init, const double* readings) const {
    double currentScore;
    double sc = score(map, init, readings);
    OrientedPoint start = init;
    pnew = init;
    int iterations = 0;
    do {
        currentScore = sc;
        sc = icpStep(pnew, map, start, readings);
        start = pnew;
        iterations++;
    } while (sc > currentScore);
    std::cerr << "i=" << iterations << std::endl;
    return currentScore;
}

}  // namespace GMapping
