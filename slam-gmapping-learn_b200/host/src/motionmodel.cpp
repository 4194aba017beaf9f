// motionmodel.cpp -- odometry motion model (reference: gridfastslam/motionmodel.cpp:13-88).  Host code: its draws
// consume the process-wide drand48() stream, particle by particle, and that order is part of the parity contract.
#include <gmapping/gridfastslam/motionmodel.h>
#include <gmapping/utils/stat.h>

#include <cmath>

// regularisation added to the diagonal of the Gaussian approximation (reference: motionmodel.cpp:5-6)
#define MotionModelConditioningLinearCovariance 0.01
#define MotionModelConditioningAngularCovariance 0.001

namespace GMapping {

OrientedPoint MotionModel::drawFromMotion(const OrientedPoint& p, double linearMove, double angularMove) const {
    OrientedPoint n(p);
    const double lm = linearMove + fabs(linearMove) * sampleGaussian(srr) + fabs(angularMove) * sampleGaussian(str);
    const double am = angularMove + fabs(linearMove) * sampleGaussian(srt) + fabs(angularMove) * sampleGaussian(stt);
    n.x += lm * cos(n.theta + .5 * am);
    n.y += lm * sin(n.theta + .5 * am);
    n.theta += am;
    n.theta = atan2(sin(n.theta), cos(n.theta));
    return n;
}

// odometry step pold -> pnew in the robot frame, perturbed, then applied to the particle pose p
OrientedPoint MotionModel::drawFromMotion(const OrientedPoint& p, const OrientedPoint& pnew, const OrientedPoint& pold) const {
    const double sxy = 0.3 * srr;
    const OrientedPoint delta = absoluteDifference(pnew, pold);
    OrientedPoint noisy(delta);
    noisy.x += sampleGaussian(srr * fabs(delta.x) + str * fabs(delta.theta) + sxy * fabs(delta.y));
    noisy.y += sampleGaussian(srr * fabs(delta.y) + str * fabs(delta.theta) + sxy * fabs(delta.x));
    noisy.theta += sampleGaussian(stt * fabs(delta.theta) + srt * sqrt(delta.x * delta.x + delta.y * delta.y));
    noisy.theta = fmod(noisy.theta, 2 * M_PI);
    if (noisy.theta > M_PI) noisy.theta -= 2 * M_PI;
    return absoluteSum(p, noisy);
}

Covariance3 MotionModel::gaussianApproximation(const OrientedPoint& pnew, const OrientedPoint& pold) const {
    const OrientedPoint delta = absoluteDifference(pnew, pold);
    const double linearMove = sqrt(delta.x * delta.x + delta.y * delta.y), angularMove = fabs(delta.x);
    const double s11 = srr * srr * linearMove * linearMove;
    const double s22 = stt * stt * angularMove * angularMove;
    const double s12 = str * angularMove * srt * linearMove;
    Covariance3 cov;
    const double s = sin(pold.theta), c = cos(pold.theta);
    cov.xx = c * c * s11 + MotionModelConditioningLinearCovariance;
    cov.yy = s * s * s11 + MotionModelConditioningLinearCovariance;
    cov.tt = s22 + MotionModelConditioningAngularCovariance;
    cov.xy = s * c * s11;
    cov.xt = c * s12;
    cov.yt = s * s12;
    return cov;
}

}  // namespace GMapping
