// motionmodel.h -- odometry motion model (reference API: gridfastslam/motionmodel.h, motionmodel.cpp:40-60).
// Stays on the host: its Gaussian draws consume the global drand48() stream in particle order.
#ifndef GMB200_MOTIONMODEL_H
#define GMB200_MOTIONMODEL_H
#include <gmapping/utils/point.h>
#include <gmapping/utils/stat.h>

namespace GMapping {

struct MotionModel {
    OrientedPoint drawFromMotion(const OrientedPoint& p, double linearMove, double angularMove) const;
    OrientedPoint drawFromMotion(const OrientedPoint& p, const OrientedPoint& pnew, const OrientedPoint& pold) const;
    Covariance3 gaussianApproximation(const OrientedPoint& pnew, const OrientedPoint& pold) const;
    double srr, str, srt, stt;
};

}  // namespace GMapping
#endif
