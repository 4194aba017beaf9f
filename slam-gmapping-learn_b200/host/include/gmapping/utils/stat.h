// stat.h -- sampling helpers (reference API: include/gmapping/utils/stat.h, utils/stat.cpp:31-65).
// The global drand48() stream and the order it is consumed in are part of the parity contract, so these stay
// plain host functions.
#ifndef GMB200_STAT_H
#define GMB200_STAT_H
#include <gmapping/utils/point.h>
#include <gmapping/utils/utils_export.h>

#include <vector>

namespace GMapping {

// zero-mean Gaussian sample (polar Box-Muller over drand48).  S != 0 re-seeds with srand48(S) first;
// sigma == 0 returns 0 without consuming a draw.
double UTILS_EXPORT sampleGaussian(double sigma, unsigned long int S = 0);
double UTILS_EXPORT evalGaussian(double sigmaSquare, double delta);
double UTILS_EXPORT evalLogGaussian(double sigmaSquare, double delta);
int UTILS_EXPORT sampleUniformInt(int max);
double UTILS_EXPORT sampleUniformDouble(double min, double max);

struct UTILS_EXPORT Covariance3 {
    Covariance3 operator+(const Covariance3& cov) const;
    static Covariance3 zero;
    double xx, yy, tt, xy, xt, yt;
};

}  // namespace GMapping
#endif
