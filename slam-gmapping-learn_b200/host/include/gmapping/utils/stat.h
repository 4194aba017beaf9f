// stat.h -- sampling helpers (reference API: include/gmapping/utils/stat.h, utils/stat.cpp:31-65).
// The global drand48() stream and the order it is consumed in are part of the parity contract, so these stay
// plain host functions.
#ifndef GMB200_STAT_H
#define GMB200_STAT_H
#include <gmapping/utils/point.h>
#include <gmapping/utils/utils_export.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

namespace GMapping {

// B200 build: a window onto the process-wide drand48() stream for loops that draw thousands of samples.  It takes the
// generator's state out of libc (seed48 returns the previous state), steps the same 48-bit linear congruential sequence
// inline (X' = 0x5DEECE66D X + 0xB mod 2^48, value X' / 2^48: drand48's definition) and puts the state back when it goes
// out of scope, so the draws are the ones drand48() would have returned and every later drand48() continues from there.
// Nothing else may call the drand48 family while one is alive.
class Drand48Stream {
  public:
    Drand48Stream() {
        unsigned short zero[3] = {0, 0, 0};
        const unsigned short* prev = seed48(zero);
        m_x = (uint64_t)prev[0] | ((uint64_t)prev[1] << 16) | ((uint64_t)prev[2] << 32);
    }
    ~Drand48Stream() {
        unsigned short s[3] = {(unsigned short)(m_x & 0xffff), (unsigned short)((m_x >> 16) & 0xffff), (unsigned short)((m_x >> 32) & 0xffff)};
        seed48(s);
    }
    Drand48Stream(const Drand48Stream&) = delete;
    Drand48Stream& operator=(const Drand48Stream&) = delete;
    inline double next() {
        m_x = (m_x * 0x5DEECE66DULL + 0xBULL) & 0xFFFFFFFFFFFFULL;
        return (double)m_x * (1.0 / 281474976710656.0);  // 48 bits: exact in a double, and so is the scaling by 2^-48
    }
    // sampleGaussian(sigma) on this stream: same draws, same arithmetic (utils/stat.cpp:31-47 of the reference)
    inline double gaussian(double sigma) {
        if (sigma == 0) return 0;
        double a, b, w, u;
        do {
            do { u = next(); } while (u == 0.0);
            a = 2.0 * u - 1.0;
            do { u = next(); } while (u == 0.0);
            b = 2.0 * next() - 1.0;
            w = a * a + b * b;
        } while (w > 1.0 || w == 0.0);
        return sigma * b * std::sqrt(-2.0 * std::log(w) / w);
    }

  private:
    uint64_t m_x;
};

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
