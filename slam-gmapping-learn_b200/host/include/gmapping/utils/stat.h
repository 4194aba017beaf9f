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
    Drand48Stream() : m_x(take()), m_detached(false) {}
    // a look-ahead copy starting at `state`: the C library's generator is left alone (GridSlamProcessor draws the next
    // reading's motion samples ahead of time and adopts them only if nothing else has drawn in between)
    explicit Drand48Stream(uint64_t state) : m_x(state), m_detached(true) {}
    ~Drand48Stream() { if (!m_detached) put(m_x); }
    // the C library generator's 48-bit state: read (and leave unchanged) / replace
    static uint64_t peek() { const uint64_t x = take(); put(x); return x; }
    static void put(uint64_t x) {
        unsigned short s[3] = {(unsigned short)(x & 0xffff), (unsigned short)((x >> 16) & 0xffff), (unsigned short)((x >> 32) & 0xffff)};
        seed48(s);
    }
    uint64_t state() const { return m_x; }
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
    // the sigma-independent part of gaussian(): the same draws, the sample is then sigma * b * r
    inline void polar(double& b_out, double& r_out) {
        double a, b, w, u;
        do {
            do { u = next(); } while (u == 0.0);
            a = 2.0 * u - 1.0;
            do { u = next(); } while (u == 0.0);
            b = 2.0 * next() - 1.0;
            w = a * a + b * b;
        } while (w > 1.0 || w == 0.0);
        b_out = b; r_out = std::sqrt(-2.0 * std::log(w) / w);
    }

  private:
    static uint64_t take() {  // reads the state; the generator is left zeroed until put()
        unsigned short zero[3] = {0, 0, 0};
        const unsigned short* prev = seed48(zero);
        return (uint64_t)prev[0] | ((uint64_t)prev[1] << 16) | ((uint64_t)prev[2] << 32);
    }
    uint64_t m_x;
    bool m_detached;
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
