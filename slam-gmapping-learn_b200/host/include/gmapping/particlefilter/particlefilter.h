// particlefilter.h -- the part of the reference's particle-filter templates that the SLAM processor uses
// (include/gmapping/particlefilter/particlefilter.h:180-229, 274-286): systematic ("uniform") resampling and Neff.
// GridSlamProcessor itself runs these on the GPU (gm_resample_indexes / gm_normalize_neff); the templates below
// are the generic host versions kept for API users that instantiate them on their own particle types.
#ifndef GMB200_PARTICLEFILTER_H
#define GMB200_PARTICLEFILTER_H
#include <stdlib.h>

#include <utility>
#include <vector>

template <class Particle, class Numeric>
struct uniform_resampler {
    // systematic resampling: one uniform draw, n equally spaced targets walked over the cumulative weights
    std::vector<unsigned int> resampleIndexes(const std::vector<Particle>& particles, int nparticles = 0) const {
        Numeric total = 0;
        for (typename std::vector<Particle>::const_iterator it = particles.begin(); it != particles.end(); ++it) total += (Numeric)*it;
        unsigned int n = particles.size();
        if (nparticles > 0) n = nparticles;
        const Numeric step = total / n;
        Numeric target = step * ::drand48();
        std::vector<unsigned int> indexes(n);
        Numeric running = 0;
        unsigned int i = 0, k = 0;
        for (typename std::vector<Particle>::const_iterator it = particles.begin(); it != particles.end(); ++it, ++i) {
            running += (Numeric)*it;
            while (running > target) {
                if (k < n) indexes[k] = i;  // the reference writes past the end here when rounding emits n + 1 indexes
                k++;
                target += step;
            }
        }
        return indexes;
    }
    std::vector<Particle> resample(const std::vector<Particle>& particles, int nparticles = 0) const {
        const std::vector<unsigned int> idx = resampleIndexes(particles, nparticles);
        std::vector<Particle> out;
        out.reserve(idx.size());
        for (size_t k = 0; k < idx.size(); k++) out.push_back(particles[idx[k]]);
        return out;
    }
    Numeric neff(const std::vector<Particle>& particles) const {
        double total = 0, sq = 0;
        for (typename std::vector<Particle>::const_iterator it = particles.begin(); it != particles.end(); ++it) {
            const Numeric w = (Numeric)*it;
            total += w;
            sq += w * w;
        }
        return total * total / sq;
    }
};

template <class WeightVector>
double neff(WeightVector begin, WeightVector end) {
    double sum = 0;
    for (WeightVector it = begin; it != end; ++it) sum += *it;
    double sq = 0;
    for (WeightVector it = begin; it != end; ++it) {
        const double w = *it / sum;
        sq += w * w;
    }
    return 1. / sq;
}
#endif
