// utils.cpp -- sampling helpers, sensors and readings (host-side plain data; reference: utils/stat.cpp,
// sensor/sensor_base/*.cpp, sensor/sensor_range/*.cpp, sensor/sensor_odometry/*.cpp)
#include <gmapping/sensor/sensor_odometry/odometryreading.h>
#include <gmapping/sensor/sensor_range/rangereading.h>
#include <gmapping/utils/stat.h>
#include <stdlib.h>

#include <cassert>
#include <cmath>
#include <limits>

namespace GMapping {

// polar Box-Muller over drand48().  The stream positions matter (stat.cpp:31-47): the first coordinate takes one
// non-zero draw, then one draw is discarded (its loop only rejects zeros) and the second coordinate takes a fresh one.
static double boxMullerDrand48(double sigma) {
    double a, b, w, u;
    do {
        do { u = drand48(); } while (u == 0.0);
        a = 2.0 * u - 1.0;
        do { u = drand48(); } while (u == 0.0);
        b = 2.0 * drand48() - 1.0;
        w = a * a + b * b;
    } while (w > 1.0 || w == 0.0);
    return sigma * b * sqrt(-2.0 * log(w) / w);
}

double sampleGaussian(double sigma, unsigned long int S) {
    if (S != 0) srand48(S);
    if (sigma == 0) return 0;
    return boxMullerDrand48(sigma);
}

double evalGaussian(double sigmaSquare, double delta) {
    if (sigmaSquare <= 0) sigmaSquare = 1e-4;
    return exp(-.5 * delta * delta / sigmaSquare) / sqrt(2 * M_PI * sigmaSquare);
}

double evalLogGaussian(double sigmaSquare, double delta) {
    if (sigmaSquare <= 0) sigmaSquare = 1e-4;
    return -.5 * delta * delta / sigmaSquare - .5 * log(2 * M_PI * sigmaSquare);
}

int sampleUniformInt(int max) { return (int)(max * drand48()); }
double sampleUniformDouble(double min, double max) { return min + (max - min) * drand48(); }

Covariance3 Covariance3::zero = {0., 0., 0., 0., 0., 0.};
Covariance3 Covariance3::operator+(const Covariance3& c) const {
    Covariance3 r(*this);
    r.xx += c.xx; r.yy += c.yy; r.tt += c.tt; r.xy += c.xy; r.yt += c.yt; r.xt += c.xt;
    return r;
}

Sensor::Sensor(const std::string& name) : m_name(name) {}
Sensor::~Sensor() {}
SensorReading::SensorReading(const Sensor* s, double t) : m_time(t), m_sensor(s) {}
SensorReading::~SensorReading() {}

OdometrySensor::OdometrySensor(const std::string& name, bool ideal) : Sensor(name), m_ideal(ideal) {}
OdometryReading::OdometryReading(const OdometrySensor* odo, double time) : SensorReading(odo, time) {}

RangeSensor::RangeSensor(std::string name) : Sensor(name), newFormat(false) {}
RangeSensor::RangeSensor(std::string name, unsigned int beams_num, double res, const OrientedPoint& position, double span, double maxrange)
    : Sensor(name), newFormat(false), m_beams(beams_num), m_pose(position) {
    double angle = -.5 * res * beams_num;  // accumulated, not i * res: the table's rounding is part of the matcher's input
    for (unsigned int i = 0; i < beams_num; i++, angle += res) {
        Beam& b = m_beams[i];
        b.span = span;
        b.pose = OrientedPoint(0, 0, angle);
        b.maxRange = maxrange;
    }
    updateBeamsLookup();
}
void RangeSensor::updateBeamsLookup() {
    for (size_t i = 0; i < m_beams.size(); i++) {
        m_beams[i].s = sin(m_beams[i].pose.theta);
        m_beams[i].c = cos(m_beams[i].pose.theta);
    }
}

RangeReading::RangeReading(const RangeSensor* rs, double time) : SensorReading(rs, time) {}
RangeReading::RangeReading(unsigned int n_beams, const double* d, const RangeSensor* rs, double time) : SensorReading(rs, time) {
    assert(n_beams == rs->beams().size());
    assign(d, d + n_beams);
}
RangeReading::~RangeReading() {}

// shared walk of rawView()/activeBeams(): a beam is "active" when its endpoint is at least `density` away from
// the last active one
template <class F>
static void walkBeams(const RangeReading& r, double density, F visit) {
    const RangeSensor* rs = dynamic_cast<const RangeSensor*>(r.getSensor());
    assert(rs);
    Point last(0, 0);
    for (unsigned int i = 0; i < r.size(); i++) {
        const Point lp(cos(rs->beams()[i].pose.theta) * r[i], sin(rs->beams()[i].pose.theta) * r[i]);
        const Point d = last - lp;
        const bool active = !(sqrt(d * d) < density);
        if (active) last = lp;
        visit(i, active);
    }
}

unsigned int RangeReading::rawView(double* v, double density) const {
    if (density == 0) {
        for (unsigned int i = 0; i < size(); i++) v[i] = (*this)[i];
    } else {
        walkBeams(*this, density, [&](unsigned int i, bool active) { v[i] = active ? (*this)[i] : std::numeric_limits<double>::max(); });
    }
    return static_cast<unsigned int>(size());
}

unsigned int RangeReading::activeBeams(double density) const {
    if (density == 0.) return size();
    unsigned int n = 0;
    walkBeams(*this, density, [&](unsigned int, bool active) { n += active ? 1 : 0; });
    return n;
}

std::vector<Point> RangeReading::cartesianForm(double maxRange) const {
    const RangeSensor* rs = dynamic_cast<const RangeSensor*>(getSensor());
    assert(rs && rs->beams().size());
    const unsigned int n = rs->beams().size();
    std::vector<Point> out(n);
    for (unsigned int i = 0; i < n; i++) {
        const double rho = (*this)[i], theta = rs->beams()[i].pose.theta;
        out[i] = rho > maxRange ? Point(0, 0) : Point(rho * cos(theta), rho * sin(theta));
    }
    return out;
}

}  // namespace GMapping
