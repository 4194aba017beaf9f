// rangereading.h -- one laser scan: a vector of ranges + the odometry pose it was taken at
// (reference API: sensor/sensor_range/rangereading.h:21-34)
#ifndef GMB200_RANGEREADING_H
#define GMB200_RANGEREADING_H
#include <gmapping/sensor/sensor_base/sensorreading.h>
#include <gmapping/sensor/sensor_range/rangesensor.h>
#include <gmapping/sensor/sensor_range/sensor_range_export.h>

#include <vector>

namespace GMapping {

class SENSOR_RANGE_EXPORT RangeReading : public SensorReading, public std::vector<double> {
  public:
    RangeReading(const RangeSensor* rs, double time = 0);
    RangeReading(unsigned int n_beams, const double* d, const RangeSensor* rs, double time = 0);
    virtual ~RangeReading();
    inline const OrientedPoint& getPose() const { return m_pose; }
    inline void setPose(const OrientedPoint& pose) { m_pose = pose; }
    // copy the ranges into v; with density > 0 beams closer than `density` to the last kept endpoint are suppressed (set to DBL_MAX)
    unsigned int rawView(double* v, double density = 0.) const;
    std::vector<Point> cartesianForm(double maxRange = 1e6) const;
    unsigned int activeBeams(double density = 0.) const;

  protected:
    OrientedPoint m_pose;
};

}  // namespace GMapping
#endif
