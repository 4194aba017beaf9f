// rangesensor.h -- planar range finder: a table of beams (reference API: sensor/sensor_range/rangesensor.h,
// beam table construction sensor/sensor_range/rangesensor.cpp:20-38)
#ifndef GMB200_RANGESENSOR_H
#define GMB200_RANGESENSOR_H
#include <gmapping/sensor/sensor_base/sensor.h>
#include <gmapping/sensor/sensor_range/sensor_range_export.h>
#include <gmapping/utils/point.h>

#include <vector>

namespace GMapping {

class SENSOR_RANGE_EXPORT RangeSensor : public Sensor {
    friend class Configuration;
    friend class CarmenConfiguration;
    friend class CarmenWrapper;

  public:
    struct Beam {
        OrientedPoint pose;  // relative to the sensor centre; theta is the beam angle
        double span;         // 0 = line-like beam
        double maxRange;
        double s, c;         // sin / cos of pose.theta
    };
    RangeSensor(std::string name);
    // `beams` beams, `res` radians apart, starting at -res*beams/2 and accumulated with += (not i*res)
    RangeSensor(std::string name, unsigned int beams, double res, const OrientedPoint& position = OrientedPoint(0, 0, 0),
                double span = 0, double maxrange = 89.0);
    inline const std::vector<Beam>& beams() const { return m_beams; }
    inline std::vector<Beam>& beams() { return m_beams; }
    inline OrientedPoint getPose() const { return m_pose; }
    void updateBeamsLookup();
    bool newFormat;

  protected:
    OrientedPoint m_pose;
    std::vector<Beam> m_beams;
};

}  // namespace GMapping
#endif
