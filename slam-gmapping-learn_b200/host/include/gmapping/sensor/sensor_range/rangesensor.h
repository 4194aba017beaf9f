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
  public:
    // one beam: where it leaves the sensor (pose.theta is the beam angle), its opening (0 = a line), its reach,
    // and the cached sine / cosine of the angle
    struct Beam {
        OrientedPoint pose;
        double span;
        double maxRange;
        double s, c;
    };
    typedef std::vector<Beam> BeamVector;

    // a sensor whose beam table is filled in by a configuration reader (the friends below)
    explicit RangeSensor(std::string name);
    // `beams` beams, `res` radians apart: the first at -res*beams/2, each next one by += res (an accumulated sum, not
    // i*res -- the table's rounding is part of what the scan matcher sees)
    RangeSensor(std::string name, unsigned int beams, double res, const OrientedPoint& position = OrientedPoint(0, 0, 0),
                double span = 0, double maxrange = 89.0);

    // recompute Beam::s / Beam::c from the beam angles
    void updateBeamsLookup();

    inline OrientedPoint getPose() const { return m_pose; }
    inline BeamVector& beams() { return m_beams; }
    inline const BeamVector& beams() const { return m_beams; }

    bool newFormat;  // CARMEN ROBOTLASERx message layout instead of FLASER/RLASER

  protected:
    BeamVector m_beams;
    OrientedPoint m_pose;

    friend class CarmenWrapper;
    friend class CarmenConfiguration;
    friend class Configuration;
};

}  // namespace GMapping
#endif
