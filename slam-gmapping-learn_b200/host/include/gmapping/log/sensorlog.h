// sensorlog.h -- a whole log in memory (reference API: include/gmapping/log/sensorlog.h)
#ifndef GMB200_SENSORLOG_H
#define GMB200_SENSORLOG_H
#include <gmapping/log/log_export.h>
#include <gmapping/sensor/sensor_base/sensor.h>
#include <gmapping/sensor/sensor_base/sensorreading.h>
#include <gmapping/sensor/sensor_odometry/odometryreading.h>
#include <gmapping/sensor/sensor_range/rangereading.h>
#include <gmapping/utils/point.h>

#include <istream>
#include <list>

namespace GMapping {
class LOG_EXPORT SensorLog : public std::list<SensorReading*> {
  public:
    SensorLog(const SensorMap&);
    ~SensorLog();
    std::istream& load(std::istream& is);
    // bounding box of the scan endpoints at the odometry poses; returns the first pose
    OrientedPoint boundingBox(double& xmin, double& ymin, double& xmax, double& ymax) const;

  protected:
    const SensorMap& m_sensorMap;
};
}  // namespace GMapping
#endif
