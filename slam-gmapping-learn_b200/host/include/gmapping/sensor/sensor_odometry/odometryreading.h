// odometryreading.h (reference API: sensor/sensor_odometry/odometryreading.h)
#ifndef GMB200_ODOMETRYREADING_H
#define GMB200_ODOMETRYREADING_H
#include <gmapping/sensor/sensor_base/sensorreading.h>
#include <gmapping/sensor/sensor_odometry/odometrysensor.h>
#include <gmapping/sensor/sensor_odometry/sensor_odometry_export.h>
#include <gmapping/utils/point.h>

namespace GMapping {

class SENSOR_ODOMETRY_EXPORT OdometryReading : public SensorReading {
  public:
    OdometryReading(const OdometrySensor* odo, double time = 0);
    inline const OrientedPoint& getPose() const { return m_pose; }
    inline const OrientedPoint& getSpeed() const { return m_speed; }
    inline const OrientedPoint& getAcceleration() const { return m_acceleration; }
    inline void setPose(const OrientedPoint& pose) { m_pose = pose; }
    inline void setSpeed(const OrientedPoint& speed) { m_speed = speed; }
    inline void setAcceleration(const OrientedPoint& acceleration) { m_acceleration = acceleration; }

  protected:
    OrientedPoint m_pose, m_speed, m_acceleration;
};

}  // namespace GMapping
#endif
