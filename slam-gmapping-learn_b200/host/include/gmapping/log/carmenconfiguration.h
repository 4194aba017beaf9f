// carmenconfiguration.h -- PARAM lines of a CARMEN log -> sensors (reference API: include/gmapping/log/carmenconfiguration.h).
// Covers what the particle-update path consumes: ODOM / TRUEPOS odometry and the front laser (FLASER / ROBOTLASER1).
#ifndef GMB200_CARMENCONFIGURATION_H
#define GMB200_CARMENCONFIGURATION_H
#include <gmapping/log/configuration.h>
#include <gmapping/log/log_export.h>
#include <gmapping/sensor/sensor_base/sensor.h>

#include <istream>
#include <map>
#include <string>
#include <vector>

namespace GMapping {
class LOG_EXPORT CarmenConfiguration : public Configuration, public std::map<std::string, std::vector<std::string> > {
  public:
    virtual std::istream& load(std::istream& is);
    virtual SensorMap computeSensorMap() const;
};
}  // namespace GMapping
#endif
