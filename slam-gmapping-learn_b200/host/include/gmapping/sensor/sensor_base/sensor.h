// sensor.h -- named sensor base class and the name -> sensor registry (reference API: sensor/sensor_base/sensor.h)
#ifndef GMB200_SENSOR_H
#define GMB200_SENSOR_H
#include <gmapping/sensor/sensor_base/sensor_base_export.h>

#include <map>
#include <string>

namespace GMapping {

class SENSOR_BASE_EXPORT Sensor {
  public:
    Sensor(const std::string& name = "");
    virtual ~Sensor();
    inline std::string getName() const { return m_name; }
    inline void setName(const std::string& name) { m_name = name; }

  protected:
    std::string m_name;
};

typedef std::map<std::string, Sensor*> SensorMap;

}  // namespace GMapping
#endif
