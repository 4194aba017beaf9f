// sensorstream.h -- reading-by-reading access to a CARMEN log (reference API: include/gmapping/log/sensorstream.h)
#ifndef GMB200_SENSORSTREAM_H
#define GMB200_SENSORSTREAM_H
#include <gmapping/log/log_export.h>
#include <gmapping/log/sensorlog.h>

#include <istream>

namespace GMapping {

// abstract source of readings; `stream >> reading` hands out the next one (null for lines of unknown sensors)
class LOG_EXPORT SensorStream {
  public:
    SensorStream(const SensorMap& sensorMap);
    virtual ~SensorStream();

    inline const SensorMap& getSensorMap() const { return m_sensorMap; }
    virtual SensorStream& operator>>(const SensorReading*&) = 0;
    virtual bool rewind() = 0;
    virtual operator bool() const = 0;

    // line parsers (the caller owns the result).  FLASER carries two pose triples; the second one is the odometry pose.
    static RangeReading* parseRange(std::istream& is, const RangeSensor*);
    static OdometryReading* parseOdometry(std::istream& is, const OdometrySensor*);
    static SensorReading* parseReading(std::istream& is, const SensorMap& smap);

  protected:
    const SensorMap& m_sensorMap;
};

// readings parsed on the fly from a text stream; cannot rewind
class LOG_EXPORT InputSensorStream : public SensorStream {
  public:
    InputSensorStream(const SensorMap& sensorMap, std::istream& is);
    virtual SensorStream& operator>>(const SensorReading*&);
    virtual bool rewind();
    virtual operator bool() const;

  protected:
    std::istream& m_inputStream;
};

// readings handed out from a SensorLog held in memory (the log keeps ownership)
class LOG_EXPORT LogSensorStream : public SensorStream {
  public:
    LogSensorStream(const SensorMap& sensorMap, const SensorLog* log);
    virtual SensorStream& operator>>(const SensorReading*&);
    virtual bool rewind();
    virtual operator bool() const;

  protected:
    SensorLog::const_iterator m_cursor;
    const SensorLog* m_log;
};

}  // namespace GMapping
#endif
