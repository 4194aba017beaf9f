// log.cpp -- CARMEN log reading for the log-driven drivers (reference: log/carmenconfiguration.cpp, log/sensorstream.cpp,
// log/sensorlog.cpp).  I/O only; nothing here is on the particle-update path.
#include <gmapping/log/carmenconfiguration.h>
#include <gmapping/log/sensorlog.h>
#include <gmapping/log/sensorstream.h>
#include <gmapping/sensor/sensor_odometry/odometrysensor.h>
#include <gmapping/sensor/sensor_range/rangesensor.h>

#include <cassert>
#include <cmath>
#include <cstdlib>
#include <sstream>
#include <string>

namespace GMapping {

Configuration::~Configuration() {}

// collect `PARAM name values...` lines; remember the beam count of the first laser message of each kind
std::istream& CarmenConfiguration::load(std::istream& is) {
    clear();
    std::string line, frontBeams, rearBeams;
    bool front = false, rear = false;
    while (std::getline(is, line)) {
        std::istringstream ls(line);
        std::string tag;
        if (!(ls >> tag)) continue;
        if (tag == "FLASER") { front = true; ls >> frontBeams; }
        else if (tag == "RLASER") { rear = true; ls >> rearBeams; }
        else if (tag == "ROBOTLASER1" || tag == "ROBOTLASER2") {
            std::string skip;
            for (int k = 0; k < 7; k++) ls >> skip;  // type, start angle, fov, resolution, max range, accuracy, remission mode
            if (tag == "ROBOTLASER1") { front = true; ls >> frontBeams; } else { rear = true; ls >> rearBeams; }
        } else if (tag == "PARAM") {
            std::string name, v;
            if (!(ls >> name)) continue;
            std::vector<std::string> values;
            while (ls >> v) values.push_back(v);
            insert(std::make_pair(name, values));
        }
    }
    if (front) {
        insert(std::make_pair(std::string("laser_beams"), std::vector<std::string>(1, frontBeams)));
        insert(std::make_pair(std::string("robot_use_laser"), std::vector<std::string>(1, "on")));
    }
    if (rear) {
        insert(std::make_pair(std::string("rear_laser_beams"), std::vector<std::string>(1, rearBeams)));
        insert(std::make_pair(std::string("robot_use_rear_laser"), std::vector<std::string>(1, "on")));
    }
    return is;
}

// beam table symmetric about the centre beam, angles accumulated with += step (reference: carmenconfiguration.cpp:218-236);
// the accumulated rounding is what the matcher sees, so it is reproduced, not "improved"
static void symmetricBeams(std::vector<RangeSensor::Beam>& beams, unsigned int n, double resolutionDeg, double maxrange) {
    beams.resize(n);
    const double half = (double)n / 2.;
    const unsigned int low = (unsigned int)floor(half), up = (unsigned int)ceil(half);
    const double step = resolutionDeg * M_PI / 180.;
    double angle = n % 2 ? 0 : step;
    RangeSensor::Beam b;
    b.span = 0; b.maxRange = maxrange; b.s = 1; b.c = 1;
    for (unsigned int i = n % 2 ? 0 : 1; i < low + 1; i++, angle += step) {
        b.pose = OrientedPoint(0, 0, -angle);
        beams[low - i] = b;
        b.pose = OrientedPoint(0, 0, angle);
        beams[up + i - 1] = b;
    }
}

SensorMap CarmenConfiguration::computeSensorMap() const {
    SensorMap smap;
    OdometrySensor* odometry = new OdometrySensor("ODOM");
    OdometrySensor* truepos = new OdometrySensor("TRUEPOS", true);
    smap.insert(std::make_pair(odometry->getName(), odometry));
    smap.insert(std::make_pair(truepos->getName(), truepos));
    const_iterator key = find("robot_use_laser");
    if (key != end() && key->second.front() == "on") {
        for (int fmt = 0; fmt < 2; fmt++) {  // the same laser under its old (FLASER) and new (ROBOTLASER1) message name
            RangeSensor* laser = new RangeSensor(fmt ? "ROBOTLASER1" : "FLASER");
            laser->newFormat = fmt != 0;
            laser->m_pose = OrientedPoint(0, 0, 0);
            key = find("robot_frontlaser_offset");
            if (key != end()) laser->m_pose.x = atof(key->second.front().c_str());
            unsigned int n = 180;
            key = find("laser_beams");
            if (key != end()) n = atoi(key->second.front().c_str());
            double maxrange = 50, resolution = 1.;
            if (n == 180 || n == 181) resolution = 1.;
            else if (n == 360 || n == 361 || n == 540 || n == 541) resolution = .5;
            else if (n == 769 || n == 682) { resolution = 360. / 1024.; maxrange = 4.1; }
            else if (n == 683) { resolution = 360. / 1024.; maxrange = 5.5; }
            else {
                key = find("laser_front_laser_resolution");
                if (key != end()) resolution = atof(key->second.front().c_str());
            }
            symmetricBeams(laser->m_beams, n, resolution, maxrange);
            laser->updateBeamsLookup();
            smap.insert(std::make_pair(laser->getName(), laser));
        }
    }
    return smap;
}

// ------------------------------------------------------------------------------------------ streams
SensorStream::SensorStream(const SensorMap& sensorMap) : m_sensorMap(sensorMap) {}
SensorStream::~SensorStream() {}

SensorReading* SensorStream::parseReading(std::istream& is, const SensorMap& smap) {
    std::string line;
    if (!std::getline(is, line)) return 0;
    std::istringstream ls(line);
    std::string name;
    if (!(ls >> name)) return 0;
    SensorMap::const_iterator it = smap.find(name);
    if (it == smap.end()) return 0;
    if (const OdometrySensor* odo = dynamic_cast<const OdometrySensor*>(it->second)) return parseOdometry(ls, odo);
    if (const RangeSensor* range = dynamic_cast<const RangeSensor*>(it->second)) return parseRange(ls, range);
    return 0;
}

// ODOM x y theta tv rv accel timestamp host logger_timestamp
OdometryReading* SensorStream::parseOdometry(std::istream& is, const OdometrySensor* osen) {
    OdometryReading* reading = new OdometryReading(osen);
    OrientedPoint pose, speed(0, 0, 0), accel(0, 0, 0);
    is >> pose.x >> pose.y >> pose.theta >> speed.x >> speed.theta >> accel.x;
    reading->setPose(pose); reading->setSpeed(speed); reading->setAcceleration(accel);
    double t = 0, rel = 0;
    std::string host;
    is >> t >> host >> rel;
    reading->setTime(t);
    return reading;
}

// FLASER n r_0..r_{n-1} laser_x laser_y laser_theta odom_x odom_y odom_theta timestamp host logger_timestamp
// (the SECOND pose triple is the odometry pose the filter uses); ROBOTLASER1 carries extra fields around the ranges
RangeReading* SensorStream::parseRange(std::istream& is, const RangeSensor* rs) {
    std::string skip;
    if (rs->newFormat) for (int k = 0; k < 7; k++) is >> skip;
    unsigned int n = 0;
    is >> n;
    assert(n == rs->beams().size());
    RangeReading* reading = new RangeReading(rs);
    reading->resize(n);
    for (unsigned int i = 0; i < n; i++) is >> (*reading)[i];
    if (rs->newFormat) {
        int remissions = 0;
        is >> remissions;
        double v;
        for (int i = 0; i < remissions; i++) is >> v;
    }
    OrientedPoint laserPose, pose;
    is >> laserPose.x >> laserPose.y >> laserPose.theta >> pose.x >> pose.y >> pose.theta;
    reading->setPose(pose);
    if (rs->newFormat) for (int k = 0; k < 5; k++) is >> skip;
    double t = 0, rel = 0;
    std::string host;
    is >> t >> host >> rel;
    reading->setTime(t);
    return reading;
}

InputSensorStream::InputSensorStream(const SensorMap& sensorMap, std::istream& is) : SensorStream(sensorMap), m_inputStream(is) {}
InputSensorStream::operator bool() const { return (bool)m_inputStream; }
bool InputSensorStream::rewind() { return false; }
SensorStream& InputSensorStream::operator>>(const SensorReading*& reading) {
    reading = parseReading(m_inputStream, m_sensorMap);
    return *this;
}

LogSensorStream::LogSensorStream(const SensorMap& sensorMap, const SensorLog* log) : SensorStream(sensorMap), m_cursor(log->begin()), m_log(log) {
    assert(m_log);
}
LogSensorStream::operator bool() const { return m_cursor == m_log->end(); }  // sic: the reference reports "true" at the end
bool LogSensorStream::rewind() { m_cursor = m_log->begin(); return true; }
SensorStream& LogSensorStream::operator>>(const SensorReading*& rd) {
    rd = *m_cursor;
    m_cursor++;
    return *this;
}

// ------------------------------------------------------------------------------------------ whole-log container
SensorLog::SensorLog(const SensorMap& sm) : m_sensorMap(sm) {}
SensorLog::~SensorLog() {
    for (iterator it = begin(); it != end(); ++it) delete *it;
}
std::istream& SensorLog::load(std::istream& is) {
    for (iterator it = begin(); it != end(); ++it) delete *it;
    clear();
    while (is) {
        SensorReading* r = SensorStream::parseReading(is, m_sensorMap);
        if (r) push_back(r);
    }
    return is;
}
// Extent of the poses in the log and the pose of the first range reading (reference: log/sensorlog.cpp:121-150), with the
// reference's behaviour kept where it is odd, because gfs_nogui-style drivers size their maps from it (-autosize):
//   * odometry readings and range readings both count; a reading of any other kind counts as the point (0, 0);
//   * ymin is NOT a minimum: the reference assigns the current y on both branches of its comparison, so ymin ends up as
//     the y of the last reading of the log.
OrientedPoint SensorLog::boundingBox(double& xmin, double& ymin, double& xmax, double& ymax) const {
    xmin = ymin = 1e6;
    xmax = ymax = -1e6;
    OrientedPoint start;
    bool haveStart = false;
    for (const_iterator it = begin(); it != end(); ++it) {
        double x = 0., y = 0.;
        if (const OdometryReading* o = dynamic_cast<const OdometryReading*>(*it)) { x = o->getPose().x; y = o->getPose().y; }
        if (const RangeReading* rr = dynamic_cast<const RangeReading*>(*it)) {
            x = rr->getPose().x; y = rr->getPose().y;
            if (!haveStart) { start = rr->getPose(); haveStart = true; }
        }
        if (x < xmin) xmin = x;
        if (x > xmax) xmax = x;
        ymin = y;  // sic
        if (y > ymax) ymax = y;
    }
    return start;
}

}  // namespace GMapping
