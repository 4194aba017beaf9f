// tests/stubs/ros1/stub_all.h -- TEST INFRASTRUCTURE.  Just enough of the ROS1 / tf / rosbag / boost surface that
// gmapping/src/slam_gmapping.cpp of the reference COMPILES (it is never linked or run): declarations with the right shapes, no
// behaviour.  The point of the test that uses it (tests/test_consumers_cpu.py) is the other side of that translation unit: every
// call the ROS node makes into GMapping:: must compile, unchanged, against slam-gmapping-learn_b200/host/include.
#pragma once
#include <cmath>
#include <cstdio>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <queue>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost {
template <class T> using shared_ptr = std::shared_ptr<T>;
struct mutex { struct scoped_lock { explicit scoped_lock(mutex&) {} }; void lock() {} void unlock() {} };
struct thread { template <class F> explicit thread(F) {} void join() {} };
template <class F, class... A> auto bind(F f, A... a) -> decltype(std::bind(f, a...)) { return std::bind(f, a...); }
}  // namespace boost
using namespace std::placeholders;  // _1

#define BOOST_FOREACH(decl, range) for (decl : range)
#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_DEBUG(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_WARN_STREAM(x) ((void)0)
#define ROS_ASSERT(x) ((void)(x))

namespace ros {
struct Duration {
    Duration() {} explicit Duration(double) {}
    Duration& fromSec(double) { return *this; } double toSec() const { return 0; }
    void sleep() {}
};
struct WallDuration : Duration { using Duration::Duration; };
struct Time {
    Time() {} Time(unsigned, unsigned) {} explicit Time(double) {}
    static Time now() { return Time(); }
    bool is_zero() const { return false; } double toSec() const { return 0; }
    Time operator+(const Duration&) const { return *this; } Duration operator-(const Time&) const { return Duration(); }
    bool operator>(const Time&) const { return false; } bool operator<(const Time&) const { return false; }
};
inline bool operator>(const Duration&, const Duration&) { return false; }
struct Rate { explicit Rate(double) {} void sleep() {} };
inline bool ok() { return true; }
struct Publisher { template <class M> void publish(const M&) const {} };
struct ServiceServer {};
struct NodeHandle {
    NodeHandle() {} explicit NodeHandle(const std::string&) {}
    template <class T> bool getParam(const std::string&, T&) const { return false; }
    template <class T, class D> void param(const std::string&, T& v, const D& d) const { v = d; }
    template <class M> Publisher advertise(const std::string&, int, bool = false) { return Publisher(); }
    template <class C, class Q, class R> ServiceServer advertiseService(const std::string&, bool (C::*)(Q&, R&), C*) { return ServiceServer(); }
};
}  // namespace ros

namespace std_msgs {
struct Header { unsigned seq = 0; ros::Time stamp; std::string frame_id; };
struct Float64 { double data = 0; };
}  // namespace std_msgs
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct TransformStamped { std_msgs::Header header; std::string child_frame_id; };
}  // namespace geometry_msgs
namespace sensor_msgs {
struct LaserScan {
    typedef std::shared_ptr<const LaserScan> ConstPtr;
    std_msgs::Header header;
    float angle_min = 0, angle_max = 0, angle_increment = 0, time_increment = 0, scan_time = 0, range_min = 0, range_max = 0;
    std::vector<float> ranges, intensities;
};
}  // namespace sensor_msgs
namespace nav_msgs {
struct MapMetaData { ros::Time map_load_time; float resolution = 0; unsigned width = 0, height = 0; geometry_msgs::Pose origin; };
struct OccupancyGrid { std_msgs::Header header; MapMetaData info; std::vector<signed char> data; };
struct GetMap {
    struct Request {};
    struct Response { OccupancyGrid map; };
};
}  // namespace nav_msgs

namespace tf2 { struct TransformException : std::runtime_error { using std::runtime_error::runtime_error; }; }
namespace tf {
typedef tf2::TransformException TransformException;
struct Vector3 {
    Vector3() {} Vector3(double, double, double) {}
    double x() const { return 0; } double y() const { return 0; } double z() const { return 0; }
    void setValue(double, double, double) {}
};
typedef Vector3 Point;
struct Quaternion { Quaternion() {} };
inline Quaternion createQuaternionFromRPY(double, double, double) { return Quaternion(); }
inline double getYaw(const Quaternion&) { return 0; }
struct Transform {
    Transform() {} Transform(const Quaternion&, const Vector3& = Vector3()) {}
    Transform inverse() const { return *this; } Transform operator*(const Transform&) const { return *this; }
    Vector3 getOrigin() const { return Vector3(); } Quaternion getRotation() const { return Quaternion(); }
    void setIdentity() {}
};
typedef Transform Pose;
template <class T> struct Stamped : T {
    Stamped() {} Stamped(const T& t, const ros::Time&, const std::string&) : T(t) {}
    ros::Time stamp_; std::string frame_id_;
};
struct StampedTransform : Transform {
    StampedTransform() {} StampedTransform(const Transform& t, const ros::Time&, const std::string&, const std::string&) : Transform(t) {}
};
struct tfMessage { typedef std::shared_ptr<const tfMessage> ConstPtr; std::vector<geometry_msgs::TransformStamped> transforms; };
inline void transformStampedMsgToTF(const geometry_msgs::TransformStamped&, StampedTransform&) {}
struct TransformListener {
    TransformListener() {} explicit TransformListener(const ros::Duration&) {}
    template <class A, class B> void transformPose(const std::string&, const A&, B&) const {}
    template <class A, class B> void transformPoint(const std::string&, const A&, B&) const {}
    void lookupTransform(const std::string&, const std::string&, const ros::Time&, StampedTransform&) const {}
    void setTransform(const StampedTransform&) {}
    std::string resolve(const std::string& s) const { return s; }
};
struct TransformBroadcaster { void sendTransform(const StampedTransform&) {} };
}  // namespace tf
namespace message_filters {
template <class M> struct Subscriber { Subscriber(ros::NodeHandle&, const std::string&, int) {} };
}
namespace tf {
template <class M> struct MessageFilter {
    template <class S> MessageFilter(S&, TransformListener&, const std::string&, int) {}
    template <class F> void registerCallback(F) {}
};
}  // namespace tf
namespace rosbag {
namespace bagmode { enum Mode { Read = 1 }; }
struct Bag { void open(const std::string&, int) {} void close() {} };
struct TopicQuery { explicit TopicQuery(const std::vector<std::string>&) {} };
struct MessageInstance { template <class M> std::shared_ptr<const M> instantiate() const { return std::shared_ptr<const M>(); } };
struct View {
    View(Bag&, const TopicQuery&) {}
    const MessageInstance* begin() const { return nullptr; } const MessageInstance* end() const { return nullptr; }
};
}  // namespace rosbag
