// point.h -- 2-D points and poses (reference API: include/gmapping/utils/point.h).
// The arithmetic forms matter for parity: `point * point` is the dot product x1*x2 + y1*y2, `scalar * point`
// scales both coordinates, absoluteDifference / absoluteSum compose poses in the order of point.h:117-134.
#ifndef GMB200_POINT_H
#define GMB200_POINT_H
#include <assert.h>
#include <math.h>

#include <iostream>

#include "gmapping/utils/gvalues.h"

namespace GMapping {

template <class T>
struct point {
    inline point() : x(0), y(0) {}
    inline point(T _x, T _y) : x(_x), y(_y) {}
    T x, y;
};

template <class T> inline point<T> operator+(const point<T>& a, const point<T>& b) { return point<T>(a.x + b.x, a.y + b.y); }
template <class T> inline point<T> operator-(const point<T>& a, const point<T>& b) { return point<T>(a.x - b.x, a.y - b.y); }
template <class T> inline point<T> operator*(const point<T>& p, const T& v) { return point<T>(p.x * v, p.y * v); }
template <class T> inline point<T> operator*(const T& v, const point<T>& p) { return point<T>(p.x * v, p.y * v); }
template <class T> inline T operator*(const point<T>& a, const point<T>& b) { return a.x * b.x + a.y * b.y; }

template <class T, class A>
struct orientedpoint : public point<T> {
    inline orientedpoint() : point<T>(0, 0), theta(0) {}
    inline orientedpoint(const point<T>& p) : point<T>(p.x, p.y), theta(0) {}
    inline orientedpoint(T x, T y, A _theta) : point<T>(x, y), theta(_theta) {}
    // bring theta into [-pi, pi)
    inline void normalize() {
        if (theta >= -M_PI && theta < M_PI) return;
        const int turns = (int)(theta / (2 * M_PI));
        theta = theta - turns * 2 * M_PI;
        if (theta >= M_PI) theta -= 2 * M_PI;
        if (theta < -M_PI) theta += 2 * M_PI;
    }
    inline orientedpoint<T, A> rotate(A alpha) {
        const T s = sin(alpha), c = cos(alpha);
        A a = alpha + theta;
        a = atan2(sin(a), cos(a));
        return orientedpoint(c * this->x - s * this->y, s * this->x + c * this->y, a);
    }
    A theta;
};

template <class T, class A> orientedpoint<T, A> operator+(const orientedpoint<T, A>& a, const orientedpoint<T, A>& b) {
    return orientedpoint<T, A>(a.x + b.x, a.y + b.y, a.theta + b.theta);
}
template <class T, class A> orientedpoint<T, A> operator-(const orientedpoint<T, A>& a, const orientedpoint<T, A>& b) {
    return orientedpoint<T, A>(a.x - b.x, a.y - b.y, a.theta - b.theta);
}
template <class T, class A> orientedpoint<T, A> operator*(const orientedpoint<T, A>& p, const T& v) {
    return orientedpoint<T, A>(p.x * v, p.y * v, p.theta * v);
}
template <class T, class A> orientedpoint<T, A> operator*(const T& v, const orientedpoint<T, A>& p) {
    return orientedpoint<T, A>(p.x * v, p.y * v, p.theta * v);
}

// p1 expressed in the frame of p2 (point.h:117-124)
template <class T, class A> orientedpoint<T, A> absoluteDifference(const orientedpoint<T, A>& p1, const orientedpoint<T, A>& p2) {
    orientedpoint<T, A> d = p1 - p2;
    d.theta = atan2(sin(d.theta), cos(d.theta));
    const double s = sin(p2.theta), c = cos(p2.theta);
    return orientedpoint<T, A>(c * d.x + s * d.y, -s * d.x + c * d.y, d.theta);
}
// p2 (given in the frame of p1) expressed in the frame p1 lives in (point.h:126-134)
template <class T, class A> orientedpoint<T, A> absoluteSum(const orientedpoint<T, A>& p1, const orientedpoint<T, A>& p2) {
    const double s = sin(p1.theta), c = cos(p1.theta);
    return orientedpoint<T, A>(c * p2.x - s * p2.y, s * p2.x + c * p2.y, p2.theta) + p1;
}
template <class T, class A> point<T> absoluteSum(const orientedpoint<T, A>& p1, const point<T>& p2) {
    const double s = sin(p1.theta), c = cos(p1.theta);
    return point<T>(c * p2.x - s * p2.y, s * p2.x + c * p2.y) + (point<T>)p1;
}

template <class T> struct pointcomparator {
    bool operator()(const point<T>& a, const point<T>& b) const { return a.x < b.x || (a.x == b.x && a.y < b.y); }
};
template <class T> struct pointradialcomparator {
    point<T> origin;
    bool operator()(const point<T>& a, const point<T>& b) const {
        const point<T> da = a - origin, db = b - origin;
        return atan2(da.y, da.x) < atan2(db.y, db.x);
    }
};

template <class T> inline point<T> max(const point<T>& a, const point<T>& b) { return point<T>(a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y); }
template <class T> inline point<T> min(const point<T>& a, const point<T>& b) { return point<T>(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y); }
template <class T, class F> inline point<T> interpolate(const point<T>& p1, const F& t1, const point<T>& p2, const F& t2, const F& t3) {
    const F gain = (t3 - t1) / (t2 - t1);
    return p1 + (p2 - p1) * gain;
}
template <class T, class A, class F>
inline orientedpoint<T, A> interpolate(const orientedpoint<T, A>& p1, const F& t1, const orientedpoint<T, A>& p2, const F& t2, const F& t3) {
    const F gain = (t3 - t1) / (t2 - t1);
    orientedpoint<T, A> p;
    p.x = p1.x + (p2.x - p1.x) * gain;
    p.y = p1.y + (p2.y - p1.y) * gain;
    const double s = sin(p1.theta) + (sin(p2.theta) - sin(p1.theta)) * gain, c = cos(p1.theta) + (cos(p2.theta) - cos(p1.theta)) * gain;
    p.theta = atan2(s, c);
    return p;
}
template <class T> inline double euclidianDist(const point<T>& a, const point<T>& b) { return hypot(a.x - b.x, a.y - b.y); }
template <class T, class A> inline double euclidianDist(const orientedpoint<T, A>& a, const orientedpoint<T, A>& b) { return hypot(a.x - b.x, a.y - b.y); }
template <class T, class A> inline double euclidianDist(const orientedpoint<T, A>& a, const point<T>& b) { return hypot(a.x - b.x, a.y - b.y); }
template <class T, class A> inline double euclidianDist(const point<T>& a, const orientedpoint<T, A>& b) { return hypot(a.x - b.x, a.y - b.y); }

typedef point<int> IntPoint;
typedef point<double> Point;
typedef orientedpoint<double, double> OrientedPoint;

}  // namespace GMapping
#endif
