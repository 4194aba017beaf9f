// smmap.h -- the matcher's map cell and map type (reference API: include/gmapping/scanmatcher/smmap.h).
// PointAccumulator is the 16-byte cell {float acc.x, acc.y; int n, visits}: bit-identical to gm_cell in
// include/gm_b200.h, which is how patches travel between HBM and this host mirror.
#ifndef GMB200_SMMAP_H
#define GMB200_SMMAP_H
#include <gmapping/grid/harray2d.h>
#include <gmapping/grid/map.h>
#include <gmapping/utils/point.h>
#define SIGHT_INC 1

namespace GMapping {

struct PointAccumulator {
    typedef point<float> FloatPoint;
    PointAccumulator() : acc(0, 0), n(0), visits(0) {}
    PointAccumulator(int i) : acc(0, 0), n(0), visits(0) { assert(i == -1); (void)i; }
    // value == true: an endpoint fell into the cell; false: a ray passed through it
    inline void update(bool value, const Point& p = Point(0, 0)) {
        if (value) {
            acc.x += static_cast<float>(p.x);
            acc.y += static_cast<float>(p.y);
            n++;
            visits += SIGHT_INC;
        } else
            visits++;
    }
    inline Point mean() const { return 1. / n * Point(acc.x, acc.y); }
    // occupancy: hits / visits, -1 when never visited
    inline operator double() const { return visits ? (double)n * SIGHT_INC / (double)visits : -1; }
    inline void add(const PointAccumulator& p) {
        acc = acc + p.acc;
        n += p.n;
        visits += p.visits;
    }
    static const PointAccumulator& Unknown();
    static PointAccumulator* unknown_ptr;
    FloatPoint acc;
    int n, visits;
    inline double entropy() const {
        if (!visits) return -log(.5);
        if (n == visits || n == 0) return 0;
        const double x = (double)n * SIGHT_INC / (double)visits;
        return -(x * log(x) + (1 - x) * log(1 - x));
    }
};
static_assert(sizeof(PointAccumulator) == 16, "PointAccumulator must match the 16-byte device cell");

typedef Map<PointAccumulator, HierarchicalArray2D<PointAccumulator> > ScanMatcherMap;

}  // namespace GMapping
#endif
