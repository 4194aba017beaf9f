// gridlinetraversal.h -- the grid line of the reference's ray casting as a host utility
// (API of include/gmapping/scanmatcher/gridlinetraversal.h:8-19).  The B200 kernels do not walk lines on the host: k_register
// evaluates the same Bresenham state in closed form per 32-cell chunk.  This header exists for consumers that include it,
// and as the host-side statement of which cells a beam crosses: the cells of an integer Bresenham line that always advances
// along the major axis from the smaller coordinate, reported in start -> end order (the last cell is the beam's end point).
#ifndef GMB200_GRIDLINETRAVERSAL_H
#define GMB200_GRIDLINETRAVERSAL_H
#include <gmapping/utils/point.h>

#include <cstdlib>

namespace GMapping {

typedef struct {
    int num_points;
    IntPoint* points;  // caller-owned, large enough for max(|dx|, |dy|) + 1 cells
} GridLineTraversalLine;

struct GridLineTraversal {
    // cells in the order the walk produces them: from the end with the smaller major-axis coordinate upwards
    static inline void gridLineCore(IntPoint start, IntPoint end, GridLineTraversalLine* line) {
        const int dx = std::abs(end.x - start.x), dy = std::abs(end.y - start.y);
        const bool steep = dy > dx;  // major axis: y when steep, else x
        const int major0 = steep ? start.y : start.x, major1 = steep ? end.y : end.x;
        const int minor0 = steep ? start.x : start.y, minor1 = steep ? end.x : end.y;
        const int dMajor = steep ? dy : dx, dMinor = steep ? dx : dy;
        const bool fromEnd = major0 > major1;
        int a = fromEnd ? major1 : major0, b = fromEnd ? minor1 : minor0;
        const int last = fromEnd ? major0 : major1;
        const int minorStep = ((minor1 - minor0) * (fromEnd ? -1 : 1)) > 0 ? 1 : -1;
        int err = 2 * dMinor - dMajor;
        int n = 0;
        line->points[n++] = steep ? IntPoint(b, a) : IntPoint(a, b);
        while (a < last) {
            a++;
            if (err < 0) err += 2 * dMinor;
            else { b += minorStep; err += 2 * (dMinor - dMajor); }
            line->points[n++] = steep ? IntPoint(b, a) : IntPoint(a, b);
        }
        line->num_points = n;
    }
    // the same cells from `start` to `end`
    static inline void gridLine(IntPoint start, IntPoint end, GridLineTraversalLine* line) {
        gridLineCore(start, end, line);
        if (line->num_points && (line->points[0].x != start.x || line->points[0].y != start.y))
            for (int i = 0, j = line->num_points - 1; i < j; i++, j--) {
                const IntPoint t = line->points[i];
                line->points[i] = line->points[j];
                line->points[j] = t;
            }
    }
};

}  // namespace GMapping
#endif
