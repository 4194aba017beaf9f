// array2d.h -- dense 2-D array, m_cells[x][y] (reference API: include/gmapping/grid/array2d.h; m_cells is public there
// and consumers index it directly).  Host-side container only: the filter's maps live in HBM (csrc/), these objects
// are the host mirror handed to API users.
#ifndef GMB200_ARRAY2D_H
#define GMB200_ARRAY2D_H
#include <assert.h>
#include <gmapping/grid/accessstate.h>
#include <gmapping/utils/point.h>

namespace GMapping {

template <class Cell, const bool debug = false>
class Array2D {
  public:
    Array2D(int xsize = 0, int ysize = 0) : m_cells(0), m_xsize(0), m_ysize(0) { alloc(xsize, ysize); }
    Array2D(const Array2D<Cell, debug>& g) : m_cells(0), m_xsize(0), m_ysize(0) { copyFrom(g); }
    Array2D& operator=(const Array2D& g) {
        if (this != &g) { clear(); copyFrom(g); }
        return *this;
    }
    ~Array2D() { clear(); }
    void clear() {
        for (int x = 0; x < m_xsize; x++) delete[] m_cells[x];
        delete[] m_cells;
        m_cells = 0;
        m_xsize = m_ysize = 0;
    }
    // keep the overlap of the old extent with [xmin,xmax) x [ymin,ymax), re-based at (xmin, ymin)
    void resize(int xmin, int ymin, int xmax, int ymax) {
        const int nx = xmax - xmin, ny = ymax - ymin;
        Cell** nc = new Cell*[nx];
        for (int x = 0; x < nx; x++) nc[x] = new Cell[ny];
        const int x0 = xmin < 0 ? 0 : xmin, y0 = ymin < 0 ? 0 : ymin;
        const int x1 = xmax < m_xsize ? xmax : m_xsize, y1 = ymax < m_ysize ? ymax : m_ysize;
        for (int x = x0; x < x1; x++)
            for (int y = y0; y < y1; y++) nc[x - xmin][y - ymin] = m_cells[x][y];
        clear();
        m_cells = nc;
        m_xsize = nx;
        m_ysize = ny;
    }
    inline bool isInside(int x, int y) const { return x >= 0 && y >= 0 && x < m_xsize && y < m_ysize; }
    inline const Cell& cell(int x, int y) const { assert(isInside(x, y)); return m_cells[x][y]; }
    inline Cell& cell(int x, int y) { assert(isInside(x, y)); return m_cells[x][y]; }
    inline AccessibilityState cellState(int x, int y) const { return (AccessibilityState)(isInside(x, y) ? (Inside | Allocated) : Outside); }
    inline bool isInside(const IntPoint& p) const { return isInside(p.x, p.y); }
    inline const Cell& cell(const IntPoint& p) const { return cell(p.x, p.y); }
    inline Cell& cell(const IntPoint& p) { return cell(p.x, p.y); }
    inline AccessibilityState cellState(const IntPoint& p) const { return cellState(p.x, p.y); }
    inline int getPatchSize() const { return 0; }
    inline int getPatchMagnitude() const { return 0; }
    inline int getXSize() const { return m_xsize; }
    inline int getYSize() const { return m_ysize; }
    inline Cell** cells() { return m_cells; }
    Cell** m_cells;

  protected:
    int m_xsize, m_ysize;

  private:
    void alloc(int xsize, int ysize) {
        if (xsize <= 0 || ysize <= 0) return;
        m_xsize = xsize;
        m_ysize = ysize;
        m_cells = new Cell*[xsize];
        for (int x = 0; x < xsize; x++) m_cells[x] = new Cell[ysize];
    }
    void copyFrom(const Array2D& g) {
        alloc(g.m_xsize, g.m_ysize);
        for (int x = 0; x < m_xsize; x++)
            for (int y = 0; y < m_ysize; y++) m_cells[x][y] = g.m_cells[x][y];
    }
};

}  // namespace GMapping
#endif
