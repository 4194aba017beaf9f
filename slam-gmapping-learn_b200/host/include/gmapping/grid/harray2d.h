// harray2d.h -- two-level grid: a directory of reference-counted 2^m x 2^m patches (reference API:
// include/gmapping/grid/harray2d.h).  Copying shares patches; allocActiveArea() gives the patches of the active
// area a private copy before they are written.  This is the host mirror; the device layout is csrc/gm_device.cuh.
#ifndef GMB200_HARRAY2D_H
#define GMB200_HARRAY2D_H
#include <gmapping/grid/array2d.h>
#include <gmapping/utils/autoptr.h>
#include <gmapping/utils/point.h>

#include <set>

namespace GMapping {

template <class Cell>
class HierarchicalArray2D : public Array2D<autoptr<Array2D<Cell> > > {
    typedef Array2D<autoptr<Array2D<Cell> > > Base;

  public:
    typedef std::set<point<int>, pointcomparator<int> > PointSet;
    // xsize / ysize in cells; the directory has (size >> patchMagnitude) entries per axis
    HierarchicalArray2D(int xsize, int ysize, int patchMagnitude = 5)
        : Base(xsize >> patchMagnitude, ysize >> patchMagnitude), m_patchMagnitude(patchMagnitude), m_patchSize(1 << patchMagnitude) {}
    // sharing copy: directory entries are copied (reference counts go up), the active area is not
    HierarchicalArray2D(const HierarchicalArray2D& hg) : Base(hg), m_patchMagnitude(hg.m_patchMagnitude), m_patchSize(hg.m_patchSize) {}
    HierarchicalArray2D& operator=(const HierarchicalArray2D& hg) {
        if (this != &hg) {
            Base::operator=(hg);
            m_activeArea.clear();
            m_patchMagnitude = hg.m_patchMagnitude;
            m_patchSize = hg.m_patchSize;
        }
        return *this;
    }
    virtual ~HierarchicalArray2D() {}
    // directory resize in PATCH coordinates
    void resize(int ixmin, int iymin, int ixmax, int iymax) { Base::resize(ixmin, iymin, ixmax, iymax); }
    inline int getPatchSize() const { return m_patchMagnitude; }
    inline int getPatchMagnitude() const { return m_patchMagnitude; }

    inline IntPoint patchIndexes(int x, int y) const {
        if (x >= 0 && y >= 0) return IntPoint(x >> m_patchMagnitude, y >> m_patchMagnitude);
        return IntPoint(-1, -1);
    }
    inline bool isAllocated(int x, int y) const {
        const IntPoint c = patchIndexes(x, y);
        return Base::isInside(c.x, c.y) && (int)this->m_cells[c.x][c.y];
    }
    inline AccessibilityState cellState(int x, int y) const {
        const IntPoint c = patchIndexes(x, y);
        if (!Base::isInside(c.x, c.y)) return Outside;
        return (AccessibilityState)(((int)this->m_cells[c.x][c.y]) ? (Inside | Allocated) : Inside);
    }
    inline const Cell& cell(int x, int y) const {
        assert(isAllocated(x, y));
        const IntPoint c = patchIndexes(x, y);
        return (*this->m_cells[c.x][c.y]).cell(x - (c.x << m_patchMagnitude), y - (c.y << m_patchMagnitude));
    }
    // non-const access creates the patch on demand
    inline Cell& cell(int x, int y) {
        const IntPoint c = patchIndexes(x, y);
        assert(Base::isInside(c.x, c.y));
        autoptr<Array2D<Cell> >& slot = this->m_cells[c.x][c.y];
        if (!slot) slot = autoptr<Array2D<Cell> >(createPatch(IntPoint(x, y)));
        return (*slot).cell(x - (c.x << m_patchMagnitude), y - (c.y << m_patchMagnitude));
    }
    inline const Cell& cell(const IntPoint& p) const { return cell(p.x, p.y); }
    inline Cell& cell(const IntPoint& p) { return cell(p.x, p.y); }
    inline bool isAllocated(const IntPoint& p) const { return isAllocated(p.x, p.y); }
    inline AccessibilityState cellState(const IntPoint& p) const { return cellState(p.x, p.y); }
    inline IntPoint patchIndexes(const IntPoint& p) const { return patchIndexes(p.x, p.y); }

    inline void setActiveArea(const PointSet& aa, bool patchCoords = false) {
        m_activeArea.clear();
        for (typename PointSet::const_iterator it = aa.begin(); it != aa.end(); ++it)
            m_activeArea.insert(patchCoords ? *it : patchIndexes(*it));
    }
    const PointSet& getActiveArea() const { return m_activeArea; }
    // every active patch becomes private to this grid: copied if it exists, created if it does not
    inline void allocActiveArea() {
        for (typename PointSet::const_iterator it = m_activeArea.begin(); it != m_activeArea.end(); ++it) {
            autoptr<Array2D<Cell> >& slot = this->m_cells[it->x][it->y];
            Array2D<Cell>* patch = slot ? new Array2D<Cell>(*slot) : createPatch(*it);
            slot = autoptr<Array2D<Cell> >(patch);
        }
    }

  protected:
    virtual Array2D<Cell>* createPatch(const IntPoint&) const { return new Array2D<Cell>(1 << m_patchMagnitude, 1 << m_patchMagnitude); }
    PointSet m_activeArea;
    int m_patchMagnitude;
    int m_patchSize;
};

}  // namespace GMapping
#endif
