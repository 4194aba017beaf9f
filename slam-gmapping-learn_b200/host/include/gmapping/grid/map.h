// map.h -- metric grid map: world <-> cell transform over a storage class (reference API: include/gmapping/grid/map.h).
//
// B200 addition: a map can be *device-resident*.  Then a MapResidency object is attached; the cells live in HBM
// (a particle slot of a gm_ctx, include/gm_b200.h) and this object is only the lazily refreshed host mirror:
// every accessor first calls sync(), which downloads geometry + patches when the device copy changed.  Plain
// host maps (no residency attached) behave exactly like the reference's.
#ifndef GMB200_MAP_H
#define GMB200_MAP_H
#include <assert.h>
#include <gmapping/grid/accessstate.h>
#include <gmapping/grid/array2d.h>
#include <gmapping/grid/harray2d.h>
#include <gmapping/utils/point.h>

#include <memory>

namespace GMapping {

typedef Array2D<double> DoubleArray2D;

// link between a host Map object and its device-resident copy (implemented in scanmatcher/devicemap.cpp)
struct MapResidency {
    virtual ~MapResidency() {}
    virtual void pull(void* hostMap) = 0;        // refresh the host mirror from the device
    virtual void push(const void* hostMap) = 0;  // replace the device copy by the host mirror
    bool stale = false;                          // device copy is newer than the host mirror
    // the mirror was handed out for writing (non-const cell() / storage() / resize() / grow()) since the last transfer: the
    // next device call on this map uploads it first, so that a host-side edit is never silently lost
    bool hostDirty = false;
};

template <class Cell, class Storage, const bool isClass = true>
class Map {
  public:
    Map(int mapSizeX, int mapSizeY, double delta);
    Map(const Point& center, double worldSizeX, double worldSizeY, double delta);
    Map(const Point& center, double xmin, double ymin, double xmax, double ymax, double delta);
    // copies are plain host maps: the mirror is refreshed first, the residency is not inherited
    Map(const Map& m)
        : m_center((m.sync(), m.m_center)), m_worldSizeX(m.m_worldSizeX), m_worldSizeY(m.m_worldSizeY), m_delta(m.m_delta),
          m_storage(m.m_storage), m_mapSizeX(m.m_mapSizeX), m_mapSizeY(m.m_mapSizeY), m_sizeX2(m.m_sizeX2), m_sizeY2(m.m_sizeY2) {}
    Map& operator=(const Map& m) {
        if (this == &m) return *this;
        m.sync();
        m_center = m.m_center; m_worldSizeX = m.m_worldSizeX; m_worldSizeY = m.m_worldSizeY; m_delta = m.m_delta;
        m_storage = m.m_storage; m_mapSizeX = m.m_mapSizeX; m_mapSizeY = m.m_mapSizeY; m_sizeX2 = m.m_sizeX2; m_sizeY2 = m.m_sizeY2;
        m_residency.reset();
        return *this;
    }

    void resize(double xmin, double ymin, double xmax, double ymax);
    void grow(double xmin, double ymin, double xmax, double ymax);

    // ix = (int)round((x - center.x) / delta) + sizeX2, round() half away from zero (map.h:283-287 in the reference)
    inline IntPoint world2map(const Point& p) const {
        sync();
        return IntPoint((int)round((p.x - m_center.x) / m_delta) + m_sizeX2, (int)round((p.y - m_center.y) / m_delta) + m_sizeY2);
    }
    inline Point map2world(const IntPoint& p) const {
        sync();
        return Point((p.x - m_sizeX2) * m_delta, (p.y - m_sizeY2) * m_delta) + m_center;
    }
    inline IntPoint world2map(double x, double y) const { return world2map(Point(x, y)); }
    inline Point map2world(int x, int y) const { return map2world(IntPoint(x, y)); }

    inline Point getCenter() const { return m_center; }
    inline double getWorldSizeX() const { sync(); return m_worldSizeX; }
    inline double getWorldSizeY() const { sync(); return m_worldSizeY; }
    inline int getMapSizeX() const { sync(); return m_mapSizeX; }
    inline int getMapSizeY() const { sync(); return m_mapSizeY; }
    inline double getDelta() const { return m_delta; }
    inline double getMapResolution() const { return m_delta; }
    inline double getResolution() const { return m_delta; }
    inline void getSize(double& xmin, double& ymin, double& xmax, double& ymax) const {
        const Point lo = map2world(0, 0), hi = map2world(IntPoint(m_mapSizeX - 1, m_mapSizeY - 1));
        xmin = lo.x, ymin = lo.y, xmax = hi.x, ymax = hi.y;
    }

    inline Cell& cell(int x, int y) { return cell(IntPoint(x, y)); }
    inline Cell& cell(const IntPoint& p) {
        syncForWrite();
        assert(m_storage.cellState(p) & Inside);
        return m_storage.cell(p);
    }
    inline const Cell& cell(int x, int y) const { return cell(IntPoint(x, y)); }
    // cells outside the map or in unallocated patches read as the static unknown cell
    inline const Cell& cell(const IntPoint& p) const {
        sync();
        if (m_storage.cellState(p) & Allocated) return m_storage.cell(p);
        return m_unknown;
    }
    inline Cell& cell(double x, double y) { return cell(Point(x, y)); }
    inline Cell& cell(const Point& p) { return cell(world2map(p)); }
    inline const Cell& cell(double x, double y) const { return cell(Point(x, y)); }
    inline const Cell& cell(const Point& p) const { return cell(world2map(p)); }

    inline bool isInside(int x, int y) const { sync(); return m_storage.cellState(IntPoint(x, y)) & Inside; }
    inline bool isInside(const IntPoint& p) const { sync(); return m_storage.cellState(p) & Inside; }
    inline bool isInside(double x, double y) const { return m_storage.cellState(world2map(x, y)) & Inside; }
    inline bool isInside(const Point& p) const { return m_storage.cellState(world2map(p)) & Inside; }

    inline Storage& storage() { syncForWrite(); return m_storage; }
    inline const Storage& storage() const { sync(); return m_storage; }
    DoubleArray2D* toDoubleArray() const;
    Map<double, DoubleArray2D, false>* toDoubleMap() const;

    // ---- device residency (B200 addition)
    inline void sync() const {
        if (m_residency && m_residency->stale) m_residency->pull(const_cast<Map*>(this));
    }
    inline void syncForWrite() {
        sync();
        if (m_residency) m_residency->hostDirty = true;
    }
    inline const std::shared_ptr<MapResidency>& residency() const { return m_residency; }
    inline void setResidency(const std::shared_ptr<MapResidency>& r) const { m_residency = r; }
    // used by the residency to overwrite the mirror's geometry after a device-side Map::resize
    inline void setGeometry(int sizeX2, int sizeY2, int patchesX, int patchesY) {
        m_sizeX2 = sizeX2; m_sizeY2 = sizeY2;
        m_mapSizeX = patchesX << m_storage.getPatchSize(); m_mapSizeY = patchesY << m_storage.getPatchSize();
        m_worldSizeX = m_mapSizeX * m_delta; m_worldSizeY = m_mapSizeY * m_delta;
    }
    inline int getSizeX2() const { sync(); return m_sizeX2; }
    inline int getSizeY2() const { sync(); return m_sizeY2; }
    inline Storage& rawStorage() { return m_storage; }  // no sync: for the residency itself

  protected:
    Point m_center;
    double m_worldSizeX, m_worldSizeY, m_delta;
    Storage m_storage;
    int m_mapSizeX, m_mapSizeY;
    int m_sizeX2, m_sizeY2;
    static const Cell m_unknown;
    mutable std::shared_ptr<MapResidency> m_residency;
};

typedef Map<double, DoubleArray2D, false> DoubleMap;

template <class Cell, class Storage, const bool isClass>
const Cell Map<Cell, Storage, isClass>::m_unknown = Cell(-1);

template <class Cell, class Storage, const bool isClass>
Map<Cell, Storage, isClass>::Map(int mapSizeX, int mapSizeY, double delta)
    : m_center(0.5 * mapSizeX * delta, 0.5 * mapSizeY * delta), m_worldSizeX(mapSizeX * delta), m_worldSizeY(mapSizeY * delta), m_delta(delta),
      m_storage(mapSizeX, mapSizeY), m_mapSizeX(mapSizeX), m_mapSizeY(mapSizeY), m_sizeX2(mapSizeX >> 1), m_sizeY2(mapSizeY >> 1) {}

// storage of ceil(world / delta) cells; with a patch storage the map size is rounded DOWN to whole patches
template <class Cell, class Storage, const bool isClass>
Map<Cell, Storage, isClass>::Map(const Point& center, double worldSizeX, double worldSizeY, double delta)
    : m_center(center), m_worldSizeX(worldSizeX), m_worldSizeY(worldSizeY), m_delta(delta),
      m_storage((int)ceil(worldSizeX / delta), (int)ceil(worldSizeY / delta)) {
    m_mapSizeX = m_storage.getXSize() << m_storage.getPatchSize();
    m_mapSizeY = m_storage.getYSize() << m_storage.getPatchSize();
    m_sizeX2 = m_mapSizeX >> 1;
    m_sizeY2 = m_mapSizeY >> 1;
}

// as above, but the cell of the centre is placed from the given lower corner
template <class Cell, class Storage, const bool isClass>
Map<Cell, Storage, isClass>::Map(const Point& center, double xmin, double ymin, double xmax, double ymax, double delta)
    : m_center(center), m_worldSizeX(xmax - xmin), m_worldSizeY(ymax - ymin), m_delta(delta),
      m_storage((int)ceil((xmax - xmin) / delta), (int)ceil((ymax - ymin) / delta)) {
    m_mapSizeX = m_storage.getXSize() << m_storage.getPatchSize();
    m_mapSizeY = m_storage.getYSize() << m_storage.getPatchSize();
    m_sizeX2 = (int)round((m_center.x - xmin) / m_delta);
    m_sizeY2 = (int)round((m_center.y - ymin) / m_delta);
}

// re-base the storage so that it covers [xmin,xmax] x [ymin,ymax] in whole patches; the centre never moves,
// only its cell index does
template <class Cell, class Storage, const bool isClass>
void Map<Cell, Storage, isClass>::resize(double xmin, double ymin, double xmax, double ymax) {
    syncForWrite();
    const IntPoint imin = world2map(xmin, ymin), imax = world2map(xmax, ymax);
    const int unit = 1 << m_storage.getPatchMagnitude();
    const int pxmin = (int)floor((float)imin.x / unit), pxmax = (int)ceil((float)imax.x / unit);
    const int pymin = (int)floor((float)imin.y / unit), pymax = (int)ceil((float)imax.y / unit);
    m_storage.resize(pxmin, pymin, pxmax, pymax);
    m_mapSizeX = m_storage.getXSize() << m_storage.getPatchSize();
    m_mapSizeY = m_storage.getYSize() << m_storage.getPatchSize();
    m_worldSizeX = xmax - xmin;
    m_worldSizeY = ymax - ymin;
    m_sizeX2 -= pxmin * unit;
    m_sizeY2 -= pymin * unit;
}

// like resize(), but never shrinks
template <class Cell, class Storage, const bool isClass>
void Map<Cell, Storage, isClass>::grow(double xmin, double ymin, double xmax, double ymax) {
    syncForWrite();
    const IntPoint imin = world2map(xmin, ymin), imax = world2map(xmax, ymax);
    if (isInside(imin) && isInside(imax)) return;
    const IntPoint lo = min(imin, IntPoint(0, 0)), hi = max(imax, IntPoint(m_mapSizeX - 1, m_mapSizeY - 1));
    const int unit = 1 << m_storage.getPatchMagnitude();
    const int pxmin = (int)floor((float)lo.x / unit), pxmax = (int)ceil((float)hi.x / unit);
    const int pymin = (int)floor((float)lo.y / unit), pymax = (int)ceil((float)hi.y / unit);
    m_storage.resize(pxmin, pymin, pxmax, pymax);
    m_mapSizeX = m_storage.getXSize() << m_storage.getPatchSize();
    m_mapSizeY = m_storage.getYSize() << m_storage.getPatchSize();
    m_worldSizeX = xmax - xmin;
    m_worldSizeY = ymax - ymin;
    m_sizeX2 -= pxmin * unit;
    m_sizeY2 -= pymin * unit;
}

template <class Cell, class Storage, const bool isClass>
DoubleArray2D* Map<Cell, Storage, isClass>::toDoubleArray() const {
    sync();
    DoubleArray2D* out = new DoubleArray2D(m_mapSizeX - 1, m_mapSizeY - 1);
    for (int x = 0; x < m_mapSizeX - 1; x++)
        for (int y = 0; y < m_mapSizeY - 1; y++) out->cell(x, y) = (double)cell(IntPoint(x, y));
    return out;
}

// plain double map spanning the same world rectangle (centre = middle of the first and last cell)
template <class Cell, class Storage, const bool isClass>
Map<double, DoubleArray2D, false>* Map<Cell, Storage, isClass>::toDoubleMap() const {
    sync();
    const Point lo = map2world(IntPoint(0, 0)), hi = map2world(m_mapSizeX - 1, m_mapSizeY - 1);
    const Point span = hi - lo, center = (hi + lo) * 0.5;
    Map<double, DoubleArray2D, false>* out = new Map<double, DoubleArray2D, false>(center, span.x, span.y, m_delta);
    for (int x = 0; x < m_mapSizeX - 1; x++)
        for (int y = 0; y < m_mapSizeY - 1; y++) out->cell(IntPoint(x, y)) = (double)cell(IntPoint(x, y));
    return out;
}

}  // namespace GMapping
#endif
