// devicemap.cpp -- device residency of ScanMatcherMap objects (see grid/map.h, devicemap.h)
#include "devicemap.h"

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

namespace GMapping {
namespace b200 {

void fatal(const std::string& what) {
    std::cerr << "gmapping-b200: " << what << std::endl;
    std::abort();
}

int deviceOrdinal() {
    if (const char* e = std::getenv("GM_B200_DEVICE")) return std::atoi(e);
    if (const char* e = std::getenv("LOCAL_RANK")) return std::atoi(e);
    return 0;
}

DeviceContext::DeviceContext(unsigned int maxParticles, unsigned long long poolPatches, unsigned int globalParticles) {
    gm_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.device = deviceOrdinal();
    cfg.max_particles = maxParticles;
    cfg.pool_patches = poolPatches;
    cfg.global_particles = globalParticles;
    if (const char* e = std::getenv("GM_B200_POOL_PATCHES")) cfg.pool_patches = std::strtoull(e, nullptr, 10);
    const int rc = gm_create(&h, &cfg);
    if (rc) fatal(std::string("gm_create: ") + gm_last_error(nullptr));
    std::memset(&m_params, 0, sizeof m_params);
}

DeviceContext::DeviceContext(const DeviceContext& o, int) : particles(o.particles) {
    const int rc = gm_clone(o.h, &h);
    if (rc) fatal(std::string("gm_clone: ") + gm_last_error(o.h));
    m_haveLaser = o.m_haveLaser; m_haveParams = o.m_haveParams; m_beams = o.m_beams;
    std::memcpy(m_angles, o.m_angles, sizeof m_angles);
    std::memcpy(m_laserPose, o.m_laserPose, sizeof m_laserPose);
    m_params = o.m_params;
}

DeviceContext::~DeviceContext() {
    if (h) gm_destroy(h);
}

void DeviceContext::check(int rc, const char* what) const {
    if (rc) fatal(std::string(what) + ": " + gm_last_error(h));
}

void DeviceContext::configure(const ScanMatcher& m) {
    const OrientedPoint lp = m.getlaserPose();
    const double lpose[3] = {lp.x, lp.y, lp.theta};
    const unsigned int nb = m.laserBeams();
    if (!m_haveLaser || nb != m_beams || std::memcmp(m_angles, m.laserAngles(), sizeof(double) * nb) || std::memcmp(m_laserPose, lpose, sizeof lpose)) {
        check(gm_set_laser(h, nb, m.laserAngles(), lpose), "gm_set_laser");
        m_beams = nb;
        std::memcpy(m_angles, m.laserAngles(), sizeof(double) * nb);
        std::memcpy(m_laserPose, lpose, sizeof lpose);
        m_haveLaser = true;
    }
    gm_match_params p;
    m.exportParams(&p);
    if (!m_haveParams || std::memcmp(&p, &m_params, sizeof p)) {
        check(gm_set_matching(h, &p), "gm_set_matching");
        m_params = p;
        m_haveParams = true;
    }
}

// refresh the host mirror: geometry + every allocated patch
void DeviceMapResidency::pull(void* hostMap) {
    ScanMatcherMap& map = *static_cast<ScanMatcherMap*>(hostMap);
    if (ctx->mapsAhead)
        b200::fatal("Particle::map was read between scan matching and registration (from onResampleUpdate?) while early registration is on: "
                    "the device map already holds the current scan there; call GridSlamProcessor::setearlyRegistration(false) to read maps in that window");
    stale = false;  // first: the accessors used below call sync() again
    gm_map_info info;
    ctx->check(gm_map_info_get(ctx->h, slot, &info), "gm_map_info_get");
    const int n = info.allocated_patches;
    std::vector<int32_t> coords((size_t)2 * n);
    std::vector<gm_cell> cells((size_t)1024 * n);
    if (n) ctx->check(gm_download_map(ctx->h, slot, &info, coords.data(), cells.data(), n), "gm_download_map");
    HierarchicalArray2D<PointAccumulator> fresh(info.patches_x << 5, info.patches_y << 5);
    for (int k = 0; k < n; k++) {
        Array2D<PointAccumulator>* patch = new Array2D<PointAccumulator>(32, 32);
        const gm_cell* src = cells.data() + (size_t)1024 * k;
        for (int lx = 0; lx < 32; lx++) std::memcpy(static_cast<void*>(patch->m_cells[lx]), src + 32 * lx, sizeof(gm_cell) * 32);
        fresh.m_cells[coords[2 * k]][coords[2 * k + 1]] = autoptr<Array2D<PointAccumulator> >(patch);
    }
    map.rawStorage() = fresh;
    map.setGeometry(info.size_x2, info.size_y2, info.patches_x, info.patches_y);
}

// the host mirror replaces the device copy (gm_upload_map): geometry + every allocated patch
static void uploadMap(DeviceContext& ctx, unsigned int slot, const ScanMatcherMap& map, const HierarchicalArray2D<PointAccumulator>& st,
                      int sizeX2, int sizeY2) {
    const int npx = st.getXSize(), npy = st.getYSize();
    std::vector<int32_t> coords;
    std::vector<gm_cell> cells;
    for (int px = 0; px < npx; px++)
        for (int py = 0; py < npy; py++) {
            if (!st.m_cells[px][py]) continue;
            const Array2D<PointAccumulator>& patch = *st.m_cells[px][py];
            coords.push_back(px); coords.push_back(py);
            const size_t at = cells.size();
            cells.resize(at + 1024);
            for (int lx = 0; lx < 32; lx++) std::memcpy(static_cast<void*>(cells.data() + at + 32 * lx), patch.m_cells[lx], sizeof(gm_cell) * 32);
        }
    const Point c = map.getCenter();
    gm_map_info info;
    std::memset(&info, 0, sizeof info);
    info.center_x = c.x; info.center_y = c.y; info.delta = map.getDelta();
    info.size_x2 = sizeX2; info.size_y2 = sizeY2; info.patches_x = npx; info.patches_y = npy;
    ctx.check(gm_upload_map(ctx.h, slot, &info, coords.data(), cells.data(), (uint32_t)(coords.size() / 2)), "gm_upload_map");
}

void DeviceMapResidency::push(const void* hostMap) {
    const ScanMatcherMap& map = *static_cast<const ScanMatcherMap*>(hostMap);
    if (stale) b200::fatal("a device-resident map was written on the host while its device copy was newer");  // cannot happen: writers sync first
    hostDirty = false;  // first: the const accessors below must not recurse
    uploadMap(*ctx, slot, map, map.storage(), map.getSizeX2(), map.getSizeY2());
}

DeviceMapResidency* residencyOf(const ScanMatcherMap& map) { return dynamic_cast<DeviceMapResidency*>(map.residency().get()); }

DeviceMapResidency* makeResident(const ScanMatcherMap& map, const ScanMatcher& m) {
    if (DeviceMapResidency* r = residencyOf(map)) {
        r->ctx->configure(m);
        r->flush(map);
        return r;
    }
    // a plain host map: give it a single-map context and upload what it holds
    const HierarchicalArray2D<PointAccumulator>& st = map.storage();
    const int npx = st.getXSize(), npy = st.getYSize();
    unsigned long long pool = 2ull * npx * npy;
    if (pool < 4096) pool = 4096;
    if (pool > 262144) pool = 262144;
    std::shared_ptr<DeviceContext> ctx(new DeviceContext(1, pool));
    ctx->configure(m);
    const Point c = map.getCenter();
    const double center[2] = {c.x, c.y};
    ctx->check(gm_init_particles(ctx->h, 1, center, map.getMapSizeX() * map.getDelta(), map.getMapSizeY() * map.getDelta(), map.getDelta()),
               "gm_init_particles");
    ctx->particles = 1;
    uploadMap(*ctx, 0, map, st, map.getSizeX2(), map.getSizeY2());
    std::shared_ptr<DeviceMapResidency> r(new DeviceMapResidency);
    r->ctx = ctx;
    r->slot = 0;
    r->stale = false;
    map.setResidency(r);
    return r.get();
}

}  // namespace b200
}  // namespace GMapping
