// devicemap.h -- glue between the host API objects and the C ABI (internal to the host library).
#pragma once
#include <gmapping/scanmatcher/scanmatcher.h>
#include <gmapping/scanmatcher/smmap.h>

#include <memory>
#include <string>

#include "gm_b200.h"

namespace GMapping {
namespace b200 {

// C-ABI failures are fatal, like the reference's asserts: message on cerr, then abort()
[[noreturn]] void fatal(const std::string& what);

// RAII owner of one gm_ctx
class DeviceContext {
  public:
    DeviceContext(unsigned int maxParticles, unsigned long long poolPatches, unsigned int globalParticles = 0);
    // second context on the same GPU with the same maps (gm_clone)
    explicit DeviceContext(const DeviceContext& other, int);
    ~DeviceContext();
    DeviceContext(const DeviceContext&) = delete;
    DeviceContext& operator=(const DeviceContext&) = delete;
    void check(int rc, const char* what) const;
    // push the matcher's laser table and parameters when they differ from what the context has
    void configure(const ScanMatcher& m);
    gm_ctx* h = nullptr;
    unsigned int particles = 0;
    // GridSlamProcessor's early registration: the device maps hold a scan the reference has not registered yet
    bool mapsAhead = false;

  private:
    bool m_haveLaser = false, m_haveParams = false;
    unsigned int m_beams = 0;
    double m_angles[LASER_MAXBEAMS];
    double m_laserPose[3];
    gm_match_params m_params;
};

// a ScanMatcherMap whose cells live in slot `slot` of `ctx`
struct DeviceMapResidency : public MapResidency {
    std::shared_ptr<DeviceContext> ctx;
    unsigned int slot = 0;
    void pull(void* hostMap) override;
    void push(const void* hostMap) override;
    // upload the mirror when the host may have written to it since the last transfer
    inline void flush(const ScanMatcherMap& map) { if (hostDirty) push(&map); }
};

int deviceOrdinal();  // GM_B200_DEVICE, else LOCAL_RANK, else 0

// residency of a map, or null for a plain host map
DeviceMapResidency* residencyOf(const ScanMatcherMap& map);
// residency of a map; a plain host map is uploaded into a single-map context first
DeviceMapResidency* makeResident(const ScanMatcherMap& map, const ScanMatcher& m);

}  // namespace b200
}  // namespace GMapping
