// gm_dist.cu -- multi-GPU particle sharding over PEER MEMORY: one process and one gm_ctx per GPU of one node, every rank's
// patch pool, patch directories and result buffers mapped into every other rank with CUDA IPC (NVLink 5 / NVSwitch
// underneath), no message-passing library on the data path.
//
// Rank r owns the global particles [r*n, (r+1)*n).  Scan matching and registration never leave the GPU that owns the
// particle.  The two exchanges of a scan:
//   (1) the "all-gather" of the scan-match results is the epilogue of k_scanmatch itself: every CTA stores its particle's
//       row {x, y, theta, score, l} into the gather buffer of EVERY rank (peer stores), the last CTA raises a flag in a small
//       host-memory segment all ranks share; a rank launches k_normalize as soon as every rank's flag shows the scan;
//   (2) when the filter resamples, a child whose parent lives on another rank PULLS the parent's map: a kernel on the
//       child's GPU reads the owner's patch directory and patches straight out of the owner's HBM (peer loads), keyed by
//       (owner, patch id) in a hash table so that a patch shared by several imported parents -- siblings of an earlier
//       resampling -- is copied once and stays shared here; the ordinary patch-reference remap then points the children at
//       the imported slots.  Two flag barriers bracket it: "my maps are complete" before anybody reads them, "I have pulled"
//       before an owner frees or rewrites patches.
// Systematic resampling keeps the index vector ordered, so few parents cross a slice boundary.
//
// The shared segment (POSIX shm, named after the 128-byte id every rank passes to gm_dist_init) carries the IPC handles and
// the flags; it is registered with the CUDA driver so that kernels can write the flags.  Two ranks may share one GPU
// (tests on a one-GPU box): nothing below depends on the peers being different devices.
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <map>
#include <set>

#include "gm_ctx.h"

namespace {

struct alignas(64) ShmRank {
    uint32_t ready;     // 1: this rank's handles are published; 2: it has mapped everybody's
    uint32_t dir_gen;   // generation of the directory allocations behind dir[] (re-opened by the peers when it changes)
    int32_t dir_cur, geom_cur, dir_cap;
    uint32_t n_local, max_particles;
    int32_t pool_cap, pool_cap_h;  // patch ids, and how many of them are hit-capable (the rest: visit counters only)
    cudaIpcMemHandle_t cells, visits, gather[2], dir[2], geom[2];
    alignas(64) uint32_t flag_rows;    // written by this rank's GPU: number of the last scan whose rows are in every gather buffer
    alignas(64) uint32_t flag_maps;    // host: this rank's maps are complete and published for resampling #k
    alignas(64) uint32_t flag_pulled;  // host: this rank has finished pulling for resampling #k
    alignas(64) uint32_t flag_error;   // a rank that fails says so here: its peers stop waiting
};
struct Shm {
    uint32_t magic, world;
    ShmRank r[gm::kMaxRanks];
};
constexpr uint32_t kShmMagic = 0x42323030u;

inline uint32_t load_acquire(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void store_release(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

}  // namespace

struct PeerMap {  // a peer's allocations as seen from this process
    int4* cells = nullptr; int* visits = nullptr; double* gather[2] = {nullptr, nullptr};
    int* dir[2] = {nullptr, nullptr}; gm::Geom* geom[2] = {nullptr, nullptr};
    uint32_t dir_gen = 0; bool dir_open = false;
};

struct gm_dist {
    int rank = 0, world = 1;
    Shm* shm = nullptr; Shm* shm_dev = nullptr; bool registered = false;
    char shm_name[64] = {0};
    double* gather[2] = {nullptr, nullptr};
    PeerMap peer[gm::kMaxRanks];
    uint32_t rows_seq = 0, resample_seq = 0;
    // what this rank last exported for its directories (the two buffers swap at every remap and are re-allocated on growth)
    int* exp_dir[2] = {nullptr, nullptr}; gm::Geom* exp_geom[2] = {nullptr, nullptr};
    // pull state
    unsigned long long* hkey = nullptr; int* hval = nullptr; unsigned hcap = 0;
    int *uniq_list = nullptr, *uniq_n = nullptr;
    uint2* d_imports = nullptr; size_t imports_cap = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    double timeout_ms = 120e3;
};

namespace gm {

struct PeerView { const int* dir; const Geom* geom; const int4* cells; const int* visits; int dir_cap, cap_h; };
struct PullArgs {
    PeerView peer[kMaxRanks];
    const uint2* imports;  // {owner rank, slot on the owner} per imported parent
    int first_slot;        // local slot of import 0
    Geom* geom; int* dir; int dir_cap;
    Pool pool; Counters* ctr;
    unsigned long long* hkey; int* hval; unsigned hmask;
    int *uniq_list, *uniq_n; int uniq_cap;
};
constexpr unsigned long long kEmptyKey = ~0ull;

// pass 1, one CTA per imported parent: geometry, and every allocated directory entry of the owner becomes a reference to the
// hash slot of (owner, patch id) -- encoded as -2 - slot until pass 3 -- so that equal patches meet in one slot
__global__ void __launch_bounds__(256) k_pull_scan(PullArgs A) {
    const int k = blockIdx.x;
    const uint2 imp = A.imports[k];
    const PeerView V = A.peer[imp.x];
    const Geom g = V.geom[imp.y];
    const int slot = A.first_slot + k;
    int* dirp = A.dir + (size_t)slot * A.dir_cap;
    const int nent = g.npx * g.npy;
    if (nent > A.dir_cap) { if (threadIdx.x == 0) atomicCAS(&A.ctr->error, 0, -2); return; }
    const int* rdir = V.dir + (size_t)imp.y * V.dir_cap;
    for (int e = threadIdx.x; e < A.dir_cap; e += blockDim.x) {
        int out = -1;
        const int rid = e < nent ? __ldcg(rdir + e) : -1;
        if (rid >= 0) {
            const unsigned long long key = ((unsigned long long)imp.x << 32) | (unsigned)rid;
            unsigned h = (unsigned)mix64(key) & A.hmask;
            for (unsigned probe = 0; probe <= A.hmask; probe++) {
                const unsigned long long prev = atomicCAS(A.hkey + h, kEmptyKey, key);
                if (prev == kEmptyKey) {
                    const int u = atomicAdd(A.uniq_n, 1);
                    if (u < A.uniq_cap) A.uniq_list[u] = (int)h; else atomicCAS(&A.ctr->error, 0, -3);  // more patches than the pool holds
                    out = -2 - (int)h;
                    break;
                }
                if (prev == key) { out = -2 - (int)h; break; }
                h = (h + 1) & A.hmask;
            }
            if (out == -1) atomicCAS(&A.ctr->error, 0, -3);  // table full: more distinct patches than the pool could hold anyway
        }
        dirp[e] = out;
    }
    if (threadIdx.x == 0) A.geom[slot] = g;
}

// pass 2, CTAs stride over the distinct patches: take a local patch of the owner's kind, copy the visit counters -- and, for a
// hit-capable patch, the cells -- out of the owner's HBM, rebuild the derived planes (means where n > 0, occupancy bitmap)
__global__ void __launch_bounds__(256) k_pull_copy(PullArgs A) {
    __shared__ int s_id;
    const int total = min(*A.uniq_n, A.uniq_cap);
    for (int u = blockIdx.x; u < total; u += gridDim.x) {
        const int h = A.uniq_list[u];
        const unsigned long long key = A.hkey[h];
        const PeerView V = A.peer[(unsigned)(key >> 32)];
        const unsigned rid = (unsigned)key;
        __syncthreads();
        const bool src_h = (int)rid < V.cap_h;
        if (threadIdx.x == 0) {
            s_id = pool_alloc(A.pool, A.ctr, src_h);
            A.hval[h] = s_id;
            if (s_id >= 0) A.pool.refcnt[s_id] = 0;
        }
        __syncthreads();
        const int id = s_id;
        if (id < 0) continue;
        const bool dst_h = id < A.pool.cap_h;  // (a free-space patch lands on a hit-capable id when the other stack is empty)
        const int4* src = V.cells + (size_t)rid * kPatchCells;
        const int* vsrc = V.visits + (size_t)rid * kPatchCells;
        int4 c[4]; int v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            c[i] = src_h ? __ldcg(src + i * 256 + threadIdx.x) : make_int4(0, 0, 0, 0);
            v[i] = __ldcg(vsrc + i * 256 + threadIdx.x);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int ci = i * 256 + threadIdx.x;  // a warp covers one patch row per trip
            A.pool.visits[(size_t)id * kPatchCells + ci] = v[i];
            if (!dst_h) continue;
            A.pool.cells[(size_t)id * kPatchCells + ci] = make_int4(c[i].x, c[i].y, c[i].z, 0);
            if (c[i].z > 0) A.pool.means[(size_t)id * kPatchCells + ci] = cell_mean(c[i].x, c[i].y, c[i].z);
            const unsigned bits = __ballot_sync(0xffffffffu, occ_default(c[i].z, v[i]));
            if ((threadIdx.x & 31) == 0) A.pool.occ[(size_t)id * kPatch + (ci >> kPatchMag)] = bits;
        }
    }
}

// pass 3: directory entries become patch ids, reference counts are counted
__global__ void __launch_bounds__(256) k_pull_fix(PullArgs A) {
    int* dirp = A.dir + (size_t)(A.first_slot + blockIdx.x) * A.dir_cap;
    for (int e = threadIdx.x; e < A.dir_cap; e += blockDim.x) {
        const int v = dirp[e];
        if (v > -2) continue;
        const int id = A.hval[-2 - v];
        dirp[e] = id;
        if (id >= 0) atomicAdd(&A.pool.refcnt[id], 1);
    }
}

// leave the hash table empty for the next resampling (only the slots that were used)
// (and count what was read from the peers: 4 KB of visit counters per patch, 16 KB of cells more for a hit-capable one;
// *bytes_out is zeroed by the host)
__global__ void k_pull_clear(PullArgs A, unsigned long long* bytes_out) {
    const int total = min(*A.uniq_n, A.uniq_cap);
    unsigned long long bytes = 0;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < total; u += gridDim.x * blockDim.x) {
        const int h = A.uniq_list[u];
        const unsigned long long key = A.hkey[h];
        bytes += (int)(unsigned)key < A.peer[(unsigned)(key >> 32)].cap_h ? kPatchCells * 20ull : kPatchCells * 4ull;
        A.hkey[h] = kEmptyKey;
    }
    if (bytes) atomicAdd(bytes_out, bytes);
}
__global__ void k_pull_reset(int* uniq_n) { *uniq_n = 0; }

}  // namespace gm

// the parents rank `rank` must import for idx_global, ascending (which groups them by owner), and the slot each
// local child takes its map from: a local slot, or n_local + position in the import list
static void plan_imports(uint32_t n, int rank, const uint32_t* idx, std::vector<uint32_t>& imports, std::vector<uint32_t>& local_idx) {
    const uint32_t first = (uint32_t)rank * n;
    std::set<uint32_t> need;
    for (uint32_t j = 0; j < n; j++) {
        const uint32_t p = idx[first + j];
        if (p < first || p >= first + n) need.insert(p);
    }
    imports.assign(need.begin(), need.end());
    std::map<uint32_t, uint32_t> pos;
    for (uint32_t k = 0; k < imports.size(); k++) pos[imports[k]] = n + k;
    local_idx.resize(n);
    for (uint32_t j = 0; j < n; j++) {
        const uint32_t p = idx[first + j];
        local_idx[j] = (p >= first && p < first + n) ? p - first : pos[p];
    }
}

// ---- flags
static int dist_fail(gm_ctx* ctx, int code, const char* fmt, ...) {
    char buf[400];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    gm_dist* d = ctx->dist;
    if (d && d->shm) store_release(&d->shm->r[d->rank].flag_error, 1u);  // peers stop waiting for this rank
    return fail(ctx, code, "%s", buf);
}

// wait until `field` of rank g reaches `want` (sequence numbers only grow)
template <class F>
static int wait_flag(gm_ctx* ctx, int g, F field, uint32_t want, const char* what) {
    gm_dist* d = ctx->dist;
    const double t0 = now_ms();
    for (unsigned spin = 0;; spin++) {
        if ((int32_t)(load_acquire(field(d->shm->r[g])) - want) >= 0) return GM_OK;
        if (load_acquire(&d->shm->r[g].flag_error)) return dist_fail(ctx, GM_ERR_CUDA, "%s: rank %d reported an error", what, g);
        if ((spin & 63) == 63) {
            sched_yield();
            if (now_ms() - t0 > d->timeout_ms) return dist_fail(ctx, GM_ERR_CUDA, "%s: rank %d did not arrive within %.0f s", what, g, d->timeout_ms / 1e3);
        }
    }
}

// export (or re-export) this rank's directory / geometry allocations; dir and dir_alt swap at every remap
static int publish_dirs(gm_ctx* ctx) {
    gm_dist* d = ctx->dist;
    ShmRank& me = d->shm->r[d->rank];
    auto index_of = [](int* const* have, const int* p) { return have[0] == p ? 0 : have[1] == p ? 1 : -1; };
    int cur = index_of(d->exp_dir, ctx->dir);
    if (cur < 0 || index_of(d->exp_dir, ctx->dir_alt) < 0) {  // first time, or re-allocated by a directory growth
        d->exp_dir[0] = ctx->dir; d->exp_dir[1] = ctx->dir_alt;
        CK(cudaIpcGetMemHandle(&me.dir[0], ctx->dir));
        CK(cudaIpcGetMemHandle(&me.dir[1], ctx->dir_alt));
        me.dir_gen++;
        cur = 0;
    }
    me.dir_cur = cur;
    me.dir_cap = ctx->dir_cap;
    if (!d->exp_geom[0]) {
        d->exp_geom[0] = ctx->geom; d->exp_geom[1] = ctx->geom_alt;
        CK(cudaIpcGetMemHandle(&me.geom[0], ctx->geom));
        CK(cudaIpcGetMemHandle(&me.geom[1], ctx->geom_alt));
    }
    me.geom_cur = d->exp_geom[0] == ctx->geom ? 0 : 1;
    return GM_OK;
}

static int open_peer_dirs(gm_ctx* ctx, int g) {
    gm_dist* d = ctx->dist;
    PeerMap& pm = d->peer[g];
    const ShmRank& sr = d->shm->r[g];
    if (pm.dir_open && pm.dir_gen == sr.dir_gen) return GM_OK;
    if (pm.dir_open) { cudaIpcCloseMemHandle(pm.dir[0]); cudaIpcCloseMemHandle(pm.dir[1]); pm.dir_open = false; }
    for (int k = 0; k < 2; k++) CK(cudaIpcOpenMemHandle(reinterpret_cast<void**>(&pm.dir[k]), sr.dir[k], cudaIpcMemLazyEnablePeerAccess));
    if (!pm.geom[0])
        for (int k = 0; k < 2; k++) CK(cudaIpcOpenMemHandle(reinterpret_cast<void**>(&pm.geom[k]), sr.geom[k], cudaIpcMemLazyEnablePeerAccess));
    pm.dir_gen = sr.dir_gen; pm.dir_open = true;
    return GM_OK;
}

// the parents of `idx_global` that live on other ranks are pulled into spare slots, then the ordinary patch-reference remap
static int dist_migrate_and_remap(gm_ctx* ctx, const uint32_t* idx_global, const char* who) {
    gm_dist* d = ctx->dist;
    const uint32_t n = ctx->n;
    const int world = d->world, rank = d->rank;
    const uint32_t N = n * (uint32_t)world;
    bool any_cross = false;
    for (uint32_t i = 0; i < N; i++) {
        if (idx_global[i] >= N) return dist_fail(ctx, GM_ERR_ARG, "%s: parent index %u >= %u", who, idx_global[i], N);
        any_cross = any_cross || idx_global[i] / n != i / n;
    }
    std::vector<uint32_t> imports, local_idx;
    plan_imports(n, rank, idx_global, imports, local_idx);
    ctx->stats.ms_migrate = 0.f; ctx->stats.migrated_particles = 0; ctx->stats.migrated_bytes = 0;
    if (!any_cross) return gmi_remap(ctx, local_idx.data(), n);  // every rank sees the same index vector: nobody waits for anybody
    const size_t nimp = imports.size();
    if (n + nimp > ctx->max_particles)
        return dist_fail(ctx, GM_ERR_ARG, "%s: %zu imported parents do not fit the %u spare particle slots", who, nimp, ctx->max_particles - n);
    std::vector<char> imports_from_me(world, 0);  // ranks that read this rank's maps
    for (uint32_t i = 0; i < N; i++)
        if (idx_global[i] / n == (uint32_t)rank && i / n != (uint32_t)rank) imports_from_me[i / n] = 1;

    // ---- barrier A: this rank's maps are complete (the caller finished any registration in flight) and published
    const uint32_t k = ++d->resample_seq;
    if (int r = publish_dirs(ctx)) return r;
    store_release(&d->shm->r[rank].flag_maps, k);
    int need_cap = ctx->dir_cap;
    std::set<int> owners;
    for (uint32_t p : imports) owners.insert((int)(p / n));
    for (int g : owners) {
        if (int r = wait_flag(ctx, g, [](ShmRank& s) { return &s.flag_maps; }, k, who)) return r;
        if (int r = open_peer_dirs(ctx, g)) return r;
        need_cap = std::max(need_cap, (int)d->shm->r[g].dir_cap);
    }
    CK(cudaEventRecord(d->e0, ctx->st));
    if (nimp) {
        if (need_cap > ctx->dir_cap) {
            if (int r = gmi_ensure_dir_cap(ctx, need_cap)) return r;  // (re-published at the next resampling; nobody reads it now)
        }
        // ---- pull: the owners' directories and patches are read in place, over NVLink
        if (nimp > d->imports_cap) { CK(dalloc(d->d_imports, nimp * 2)); d->imports_cap = nimp * 2; }
        std::vector<uint2> imp(nimp);
        for (size_t i = 0; i < nimp; i++) imp[i] = make_uint2(imports[i] / n, imports[i] % n);
        CK(cudaMemcpyAsync(d->d_imports, imp.data(), sizeof(uint2) * nimp, cudaMemcpyHostToDevice, ctx->st));
        if (int r = zero_counters(ctx)) return r;
        gm::PullArgs a{};
        for (int g : owners) {
            const ShmRank& sr = d->shm->r[g];
            const PeerMap& pm = d->peer[g];
            a.peer[g].dir = pm.dir[sr.dir_cur]; a.peer[g].geom = pm.geom[sr.geom_cur];
            a.peer[g].cells = pm.cells; a.peer[g].visits = pm.visits; a.peer[g].dir_cap = sr.dir_cap; a.peer[g].cap_h = sr.pool_cap_h;
        }
        a.imports = d->d_imports; a.first_slot = (int)n;
        a.geom = ctx->geom; a.dir = ctx->dir; a.dir_cap = ctx->dir_cap; a.pool = ctx->pool; a.ctr = ctx->d_counters;
        a.hkey = d->hkey; a.hval = d->hval; a.hmask = d->hcap - 1; a.uniq_list = d->uniq_list; a.uniq_n = d->uniq_n; a.uniq_cap = ctx->pool.cap;
        CK(cudaMemsetAsync(ctx->d_digest, 0, sizeof(unsigned long long), ctx->st));  // bytes read from the peers (k_pull_clear)
        gm::k_pull_scan<<<(unsigned)nimp, 256, 0, ctx->st>>>(a);
        gm::k_pull_copy<<<148 * 8, 256, 0, ctx->st>>>(a);
        gm::k_pull_fix<<<(unsigned)nimp, 256, 0, ctx->st>>>(a);
        gm::k_pull_clear<<<64, 256, 0, ctx->st>>>(a, reinterpret_cast<unsigned long long*>(ctx->d_digest));
        gm::k_pull_reset<<<1, 1, 0, ctx->st>>>(d->uniq_n);
        ctx->stats.launches += 5;
        if (int rc = check_launch(ctx, "patch migration")) return rc;
    }
    CK(cudaEventRecord(d->e1, ctx->st));
    // ---- pool exhaustion is reported here, before any child is pointed at a map with holes
    unsigned long long uniq = 0;
    if (nimp) {
        Counters h;
        CK(cudaMemcpyAsync(&uniq, ctx->d_digest, sizeof uniq, cudaMemcpyDeviceToHost, ctx->st));
        if (int r = read_counters(ctx, h)) return r;
        if (h.error == GM_ERR_POOL) return dist_fail(ctx, GM_ERR_POOL, "%s: patch pool exhausted (%d patch ids, %d of them able to hold end points) while importing %zu parents", who, ctx->pool.cap, ctx->pool.cap_h, nimp);
        if (h.error) return dist_fail(ctx, h.error, "%s: device-side error %d while importing parents", who, h.error);
    } else CK(cudaStreamSynchronize(ctx->st));
    // ---- barrier B: nobody frees or rewrites a patch while a peer may still be reading it
    store_release(&d->shm->r[rank].flag_pulled, k);
    for (int g = 0; g < world; g++)
        if (imports_from_me[g])
            if (int r = wait_flag(ctx, g, [](ShmRank& s) { return &s.flag_pulled; }, k, who)) return r;
    ctx->stats.migrated_particles = nimp;
    ctx->stats.migrated_bytes = uniq + (unsigned long long)nimp * ctx->dir_cap * 4;  // patch planes + directories
    ctx->stats.total_migrated_particles += ctx->stats.migrated_particles; ctx->stats.total_migrated_bytes += ctx->stats.migrated_bytes;
    // ---- the ordinary remap, children of imported parents read from the spare slots
    if (int r = gmi_remap(ctx, local_idx.data(), n + (uint32_t)nimp)) return r;
    CK(cudaEventElapsedTime(&ctx->stats.ms_migrate, d->e0, d->e1));
    return GM_OK;
}

extern "C" {

void gmi_dist_destroy(gm_ctx* ctx) {
    gm_dist* d = ctx->dist;
    if (!d) return;
    for (int g = 0; g < d->world; g++) {
        if (g == d->rank) continue;
        PeerMap& pm = d->peer[g];
        if (pm.cells) cudaIpcCloseMemHandle(pm.cells);
        if (pm.visits) cudaIpcCloseMemHandle(pm.visits);
        for (int k = 0; k < 2; k++) {
            if (pm.gather[k]) cudaIpcCloseMemHandle(pm.gather[k]);
            if (pm.dir[k]) cudaIpcCloseMemHandle(pm.dir[k]);
            if (pm.geom[k]) cudaIpcCloseMemHandle(pm.geom[k]);
        }
    }
    cudaFree(d->gather[0]); cudaFree(d->gather[1]);
    cudaFree(d->hkey); cudaFree(d->hval); cudaFree(d->uniq_list); cudaFree(d->uniq_n); cudaFree(d->d_imports);
    if (d->e0) cudaEventDestroy(d->e0);
    if (d->e1) cudaEventDestroy(d->e1);
    if (d->shm) {
        if (d->registered) cudaHostUnregister(d->shm);
        munmap(d->shm, sizeof(Shm));
        if (d->shm_name[0]) shm_unlink(d->shm_name);  // normally gone already (rank 0 unlinks once everybody is attached)
    }
    delete d;
    ctx->dist = nullptr;
    ctx->fail_hook = nullptr;
}

// 128 random bytes that name the ranks' shared segment; call on one rank and hand the bytes to all of them
int gm_dist_unique_id(void* id_out) {
    gm_ctx* ctx = nullptr;
    if (!id_out) return fail(ctx, GM_ERR_ARG, "gm_dist_unique_id: null argument");
    unsigned char* b = static_cast<unsigned char*>(id_out);
    FILE* f = fopen("/dev/urandom", "rb");
    const size_t got = f ? fread(b, 1, GM_DIST_ID_BYTES, f) : 0;
    if (f) fclose(f);
    if (got != GM_DIST_ID_BYTES) {
        unsigned long long x = (unsigned long long)getpid() * 0x9E3779B97F4A7C15ull ^ (unsigned long long)(now_ms() * 1e3);
        for (int i = 0; i < GM_DIST_ID_BYTES; i++) { x = x * 6364136223846793005ull + 1442695040888963407ull; b[i] = (unsigned char)(x >> 56); }
    }
    return GM_OK;
}

int gm_dist_init(gm_ctx* ctx, int rank, int world, const void* id_bytes) {
    if (!ctx) return GM_ERR_ARG;
    if (!id_bytes || world < 1 || world > gm::kMaxRanks || rank < 0 || rank >= world)
        return fail(ctx, GM_ERR_ARG, "gm_dist_init: bad argument (rank %d of %d, at most %d ranks)", rank, world, gm::kMaxRanks);
    if (!ctx->inited) return fail(ctx, GM_ERR_ARG, "gm_dist_init: call gm_init_particles first (the directories are exported)");
    if ((uint64_t)ctx->n * world > ctx->weights_cap) return fail(ctx, GM_ERR_ARG, "gm_dist_init: %u x %d particles, gm_config.global_particles is %u", ctx->n, world, ctx->weights_cap);
    if (int r = bind(ctx)) return r;
    gmi_dist_destroy(ctx);
    gm_dist* d = new gm_dist();
    d->rank = rank; d->world = world;
    if (const char* e = getenv("GM_B200_DIST_TIMEOUT_S")) d->timeout_ms = atof(e) * 1e3;
    ctx->dist = d;
    ctx->fail_hook = [](gm_ctx* c) { if (c->dist && c->dist->shm) store_release(&c->dist->shm->r[c->dist->rank].flag_error, 1u); };
    // ---- the shared segment
    const unsigned char* b = static_cast<const unsigned char*>(id_bytes);
    int len = snprintf(d->shm_name, sizeof d->shm_name, "/gmb200_");
    for (int i = 0; i < 16; i++) len += snprintf(d->shm_name + len, sizeof d->shm_name - len, "%02x", b[i]);
    const int fd = shm_open(d->shm_name, O_CREAT | O_RDWR, 0600);
    if (fd < 0) return fail(ctx, GM_ERR_CUDA, "gm_dist_init: shm_open(%s) failed", d->shm_name);
    if (ftruncate(fd, sizeof(Shm))) { close(fd); return fail(ctx, GM_ERR_CUDA, "gm_dist_init: ftruncate(%s) failed", d->shm_name); }
    void* m = mmap(nullptr, sizeof(Shm), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return fail(ctx, GM_ERR_CUDA, "gm_dist_init: mmap(%s) failed", d->shm_name);
    d->shm = static_cast<Shm*>(m);  // a fresh segment is all zeros
    CK(cudaHostRegister(d->shm, sizeof(Shm), cudaHostRegisterMapped | cudaHostRegisterPortable));
    d->registered = true;
    CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d->shm_dev), d->shm, 0));
    // ---- this rank's buffers and handles
    const size_t rows = (size_t)ctx->weights_cap * 5;
    for (int k = 0; k < 2; k++) { CK(dalloc(d->gather[k], rows)); CK(cudaMemsetAsync(d->gather[k], 0, rows * sizeof(double), ctx->st)); }
    d->hcap = 1024;
    while (d->hcap < 2u * (unsigned)ctx->pool.cap && d->hcap < (1u << 30)) d->hcap <<= 1;
    CK(dalloc(d->hkey, d->hcap)); CK(dalloc(d->hval, d->hcap)); CK(dalloc(d->uniq_list, (size_t)ctx->pool.cap + 1)); CK(dalloc(d->uniq_n, 1));
    CK(cudaMemsetAsync(d->hkey, 0xff, sizeof(unsigned long long) * d->hcap, ctx->st));
    CK(cudaMemsetAsync(d->uniq_n, 0, sizeof(int), ctx->st));
    CK(cudaEventCreate(&d->e0)); CK(cudaEventCreate(&d->e1));
    CK(cudaStreamSynchronize(ctx->st));
    ShmRank& me = d->shm->r[rank];
    me.n_local = ctx->n; me.max_particles = ctx->max_particles; me.pool_cap = ctx->pool.cap; me.pool_cap_h = ctx->pool.cap_h;
    CK(cudaIpcGetMemHandle(&me.cells, ctx->pool.cells));
    CK(cudaIpcGetMemHandle(&me.visits, ctx->pool.visits));
    for (int k = 0; k < 2; k++) CK(cudaIpcGetMemHandle(&me.gather[k], d->gather[k]));
    if (int r = publish_dirs(ctx)) return r;
    if (rank == 0) { d->shm->magic = kShmMagic; d->shm->world = (uint32_t)world; }
    store_release(&me.ready, 1u);
    // ---- everybody's handles
    for (int g = 0; g < world; g++) {
        if (int r = wait_flag(ctx, g, [](ShmRank& s) { return &s.ready; }, 1u, "gm_dist_init")) return r;
        if (g == rank) continue;
        const ShmRank& sr = d->shm->r[g];
        if (sr.n_local != ctx->n) return dist_fail(ctx, GM_ERR_ARG, "gm_dist_init: rank %d holds %u particles, this rank %u", g, sr.n_local, ctx->n);
        PeerMap& pm = d->peer[g];
        CK(cudaIpcOpenMemHandle(reinterpret_cast<void**>(&pm.cells), sr.cells, cudaIpcMemLazyEnablePeerAccess));
        CK(cudaIpcOpenMemHandle(reinterpret_cast<void**>(&pm.visits), sr.visits, cudaIpcMemLazyEnablePeerAccess));
        for (int k = 0; k < 2; k++) CK(cudaIpcOpenMemHandle(reinterpret_cast<void**>(&pm.gather[k]), sr.gather[k], cudaIpcMemLazyEnablePeerAccess));
        if (int r = open_peer_dirs(ctx, g)) return r;
    }
    store_release(&me.ready, 2u);
    for (int g = 0; g < world; g++)
        if (int r = wait_flag(ctx, g, [](ShmRank& s) { return &s.ready; }, 2u, "gm_dist_init")) return r;
    if (rank == 0) shm_unlink(d->shm_name);  // every rank is attached: the name can go, the mapping lives on
    d->shm_name[0] = 0;
    return GM_OK;
}

// where k_scanmatch's epilogue stores the result rows of this scan, and the flag it raises when the slice is complete
int gmi_dist_rows(gm_ctx* ctx, ScanMatchArgs* a) {
    gm_dist* d = ctx->dist;
    const uint32_t seq = ++d->rows_seq;
    const int buf = (int)(seq & 1u);  // a rank can be at most one scan ahead of a peer that still reads the previous rows
    a->nranks = d->world;
    for (int g = 0; g < d->world; g++) a->rows[g] = g == d->rank ? d->gather[buf] : d->peer[g].gather[buf];
    a->first_global = (unsigned)d->rank * ctx->n;
    a->done_counter = ctx->d_done;
    a->flag = &d->shm_dev->r[d->rank].flag_rows;
    a->seq = seq;
    return GM_OK;
}

int gmi_dist_wait_rows(gm_ctx* ctx, const double** rows) {
    gm_dist* d = ctx->dist;
    const uint32_t seq = d->rows_seq;
    if (int r = wait_flag(ctx, d->rank, [](ShmRank& s) { return &s.flag_rows; }, seq, "gm_scanmatch_weights")) return r;
    const double t0 = now_ms();
    for (int g = 0; g < d->world; g++)
        if (g != d->rank)
            if (int r = wait_flag(ctx, g, [](ShmRank& s) { return &s.flag_rows; }, seq, "gm_scanmatch_weights")) return r;
    // not device time: how long this rank waited, after its own kernel, for the slowest rank's rows
    ctx->stats.ms_allgather = (float)(now_ms() - t0);
    ctx->stats.total_ms[5] += ctx->stats.ms_allgather;
    *rows = d->gather[seq & 1u];
    return GM_OK;
}

int gm_dist_plan(uint32_t n_local, int rank, int world, const uint32_t* idx_global, uint32_t* parents_out, uint32_t* local_idx_out) {
    if (!idx_global || !n_local || rank < 0 || rank >= world) return GM_ERR_ARG;
    std::vector<uint32_t> imports, local_idx;
    plan_imports(n_local, rank, idx_global, imports, local_idx);
    if (parents_out) memcpy(parents_out, imports.data(), sizeof(uint32_t) * imports.size());
    if (local_idx_out) memcpy(local_idx_out, local_idx.data(), sizeof(uint32_t) * n_local);
    return (int)imports.size();
}

int gm_dist_register(gm_ctx* ctx, const double* ranges, const double* poses, const uint32_t* idx_global) {
    if (int r = ready(ctx, true)) return r;
    if (!ctx->dist) return fail(ctx, GM_ERR_ARG, "gm_dist_register: gm_dist_init has not been called");
    if (!ranges || !poses) return fail(ctx, GM_ERR_ARG, "gm_dist_register: null argument");
    ctx->stats.ms_remap = 0.f; ctx->stats.ms_migrate = 0.f; ctx->stats.migrated_particles = 0; ctx->stats.migrated_bytes = 0;
    if (idx_global)
        if (int r = dist_migrate_and_remap(ctx, idx_global, "gm_dist_register")) return r;
    if (int r = gmi_stage(ctx, ranges, poses)) return r;
    CK(cudaEventRecord(ctx->evr0, ctx->st));
    if (int rc = gmi_bounds_and_resize(ctx)) return rc;
    return gmi_register_update(ctx);
}

int gm_dist_copy_particles(gm_ctx* ctx, const uint32_t* idx_global) {
    if (int r = ready(ctx, true)) return r;  // completes a registration in flight: the copies take the registered maps
    if (!ctx->dist) return fail(ctx, GM_ERR_ARG, "gm_dist_copy_particles: gm_dist_init has not been called");
    if (!idx_global) return fail(ctx, GM_ERR_ARG, "gm_dist_copy_particles: null argument");
    ctx->stats.ms_remap = 0.f;
    if (int r = dist_migrate_and_remap(ctx, idx_global, "gm_dist_copy_particles")) return r;
    ctx->stats.total_ms[2] += ctx->stats.ms_remap; ctx->stats.total_ms[4] += ctx->stats.ms_migrate;
    return GM_OK;
}

}  // extern "C"
