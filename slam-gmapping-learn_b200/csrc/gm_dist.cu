// gm_dist.cu -- multi-GPU particle sharding: NCCL all-gather of per-particle results and migration of resampled
// parents' maps between ranks (one process and one gm_ctx per GPU, NVLink 5 / NVSwitch underneath NCCL).
//
// Rank r owns the global particles [r*n, (r+1)*n).  Scan matching and registration never leave the GPU that owns
// the particle.  The only exchange is (1) an all-gather of poses / scores / log-likelihoods (5 doubles per particle)
// and (2), when the filter resamples, the maps of parents whose children live on another rank: the owner packs the
// parent's patch directory and its patches into one buffer per destination, ncclSend/ncclRecv moves it, the receiver
// installs it in a spare particle slot; the ordinary patch-reference remap then points the children at it.
// Systematic resampling keeps indexes ordered, so few parents cross a slice boundary.
//
// NCCL is resolved with dlopen at run time (torch's bundled libnccl.so.2 when the process already loaded it, the
// system one otherwise): single-GPU users do not need it.
#include <dlfcn.h>
#include <nccl.h>

#include <map>
#include <set>

#include "gm_ctx.h"

namespace {

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string error;
    bool load() {
        if (lib) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { error = std::string("dlopen libnccl.so.2: ") + dlerror(); return false; }
#define SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(lib, name)); if (!field) { error = std::string("dlsym ") + name; return false; }
        SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
        SYM(AllGather, "ncclAllGather") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart")
        SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
        return true;
    }
};
NcclApi g_nccl;

}  // namespace

struct gm_dist {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    double *d_gin = nullptr, *d_gout = nullptr; size_t gather_cap = 0;
    unsigned char *d_send = nullptr, *d_recv = nullptr; size_t send_cap = 0, recv_cap = 0;
    int *d_slots = nullptr, *d_meta_send = nullptr, *d_meta_recv = nullptr; size_t meta_cap = 0;
    unsigned long long *d_off = nullptr; size_t off_cap = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

#define NK(call)                                                                                                      \
    do {                                                                                                              \
        ncclResult_t r_ = (call);                                                                                     \
        if (r_ != ncclSuccess) return fail(ctx, GM_ERR_CUDA, "%s: %s", #call, g_nccl.GetErrorString(r_));             \
    } while (0)

namespace gm {

constexpr int kMetaInts = 4;      // per migrated parent: patch count, npx, npy, reserved
constexpr int kHeaderBytes = 32;  // packed parent: int[8] = sx2, sy2, npx, npy, count, 0, 0, 0

__host__ __device__ inline unsigned long long packed_bytes(int count) {
    return kHeaderBytes + (((unsigned long long)count * 4 + 15) & ~15ull) + (unsigned long long)count * kPatchCells * 16;
}

// per parent to send: how many patches it holds and its directory shape
__global__ void k_dist_count(const int* __restrict__ slots, const Geom* __restrict__ geom, const int* __restrict__ dir, int dir_cap, int* __restrict__ meta) {
    __shared__ int s_count;
    const int s = blockIdx.x, slot = slots[s];
    const Geom g = geom[slot];
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    int c = 0;
    for (int e = threadIdx.x; e < g.npx * g.npy; e += blockDim.x) c += dir[(size_t)slot * dir_cap + e] >= 0;
    if (c) atomicAdd(&s_count, c);
    __syncthreads();
    if (threadIdx.x == 0) { meta[kMetaInts * s] = s_count; meta[kMetaInts * s + 1] = g.npx; meta[kMetaInts * s + 2] = g.npy; meta[kMetaInts * s + 3] = 0; }
}

// pack one parent per block: header, directory entries that are allocated, their patches
__global__ void __launch_bounds__(512) k_dist_pack(const int* __restrict__ slots, const Geom* __restrict__ geom, const int* __restrict__ dir, int dir_cap,
                                                   const int4* __restrict__ pool, const int* __restrict__ visits, const unsigned long long* __restrict__ off, unsigned char* __restrict__ buf) {
    __shared__ int s_n;
    const int s = blockIdx.x, slot = slots[s];
    const Geom g = geom[slot];
    unsigned char* base = buf + off[s];
    int* header = reinterpret_cast<int*>(base);
    int* entries = reinterpret_cast<int*>(base + kHeaderBytes);
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const int* dirp = dir + (size_t)slot * dir_cap;
    for (int e = threadIdx.x; e < g.npx * g.npy; e += blockDim.x)
        if (dirp[e] >= 0) entries[atomicAdd(&s_n, 1)] = e;
    __syncthreads();
    const int count = s_n;
    if (threadIdx.x == 0) { header[0] = g.sx2; header[1] = g.sy2; header[2] = g.npx; header[3] = g.npy; header[4] = count; header[5] = header[6] = header[7] = 0; }
    int4* cells = reinterpret_cast<int4*>(base + kHeaderBytes + (((unsigned long long)count * 4 + 15) & ~15ull));
    for (long long i = threadIdx.x; i < (long long)count * kPatchCells; i += blockDim.x) {
        const int k = (int)(i >> 10);
        const size_t ci = (size_t)dirp[entries[k]] * kPatchCells + (i & 1023);
        int4 c = __ldcg(pool + ci);
        c.w = __ldcg(visits + ci);  // on the wire a cell is the reference's PointAccumulator: {acc.x, acc.y, n, visits}
        cells[i] = c;
    }
}

// install one received parent per block into particle slot first_slot + block: geometry, a fresh directory, private patches
__global__ void __launch_bounds__(512) k_dist_install(int first_slot, Geom* __restrict__ geom, int* __restrict__ dir, int dir_cap, Pool pool,
                                                      const unsigned long long* __restrict__ off, const unsigned char* __restrict__ buf, Counters* ctr) {
    extern __shared__ int s_ids[];  // patch id per entry (count ints)
    const int k = blockIdx.x, slot = first_slot + k;
    const unsigned char* base = buf + off[k];
    const int* header = reinterpret_cast<const int*>(base);
    const int* entries = reinterpret_cast<const int*>(base + kHeaderBytes);
    const int count = header[4];
    int* dirp = dir + (size_t)slot * dir_cap;
    for (int e = threadIdx.x; e < dir_cap; e += blockDim.x) dirp[e] = -1;
    if (threadIdx.x == 0) { Geom g; g.sx2 = header[0]; g.sy2 = header[1]; g.npx = header[2]; g.npy = header[3]; geom[slot] = g; }
    __syncthreads();
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const int id = pool_alloc(pool, ctr);
        s_ids[i] = id;
        if (id >= 0) { dirp[entries[i]] = id; pool.refcnt[id] = 1; }
    }
    __syncthreads();
    const int4* cells = reinterpret_cast<const int4*>(base + kHeaderBytes + (((unsigned long long)count * 4 + 15) & ~15ull));
    for (long long i = threadIdx.x; i < (long long)count * kPatchCells; i += blockDim.x) {
        const int id = s_ids[i >> 10];
        const int4 c = cells[i];
        if (id >= 0) { pool.cells[(size_t)id * kPatchCells + (i & 1023)] = make_int4(c.x, c.y, c.z, 0); pool.visits[(size_t)id * kPatchCells + (i & 1023)] = c.w; pool.means[(size_t)id * kPatchCells + (i & 1023)] = cell_mean(c.x, c.y, c.z); }
        // 512 threads, count * 1024 cells: every warp covers one whole patch row per trip
        const unsigned bits = __ballot_sync(0xffffffffu, occ_default(c.z, c.w));
        if (id >= 0 && (i & 31) == 0) pool.occ[(size_t)id * kPatch + ((i & 1023) >> kPatchMag)] = bits;
    }
}

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

// the parents of `idx_global` that live on other ranks migrate into spare slots, then the ordinary patch-reference remap
static int dist_migrate_and_remap(gm_ctx* ctx, const uint32_t* idx_global, const char* who) {
    gm_dist* d = ctx->dist;
    const uint32_t n = ctx->n;
    const int world = d->world, rank = d->rank;
    const uint32_t N = n * (uint32_t)world;
    for (uint32_t i = 0; i < N; i++)
        if (idx_global[i] >= N) return fail(ctx, GM_ERR_ARG, "%s: parent index %u >= %u", who, idx_global[i], N);
    // ---- plan: imports of this rank (grouped by owner) and, for every other rank, what it imports from us
    std::vector<uint32_t> imports, local_idx;
    plan_imports(n, rank, idx_global, imports, local_idx);
    if (n + imports.size() > ctx->max_particles)
        return fail(ctx, GM_ERR_ARG, "%s: %zu imported parents do not fit the %u spare particle slots", who, imports.size(), ctx->max_particles - n);
    std::vector<std::vector<int> > sends(world);  // local slots to send to each rank, ascending
    size_t nsend = 0;
    for (int g = 0; g < world; g++) {
        if (g == rank) continue;
        std::vector<uint32_t> imp_g, unused;
        plan_imports(n, g, idx_global, imp_g, unused);
        for (uint32_t p : imp_g)
            if (p / n == (uint32_t)rank) sends[g].push_back((int)(p - (uint32_t)rank * n));
        nsend += sends[g].size();
    }
    std::vector<size_t> recv_from(world, 0);  // imports per owner; the import list is ascending, hence grouped by owner
    for (uint32_t p : imports) recv_from[p / n]++;
    const size_t nrecv = imports.size();

    CK(cudaEventRecord(d->e0, ctx->st));
    if (nsend + nrecv) {
        if (nsend + nrecv > d->meta_cap) {
            const size_t cap = (nsend + nrecv) * 2;
            CK(dalloc(d->d_slots, cap)); CK(dalloc(d->d_meta_send, cap * gm::kMetaInts)); CK(dalloc(d->d_meta_recv, cap * gm::kMetaInts));
            CK(dalloc(d->d_off, cap));
            d->meta_cap = cap;
        }
        // ---- 1. patch counts of what we send; exchange the counts
        std::vector<int> slots;
        for (int g = 0; g < world; g++) slots.insert(slots.end(), sends[g].begin(), sends[g].end());
        std::vector<int> meta_send(nsend * gm::kMetaInts), meta_recv(nrecv * gm::kMetaInts);
        if (nsend) {
            CK(cudaMemcpyAsync(d->d_slots, slots.data(), sizeof(int) * nsend, cudaMemcpyHostToDevice, ctx->st));
            gm::k_dist_count<<<(unsigned)nsend, 256, 0, ctx->st>>>(d->d_slots, ctx->geom, ctx->dir, ctx->dir_cap, d->d_meta_send);
            ctx->stats.launches++;
        }
        NK(g_nccl.GroupStart());
        {
            size_t so = 0, ro = 0;
            for (int g = 0; g < world; g++) {
                if (g == rank) continue;
                if (!sends[g].empty()) NK(g_nccl.Send(d->d_meta_send + so * gm::kMetaInts, sends[g].size() * gm::kMetaInts, ncclInt32, g, d->comm, ctx->st));
                if (recv_from[g]) NK(g_nccl.Recv(d->d_meta_recv + ro * gm::kMetaInts, recv_from[g] * gm::kMetaInts, ncclInt32, g, d->comm, ctx->st));
                so += sends[g].size(); ro += recv_from[g];
            }
        }
        NK(g_nccl.GroupEnd());
        if (nsend) CK(cudaMemcpyAsync(meta_send.data(), d->d_meta_send, sizeof(int) * meta_send.size(), cudaMemcpyDeviceToHost, ctx->st));
        if (nrecv) CK(cudaMemcpyAsync(meta_recv.data(), d->d_meta_recv, sizeof(int) * meta_recv.size(), cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        // ---- 2. byte offsets, buffers, pack, exchange the data
        std::vector<unsigned long long> send_off(nsend + 1, 0), recv_off(nrecv + 1, 0);
        for (size_t s = 0; s < nsend; s++) send_off[s + 1] = send_off[s] + gm::packed_bytes(meta_send[gm::kMetaInts * s]);
        int need_cap = 0, max_count = 0;
        for (size_t k = 0; k < nrecv; k++) {
            recv_off[k + 1] = recv_off[k] + gm::packed_bytes(meta_recv[gm::kMetaInts * k]);
            need_cap = std::max(need_cap, meta_recv[gm::kMetaInts * k + 1] * meta_recv[gm::kMetaInts * k + 2]);
            max_count = std::max(max_count, meta_recv[gm::kMetaInts * k]);
        }
        if (send_off[nsend] > d->send_cap) { CK(dalloc(d->d_send, (size_t)send_off[nsend])); d->send_cap = send_off[nsend]; }
        if (recv_off[nrecv] > d->recv_cap) { CK(dalloc(d->d_recv, (size_t)recv_off[nrecv])); d->recv_cap = recv_off[nrecv]; }
        if (nsend) {
            CK(cudaMemcpyAsync(d->d_off, send_off.data(), sizeof(unsigned long long) * nsend, cudaMemcpyHostToDevice, ctx->st));
            gm::k_dist_pack<<<(unsigned)nsend, 512, 0, ctx->st>>>(d->d_slots, ctx->geom, ctx->dir, ctx->dir_cap, ctx->pool.cells, ctx->pool.visits, d->d_off, d->d_send);
            ctx->stats.launches++;
        }
        NK(g_nccl.GroupStart());
        {
            size_t so = 0, ro = 0;
            for (int g = 0; g < world; g++) {
                if (g == rank) continue;
                const size_t ns = sends[g].size(), nr = recv_from[g];
                if (ns) NK(g_nccl.Send(d->d_send + send_off[so], (size_t)(send_off[so + ns] - send_off[so]), ncclChar, g, d->comm, ctx->st));
                if (nr) NK(g_nccl.Recv(d->d_recv + recv_off[ro], (size_t)(recv_off[ro + nr] - recv_off[ro]), ncclChar, g, d->comm, ctx->st));
                so += ns; ro += nr;
            }
        }
        NK(g_nccl.GroupEnd());
        // ---- 3. install the received parents in the spare slots n .. n + nrecv - 1
        if (nrecv) {
            CK(cudaStreamSynchronize(ctx->st));  // d_off is reused below
            if (int r = gmi_ensure_dir_cap(ctx, need_cap)) return r;
            CK(cudaMemcpyAsync(d->d_off, recv_off.data(), sizeof(unsigned long long) * nrecv, cudaMemcpyHostToDevice, ctx->st));
            const size_t smem = (size_t)std::max(max_count, 1) * sizeof(int);
            if (smem > 48 * 1024) CK(cudaFuncSetAttribute(gm::k_dist_install, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            gm::k_dist_install<<<(unsigned)nrecv, 512, smem, ctx->st>>>((int)n, ctx->geom, ctx->dir, ctx->dir_cap, ctx->pool, d->d_off, d->d_recv, ctx->d_counters);
            ctx->stats.launches++;
        }
        if (int rc = check_launch(ctx, "patch migration")) return rc;
        ctx->stats.migrated_particles = nrecv;
        ctx->stats.migrated_bytes = recv_off[nrecv];
    }
    CK(cudaEventRecord(d->e1, ctx->st));
    // ---- 4. the ordinary remap, children of imported parents read from the spare slots
    if (int r = gmi_remap(ctx, local_idx.data(), n + (uint32_t)nrecv)) return r;
    CK(cudaEventElapsedTime(&ctx->stats.ms_migrate, d->e0, d->e1));
    return GM_OK;
}

extern "C" {

void gmi_dist_destroy(gm_ctx* ctx) {
    gm_dist* d = ctx->dist;
    if (!d) return;
    if (d->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm);
    cudaFree(d->d_gin); cudaFree(d->d_gout); cudaFree(d->d_send); cudaFree(d->d_recv);
    cudaFree(d->d_slots); cudaFree(d->d_meta_send); cudaFree(d->d_meta_recv); cudaFree(d->d_off);
    if (d->e0) cudaEventDestroy(d->e0);
    if (d->e1) cudaEventDestroy(d->e1);
    delete d;
    ctx->dist = nullptr;
}

int gm_dist_unique_id(void* id_out) {
    gm_ctx* ctx = nullptr;
    if (!id_out) return fail(ctx, GM_ERR_ARG, "gm_dist_unique_id: null argument");
    if (!g_nccl.load()) return fail(ctx, GM_ERR_CUDA, "%s", g_nccl.error.c_str());
    static_assert(sizeof(ncclUniqueId) == GM_DIST_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return GM_OK;
}

int gm_dist_init(gm_ctx* ctx, int rank, int world, const void* id_bytes) {
    if (!ctx) return GM_ERR_ARG;
    if (!id_bytes || world < 1 || rank < 0 || rank >= world) return fail(ctx, GM_ERR_ARG, "gm_dist_init: bad argument");
    if (int r = bind(ctx)) return r;
    if (!g_nccl.load()) return fail(ctx, GM_ERR_CUDA, "%s", g_nccl.error.c_str());
    gmi_dist_destroy(ctx);
    gm_dist* d = new gm_dist();
    d->rank = rank; d->world = world;
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    ctx->dist = d;
    NK(g_nccl.CommInitRank(&d->comm, world, id, rank));
    CK(cudaEventCreate(&d->e0)); CK(cudaEventCreate(&d->e1));
    return GM_OK;
}

int gm_dist_allgather(gm_ctx* ctx, const double* local, uint32_t count, double* all) {
    if (!ctx || !ctx->dist) return fail(ctx, GM_ERR_ARG, "gm_dist_allgather: gm_dist_init has not been called");
    if (!local || !all || !count) return fail(ctx, GM_ERR_ARG, "gm_dist_allgather: bad argument");
    if (int r = bind(ctx)) return r;
    if (int r = gmi_register_finish(ctx)) return r;
    gm_dist* d = ctx->dist;
    const size_t total = (size_t)count * d->world;
    if (total > d->gather_cap) { CK(dalloc(d->d_gin, count)); CK(dalloc(d->d_gout, total)); d->gather_cap = total; }
    CK(cudaMemcpyAsync(d->d_gin, local, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaEventRecord(d->e0, ctx->st));
    NK(g_nccl.AllGather(d->d_gin, d->d_gout, count, ncclDouble, d->comm, ctx->st));
    CK(cudaEventRecord(d->e1, ctx->st));
    CK(cudaMemcpyAsync(all, d->d_gout, sizeof(double) * total, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaEventElapsedTime(&ctx->stats.ms_allgather, d->e0, d->e1));
    ctx->stats.total_ms[5] += ctx->stats.ms_allgather;
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
    if (int r = gmi_stage(ctx, ranges, poses)) return r;
    ctx->stats.ms_migrate = 0.f; ctx->stats.migrated_particles = 0; ctx->stats.migrated_bytes = 0;
    if (idx_global)
        if (int r = dist_migrate_and_remap(ctx, idx_global, "gm_dist_register")) return r;
    CK(cudaEventRecord(ctx->ev0, ctx->st));
    if (int rc = gmi_bounds_and_resize(ctx)) return rc;
    return gmi_register_update(ctx);
}

int gm_dist_copy_particles(gm_ctx* ctx, const uint32_t* idx_global) {
    if (int r = ready(ctx, true)) return r;  // completes a registration in flight: the copies take the registered maps
    if (!ctx->dist) return fail(ctx, GM_ERR_ARG, "gm_dist_copy_particles: gm_dist_init has not been called");
    if (!idx_global) return fail(ctx, GM_ERR_ARG, "gm_dist_copy_particles: null argument");
    ctx->stats.ms_remap = 0.f; ctx->stats.ms_migrate = 0.f; ctx->stats.migrated_particles = 0; ctx->stats.migrated_bytes = 0;
    if (int r = dist_migrate_and_remap(ctx, idx_global, "gm_dist_copy_particles")) return r;
    ctx->stats.total_ms[2] += ctx->stats.ms_remap; ctx->stats.total_ms[4] += ctx->stats.ms_migrate;
    return GM_OK;
}

}  // extern "C"
