// gm_kernels.h -- kernel argument blocks and launchers (internal; the public boundary is include/gm_b200.h)
#pragma once
#include "gm_device.cuh"

namespace gm {

struct ScanMatchArgs {
    DevParams P;
    const double* ranges; const double* angles;
    double* poses;
    const Geom* geom; const int* dir; int dir_cap;
    const int4* pool; const int* visits; const double2* means;
    const unsigned* occ;  // occupancy bitmap of the pool (gm_device.cuh)
    const double* inv_n;
    double minimum_score;
    double *score, *s, *l; unsigned *c, *evals;
    Counters* counters;
    unsigned valid_score_beams;
    // standalone queries: query = 0 score(), 1 likelihoodAndScore() of pose q on particle_idx[q]'s map; -1 = scanMatch
    int query;
    const unsigned* particle_idx;
    int zero_patch;  // pool index of the all-zero patch
    double* poses_out;  // query 2 (standalone optimize): climbed poses
    // fused result rows (gm_scanmatch_weights): particle p writes {x, y, theta, optimize() score, l} to row first_global + p
    // of rows[r] for every rank r -- rows[r] is rank r's gather buffer, its own HBM or a peer's mapped through CUDA IPC
    // (stores over NVLink).  The last CTA to finish then publishes `seq` in *flag (host memory every rank can poll).
    double* rows[kMaxRanks];
    int nranks;             // 0: no rows
    unsigned first_global;
    unsigned* done_counter; // CTAs finished (device memory, left at 0)
    unsigned* flag; unsigned seq;
};
struct BoundsArgs {
    DevParams P;
    const double* ranges; const double* angles; const double* poses;
    const Geom* geom; Geom* geom_new; int4* shift;
    Counters* counters;
};
struct RelayoutArgs {
    const Counters* counters;  // the kernels stand down when k_bounds asked for a longer directory stride than dir_cap_old
    Geom* geom; const Geom* geom_new; const int4* shift;
    const int* dir; int* dir_alt; int* dir_out;
    int dir_cap_old, dir_cap_new;
    Pool pool;
};
struct RegisterArgs {
    DevParams P;
    const double* ranges; const double* angles; const double* poses;
    const Geom* geom; int* dir; int dir_cap;
    Pool pool;
    Counters* counters;
    int bitmap_words, sort_n;
    int tile_cap;  // visit-count tiles (active patches) per pass of the free-space update
};
struct RemapArgs {
    const unsigned* idx;
    const Geom* geom; Geom* geom_alt;
    const int* dir; int* dir_alt; int dir_cap;
    int* refcnt;  // the pool's reference counts
};
struct DigestArgs {
    const Geom* geom; const int* dir; int dir_cap; const int4* pool; const int* visits; const double2* means; const unsigned* occ;
    unsigned first; unsigned long long* out;  // 7 words per particle: the six of gm_digest + mismatching occupancy-bitmap rows
    int cap_h;                                // Pool::cap_h: ids at or above it are free-space patches (visit counters only)
};

size_t scanmatch_smem_bytes(unsigned nbeams);
cudaError_t scanmatch_prepare(size_t smem);
void launch_scanmatch(const ScanMatchArgs& a, unsigned n, bool wide, cudaStream_t st);
void launch_bounds(const BoundsArgs& a, unsigned n, unsigned threads, cudaStream_t st);
void launch_relayout(const RelayoutArgs& a, unsigned n, cudaStream_t st);
void launch_relayout_commit(const RelayoutArgs& a, unsigned n, cudaStream_t st);
size_t register_smem_bytes(int bitmap_words, int sort_n, int nbeams, int tile_cap);
cudaError_t register_prepare(size_t smem);
void launch_register(const RegisterArgs& a, unsigned n, size_t smem, cudaStream_t st);
void launch_remap(const RemapArgs& a, unsigned n, cudaStream_t st);
void launch_release_slots(const Geom* geom, const int* dir, int dir_cap, unsigned first, unsigned count, const Pool& pool, cudaStream_t st);
void launch_addref_slots(const Geom* geom, const int* dir, int dir_cap, unsigned count, const Pool& pool, cudaStream_t st);
void launch_merge_free(const Pool& p, cudaStream_t st);
void launch_pool_init(const Pool& p, cudaStream_t st);
void launch_recount(const Geom* geom, const int* dir, int dir_cap, const Pool& pool, int* recount, unsigned long long* out, unsigned n, cudaStream_t st);
void launch_recount_compare(const Pool& pool, int* recount, unsigned long long* out, cudaStream_t st);
void launch_fill_int(int* p, size_t n, int v, cudaStream_t st);
void launch_fill_geom(Geom* g, unsigned n, Geom v, cudaStream_t st);
void launch_normalize(const double* logw, const double* rows, unsigned n, double gain, double* w, double* neff, cudaStream_t st);
void launch_resample(const double* w, unsigned n, unsigned n_out, double u01, double* cum, double* tgt, unsigned* idx, unsigned* emitted, cudaStream_t st);
void launch_digest(const DigestArgs& a, unsigned n, cudaStream_t st);
void launch_entropy(const DigestArgs& a, double* out, unsigned n, cudaStream_t st);
void launch_export_occupancy(const int* dirp, Geom g, const int4* pool, const int* visits, int cap_h, double occ_thresh, signed char* out, int width, int height, cudaStream_t st);
void launch_release_particle(int* dirp, int nent, const Pool& pool, cudaStream_t st);
void launch_install_patches(int* dirp, Geom g, const int* coords, const int4* staged, const Pool& pool, Counters* ctr, int n, cudaStream_t st);
void launch_export_list(const int* dirp, Geom g, int* list, int* count, cudaStream_t st);
void launch_export_gather(const int4* pool, const int* visits, int cap_h, const int* list, int4* out, int n, cudaStream_t st);

}  // namespace gm
