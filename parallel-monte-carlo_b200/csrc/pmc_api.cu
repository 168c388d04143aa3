// pmc_api.cu -- the C-ABI (include/pmc.h): handle, geometry, host-side per-sweep randomness,
// the sweep protocol of start.cu:237-260, observables glue, host I/O and the NCCL slab ring.
#include "pmc_internal.cuh"
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <utility>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

// ------------------------------------------------------------------ minimal NCCL surface
// libnccl.so.2 (the copy torch ships) is dlopen'ed on first use so that single-GPU users
// need no NCCL at all.  Declarations restated from nccl.h 2.27 (public API, stable ABI).
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0 };
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
};
static NcclApi g_nccl;

static bool nccl_load()
{
    if (g_nccl.lib) return true;
    const char *names[] = { "libnccl.so.2", "libnccl.so" };
    void *lib = nullptr;
    for (const char *nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) {
        const char *env = getenv("PMC_NCCL_LIB");
        if (env) lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!lib) return false;
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(lib, "ncclCommDestroy");
    g_nccl.Send = (decltype(g_nccl.Send))dlsym(lib, "ncclSend");
    g_nccl.Recv = (decltype(g_nccl.Recv))dlsym(lib, "ncclRecv");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(lib, "ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(lib, "ncclGroupEnd");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.Send || !g_nccl.Recv ||
        !g_nccl.GroupStart || !g_nccl.GroupEnd || !g_nccl.CommDestroy) { dlclose(lib); return false; }
    g_nccl.lib = lib;
    return true;
}

// ------------------------------------------------------------------ handle
constexpr int kMaxBands = 16;
constexpr int kGhostRows = 5;   // fused sweep halo (4) + 1 upstream row of the pending shift
constexpr int kFlagRows = 3;    // rows of 2 x 2 flag blocks that cover kMY = 5 rows starting at an odd or even row

struct pmc_handle {
    pmc_params p;
    pmc_geometry pg;
    DevGeom g;
    int device;
    cudaStream_t stream;
    bool own_stream;
    int blocking;
    Counters *d_ctr;
    float4 *scratch_disk;
    int16_t *scratch_n;
    long long *d_out4;
    unsigned *d_min;
    unsigned long long *d_hist;
    int hist_cap;
    // pmc_run_host buffers
    float *run_r;
    float4 *run_disk;
    int16_t *run_n;
    // slab ring
    ncclComm_t comm;
    // fast fused sweep (pmc_sweep4.cu): internal-layout ping-pong buffers + their TMA descriptors
    int v4_ok;                  // parameters qualify (n_M == 4, w >= 2 sigma, cps >= 48)
    int v4_padded;              // both buffers hold valid cells everywhere (padding included)
    Geom4 g4;
    float4 *v4_buf[2];
    // device time of the fused sweep kernels alone (bench.py roofline): one event pair per pmc_sweep call
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> *ktime_pending;
    double ktime_ms;
    long long ktime_launches, ktime_pending_launches;
    // slab runs: ghost rows travel on a side stream while the interior tile rows are computed
    cudaStream_t comm_stream;
    cudaEvent_t ev_interior[2], ev_exchanged[2];    // ping-pong by sweep parity
    // single GPU: bands of tile rows on their own streams, so that the tail of sweep t overlaps the head of t+1
    cudaStream_t band_stream[kMaxBands];
    cudaEvent_t ev_band[2][kMaxBands], ev_band_start;
    // crowded-cell flags of the two internal buffers (one word per 2 x 2 cells, epoch-stamped)
    unsigned *v4_flags[2];
    unsigned *v4_flag_stage;    // slab ring: the neighbours' flag rows land here and are merged
    unsigned v4_epoch[2], v4_epoch_next;
    long long launches;         // kernels launched by this handle since the last pmc_reset_counters
    // result-invariant tuning knobs (pmc_set_tuning)
    int tune_bands, tune_prefetch, tune_overlap, tune_generic, tune_force, tune_four_plane, tune_tile_rows;
    unsigned status_sticky;     // status bits already handed to the caller as a return code (pmc_get_counters ORs them back in)
    Counters *h_ctr;            // pinned host mirror for the status read of blocking calls
    alignas(64) unsigned char v4_tmap[2][2][128];   // [buffer][0: full-tile box, 1: half-height box]
};

// every entry point runs on the handle's device and leaves the caller's current device as it found it
struct DevGuard {
    int prev = -1, dev;
    explicit DevGuard(int d) : dev(d) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; if (prev != dev) cudaSetDevice(dev); }
    ~DevGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};
#define GUARD(h) DevGuard guard_((h)->device)

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

// Blocking calls end here: synchronise, then turn a cell overflow / lost particles that happened since the
// last check into the return code pmc.h promises (the reference writes past nmax silently).  The arrays are
// still complete and consistent except for the dropped particles; the counters keep the totals.  Non-blocking
// callers poll pmc_get_counters.
static int finish(pmc_handle *h)
{
    if (!h->blocking) return 0;
    CK(cudaMemcpyAsync(h->h_ctr, h->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const unsigned st = h->h_ctr->status;
    if (!st) return 0;
    CK(cudaMemsetAsync(&h->d_ctr->status, 0, sizeof(unsigned), h->stream));     // reported once; sticky on the host
    h->status_sticky |= st;
    return (st & PMC_STATUS_OVERFLOW) ? PMC_E_OVERFLOW : PMC_E_LOST;
}

static void host_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed, uint32_t out[4])
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// start.cu:14-27 as runtime values (identical arithmetic to oracle_make_geom); pure host maths
static int derive_geometry(const pmc_params &p_in, pmc_params *p_out, DevGeom *g_out, pmc_geometry *pg_out)
{
    pmc_params p = p_in;
    if (p.n_particles <= 0 || !(p.phi > 0.0f) || !(p.sigma_d > 0.0f) || !(p.cell_w >= p.sigma_d) ||
        p.n_M < 1 || p.n_M > 64 || !(p.move_delta > 0.0f)) return PMC_E_INVALID;
    if (p.nmax != PMC_NMAX) return PMC_E_UNSUPPORTED;
    if (p.proposal != PMC_PROPOSAL_UNIFORM && p.proposal != PMC_PROPOSAL_GAUSSIAN) return PMC_E_INVALID;
    // a cell of width w holds pi/4 * w^2 * phi / (pi sigma^2 / 4) disks on average; with less than two slots
    // of head-room above that mean, overflow is the rule rather than a 1e-5 event (cell_w = 3 sigma at
    // phi = 0.7: mean 8.0).  Overflow at run time is still detected and reported (PMC_E_OVERFLOW).
    if ((double)p.phi * (double)p.cell_w * (double)p.cell_w * 4.0 / (M_PI * (double)p.sigma_d * (double)p.sigma_d) > PMC_NMAX - 2)
        return PMC_E_UNSUPPORTED;
    if (p.cps_multiple < 2) p.cps_multiple = 2;
    if (p.cps_multiple & 1) return PMC_E_INVALID;
    if (p.n_ranks < 1) p.n_ranks = 1;
    if (p.rank < 0 || p.rank >= p.n_ranks) return PMC_E_INVALID;
    double L_d = sqrt((double)p.n_particles * M_PI * (double)p.sigma_d * (double)p.sigma_d / (4.0 * (double)p.phi));
    long long cps = (long long)floor(L_d / ((double)p.cps_multiple * (double)p.cell_w)) * p.cps_multiple;
    if (cps < 4 || cps > 46340) return PMC_E_INVALID;
    double w_d = L_d / (double)cps;
    // coordinate grid (pmc.h): q = 2^e with every multiple of q below 2^(e+24) > 2w representable
    int ex;
    (void)frexp(2.0 * w_d, &ex);
    const double q = ldexp(1.0, ex - 24);
    const double K = nearbyint(w_d / q), M = floor((double)p.move_delta / q);
    // the uniform proposal has 4096 levels (2k - 4095) * A * q per axis, A = floor(M / 4095) >= 1
    if (M < 4095.0 || M >= 4194304.0 || (double)p.move_delta > w_d) return PMC_E_INVALID;
    const double A = floor(M / 4095.0);
    if (2.0 * A * q * ldexp(1.0, 137) >= ldexp(1.0, 127)) return PMC_E_INVALID;     // dstep must be a finite float (w >= 4 with delta ~ w/2)
    DevGeom g;
    memset(&g, 0, sizeof(g));
    g.cps = (int)cps;
    g.w = (float)(K * q);
    g.K = (int)K; g.M = (int)M; g.A = (int)A; g.dstep = (float)(2.0 * A * q * ldexp(1.0, 137)); g.doff = (float)(4095.0 * A * q);
    g.L_box = (double)cps * (double)g.w;
    g.L = (float)g.L_box;
    g.half_L = g.L / 2.0f;
    g.sigma = p.sigma_d;
    g.sigma2 = p.sigma_d * p.sigma_d;
    g.dscale = (float)q;
    g.proposal = p.proposal;
    g.n_M = p.n_M;
    g.seed_lo = (unsigned)p.seed;
    g.seed_hi = (unsigned)(p.seed >> 32);
    g.n_particles = p.n_particles;
    if (p.n_ranks == 1) {
        g.row0 = 0; g.rows = g.cps; g.ghost = 0; g.wrap_y = 1;
    } else {
        // 1-D slabs of whole cell rows, even row count per slab so colours stay aligned
        if (g.cps % (2 * p.n_ranks) != 0) return PMC_E_INVALID;
        g.rows = g.cps / p.n_ranks;
        g.row0 = p.rank * g.rows;
        g.ghost = kGhostRows;
        g.wrap_y = 0;
        if (g.rows < 2 * kGhostRows) return PMC_E_INVALID;
    }
    g.local_rows = g.rows + 2 * g.ghost;
    pmc_geometry pg;
    memset(&pg, 0, sizeof(pg));
    pg.n_particles = p.n_particles; pg.cps = g.cps; pg.n_cells = cps * cps; pg.nmax = PMC_NMAX;
    pg.n_M = p.n_M; pg.w = g.w; pg.L = g.L; pg.sigma_d = p.sigma_d; pg.move_delta = (float)((p.proposal == PMC_PROPOSAL_GAUSSIAN ? M : 4095.0 * A) * q); pg.grid_q = (float)q;
    pg.row0 = g.row0; pg.rows = g.rows; pg.ghost_rows = g.ghost;
    pg.local_cells = (long long)g.local_rows * g.cps;
    if (p_out) *p_out = p;
    if (g_out) *g_out = g;
    if (pg_out) *pg_out = pg;
    return 0;
}

extern "C" {

const char *pmc_error_string(int code)
{
    switch (code) {
    case 0: return "success";
    case PMC_E_INVALID: return "pmc: invalid argument";
    case PMC_E_UNSUPPORTED: return "pmc: unsupported parameter (this build: nmax == 8; RSA: phi above the jamming density)";
    case PMC_E_OVERFLOW: return "pmc: a cell exceeded nmax particles";
    case PMC_E_LOST: return "pmc: particles outside the box were dropped";
    case PMC_E_NOT_SQUARE: return "pmc: init_r needs a perfect-square particle count";
    case PMC_E_COMM: return "pmc: communicator error";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "pmc: unknown error";
    }
}

int pmc_geometry_from_params(const pmc_params *pp, pmc_geometry *out)
{
    if (!pp || !out) return PMC_E_INVALID;
    return derive_geometry(*pp, nullptr, nullptr, out);
}

int pmc_create(const pmc_params *pp, pmc_handle **out)
{
    if (!pp || !out) return PMC_E_INVALID;
    pmc_params p;
    DevGeom g0;
    pmc_geometry pg0;
    int grc = derive_geometry(*pp, &p, &g0, &pg0);
    if (grc) return grc;
    // the handle embeds CUtensorMap storage: 64-byte alignment (calloc guarantees 16)
    pmc_handle *h = (pmc_handle *)aligned_alloc(64, (sizeof(pmc_handle) + 63) / 64 * 64);
    if (!h) return PMC_E_INVALID;
    memset(h, 0, sizeof(pmc_handle));
    h->p = p;
    h->g = g0;
    h->pg = pg0;
    DevGeom &g = h->g;
    {   // fast path eligibility: the 3-neighbour-cell argument needs w >= 2 sigma with a margin
        // far above float rounding; tiny boxes would need more than one periodic image
        Geom4 &q = h->g4;
        q.cps = g.cps; q.row0 = g.row0; q.rows = g.rows; q.wrap_y = g.wrap_y;
        pmc4_alloc_shape(g.cps, g.rows, &q.CH, &q.ROWS, &q.FW, &q.FH);
        q.w = g.w; q.hw = 0.5f * g.w; q.sigma2 = g.sigma2; q.dscale = g.dscale; q.dstep = g.dstep; q.doff = g.doff;
        q.seed_lo = g.seed_lo; q.seed_hi = g.seed_hi;
        q.try_ns4 = (double)p.n_particles / ((double)g.cps * (double)g.cps) < 2.5;   // a performance hint only
        for (int r = 0; r < 10; r++) { q.pk0[r] = g.seed_lo + (unsigned)r * 0x9E3779B9u; q.pk1[r] = g.seed_hi + (unsigned)r * 0xBB67AE85u; }
        h->v4_ok = (p.n_M == 4) && (p.proposal == PMC_PROPOSAL_UNIFORM) &&
                   ((double)g.w >= 2.0 * (double)p.sigma_d * (1.0 + 1e-5)) && g.cps >= 48 &&
                   g.rows >= 2 * kMY && (p.n_ranks == 1 || kGhostRows == kMY);
    }
    // tuning defaults; PMC_BANDS / PMC_PREFETCH remain as documented environment overrides of the two
    // performance knobs (they cannot change results)
    h->tune_bands = [] { const char *e = getenv("PMC_BANDS"); return e ? atoi(e) : 6; }();
    h->tune_prefetch = [] { const char *e = getenv("PMC_PREFETCH"); return e ? atoi(e) : 296; }();
    h->tune_overlap = 1;
    int caller_dev = -1;
    cudaGetDevice(&caller_dev);
    if (p.device >= 0) { cudaError_t e = cudaSetDevice(p.device); if (e != cudaSuccess) { free(h); return (int)e; } }
    cudaError_t e = cudaGetDevice(&h->device);
    if (e != cudaSuccess) { free(h); return (int)e; }
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore_{ caller_dev };
    {   // stream-ordered scratch (cell-list build) stays cached in the device's default pool
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, h->device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { free(h); return (int)e; }
    h->own_stream = true;
    h->blocking = 1;
    e = cudaMalloc(&h->d_ctr, sizeof(Counters));
    if (e == cudaSuccess) e = cudaMemset(h->d_ctr, 0, sizeof(Counters));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_out4, 4 * sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_min, sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMallocHost(&h->h_ctr, sizeof(Counters));
    if (e != cudaSuccess) { pmc_destroy(h); return (int)e; }
    *out = h;
    return 0;
}

int pmc_destroy(pmc_handle *h)
{
    if (!h) return PMC_E_INVALID;
    GUARD(h);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    if (h->h_ctr) cudaFreeHost(h->h_ctr);
    cudaFree(h->d_ctr); cudaFree(h->scratch_disk); cudaFree(h->scratch_n);
    cudaFree(h->d_out4); cudaFree(h->d_min); cudaFree(h->d_hist);
    cudaFree(h->run_r); cudaFree(h->run_disk); cudaFree(h->run_n);
    cudaFree(h->v4_buf[0]); cudaFree(h->v4_buf[1]);
    cudaFree(h->v4_flags[0]); cudaFree(h->v4_flags[1]); cudaFree(h->v4_flag_stage);
    if (h->ktime_pending) {
        for (auto &pr : *h->ktime_pending) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
        delete h->ktime_pending;
    }
    for (int b = 0; b < kMaxBands; b++)
        if (h->band_stream[b]) {
            cudaStreamDestroy(h->band_stream[b]);
            cudaEventDestroy(h->ev_band[0][b]); cudaEventDestroy(h->ev_band[1][b]);
        }
    if (h->ev_band_start) cudaEventDestroy(h->ev_band_start);
    if (h->comm_stream) {
        cudaStreamDestroy(h->comm_stream);
        for (int b = 0; b < 2; b++) { cudaEventDestroy(h->ev_interior[b]); cudaEventDestroy(h->ev_exchanged[b]); }
    }
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    free(h);
    return 0;
}

// Result-invariant knobs: which kernel / schedule computes the (identical) result.
//   "bands" (1..8)        tile-row bands of a sweep on their own streams (default 6)
//   "prefetch" (>= 0)     L2 prefetch distance in CTAs (default 296)
//   "overlap" (0/1)       slab runs: boundary rows + NCCL ring on a side stream (default 1)
//   "generic" (0/1)       use the generic fused kernel (pmc_sweep.cu) even where the fast path qualifies
//   "four_plane" (0/1)    fast path with all four planes staged and no crowded-cell flag lookup (3 CTAs / SM)
//   "tile_rows" (0, 2..28 even)   cap on the owned rows of a tile (0 = automatic: 16 for systems of less than ~1.5 waves)
//   "force_crowded", "no_ns4", "full_halo" (0/1)   drive the rare paths of the fast kernel on ordinary tiles
int pmc_set_tuning(pmc_handle *h, const char *name, int value)
{
    if (!h || !name) return PMC_E_INVALID;
    auto flag = [&](int bit) { h->tune_force = value ? (h->tune_force | bit) : (h->tune_force & ~bit); return 0; };
    if (!strcmp(name, "bands")) { if (value < 1 || value > kMaxBands) return PMC_E_INVALID; h->tune_bands = value; return 0; }
    if (!strcmp(name, "prefetch")) { if (value < 0) return PMC_E_INVALID; h->tune_prefetch = value; return 0; }
    if (!strcmp(name, "overlap")) { h->tune_overlap = value ? 1 : 0; return 0; }
    if (!strcmp(name, "generic")) { h->tune_generic = value ? 1 : 0; return 0; }
    if (!strcmp(name, "four_plane")) { h->tune_four_plane = value ? 1 : 0; return 0; }
    if (!strcmp(name, "tile_rows")) { if (value < 0 || value > 28 || (value & 1)) return PMC_E_INVALID; h->tune_tile_rows = value; return 0; }
    if (!strcmp(name, "force_crowded")) return flag(8);
    if (!strcmp(name, "no_ns4")) return flag(16);
    if (!strcmp(name, "full_halo")) return flag(64);
    return PMC_E_INVALID;
}

int pmc_get_geometry(const pmc_handle *h, pmc_geometry *g)
{
    if (!h || !g) return PMC_E_INVALID;
    *g = h->pg;
    return 0;
}

size_t pmc_r_bytes(const pmc_handle *h) { return h ? (size_t)h->p.n_particles * 2 * sizeof(float) : 0; }
size_t pmc_disk_bytes(const pmc_handle *h) { return h ? (size_t)h->pg.local_cells * 2 * PMC_NMAX * sizeof(float) : 0; }
size_t pmc_n_bytes(const pmc_handle *h) { return h ? (size_t)h->pg.local_cells * sizeof(int16_t) : 0; }

int pmc_set_stream(pmc_handle *h, void *cuda_stream)
{
    if (!h) return PMC_E_INVALID;
    GUARD(h);
    CK(cudaStreamSynchronize(h->stream));
    if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
    h->stream = (cudaStream_t)cuda_stream;
    return 0;
}

int pmc_set_blocking(pmc_handle *h, int blocking)
{
    if (!h) return PMC_E_INVALID;
    h->blocking = blocking ? 1 : 0;
    return 0;
}

int pmc_synchronize(pmc_handle *h)
{
    if (!h) return PMC_E_INVALID;
    GUARD(h);
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ------------------------------------------------------------------ the four call sites
int pmc_init_r(pmc_handle *h, float *d_r)
{
    if (!h || !d_r) return PMC_E_INVALID;
    GUARD(h);
    long long N = h->p.n_particles;
    long long ns = (long long)floor(sqrt((double)N) + 0.5);
    if (ns * ns != N) return PMC_E_NOT_SQUARE;
    CK(pmc_launch_init_r(h->g, d_r, h->stream)); h->launches += 1;
    return finish(h);
}

int pmc_assign(pmc_handle *h, const float *d_r, float *d_disk, int16_t *d_n)
{
    if (!h || !d_r || !d_disk || !d_n) return PMC_E_INVALID;
    GUARD(h);
    CK(pmc_launch_assign(h->g, d_r, (float4 *)d_disk, d_n, h->d_ctr, h->stream)); h->launches += 2;
    return finish(h);
}

void pmc_colour_to_off(int colour, int off[2])
{
    off[1] = colour % 2;            // itoa start.cu:153-157, two bits in 2-D
    off[0] = (colour / 2) % 2;
}

int pmc_plan_sweep(const int order[4], int f, float d, int out[12])
{
    if (!order || !out || f < 0 || f > 1) return PMC_E_INVALID;
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    for (int k = 0; k < 4; k++) {
        if (order[k] < 0 || order[k] > 3) return PMC_E_INVALID;
        int off[2];
        pmc_colour_to_off(order[k], off);
        a.offx[k] = off[0]; a.offy[k] = off[1];
    }
    a.shift_on = 1; a.shift_f = f; a.shift_d = d;
    pmc4_plan_sweep(a, 0);
    out[0] = a.tx; out[1] = a.ty; out[2] = a.hx; out[3] = a.hy;
    for (int k = 0; k < 4; k++) {
        out[4 + k] = (int)((a.lo_x >> (4 * k)) & 15u);
        out[8 + k] = (int)((a.lo_y >> (4 * k)) & 15u);
    }
    return 0;
}

int pmc_schedule(const pmc_handle *h, uint64_t sweep, int order[4], int *f, float *d)
{
    if (!h || !order || !f || !d) return PMC_E_INVALID;
    uint32_t a[4], b[4];
    host_philox(0xFFFFFFFFu, (uint32_t)sweep, (uint32_t)(sweep >> 32), (1u << 16) | 0u, h->p.seed, a);
    host_philox(0xFFFFFFFFu, (uint32_t)sweep, (uint32_t)(sweep >> 32), (1u << 16) | 1u, h->p.seed, b);
    for (int i = 0; i < 4; i++) order[i] = i;
    for (int i = 3; i >= 1; i--) {                      // FY_Shuffle start.cu:34-44, unbiased
        int j = (int)(((uint64_t)a[3 - i] * (uint64_t)(i + 1)) >> 32);
        int t = order[i]; order[i] = order[j]; order[j] = t;
    }
    *f = (int)(a[3] >> 31);                             // kernel.cu:683 range, 2 axes
    // kernel.cu:684 range (-w/2, w/2], on the coordinate grid: dk * q (oracle_schedule)
    const int64_t K = h->g.K;
    const int64_t dk = (int64_t)(((uint64_t)b[0] * (uint64_t)K) >> 32) + 1 - (K + 1) / 2;
    *d = (float)dk * h->g.dscale;
    return 0;
}

static int ensure_scratch(pmc_handle *h)
{
    if (!h->scratch_disk) CK(cudaMalloc(&h->scratch_disk, pmc_disk_bytes(h)));
    if (!h->scratch_n) CK(cudaMalloc(&h->scratch_n, pmc_n_bytes(h)));
    return 0;
}

static int exchange_ghosts_async(pmc_handle *h, float4 *disk, int16_t *n)
{
    if (h->p.n_ranks <= 1) return 0;
    if (!h->comm) return PMC_E_COMM;
    const DevGeom &g = h->g;
    const int G = g.ghost, R = h->p.n_ranks;
    const int lower = (h->p.rank + R - 1) % R, upper = (h->p.rank + 1) % R;
    const size_t drow = (size_t)g.cps * 64, nrow = (size_t)g.cps * sizeof(int16_t);
    char *D = (char *)disk, *N = (char *)n;
    // owned rows are local rows [G, G+rows)
    int rc = 0;
    rc |= g_nccl.GroupStart();
    rc |= g_nccl.Send(D + (size_t)G * drow, G * drow, ncclInt8, lower, h->comm, h->stream);
    rc |= g_nccl.Send(D + (size_t)g.rows * drow, G * drow, ncclInt8, upper, h->comm, h->stream);
    rc |= g_nccl.Recv(D + (size_t)(G + g.rows) * drow, G * drow, ncclInt8, upper, h->comm, h->stream);
    rc |= g_nccl.Recv(D, G * drow, ncclInt8, lower, h->comm, h->stream);
    rc |= g_nccl.Send(N + (size_t)G * nrow, G * nrow, ncclInt8, lower, h->comm, h->stream);
    rc |= g_nccl.Send(N + (size_t)g.rows * nrow, G * nrow, ncclInt8, upper, h->comm, h->stream);
    rc |= g_nccl.Recv(N + (size_t)(G + g.rows) * nrow, G * nrow, ncclInt8, upper, h->comm, h->stream);
    rc |= g_nccl.Recv(N, G * nrow, ncclInt8, lower, h->comm, h->stream);
    rc |= g_nccl.GroupEnd();
    return rc ? PMC_E_COMM : 0;
}

int pmc_subsweep(pmc_handle *h, float *d_disk, int16_t *d_n, const int off[2], uint64_t sweep)
{
    if (!h || !d_disk || !d_n || !off) return PMC_E_INVALID;
    GUARD(h);
    if ((off[0] | off[1]) & ~1) return PMC_E_INVALID;
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.offx[0] = off[0]; a.offy[0] = off[1];
    a.offmask = (unsigned)off[0] | ((unsigned)off[1] << 1);
    a.sanitize_in = 1;
    a.sweep_lo = (unsigned)sweep; a.sweep_hi = (unsigned)(sweep >> 32);
    CK(pmc_launch_subsweep(h->g, (float4 *)d_disk, d_n, a, h->d_ctr, h->stream)); h->launches += 1;
    int rc = exchange_ghosts_async(h, (float4 *)d_disk, d_n);
    if (rc) return rc;
    return finish(h);
}

int pmc_shift_cells(pmc_handle *h, float *d_disk, int16_t *d_n, int f, float d)
{
    if (!h || !d_disk || !d_n || f < 0 || f > 1) return PMC_E_INVALID;
    GUARD(h);
    if (!(fabsf(d) <= 0.5f * h->g.w * 1.0001f)) return PMC_E_INVALID;   // shiftCells.h:7 contract
    d = (float)(nearbyint((double)d / (double)h->g.dscale) * (double)h->g.dscale);   // onto the coordinate grid
    int rc = ensure_scratch(h);
    if (rc) return rc;
    CK(pmc_launch_shift(h->g, (const float4 *)d_disk, d_n, h->scratch_disk, h->scratch_n, f, d, h->d_ctr, h->stream)); h->launches += 1;
    CK(cudaMemcpyAsync(d_disk, h->scratch_disk, pmc_disk_bytes(h), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(d_n, h->scratch_n, pmc_n_bytes(h), cudaMemcpyDeviceToDevice, h->stream));
    rc = exchange_ghosts_async(h, (float4 *)d_disk, d_n);
    if (rc) return rc;
    return finish(h);
}

extern "C" int pmc_get_kernel_time(pmc_handle *h, double *ms, long long *launches);
static size_t v4_bytes(const pmc_handle *h) { return (size_t)2 * h->g4.CH * h->g4.ROWS * 4 * sizeof(float4); }

// slab ring on the internal layout: whole rows (4 planes x 2 parities x CH chunks) are contiguous, one message
// per neighbour.  The crowded-cell flags of the same rows travel with them (3 rows of flag words per face, a
// few KB), so that the tile rows next to a slab face can take the 3-plane fast kernel like every other row:
// epochs advance identically on every rank, so a received stamp means the same thing here.  Flag blocks are
// 2 x 2 cells and kMY is odd: the block row at a face holds one ghost and one owned row, hence the received
// words are MERGED (a stamp is only ever added) by a tiny kernel instead of being received in place.
static int v4_exchange_async(pmc_handle *h, float4 *buf, unsigned *flags, unsigned epoch, cudaStream_t st)
{
    if (h->p.n_ranks <= 1) return 0;
    if (!h->comm) return PMC_E_COMM;
    const int R = h->p.n_ranks, rows = h->g4.rows, FW = h->g4.FW;
    const int lower = (h->p.rank + R - 1) % R, upper = (h->p.rank + 1) % R;
    const size_t rp = (size_t)h->g4.CH * 128, blk = (size_t)kMY * rp;
    const size_t fblk = (size_t)kFlagRows * FW * sizeof(unsigned);
    if (!h->v4_flag_stage) CK(cudaMalloc(&h->v4_flag_stage, 2 * fblk));
    char *B = (char *)buf;
    // flag rows of my lowest / highest kMY owned rows (Y = kMY .. 2 kMY - 1 and rows .. rows + kMY - 1)
    const unsigned *f_lo = flags + (size_t)(kMY >> 1) * FW, *f_hi = flags + (size_t)(rows >> 1) * FW;
    unsigned *r_lo = h->v4_flag_stage, *r_hi = h->v4_flag_stage + (size_t)kFlagRows * FW;
    int rc = 0;
    rc |= g_nccl.GroupStart();
    rc |= g_nccl.Send(B + (size_t)kMY * rp, blk, ncclInt8, lower, h->comm, st);
    rc |= g_nccl.Send(B + (size_t)rows * rp, blk, ncclInt8, upper, h->comm, st);
    rc |= g_nccl.Recv(B + (size_t)(kMY + rows) * rp, blk, ncclInt8, upper, h->comm, st);
    rc |= g_nccl.Recv(B, blk, ncclInt8, lower, h->comm, st);
    rc |= g_nccl.Send(f_lo, fblk, ncclInt8, lower, h->comm, st);
    rc |= g_nccl.Send(f_hi, fblk, ncclInt8, upper, h->comm, st);
    rc |= g_nccl.Recv(r_hi, fblk, ncclInt8, upper, h->comm, st);
    rc |= g_nccl.Recv(r_lo, fblk, ncclInt8, lower, h->comm, st);
    rc |= g_nccl.GroupEnd();
    if (rc) return PMC_E_COMM;
    // what the lower neighbour sent are its top rows = my ghost rows Y = 0 .. kMY - 1 (flag rows 0 ..);
    // what the upper neighbour sent are its bottom rows = my ghost rows Y = kMY + rows .. (flag rows (kMY + rows) >> 1 ..)
    CK(pmc4_launch_flag_merge(flags, r_lo, 0, r_hi, (kMY + rows) >> 1, kFlagRows, FW, epoch, st));
    h->launches += 1;
    return 0;
}

// start.cu:237-260 on the fast path: caller layout -> internal layout, n_sweeps x ONE kernel
// (4 colours + this sweep's shiftCells applied while the tile is stored), -> caller layout.
static int sweep_v4(pmc_handle *h, float *d_disk, int16_t *d_n, uint64_t sweep0, int n_sweeps)
{
    const size_t flag_bytes = (size_t)h->g4.FW * h->g4.FH * sizeof(unsigned);
    for (int b = 0; b < 2; b++)
        if (!h->v4_buf[b]) {
            CK(cudaMalloc(&h->v4_buf[b], v4_bytes(h)));
            CK(cudaMalloc(&h->v4_flags[b], flag_bytes));
            CK(cudaMemsetAsync(h->v4_flags[b], 0, flag_bytes, h->stream));
            for (int half = 0; half < 2; half++) {
                int rc = pmc4_make_tensor_map(h->v4_tmap[b][half], h->v4_buf[b], h->g4, half);
                if (rc) return rc;
            }
            h->v4_padded = 0;
            h->v4_epoch_next = 1;                   // 0 = "never flagged" (the memset above)
        }
    auto next_epoch = [&]() -> unsigned {
        if (h->v4_epoch_next == 0xFFFFFFFFu) {      // wrap: forget every stamp (once per 4e9 sweeps)
            cudaMemsetAsync(h->v4_flags[0], 0, flag_bytes, h->stream);
            cudaMemsetAsync(h->v4_flags[1], 0, flag_bytes, h->stream);
            h->v4_epoch_next = 1;
        }
        return h->v4_epoch_next++;
    };
    const int ghost = h->g.ghost;
    h->v4_epoch[0] = next_epoch();
    CK(pmc4_launch_import(h->g4, ghost, (const float4 *)d_disk, d_n, h->v4_buf[0], h->v4_flags[0], h->v4_epoch[0], h->stream)); h->launches += 1;
    if (!h->v4_padded) {        // cells beyond the margins are never written again: make them valid once
        h->v4_epoch[1] = next_epoch();
        CK(pmc4_launch_import(h->g4, ghost, (const float4 *)d_disk, d_n, h->v4_buf[1], h->v4_flags[1], h->v4_epoch[1], h->stream)); h->launches += 1;
        h->v4_padded = 1;
    }
    int cur = 0;
#ifdef PMC_DEBUG
    static const int dbg_env = [] { const char *e = getenv("PMC_DBG_SKIP"); return e ? atoi(e) : 0; }();
#else
    const int dbg_env = 0;
#endif
    const int dbg = (h->tune_force & kTuneForceBits) | dbg_env;
    const int overlap = h->tune_overlap;
    const int fast_ok = !h->tune_four_plane;
    if (h->p.n_ranks > 1 && !h->comm_stream) {
        int prio_lo = 0, prio_hi = 0;               // the exchange must not queue behind the interior tiles
        CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CK(cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, prio_hi));
        for (int b = 0; b < 2; b++) {
            CK(cudaEventCreateWithFlags(&h->ev_interior[b], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&h->ev_exchanged[b], cudaEventDisableTiming));
        }
    }
    const int pf_ahead = h->tune_prefetch;
    // the event pair that times the sweep kernels of this call; owned by this scope until it is queued
    struct EvPair {
        cudaEvent_t a = nullptr, b = nullptr;
        ~EvPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    } kev;
    CK(cudaEventCreate(&kev.a));
    CK(cudaEventCreate(&kev.b));
    cudaEvent_t k0 = kev.a, k1 = kev.b;
    CK(cudaEventRecord(k0, h->stream));
    int slab_split = 0, last_split_par = 0;         // the previous sweep ran on two streams
    // Single GPU: the tile rows of a sweep are cut into `bands` bands, each on its own stream.  Band b of
    // sweep t+1 reads and overwrites only what bands b-1, b, b+1 (periodic) of sweep t wrote and read, so it
    // waits for those three alone: the last CTAs of sweep t and the first of t+1 share the GPU, there is no
    // idle tail and no launch gap between sweeps.
    // tile height: what the colour order allows (24..29 rows) unless the whole system is less than ~1.5 waves of such
    // tiles (148 SMs x 4 CTAs): then 16-row tiles, whose shorter CTAs shorten the dependency chain from sweep to sweep
    // (N = 2^20: 28.9 -> 25.9 us per sweep; larger systems lose throughput to the deeper halo).  "tile_rows" overrides.
    int ty_cap = h->tune_tile_rows;
    if (!ty_cap && (long long)((h->g4.cps + 27) / 28) * ((h->g4.rows + 25) / 26) < 888) ty_cap = 16;
    const int ty_max = ty_cap ? (ty_cap < 29 ? ty_cap : 29) : 29;    // tallest tile any sweep of this call can plan
    const int bands_env = h->tune_bands;
    // at least 3 tile rows per band whatever this call's sweeps choose as tile height (<= ty_max rows), at least
    // 4 bands (with 3, every band is every other band's neighbour); fixed for the whole call
    int bands = (h->g4.rows + ty_max - 1) / ty_max / 3;
    if (bands > bands_env) bands = bands_env;
    if (bands > kMaxBands) bands = kMaxBands;
    if (bands < 4 || h->p.n_ranks != 1 || n_sweeps < 2) bands = 1;
    // slabs: the interior tile rows (all but one tile row per face) in bands likewise; only the two outer
    // bands depend on the boundary rows and their exchange
    int sbands = ((h->g4.rows + ty_max - 1) / ty_max - 4) / 3;      // 3 * sbands <= (rows - 5) / ty_max - 1 <= interior tile rows of any sweep
    if (sbands > bands_env) sbands = bands_env;
    if (sbands > kMaxBands) sbands = kMaxBands;
    if (sbands < 3 || h->p.n_ranks == 1 || !overlap || n_sweeps < 2) sbands = 1;
    int banded = 0, band_par = 0;                   // bands are in flight; parity of their latest events
    if (bands > 1 || sbands > 1) {
        for (int b = 0; b < (bands > sbands ? bands : sbands); b++)
            if (!h->band_stream[b]) {
                CK(cudaStreamCreateWithFlags(&h->band_stream[b], cudaStreamNonBlocking));
                CK(cudaEventCreateWithFlags(&h->ev_band[0][b], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&h->ev_band[1][b], cudaEventDisableTiming));
            }
        if (!h->ev_band_start) CK(cudaEventCreateWithFlags(&h->ev_band_start, cudaEventDisableTiming));
    }
    for (int t = 0; t < n_sweeps; t++) {
        const uint64_t sweep = sweep0 + (uint64_t)t;
        int order[4], f;
        float d;
        pmc_schedule(h, sweep, order, &f, &d);
        SweepArgs a;
        memset(&a, 0, sizeof(a));
        for (int k = 0; k < 4; k++) {
            int off[2];
            pmc_colour_to_off(order[k], off);
            a.offx[k] = off[0]; a.offy[k] = off[1];
            a.offmask |= ((unsigned)off[0] | ((unsigned)off[1] << 1)) << (2 * k);
        }
        a.sweep_lo = (unsigned)sweep; a.sweep_hi = (unsigned)(sweep >> 32);
        a.shift_on = 1; a.shift_f = f; a.shift_d = d;
        a.dbg_skip = dbg;
        a.prefetch_ahead = pf_ahead;
        pmc4_plan_sweep(a, dbg & 64);               // tile extent and halo from the colour order (64: always the full halo)
        pmc4_philox_prepare(a, h->g4);
        if (ty_cap && a.ty > ty_cap) a.ty = ty_cap; // small systems: shorter tiles, more CTAs per sweep
        h->v4_epoch[cur ^ 1] = next_epoch();
        a.flag_in = h->v4_flags[cur]; a.epoch_in = h->v4_epoch[cur];
        a.flag_out = h->v4_flags[cur ^ 1]; a.epoch_out = h->v4_epoch[cur ^ 1];
        const void *tm = h->v4_tmap[cur][0], *tmh = h->v4_tmap[cur][1];
        float4 *dst = h->v4_buf[cur ^ 1];
        const int gy = pmc4_tile_rows(h->g4, a);
        // tile rows that hold one of the kMY owned rows next to a slab face
        const int top0 = (h->g4.rows - kMY) / a.ty;
        if (h->p.n_ranks > 1 && overlap && top0 > 1 && top0 < gy) {
            // Two streams per sweep.  Side stream (high priority): the boundary tile rows (their boxes hold
            // ghost rows, whose crowded-cell flags arrive with the ring), then the NCCL ring with
            // their 5 owned rows and flag rows.  Main stream: the interior rows (the fast kernel), concurrently.  Sweep t's
            // boundary kernel needs the interior of sweep t-1 (it reads its rows as halo and overwrites the
            // buffer it read); sweep t's interior needs the boundary rows of sweep t-1 likewise.
            const int par = t & 1;
            const bool first_split = !slab_split;
            const bool sb = sbands > 1;              // interior in bands: the same for every sweep of a call
            // what makes the event graph below race-free: a band reads / overwrites rows of its two neighbour
            // bands only, i.e. every band holds at least 3 tile rows and its edges move by less than one tile row
            // (the halo, <= 5 rows, plus the change of tile height between consecutive sweeps, <= 4 rows)
            if (sb && top0 - 1 < 3 * sbands) return PMC_E_INVALID;
            if (first_split) {                       // first split sweep of this call: everything so far is on the main stream
                CK(cudaEventRecord(sb ? h->ev_band_start : h->ev_interior[par ^ 1], h->stream));
                slab_split = 1;
            } else if (!sb) {
                CK(cudaStreamWaitEvent(h->stream, h->ev_exchanged[par ^ 1], 0));
            }
            if (!sb) CK(cudaStreamWaitEvent(h->comm_stream, h->ev_interior[par ^ 1], 0));
            else if (first_split) CK(cudaStreamWaitEvent(h->comm_stream, h->ev_band_start, 0));
            else {                                   // the interior rows next to the faces
                CK(cudaStreamWaitEvent(h->comm_stream, h->ev_band[par ^ 1][0], 0));
                CK(cudaStreamWaitEvent(h->comm_stream, h->ev_band[par ^ 1][sbands - 1], 0));
            }
            CK(pmc4_launch_sweep(h->g4, tm, tmh, dst, a, h->d_ctr, h->comm_stream, fast_ok, 0, 1, top0, gy - top0)); h->launches += 1;
            int rc = v4_exchange_async(h, dst, h->v4_flags[cur ^ 1], h->v4_epoch[cur ^ 1], h->comm_stream);
            if (rc) return rc;
            CK(cudaEventRecord(h->ev_exchanged[par], h->comm_stream));
            if (!sb) {
                CK(pmc4_launch_sweep(h->g4, tm, tmh, dst, a, h->d_ctr, h->stream, fast_ok, 1, top0 - 1)); h->launches += 1;
                CK(cudaEventRecord(h->ev_interior[par], h->stream));
            } else {
                for (int b = 0; b < sbands; b++) {
                    cudaStream_t bs = h->band_stream[b];
                    if (first_split) CK(cudaStreamWaitEvent(bs, h->ev_band_start, 0));
                    else {
                        if (b > 0) CK(cudaStreamWaitEvent(bs, h->ev_band[par ^ 1][b - 1], 0));
                        if (b < sbands - 1) CK(cudaStreamWaitEvent(bs, h->ev_band[par ^ 1][b + 1], 0));
                        if (b == 0 || b == sbands - 1) CK(cudaStreamWaitEvent(bs, h->ev_exchanged[par ^ 1], 0));
                    }
                    const int r0 = 1 + (int)((long long)(top0 - 1) * b / sbands), r1 = 1 + (int)((long long)(top0 - 1) * (b + 1) / sbands);
                    CK(pmc4_launch_sweep(h->g4, tm, tmh, dst, a, h->d_ctr, bs, fast_ok, r0, r1 - r0)); h->launches += 1;
                    CK(cudaEventRecord(h->ev_band[par][b], bs));
                }
                banded = 2; band_par = par;           // 2: slab bands (joined like the single-GPU ones)
            }
            last_split_par = par;
        } else if (bands > 1 && gy >= 3 * bands) {
            const int par = t & 1;
            if (!banded) {                           // everything so far (import) is on the main stream
                CK(cudaEventRecord(h->ev_band_start, h->stream));
                for (int b = 0; b < bands; b++) CK(cudaStreamWaitEvent(h->band_stream[b], h->ev_band_start, 0));
            }
            for (int b = 0; b < bands; b++) {
                if (banded) {
                    CK(cudaStreamWaitEvent(h->band_stream[b], h->ev_band[par ^ 1][(b + bands - 1) % bands], 0));
                    CK(cudaStreamWaitEvent(h->band_stream[b], h->ev_band[par ^ 1][(b + 1) % bands], 0));
                }
                const int r0 = (int)((long long)gy * b / bands), r1 = (int)((long long)gy * (b + 1) / bands);
                CK(pmc4_launch_sweep(h->g4, tm, tmh, dst, a, h->d_ctr, h->band_stream[b], fast_ok, r0, r1 - r0)); h->launches += 1;
                CK(cudaEventRecord(h->ev_band[par][b], h->band_stream[b]));
            }
            banded = 1; band_par = par;
        } else {
            if (banded) {                            // back on the main stream: join every band
                for (int b = 0; b < (banded == 2 ? sbands : bands); b++) CK(cudaStreamWaitEvent(h->stream, h->ev_band[band_par][b], 0));
                banded = 0;
            }
            if (slab_split) {                        // back on one stream: join the side stream first
                CK(cudaStreamWaitEvent(h->stream, h->ev_exchanged[last_split_par], 0));
                slab_split = 0;
            }
            CK(pmc4_launch_sweep(h->g4, tm, tmh, dst, a, h->d_ctr, h->stream, fast_ok)); h->launches += 1;
            int rc = v4_exchange_async(h, dst, h->v4_flags[cur ^ 1], h->v4_epoch[cur ^ 1], h->stream);
            if (rc) return rc;
        }
        cur ^= 1;
    }
    if (slab_split) CK(cudaStreamWaitEvent(h->stream, h->ev_exchanged[last_split_par], 0));
    if (banded) for (int b = 0; b < (banded == 2 ? sbands : bands); b++) CK(cudaStreamWaitEvent(h->stream, h->ev_band[band_par][b], 0));
    CK(cudaEventRecord(k1, h->stream));
    if (!h->ktime_pending) h->ktime_pending = new std::vector<std::pair<cudaEvent_t, cudaEvent_t>>();
    h->ktime_pending->push_back(std::make_pair(k0, k1));
    kev.a = kev.b = nullptr;                        // now owned by the pending list
    h->ktime_pending_launches += n_sweeps;
    if (h->ktime_pending->size() > 4096) { double ms; long long nl; pmc_get_kernel_time(h, &ms, &nl); }
    CK(pmc4_launch_export(h->g4, ghost, h->v4_buf[cur], (float4 *)d_disk, d_n, h->stream)); h->launches += 1;
    return finish(h);
}

// start.cu:237-260, n_sweeps times.  Sweep t runs as ONE kernel that first applies the shift
// drawn at the end of sweep t-1 (while staging its tile) and then the four colours; the last
// shift is materialised by the stand-alone kernel so the caller's arrays are complete.
int pmc_sweep(pmc_handle *h, float *d_disk, int16_t *d_n, uint64_t sweep0, int n_sweeps)
{
    if (!h || !d_disk || !d_n || n_sweeps < 0) return PMC_E_INVALID;
    GUARD(h);
    if (n_sweeps == 0) return 0;
    if (h->v4_ok && !h->tune_generic) return sweep_v4(h, d_disk, d_n, sweep0, n_sweeps);
    int rc = ensure_scratch(h);
    if (rc) return rc;
    float4 *cur_d = (float4 *)d_disk, *oth_d = h->scratch_disk;
    int16_t *cur_n = d_n, *oth_n = h->scratch_n;
    int pend_on = 0, pend_f = 0;
    float pend_d = 0.0f;
    for (int t = 0; t < n_sweeps; t++) {
        uint64_t sweep = sweep0 + (uint64_t)t;
        int order[4], f;
        float d;
        pmc_schedule(h, sweep, order, &f, &d);
        SweepArgs a;
        memset(&a, 0, sizeof(a));
        for (int k = 0; k < 4; k++) {
            int off[2];
            pmc_colour_to_off(order[k], off);
            a.offx[k] = off[0]; a.offy[k] = off[1];
            a.offmask |= ((unsigned)off[0] | ((unsigned)off[1] << 1)) << (2 * k);
        }
        a.sweep_lo = (unsigned)sweep; a.sweep_hi = (unsigned)(sweep >> 32);
        a.shift_on = pend_on; a.shift_f = pend_f; a.shift_d = pend_d;
        a.sanitize_in = (t == 0);       // only the first kernel reads the caller's arrays
#ifdef PMC_DEBUG
        { const char *dbg = getenv("PMC_DBG_SKIP"); a.dbg_skip = dbg ? atoi(dbg) : 0; }
#endif
        CK(pmc_launch_fused_sweep(h->g, cur_d, cur_n, oth_d, oth_n, a, h->d_ctr, h->stream)); h->launches += 1;
        rc = exchange_ghosts_async(h, oth_d, oth_n);
        if (rc) return rc;
        float4 *td = cur_d; cur_d = oth_d; oth_d = td;
        int16_t *tn = cur_n; cur_n = oth_n; oth_n = tn;
        pend_on = 1; pend_f = f; pend_d = d;
    }
    CK(pmc_launch_shift(h->g, cur_d, cur_n, oth_d, oth_n, pend_f, pend_d, h->d_ctr, h->stream)); h->launches += 1;
    rc = exchange_ghosts_async(h, oth_d, oth_n);
    if (rc) return rc;
    if (oth_d != (float4 *)d_disk) {
        CK(cudaMemcpyAsync(d_disk, oth_d, pmc_disk_bytes(h), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(d_n, oth_n, pmc_n_bytes(h), cudaMemcpyDeviceToDevice, h->stream));
    }
    return finish(h);
}

// ------------------------------------------------------------------ counters / observables
int pmc_get_counters(pmc_handle *h, uint64_t *trials, uint64_t *accepted, uint64_t *lost, uint32_t *status)
{
    if (!h) return PMC_E_INVALID;
    GUARD(h);
    Counters c;
    CK(cudaMemcpyAsync(&c, h->d_ctr, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (trials) *trials = c.trials;
    if (accepted) *accepted = c.accepted;
    if (lost) *lost = c.lost;
    if (status) *status = c.status | h->status_sticky;
    return 0;
}

// device time spent in the fused sweep kernels (pmc_sweep fast path) since the last reset, and
// how many of them were launched: the per-launch time bench.py quotes against the roofline
int pmc_get_kernel_time(pmc_handle *h, double *ms, long long *launches)
{
    if (!h) return PMC_E_INVALID;
    GUARD(h);
    if (h->ktime_pending && !h->ktime_pending->empty()) {
        CK(cudaStreamSynchronize(h->stream));
        for (auto &pr : *h->ktime_pending) {
            float e = 0.0f;
            if (cudaEventElapsedTime(&e, pr.first, pr.second) == cudaSuccess) h->ktime_ms += e;
            cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
        }
        h->ktime_pending->clear();
        h->ktime_launches += h->ktime_pending_launches;
        h->ktime_pending_launches = 0;
    }
    if (ms) *ms = h->ktime_ms;
    if (launches) *launches = h->ktime_launches;
    return 0;
}

// kernels this handle launched since the last pmc_reset_counters (bench.py's gpu_launches)
int pmc_get_launch_count(pmc_handle *h, long long *launches)
{
    if (!h || !launches) return PMC_E_INVALID;
    *launches = h->launches;
    return 0;
}

int pmc_reset_counters(pmc_handle *h)
{
    if (!h) return PMC_E_INVALID;
    GUARD(h);
    h->launches = 0;
    h->status_sticky = 0;
    { double ms; long long nl; pmc_get_kernel_time(h, &ms, &nl); h->ktime_ms = 0.0; h->ktime_launches = 0; }
    CK(cudaMemsetAsync(h->d_ctr, 0, sizeof(Counters), h->stream));
    if (h->blocking) CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int pmc_check(pmc_handle *h, const float *d_disk, const int16_t *d_n, int64_t out[4], float *min_d2)
{
    if (!h || !d_disk || !d_n || !out || !min_d2) return PMC_E_INVALID;
    GUARD(h);
    CK(pmc_launch_check(h->g, (const float4 *)d_disk, d_n, h->d_out4, h->d_min, h->stream)); h->launches += 1;
    long long o[4];
    unsigned bits;
    CK(cudaMemcpyAsync(o, h->d_out4, sizeof(o), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&bits, h->d_min, sizeof(bits), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 4; i++) out[i] = o[i];
    float m;
    memcpy(&m, &bits, 4);
    *min_d2 = (bits == 0x7f7f7f7fu) ? 3.402823466e+38f : m;
    return 0;
}

int pmc_gr_hist(pmc_handle *h, const float *d_disk, const int16_t *d_n, float r_max, int nbins, uint64_t *hist_host)
{
    if (!h || !d_disk || !d_n || !hist_host || nbins < 1 || nbins > 4096) return PMC_E_INVALID;
    GUARD(h);
    if (!(r_max > 0.0f) || r_max > h->g.w) return PMC_E_INVALID;
    if (h->hist_cap < nbins) {
        cudaFree(h->d_hist); h->d_hist = nullptr; h->hist_cap = 0;
        CK(cudaMalloc(&h->d_hist, (size_t)nbins * sizeof(unsigned long long)));
        h->hist_cap = nbins;
    }
    CK(pmc_launch_gr_hist(h->g, (const float4 *)d_disk, d_n, r_max, nbins, h->d_hist, h->stream)); h->launches += 1;
    CK(cudaMemcpyAsync(hist_host, h->d_hist, (size_t)nbins * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// g(r) = pairs / (n_samples * N/2 * rho * shell area); g(sigma+) by a quadratic least-squares
// fit of the first bins at r >= sigma_d; beta P / rho = 1 + 2 phi g(sigma+)  (2-D virial).
int pmc_pressure_from_hist(const pmc_handle *h, const uint64_t *hist, float r_max, int nbins,
                           int64_t n_samples, double *g_of_r, double *g_contact, double *beta_p_over_rho)
{
    if (!h || !hist || nbins < 8 || n_samples < 1) return PMC_E_INVALID;
    const double N = (double)h->p.n_particles, Lb = h->g.L_box;
    const double rho = N / (Lb * Lb), dr = (double)r_max / nbins, sig = h->p.sigma_d;
    std::vector<double> gr(nbins);
    for (int b = 0; b < nbins; b++) {
        double r0 = b * dr, r1 = r0 + dr;
        double ideal = (double)n_samples * 0.5 * N * rho * M_PI * (r1 * r1 - r0 * r0);
        gr[b] = (double)hist[b] / ideal;
        if (g_of_r) g_of_r[b] = gr[b];
    }
    // first bin that lies entirely at r >= sigma
    int b0 = (int)ceil(sig / dr - 1e-9);
    int nfit = (int)ceil(0.08 * sig / dr);
    if (nfit < 4) nfit = 4;
    if (b0 + nfit > nbins) return PMC_E_INVALID;
    // least squares g = c0 + c1 t + c2 t^2, t = r_mid - sigma
    double S[5] = { 0, 0, 0, 0, 0 }, Tv[3] = { 0, 0, 0 };
    for (int k = 0; k < nfit; k++) {
        double t = (b0 + k + 0.5) * dr - sig, y = gr[b0 + k], tp = 1.0;
        for (int m = 0; m < 5; m++) { S[m] += tp; if (m < 3) Tv[m] += tp * y; tp *= t; }
    }
    double A[3][4] = { { S[0], S[1], S[2], Tv[0] }, { S[1], S[2], S[3], Tv[1] }, { S[2], S[3], S[4], Tv[2] } };
    for (int c = 0; c < 3; c++) {
        int piv = c;
        for (int r = c + 1; r < 3; r++) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
        for (int k = 0; k < 4; k++) { double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
        if (fabs(A[c][c]) < 1e-300) return PMC_E_INVALID;
        for (int r = 0; r < 3; r++) if (r != c) {
            double fct = A[r][c] / A[c][c];
            for (int k = c; k < 4; k++) A[r][k] -= fct * A[c][k];
        }
    }
    double gc = A[0][3] / A[0][0];
    double phi_eff = M_PI * sig * sig * rho / 4.0;
    if (g_contact) *g_contact = gc;
    if (beta_p_over_rho) *beta_p_over_rho = 1.0 + 2.0 * phi_eff * gc;
    return 0;
}

// ------------------------------------------------------------------ results back in the reference's order
// disk_to_r kernel.cu:497-507 on the device: d_r is SoA [2][N] global coordinates, cells in order, slots in
// order (owned rows only; a slab writes its own particles from index 0).  *n_found = particles met.
int pmc_disk_to_r(pmc_handle *h, const float *d_disk, const int16_t *d_n, float *d_r, int64_t *n_found)
{
    if (!h || !d_disk || !d_n || !d_r) return PMC_E_INVALID;
    GUARD(h);
    unsigned long long *scratch = nullptr;
    int nblocks = 0;
    CK(pmc_launch_disk_to_r(h->g, (const float4 *)d_disk, d_n, d_r, h->p.n_particles, h->p.n_particles, &scratch, &nblocks, h->stream));
    h->launches += 3;
    unsigned long long total = 0;
    cudaError_t e = cudaMemcpyAsync(&total, scratch + nblocks, sizeof(total), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFreeAsync(scratch, h->stream);
    if (e != cudaSuccess) return (int)e;
    if (n_found) *n_found = (int64_t)total;
    return 0;
}

// the same into HOST memory (start.cu:261-263 copies the cells back and prints them): converted on the device,
// then one D2H copy per coordinate through a pinned staging buffer
int pmc_disk_to_r_host(pmc_handle *h, const float *d_disk, const int16_t *d_n, float *r_host, int64_t *n_found)
{
    if (!h || !d_disk || !d_n || !r_host) return PMC_E_INVALID;
    GUARD(h);
    const long long N = h->p.n_particles;
    float *d_r = nullptr;
    CK(cudaMallocAsync(&d_r, (size_t)2 * N * sizeof(float), h->stream));
    int64_t found = 0;
    int rc = pmc_disk_to_r(h, d_disk, d_n, d_r, &found);
    if (!rc) {
        const size_t chunk = (size_t)8 << 20;       // floats per staged piece (32 MB pinned)
        float *stage = nullptr;
        cudaError_t e = cudaMallocHost(&stage, chunk * sizeof(float));
        for (size_t o = 0; e == cudaSuccess && o < (size_t)2 * N; o += chunk) {
            const size_t m = (size_t)2 * N - o < chunk ? (size_t)2 * N - o : chunk;
            e = cudaMemcpyAsync(stage, d_r + o, m * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            if (e == cudaSuccess) memcpy(r_host + o, stage, m * sizeof(float));
        }
        if (stage) cudaFreeHost(stage);
        if (e != cudaSuccess) rc = (int)e;
    }
    cudaFreeAsync(d_r, h->stream);
    if (n_found) *n_found = found;
    return rc;
}

int pmc_run_host(pmc_handle *h, const float *r_host, uint64_t sweep0, int n_sweeps, float *disk_host, int16_t *n_host)
{
    if (!h || !r_host || !disk_host || !n_host) return PMC_E_INVALID;
    GUARD(h);
    if (!h->run_r) CK(cudaMalloc(&h->run_r, pmc_r_bytes(h)));
    if (!h->run_disk) CK(cudaMalloc(&h->run_disk, pmc_disk_bytes(h)));
    if (!h->run_n) CK(cudaMalloc(&h->run_n, pmc_n_bytes(h)));
    const int was_blocking = h->blocking;
    h->blocking = 0;
    CK(cudaMemcpyAsync(h->run_r, r_host, pmc_r_bytes(h), cudaMemcpyHostToDevice, h->stream));
    int rc = pmc_assign(h, h->run_r, (float *)h->run_disk, h->run_n);
    if (!rc) rc = pmc_sweep(h, (float *)h->run_disk, h->run_n, sweep0, n_sweeps);
    h->blocking = was_blocking;
    if (rc) return rc;
    CK(cudaMemcpyAsync(disk_host, h->run_disk, pmc_disk_bytes(h), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(n_host, h->run_n, pmc_n_bytes(h), cudaMemcpyDeviceToHost, h->stream));
    // blocking handle (the default): synchronises; overflow / lost particles become the return code.
    // pmc_set_blocking(h, 0): returns with everything queued on the handle's stream - the host buffers are the
    // caller's until pmc_synchronize (two handles on two streams overlap one job's copies with the other's sweeps)
    return finish(h);
}

// ------------------------------------------------------------------ initial configurations / trajectory / checkpoint
// Random sequential addition (BASELINE north_star: "synthetic random-sequential-addition initial
// configurations").  RSA of disks jams at phi ~ 0.547, so this serves the dilute
// configurations; dense ones start from the reference's lattice (init_r).  Serial host code, a
// pure function of (params, seed): every rank of a slab run generates the same configuration.
int pmc_rsa_host(const pmc_params *pp, uint64_t seed, float *r_host, int64_t *attempts_out)
{
    if (!pp || !r_host) return PMC_E_INVALID;
    pmc_geometry pg;
    int rc = pmc_geometry_from_params(pp, &pg);
    if (rc) return rc;
    const long long N = pp->n_particles;
    const double L = (double)pg.cps * (double)pg.w, sig = (double)pp->sigma_d * (1.0 + 1e-5), sig2 = sig * sig;
    // insertion grid: cells of width >= sigma, at most 4 disks each (a 1 x 1 sigma square holds <= 4 centres)
    const int gc = (int)floor(L / sig);
    if (gc < 3) return PMC_E_INVALID;
    const double gw = L / gc;
    std::vector<int> cnt((size_t)gc * gc, 0);
    std::vector<float> gx((size_t)gc * gc * 4), gy((size_t)gc * gc * 4);
    long long placed = 0, attempts = 0;
    const long long max_attempts = N * 2000;
    while (placed < N && attempts < max_attempts) {
        uint32_t w4[4];
        host_philox(0xFFFFFFFEu, (uint32_t)attempts, (uint32_t)((uint64_t)attempts >> 32), 3u << 16, seed, w4);
        attempts++;
        // 53-bit uniforms in [0, 1) -> positions in (-L/2, L/2], stored as float
        const double u = ((double)(((uint64_t)w4[0] << 21) ^ (w4[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
        const double v = ((double)(((uint64_t)w4[2] << 21) ^ (w4[3] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
        const float xf = (float)(L * 0.5 - u * L), yf = (float)(L * 0.5 - v * L);
        const double x = xf, y = yf;
        if (!(x > -L * 0.5 + 1e-6 * L) || !(y > -L * 0.5 + 1e-6 * L) || x > L * 0.5 || y > L * 0.5) continue;
        int cx = (int)floor((x + L * 0.5) / gw), cy = (int)floor((y + L * 0.5) / gw);
        cx = cx < 0 ? 0 : (cx >= gc ? gc - 1 : cx); cy = cy < 0 ? 0 : (cy >= gc ? gc - 1 : cy);
        bool ok = true;
        for (int dy = -1; dy <= 1 && ok; dy++)
            for (int dx = -1; dx <= 1 && ok; dx++) {
                int nx = cx + dx, ny = cy + dy;
                double sx = 0.0, sy = 0.0;
                if (nx < 0) { nx += gc; sx = -L; } else if (nx >= gc) { nx -= gc; sx = L; }
                if (ny < 0) { ny += gc; sy = -L; } else if (ny >= gc) { ny -= gc; sy = L; }
                const size_t c = (size_t)ny * gc + nx;
                for (int k = 0; k < cnt[c]; k++) {
                    const double ddx = (double)gx[c * 4 + k] + sx - x, ddy = (double)gy[c * 4 + k] + sy - y;
                    if (ddx * ddx + ddy * ddy < sig2) { ok = false; break; }
                }
            }
        const size_t c = (size_t)cy * gc + cx;
        if (!ok || cnt[c] >= 4) continue;
        gx[c * 4 + cnt[c]] = xf; gy[c * 4 + cnt[c]] = yf; cnt[c]++;
        r_host[placed] = xf; r_host[placed + N] = yf;
        placed++;
    }
    if (attempts_out) *attempts_out = attempts;
    return placed == N ? 0 : PMC_E_UNSUPPORTED;     // jammed: phi too high for RSA
}

// One frame in the reference's trajectory format (create_dump kernel.cu:510-536, sample
// dumpR3.txt): LAMMPS / OVITO text, ids re-enumerated per frame by disk_to_r (kernel.cu:497-507),
// z = 0 in 2-D.  append = 0 truncates the file first.
int pmc_write_dump(pmc_handle *h, const float *d_disk, const int16_t *d_n, const char *path, int timestep, int append)
{
    if (!h || !d_disk || !d_n || !path) return PMC_E_INVALID;
    const long long N = h->p.n_particles;
    std::vector<float> r((size_t)2 * N);
    int64_t found = 0;
    int rc = pmc_disk_to_r_host(h, d_disk, d_n, r.data(), &found);
    if (rc) return rc;
    FILE *fp = fopen(path, append ? "a" : "w");
    if (!fp) return PMC_E_INVALID;
    const long long np = found < N ? found : N;
    const float hl = h->g.half_L;
    fprintf(fp, "ITEM: TIMESTEP \n%i\nITEM: NUMBER OF ATOMS\n%lld\nITEM: BOX BOUNDS\n%f %f\n%f %f\n%f %f\nITEM: ATOMS id type x y z ix iy iz\n",
            timestep, np, -hl, hl, -hl, hl, -0.5f, 0.5f);
    for (long long j = 0; j < np; j++)
        fprintf(fp, "%lld %lld %f %f %f 0 0 0\n", j + 1, j + 1, r[j], r[j + N], 0.0f);
    fclose(fp);
    return 0;
}

// Binary checkpoint / restart of (params, sweep, counters, status, disk, n): the reference has none
// (SURVEY section 5); needed for long equation-of-state runs.  The header is written field by field as
// little-endian fixed-width integers / IEEE bit patterns (no struct padding, no compiler dependence).
namespace {
constexpr uint32_t kCkptVersion = 3;
struct CkptWriter {
    std::vector<unsigned char> b;
    void u32(uint32_t v) { for (int i = 0; i < 4; i++) b.push_back((unsigned char)(v >> (8 * i))); }
    void u64(uint64_t v) { for (int i = 0; i < 8; i++) b.push_back((unsigned char)(v >> (8 * i))); }
    void f32(float f) { uint32_t v; memcpy(&v, &f, 4); u32(v); }
};
struct CkptReader {
    const unsigned char *p, *end;
    bool ok = true;
    uint32_t u32() { if (end - p < 4) { ok = false; return 0; } uint32_t v = 0; for (int i = 0; i < 4; i++) v |= (uint32_t)p[i] << (8 * i); p += 4; return v; }
    uint64_t u64() { if (end - p < 8) { ok = false; return 0; } uint64_t v = 0; for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i); p += 8; return v; }
    float f32() { uint32_t v = u32(); float f; memcpy(&f, &v, 4); return f; }
};
constexpr size_t kCkptHeaderBytes = 8 + 4 * 2 + 8 + 4 * 3 + 4 * 2 + 4 + 8 + 4 * 4 + 8 * 4 + 4 + 8 * 2;   // 124
}  // namespace

int pmc_save_checkpoint(pmc_handle *h, const float *d_disk, const int16_t *d_n, uint64_t sweep, const char *path)
{
    if (!h || !d_disk || !d_n || !path) return PMC_E_INVALID;
    GUARD(h);
    uint64_t trials, accepted, lost;
    uint32_t status;
    int rc = pmc_get_counters(h, &trials, &accepted, &lost, &status);
    if (rc) return rc;
    const pmc_params &p = h->p;
    CkptWriter wr;
    for (int i = 0; i < 8; i++) wr.b.push_back((unsigned char)"PMCB200"[i]);
    wr.u32(kCkptVersion); wr.u32(PMC_NMAX);
    wr.u64((uint64_t)p.n_particles); wr.f32(p.phi); wr.f32(p.sigma_d); wr.f32(p.cell_w);
    wr.u32((uint32_t)p.nmax); wr.u32((uint32_t)p.n_M); wr.f32(p.move_delta); wr.u64(p.seed);
    wr.u32((uint32_t)p.cps_multiple); wr.u32((uint32_t)p.rank); wr.u32((uint32_t)p.n_ranks); wr.u32((uint32_t)p.proposal);
    wr.u64(sweep); wr.u64(trials); wr.u64(accepted); wr.u64(lost); wr.u32(status);
    const uint64_t disk_bytes = pmc_disk_bytes(h), n_bytes = pmc_n_bytes(h);
    wr.u64(disk_bytes); wr.u64(n_bytes);
    if (wr.b.size() != kCkptHeaderBytes) return PMC_E_INVALID;
    std::vector<char> buf(disk_bytes + n_bytes);
    CK(cudaMemcpyAsync(buf.data(), d_disk, disk_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(buf.data() + disk_bytes, d_n, n_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    FILE *fp = fopen(path, "wb");
    if (!fp) return PMC_E_INVALID;
    bool ok = fwrite(wr.b.data(), 1, wr.b.size(), fp) == wr.b.size() && fwrite(buf.data(), 1, buf.size(), fp) == buf.size();
    ok = (fclose(fp) == 0) && ok;
    return ok ? 0 : PMC_E_INVALID;
}

// The state only continues the SAME chain under the geometry, slab layout, move size, trials per sub-sweep
// and random stream it was written with: every parameter must match (PMC_E_INVALID otherwise).
int pmc_load_checkpoint(pmc_handle *h, const char *path, float *d_disk, int16_t *d_n, uint64_t *sweep)
{
    if (!h || !d_disk || !d_n || !path) return PMC_E_INVALID;
    GUARD(h);
    FILE *fp = fopen(path, "rb");
    if (!fp) return PMC_E_INVALID;
    unsigned char hd[kCkptHeaderBytes];
    if (fread(hd, 1, sizeof(hd), fp) != sizeof(hd) || memcmp(hd, "PMCB200", 8) != 0) { fclose(fp); return PMC_E_INVALID; }
    CkptReader rd{ hd + 8, hd + sizeof(hd) };
    const uint32_t version = rd.u32(), nmax_file = rd.u32();
    pmc_params a;
    memset(&a, 0, sizeof(a));
    a.n_particles = (int64_t)rd.u64(); a.phi = rd.f32(); a.sigma_d = rd.f32(); a.cell_w = rd.f32();
    a.nmax = (int)rd.u32(); a.n_M = (int)rd.u32(); a.move_delta = rd.f32(); a.seed = rd.u64();
    a.cps_multiple = (int)rd.u32(); a.rank = (int)rd.u32(); a.n_ranks = (int)rd.u32(); a.proposal = (int)rd.u32();
    const uint64_t sw = rd.u64(), trials = rd.u64(), accepted = rd.u64(), lost = rd.u64();
    const uint32_t status = rd.u32();
    const uint64_t disk_bytes = rd.u64(), n_bytes = rd.u64();
    const pmc_params &b = h->p;
    if (!rd.ok || version != kCkptVersion || nmax_file != PMC_NMAX ||
        a.n_particles != b.n_particles || a.phi != b.phi || a.sigma_d != b.sigma_d || a.cell_w != b.cell_w ||
        a.nmax != b.nmax || a.n_M != b.n_M || a.move_delta != b.move_delta || a.seed != b.seed ||
        a.cps_multiple != b.cps_multiple || a.rank != b.rank || a.n_ranks != b.n_ranks || a.proposal != b.proposal ||
        disk_bytes != pmc_disk_bytes(h) || n_bytes != pmc_n_bytes(h)) { fclose(fp); return PMC_E_INVALID; }
    std::vector<char> buf(disk_bytes + n_bytes);
    bool ok = fread(buf.data(), 1, buf.size(), fp) == buf.size();
    fclose(fp);
    if (!ok) return PMC_E_INVALID;
    CK(cudaMemcpyAsync(d_disk, buf.data(), disk_bytes, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_n, buf.data() + disk_bytes, n_bytes, cudaMemcpyHostToDevice, h->stream));
    Counters c;
    memset(&c, 0, sizeof(c));
    c.trials = trials; c.accepted = accepted; c.lost = lost;
    CK(cudaMemcpyAsync(h->d_ctr, &c, sizeof(c), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->status_sticky = status;                      // already reported to whoever wrote the checkpoint
    if (sweep) *sweep = sw;
    return 0;
}

// ------------------------------------------------------------------ slab ring (NCCL)
int pmc_comm_unique_id(void *id128)
{
    if (!id128) return PMC_E_INVALID;
    if (!nccl_load()) return PMC_E_COMM;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id)) return PMC_E_COMM;
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int pmc_comm_init(pmc_handle *h, const void *id128)
{
    if (!h || !id128) return PMC_E_INVALID;
    if (h->p.n_ranks <= 1) return 0;
    if (!nccl_load()) return PMC_E_COMM;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    GUARD(h);
    if (g_nccl.CommInitRank(&h->comm, h->p.n_ranks, id, h->p.rank)) return PMC_E_COMM;
    return 0;
}

int pmc_exchange_ghosts(pmc_handle *h, float *d_disk, int16_t *d_n)
{
    if (!h || !d_disk || !d_n) return PMC_E_INVALID;
    GUARD(h);
    int rc = exchange_ghosts_async(h, (float4 *)d_disk, d_n);
    if (rc) return rc;
    return finish(h);
}

}  // extern "C"
