/*
 * pmc.h -- C-ABI of the B200-native hard-disk checkerboard Monte Carlo hot path.
 *
 * Drop-in boundary for the four kernel call sites of the reference driver
 * (qingye3/parallel-monte-carlo, start.cu:main, lines 169-272).  There is no FFI in the
 * reference; what a maintainer would bind are exactly the <<<>>> launches of main() plus
 * the per-sweep host randomness, so every entry point below names the line it replaces.
 * Plain pointers and sizes only; device pointers are owned by the caller exactly as main()
 * owns d_r / d_disk / d_n (start.cu:202-205, freed :266-269).
 *
 * Conventions kept from the reference:
 *   - errors are return codes, never aborts (start.cu:214-216 prints and continues):
 *       0            success
 *       > 0          a cudaError_t value
 *       < 0          PMC_E_* below
 *     A BLOCKING call during which a cell overflowed nmax or particles fell outside the box returns
 *     PMC_E_OVERFLOW / PMC_E_LOST (once per event; the arrays are complete except for the dropped
 *     particles, the `lost` counter and the status bits of pmc_get_counters keep the totals).  Non-blocking
 *     callers poll pmc_get_counters.  pmc_create refuses geometries whose mean cell occupancy leaves less
 *     than two free slots (PMC_E_UNSUPPORTED).
 *   - one host thread, calls are blocking unless pmc_set_blocking(h, 0)
 *   - array semantics [cell][dim][slot] + short counts (start.cu:135-137,144):
 *       disk : float[n_cells][2][nmax], n : int16[n_cells], cell = cx + cy*cps
 *
 * Documented deviations:
 *   - 2-D hard disks instead of 3-D Lennard-Jones; 4 checkerboard colours instead of 8.
 *   - coordinates stored in `disk` are CELL-LOCAL, in (0, w] (float32 global coordinates
 *     lose 2.4e-4 sigma at N=2^24); pmc_disk_to_r_host converts back to global coordinates.
 *     Unused slots hold x = PMC_SENTINEL, y = 0 (the reference leaves garbage).
 *   - COORDINATE GRID: every stored coordinate, the cell width w, every shift distance d and
 *     every trial displacement is an integer multiple of q = 2^e (pmc_geometry.grid_q; 2^-21
 *     for 2 <= w < 4), with e chosen so that every multiple of q below 2^(e+24) > 2w is exact
 *     in binary32.  Hence x - d, D +- w (shiftCells.h:62,97) and px -+ w (apply_PBC
 *     subsweep.h:139-151) are exact: the grid shift is an exact translation, a pair's squared
 *     distance fmaf(dx, dx, dy*dy) is the same number in every cell frame and after every
 *     later shift, and "no pair with d2 < sigma_d^2" is an EXACT invariant of a trajectory
 *     (pmc_check: overlaps == 0, min_d2 >= sigma_d^2, bit for bit).  pmc_assign snaps the
 *     incoming coordinates to the grid (<= q/2 = 2.4e-7), pmc_schedule draws d on the grid,
 *     pmc_shift_cells rounds a caller-chosen d to it, trial displacements are multiples of it (below).
 *   - compile-time #defines (start.cu:14-24) become the runtime pmc_params.
 *   - cuRAND XORWOW seeded identically on every launch (subsweep.h:259) becomes a
 *     counter-based Philox4x32-10 stream keyed on (seed, sweep, cell, trial).
 */
#ifndef PMC_H
#define PMC_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMC_SENTINEL 1.0e18f

#define PMC_E_INVALID      (-1)   /* bad argument / parameter */
#define PMC_E_UNSUPPORTED  (-2)   /* e.g. nmax != 8 */
#define PMC_E_OVERFLOW     (-3)   /* a cell exceeded nmax (reference: silent OOB write) */
#define PMC_E_LOST         (-4)   /* particles outside the box were dropped by assign */
#define PMC_E_NOT_SQUARE   (-5)   /* init_r needs a perfect-square particle count */
#define PMC_E_COMM         (-6)   /* multi-GPU communicator error */

/* status bits reported by pmc_get_counters */
#define PMC_STATUS_OVERFLOW 1u
#define PMC_STATUS_LOST     2u

/* start.cu:14-24  #define N_ATOMS / L / cellsPerSide / w / nmax / n_M / sigma / MCpasses
 * and the literal seed 1234 (subsweep.h:259). L is derived from (n_particles, phi, sigma_d). */
typedef struct pmc_params {
    int64_t  n_particles;   /* N_ATOMS */
    float    phi;           /* packing fraction; L = sqrt(N pi sigma_d^2 / (4 phi)) */
    float    sigma_d;       /* disk diameter */
    float    cell_w;        /* target cell width w (>= sigma_d); actual w = L / cellsPerSide */
    int      nmax;          /* slots per cell (this build: 8) */
    int      n_M;           /* trial moves per active cell per sub-sweep */
    float    move_delta;    /* proposal half-width (uniform square); reference `sigma` */
    uint64_t seed;          /* reference: 1234 */
    int      cps_multiple;  /* cellsPerSide is rounded down to a multiple of this (even; 0 -> 2) */
    int      device;        /* CUDA device ordinal, -1 = current */
    int      rank;          /* slab index in [0, n_ranks) */
    int      n_ranks;       /* number of slabs (GPUs); 0 or 1 = single GPU */
    int      proposal;      /* PMC_PROPOSAL_UNIFORM (0, default) or PMC_PROPOSAL_GAUSSIAN (1): see below */
} pmc_params;

/* Trial displacement (make_move subsweep.h:60-71; reference: x + curand_normal * sigma per axis):
 *   PMC_PROPOSAL_UNIFORM   uniform in the square on the coordinate grid: 4096 equally spaced levels per axis,
 *                          (2k - 4095) * A * q with k a 12-bit field of the trial's Philox word and
 *                          A = floor(move_delta / (4095 q)), i.e. half-width 4095 A q, at most 4095 q (0.002 at
 *                          w = 2) below move_delta (pmc_geometry.move_delta reports it; move_delta >= 4095 q is
 *                          required).  Exact arithmetic only: CPU oracle and GPU agree BIT FOR BIT; fast kernel;
 *   PMC_PROPOSAL_GAUSSIAN  the reference's N(0, move_delta^2) per axis (Box-Muller on the Philox words, rounded
 *                          to the grid, signs from separate bits: exactly symmetric).  Uses logf / sincospif, so
 *                          a CPU and the GPU agree statistically only (tests: acceptance and contact value within
 *                          3 sigma over seeds); served by the generic kernel. */
#define PMC_PROPOSAL_UNIFORM  0
#define PMC_PROPOSAL_GAUSSIAN 1

typedef struct pmc_geometry {
    int64_t n_particles;
    int     cps;            /* cellsPerSide */
    int64_t n_cells;        /* cps * cps (whole box) */
    int     nmax;
    int     n_M;
    float   w;              /* cell width */
    float   L;              /* box edge = cps * w */
    float   sigma_d;
    float   move_delta;
    int     row0;           /* first cell row owned by this rank */
    int     rows;           /* number of cell rows owned by this rank */
    int     ghost_rows;     /* ghost rows stored below and above the owned rows (0 if 1 rank) */
    int64_t local_cells;    /* (rows + 2*ghost_rows) * cps: cells in this rank's disk / n arrays */
    float   grid_q;         /* coordinate grid quantum (see "coordinate grid" above); move_delta and w are multiples */
} pmc_geometry;

typedef struct pmc_handle pmc_handle;

/* ---- lifetime (replaces the #define block start.cu:14-27 and cudaMalloc/cudaFree :202-205) */
int  pmc_create(const pmc_params *params, pmc_handle **out);
int  pmc_destroy(pmc_handle *h);
int  pmc_get_geometry(const pmc_handle *h, pmc_geometry *g);
/* the same derivation without a device or a handle (pure host maths) */
int  pmc_geometry_from_params(const pmc_params *params, pmc_geometry *g);
size_t pmc_r_bytes(const pmc_handle *h);      /* 2 * N * sizeof(float)            (rsize  start.cu:186) */
size_t pmc_disk_bytes(const pmc_handle *h);   /* local_cells * 2 * nmax * 4       (disksize :188) */
size_t pmc_n_bytes(const pmc_handle *h);      /* local_cells * sizeof(int16_t)    (nsize  :187) */
int  pmc_set_stream(pmc_handle *h, void *cuda_stream);   /* default: a handle-owned stream */
int  pmc_set_blocking(pmc_handle *h, int blocking);      /* default 1 (start.cu syncs after every launch) */
int  pmc_synchronize(pmc_handle *h);
/* Knobs that choose WHICH kernel / schedule computes the result, never the result itself (bit-identical for
 * every setting; tests use them to drive the rare paths): "bands" 1..16, "prefetch" >= 0, "overlap" 0/1,
 * "generic" 0/1, "four_plane" 0/1, "force_crowded" 0/1, "no_ns4" 0/1, "full_halo" 0/1, "tile_rows" 0 (automatic)
 * or an even number 2..28.  Unknown name: PMC_E_INVALID.
 * The library reads no environment variable that can change a result. */
int  pmc_set_tuning(pmc_handle *h, const char *name, int value);
const char *pmc_error_string(int code);

/* ---- the four kernel call sites */
/* init_r<<<>>>(d_r, N_cube)                 start.cu:212   r is SoA [2][N] global coordinates */
int  pmc_init_r(pmc_handle *h, float *d_r);
/* assign<<<>>>(d_r, d_disk, d_n)            start.cu:227 */
int  pmc_assign(pmc_handle *h, const float *d_r, float *d_disk, int16_t *d_n);
/* cudaMemcpy(d_off) + subsweep_kernel<<<>>>(d_disk, d_n, d_off)   start.cu:242-245 */
int  pmc_subsweep(pmc_handle *h, float *d_disk, int16_t *d_n, const int off[2], uint64_t sweep);
/* shiftCells<<<>>>(d_disk, d_n, f, d)       start.cu:255   f in {0,1}, d in (-w/2, w/2] */
int  pmc_shift_cells(pmc_handle *h, float *d_disk, int16_t *d_n, int f, float d);

/* ---- host-side per-sweep randomness: FY_Shuffle + itoa + (f, d)  start.cu:238,241,251-252 */
int  pmc_schedule(const pmc_handle *h, uint64_t sweep, int colour_order[4], int *f, float *d);
void pmc_colour_to_off(int colour, int off[2]);
/* How the fused sweep (pmc_sweep) tiles one sweep, from its colour order and shift alone (pure host
 * function, no handle, no GPU; exposed so that the halo argument can be tested independently):
 * out[0..3] = owned columns, owned rows, halo columns, halo rows of a tile;
 * out[4+k], out[8+k] = colour k (in execution order) is computed only for cells at least that
 * many columns / rows inside the staged region (>= 1). */
int  pmc_plan_sweep(const int colour_order[4], int f, float d, int out[12]);

/* ---- the loop body start.cu:237-260 as one call: n_sweeps x (4 sub-sweeps + shift),
 * sweeps numbered sweep0 .. sweep0+n_sweeps-1.  Fused fast path: one kernel per sweep and band of
 * tile rows.  The call may use a few internal CUDA streams (bands of a sweep; the slab boundary rows and
 * their NCCL ring); they fork from and join back into the handle's stream, so for the caller the whole
 * call is ordered on that stream exactly like a single kernel launch.  No host threads are created. */
int  pmc_sweep(pmc_handle *h, float *d_disk, int16_t *d_n, uint64_t sweep0, int n_sweeps);

/* ---- counters / observables (reference: unexported accept_counter kernel.cu:228,413) */
int  pmc_get_counters(pmc_handle *h, uint64_t *trials, uint64_t *accepted,
                      uint64_t *lost, uint32_t *status);
int  pmc_reset_counters(pmc_handle *h);
/* device time (CUDA events on the handle's stream) spent inside the fused sweep kernels of
 * pmc_sweep since the last pmc_reset_counters, and how many were launched (measurement aid) */
int  pmc_get_kernel_time(pmc_handle *h, double *ms, long long *launches);
/* CUDA kernels this handle launched since the last pmc_reset_counters */
int  pmc_get_launch_count(pmc_handle *h, long long *launches);
/* invariants: out[0]=sum n, out[1]=#coords outside (0,w], out[2]=#pairs closer than sigma_d,
 * out[3]=#unused slots without the sentinel; min_d2 = smallest squared pair distance */
int  pmc_check(pmc_handle *h, const float *d_disk, const int16_t *d_n, int64_t out[4], float *min_d2);
/* pair-distance histogram for g(r): r < r_max <= w, nbins <= 4096, hist is a HOST array */
int  pmc_gr_hist(pmc_handle *h, const float *d_disk, const int16_t *d_n,
                 float r_max, int nbins, uint64_t *hist_host);
/* g(r) normalisation + contact value + pressure beta*P/rho = 1 + 2 phi g(sigma+) (host maths) */
int  pmc_pressure_from_hist(const pmc_handle *h, const uint64_t *hist_host, float r_max, int nbins,
                            int64_t n_samples, double *g_of_r, double *g_contact, double *beta_p_over_rho);

/* ---- results back to the host: cudaMemcpy D2H + host_print_disk / disk_to_r
 * (start.cu:261-263, kernel.cu:497-507).  r_host is SoA [2][N] global coordinates in
 * cell-then-slot order; returns the particle count in *n_found. */
int  pmc_disk_to_r_host(pmc_handle *h, const float *d_disk, const int16_t *d_n,
                        float *r_host, int64_t *n_found);
/* the same conversion with the result left on the device (d_r: SoA [2][N], device pointer) */
int  pmc_disk_to_r(pmc_handle *h, const float *d_disk, const int16_t *d_n, float *d_r, int64_t *n_found);

/* ---- initial configurations, trajectory, checkpoint
 * Random sequential addition of n_particles disks (host code, no device needed; r_host is SoA
 * [2][N] global coordinates like init_r's output).  RSA jams near phi = 0.547: PMC_E_UNSUPPORTED
 * above that; dense configurations start from pmc_init_r (the reference's lattice). */
int  pmc_rsa_host(const pmc_params *params, uint64_t seed, float *r_host, int64_t *attempts);
/* one frame in the reference's dump format (create_dump kernel.cu:510-536, sample dumpR3.txt) */
int  pmc_write_dump(pmc_handle *h, const float *d_disk, const int16_t *d_n, const char *path,
                    int timestep, int append);
/* binary checkpoint / restart of (params, sweep, counters, disk, n); new */
int  pmc_save_checkpoint(pmc_handle *h, const float *d_disk, const int16_t *d_n, uint64_t sweep,
                         const char *path);
int  pmc_load_checkpoint(pmc_handle *h, const char *path, float *d_disk, int16_t *d_n, uint64_t *sweep);

/* ---- end-to-end with HOST buffers: H2D(r) -> assign -> n_sweeps sweeps -> D2H(disk, n).
 * This is what a start.cu-equivalent driver does around its loop (start.cu:227-262).
 * Blocking handles return when the results are in disk_host / n_host.  After pmc_set_blocking(h, 0) the call
 * returns with the copies and the sweeps queued on the handle's stream: r_host, disk_host and n_host (pinned)
 * belong to the library until pmc_synchronize(h); two handles on two streams then overlap the copies of one
 * job with the sweeps of the other. */
int  pmc_run_host(pmc_handle *h, const float *r_host, uint64_t sweep0, int n_sweeps,
                  float *disk_host, int16_t *n_host);

/* ---- multi-GPU slabs (new).  The unique id is an ncclUniqueId
 * (128 bytes) created by rank 0 and distributed by the caller (e.g. torch.distributed). */
int  pmc_comm_unique_id(void *id128);
int  pmc_comm_init(pmc_handle *h, const void *id128);
/* fill this rank's ghost rows from its ring neighbours (NCCL send/recv over NVLink) */
int  pmc_exchange_ghosts(pmc_handle *h, float *d_disk, int16_t *d_n);

#ifdef __cplusplus
}
#endif
#endif /* PMC_H */
