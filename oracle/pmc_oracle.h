/*
 * pmc_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C restatement of the hot path of qingye3/parallel-monte-carlo
 * (checkerboard cell-list Metropolis sweep: assign -> 4 x subsweep -> shiftCells),
 * specialised to 2-D hard disks exactly as SURVEY.md section 0 / section 8 prescribe.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call anything in this directory.  The product
 * (parallel-monte-carlo_b200/csrc) never links or falls back to it.
 *
 * PARITY STATUS: the reference ships no tests, no golden vectors and no CPU path
 * (SURVEY.md section 4, section 8c).  What pins this oracle:
 *   - Philox4x32-10 against the published Random123 known-answer vectors
 *     (tests/test_oracle_cpu.py);
 *   - init_r against frame 0 of the reference's own dumpR3.txt (tests/golden/);
 *   - assign / shiftCells against outputs of the reference's own kernels
 *     (kernel.cu, V2 shiftCells.h) compiled from /root/reference into oracle/_ref and
 *     executed on a B200 (tests/golden/ref_*.json, generator oracle/ref_harness.cu);
 *   - the sub-sweep's per-trial decision (oracle_trial: out_of_bound, in-cell and neighbour
 *     energies, PBC) against the reference's OWN device functions out_of_bound,
 *     get_neighbors, apply_PBC, calculate_energy_in_cell, calculate_energy_in_neighbors
 *     (subsweep.h:73-172), compiled unmodified from /root/reference and run on a B200 on
 *     (state, proposal) probes: tests/golden/ref_trials_*.json, generator
 *     oracle/ref_harness_v1.cu + tests/golden/make_trial_probes.py;
 *   - the random stream is a counter-based Philox4x32-10 instead of the reference's cuRAND
 *     XORWOW stream (which the reference re-seeds identically on every launch,
 *     subsweep.h:259), so WHICH proposals are drawn is ours; what is done with a proposal
 *     is pinned as above.
 *
 * Layout contract shared with the CUDA library (include/pmc.h):
 *   disk : float[n_cells][2][nmax]   cell-major, then dim, then slot
 *          (reference: disk[cell*3*nmax + dim*nmax + slot], start.cu:135-137)
 *          coordinates are CELL-LOCAL, in (0, w]; unused slots: x = PMC_SENTINEL, y = 0
 *   n    : int16[n_cells]            particles per cell (start.cu:144)
 *   cell : cx + cy*cps               (subsweep.h:14-16)
 */
#ifndef PMC_ORACLE_H
#define PMC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMC_SENTINEL 1.0e18f

typedef struct {
    int64_t n_particles;
    int     cps;        /* cells per side (even, >= 4) */
    int64_t n_cells;
    int     nmax;
    int     n_M;
    float   w;          /* cell width */
    float   L;          /* (float)(cps * (double)w) */
    float   half_L;     /* L / 2 (exact halving) */
    float   sigma;      /* disk diameter */
    float   sigma2;     /* fl(sigma*sigma) */
    float   delta;      /* proposal half-width, rounded down to the grid: M * q */
    float   dscale;     /* q: the coordinate grid quantum (a power of two, see oracle_make_geom) */
    double  L_box;      /* cps * (double)w */
    uint64_t seed;
    int     K;          /* w / q: cell width in grid units (< 2^23) */
    int     M;          /* delta / q (>= 4095): the Gaussian proposal's sigma in grid units */
    int     proposal;   /* 0: uniform in the square [-delta, delta]^2 (default, bit-exact on CPU and GPU);
                         * 1: the reference's Gaussian, N(0, delta^2) per axis (make_move subsweep.h:64
                         *    curand_normal * sigma), rounded to the grid; libm transcendentals, so CPU and
                         *    GPU agree statistically, not bit for bit */
    int     A;          /* uniform proposal: m = (2k - 4095) * A, k a 12-bit field; A = M / 4095, half-width 4095 A q */
} oracle_geom;

/* a1: #define block start.cu:14-27 -> runtime geometry.  cps_multiple: 2 normally. */
int oracle_make_geom(int64_t n_particles, float phi, float sigma_d, float cell_w,
                     int nmax, int n_M, float move_delta, uint64_t seed,
                     int cps_multiple, oracle_geom *g);

/* Philox4x32-10 (Salmon et al. SC'11; Random123 v1.09 philox.h). */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* a3: init_r start.cu:47-58 -> 2-D square lattice, r is SoA [2][N]. */
int oracle_init_r(const oracle_geom *g, float *r);

/* a4: assign start.cu:87-146.  Returns number of particles lost (out of box / overflow). */
int64_t oracle_assign(const oracle_geom *g, const float *r, float *disk, int16_t *n);
/* cell id by the reference membership rule lb < x <= ub (start.cu:129-134), per axis. */
int oracle_cell_of(const oracle_geom *g, float x);

/* a5..a16: subsweep_kernel subsweep.h:240-300 for one colour off = (ox, oy). */
void oracle_subsweep(const oracle_geom *g, float *disk, const int16_t *n,
                     const int off[2], uint64_t sweep,
                     uint64_t *trials, uint64_t *accepted);

/* One trial of subsweep.h:279-297 for the particle in `slot` of cell (cx, cy) proposed at
 * (px, py) (cell-local): 0 = accepted, 1 = out of bound (subsweep.h:73-88), 2 = overlap
 * (the hard-disk form of new_energy = +inf, subsweep.h:105-117,153-172).  *min_d2 = smallest
 * squared distance to any other disk of the 3 x 3 block (FLT_MAX if none; only set when in
 * bounds).  This is the function oracle_subsweep itself calls for every trial. */
int oracle_trial(const oracle_geom *g, const float *disk, const int16_t *n,
                 int cx, int cy, int slot, float px, float py, float *min_d2);

/* Optional trace of every trial oracle_subsweep / oracle_sweep (serial versions) execute:
 * the own cell as the trial sees it (after the shuffle and the earlier trials), the proposal
 * and the verdict.  Used to generate the probes fed to the reference's device functions. */
typedef struct {
    uint64_t sweep;
    int32_t  cx, cy, slot, cnt, verdict, trial;
    float    px, py;
    float    own_x[8], own_y[8];
} oracle_trial_record;
void oracle_set_trace(oracle_trial_record *buf, int64_t cap);   /* NULL: off */
int64_t oracle_trace_count(void);

/* a17: shiftCells (V2 shiftCells.h:23-112).  Returns number of particles lost.
 * d is rounded to the coordinate grid first (oracle_schedule only produces on-grid d). */
int64_t oracle_shift_cells(const oracle_geom *g, float *disk, int16_t *n, int f, float d);

/* a18: host randomness start.cu:238,251-252 made deterministic from (seed, sweep). */
void oracle_schedule(const oracle_geom *g, uint64_t sweep, int colour_order[4], int *f, float *d);
void oracle_colour_to_off(int colour, int off[2]);   /* itoa start.cu:153-157 */

/* a19: host loop start.cu:237-260: n_sweeps x (4 subsweeps + shift), starting at sweep0. */
int64_t oracle_sweep(const oracle_geom *g, float *disk, int16_t *n,
                     uint64_t sweep0, int n_sweeps,
                     uint64_t *trials, uint64_t *accepted);
/* same, OpenMP over same-colour cells (timed CPU baseline). Returns threads used. */
int oracle_sweep_omp(const oracle_geom *g, float *disk, int16_t *n,
                     uint64_t sweep0, int n_sweeps,
                     uint64_t *trials, uint64_t *accepted, int64_t *lost);

int oracle_set_threads(int n);   /* OpenMP threads of oracle_sweep_omp (n <= 0: query only) */

/* disk/n -> global coordinates in reference layout order (disk_to_r kernel.cu:497-507). */
int64_t oracle_disk_to_r(const oracle_geom *g, const float *disk, const int16_t *n, float *r);

/* invariants: out[0]=sum n, out[1]=#coords outside (0,w], out[2]=#pairs with d2 < sigma2,
 * out[3]=#unused slots whose x != sentinel; *min_d2 = min pair d2 (float arithmetic). */
void oracle_check(const oracle_geom *g, const float *disk, const int16_t *n,
                  int64_t out[4], float *min_d2);

/* g(r) pair histogram, r < r_max <= w, nbins bins, each unordered pair counted once. */
void oracle_gr_hist(const oracle_geom *g, const float *disk, const int16_t *n,
                    float r_max, int nbins, uint64_t *hist);

#ifdef __cplusplus
}
#endif
#endif
