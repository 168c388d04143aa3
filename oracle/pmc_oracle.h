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
 *   - the sub-sweep itself uses a counter-based Philox stream instead of the
 *     reference's cuRAND XORWOW stream, so its trajectories are "parity unpinned"
 *     against the reference and pinned only structurally + statistically.
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
    float   delta;      /* proposal half-width */
    float   dscale;     /* delta * 2^-24 */
    double  L_box;      /* cps * (double)w */
    uint64_t seed;
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

/* a17: shiftCells (V2 shiftCells.h:23-112).  Returns number of particles lost. */
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
