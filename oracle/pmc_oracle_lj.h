/*
 * pmc_oracle_lj.h -- CPU ORACLE of the 3-D Lennard-Jones mode (test infrastructure, NOT the product).
 * See pmc_oracle_lj.c.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it.
 * PARITY STATUS: assign / shiftCells are pinned to the outputs of the reference's own 3-D kernels
 * (tests/golden/ref_kernels_seed*.json), the pair / cell / neighbour energies and out_of_bound to the
 * reference's own device functions (tests/golden/ref_trials.json: e_full_cell, e_full_nbrs), to the tolerance
 * of the reference's sqrtf + __powf arithmetic; the random stream is ours (Philox).
 */
#ifndef PMC_ORACLE_LJ_H
#define PMC_ORACLE_LJ_H
#include <stdint.h>
#include "pmc_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int64_t n_particles;    /* N_ATOMS start.cu:14 */
    int64_t n_cells;        /* cps^3 */
    int     cps;            /* cellsPerSide start.cu:17 */
    int     nmax;           /* start.cu:19 (10), kernel.cu:24 (30); <= 32 here */
    int     n_M;            /* start.cu:21 */
    int     proposal;       /* 0 uniform cube (bit-exact), 1 Gaussian curand_normal * sigma (statistical) */
    float   L, half_L;      /* start.cu:15 */
    float   w;              /* L / cps start.cu:18; also the cut-off (subsweep.h:98) */
    float   rc2;            /* w^2 */
    float   beta;           /* start.cu:16 */
    float   sigma;          /* start.cu:22 */
    float   dscale;         /* sigma * 2^-23 */
    uint64_t seed;
} oracle_lj_geom;

int oracle_lj_make_geom(int64_t n_particles, float L, float beta, int cps, int nmax, int n_M, float sigma,
                        uint64_t seed, int proposal, oracle_lj_geom *g);
int oracle_lj_init_r(const oracle_lj_geom *g, float *r);
int oracle_lj_cell_of(const oracle_lj_geom *g, float x);
int64_t oracle_lj_assign(const oracle_lj_geom *g, const float *r, float *disk, int16_t *n);
float pmc_lj_pair(float dx, float dy, float dz, float rc2);
float pmc_exp_det(float x);
void oracle_lj_subsweep(const oracle_lj_geom *g, float *disk, const int16_t *n, const int off[3], uint64_t sweep,
                        uint64_t *trials, uint64_t *accepted, double *dE);
int64_t oracle_lj_shift_cells(const oracle_lj_geom *g, float *disk, int16_t *n, int f, float d);
void oracle_lj_schedule(const oracle_lj_geom *g, uint64_t sweep, int order[8], int *f, float *d);
void oracle_lj_colour_to_off(int colour, int off[3]);
int64_t oracle_lj_sweep(const oracle_lj_geom *g, float *disk, int16_t *n, uint64_t sweep0, int n_sweeps,
                        uint64_t *trials, uint64_t *accepted, double *trace);
double oracle_lj_energy(const oracle_lj_geom *g, const float *disk, const int16_t *n);
/* the two terms of calculate_new_energy (subsweep.h:186-191) for a proposal of particle `slot` of cell
 * (cx, cy, cz), plain left-to-right sums like the reference; *oob = out_of_bound (half-open) */
void oracle_lj_probe(const oracle_lj_geom *g, const float *disk, const int16_t *n, int cx, int cy, int cz, int slot,
                     float px, float py, float pz, int *oob, float *e_cell, float *e_nbrs);

#ifdef __cplusplus
}
#endif
#endif
