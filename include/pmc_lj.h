/*
 * pmc_lj.h -- C-ABI of the 3-D Lennard-Jones mode: the reference's ACTUAL physics behind the same call sites.
 *
 * include/pmc.h serves the path BASELINE.json names (2-D hard disks).  This header serves what
 * qingye3/parallel-monte-carlo literally simulates: N Lennard-Jones particles,
 * 4 (r^-12 - r^-6) truncated at the cell width w (calculate_pair_energy subsweep.h:90-103), in a periodic cube,
 * 8 checkerboard colours, 26 neighbour cells, Metropolis acceptance at inverse temperature beta
 * (accept_move subsweep.h:194-217), with the V2 energy accounting (d_Eblocks kernel.cu:248,415; trace
 * kernel.cu:672-680; calc_energy kernel.cu:452-470).  Arrays are exactly the reference's:
 *     r    float[3][N]                 SoA, global coordinates in (-L/2, L/2]          (start.cu:54-56)
 *     disk float[cells][3][nmax]       disk[cell*3*nmax + dim*nmax + slot], GLOBAL     (start.cu:135-137)
 *     n    int16[cells]                particles per cell                              (start.cu:144)
 *     cell = cx + cy*cps + cz*cps^2                                                    (subsweep.h:14-16)
 * so a start.cu / kernel.cu driver can swap its four <<<>>> launches for these calls one for one.
 *
 * One warp works on one active cell (V2's "block per cell", kernel.cu:209-435, at warp granularity): the 27
 * cells are gathered compactly into shared memory in make_nl order (kernel.cu:46-75, 256-278), the lanes
 * share the pair energies of a trial and reduce them with shuffles, lane 0's decision is broadcast.
 * Results are bit-identical to oracle/pmc_oracle_lj.c for the default proposal (uniform in the cube
 * [-sigma, sigma]^3); PMC_PROPOSAL_GAUSSIAN is the reference's curand_normal * sigma, checked statistically.
 * Not reproduced: the reference's bugs.  Error codes, blocking semantics and the overflow /
 * lost reporting are those of pmc.h.
 */
#ifndef PMC_LJ_H
#define PMC_LJ_H
#include "pmc.h"

#ifdef __cplusplus
extern "C" {
#endif

/* the #define block start.cu:14-24 (V2 kernel.cu:17-30) as runtime values */
typedef struct pmc_lj_params {
    int64_t  n_particles;   /* N_ATOMS                       (a perfect cube for pmc_lj_init_r) */
    float    L;             /* L: box edge, a multiple of the cell width */
    float    beta;          /* beta */
    int      cells_per_side;/* cellsPerSide: even, >= 4; w = L / cellsPerSide is also the cut-off */
    int      nmax;          /* nmax: slots per cell, 1..32   (start.cu 10, kernel.cu 30) */
    int      n_M;           /* n_M: trials per active cell per sub-sweep, 1..64 */
    float    sigma;         /* sigma: proposal width */
    uint64_t seed;          /* 1234 (subsweep.h:259) */
    int      proposal;      /* PMC_PROPOSAL_UNIFORM (cube, bit-exact) or PMC_PROPOSAL_GAUSSIAN (subsweep.h:64) */
    int      device;        /* CUDA device ordinal, -1 = current */
} pmc_lj_params;

typedef struct pmc_lj_handle pmc_lj_handle;

int  pmc_lj_create(const pmc_lj_params *params, pmc_lj_handle **out);
int  pmc_lj_destroy(pmc_lj_handle *h);
size_t pmc_lj_r_bytes(const pmc_lj_handle *h);       /* 3 * N * 4            (rsize start.cu:186) */
size_t pmc_lj_disk_bytes(const pmc_lj_handle *h);    /* cells * 3 * nmax * 4 (disksize :188) */
size_t pmc_lj_n_bytes(const pmc_lj_handle *h);       /* cells * 2            (nsize :187) */
int  pmc_lj_set_stream(pmc_lj_handle *h, void *cuda_stream);

/* init_r<<<>>>(d_r, N_cube)                              start.cu:212 */
int  pmc_lj_init_r(pmc_lj_handle *h, float *d_r);
/* assign<<<>>>(d_r, d_disk, d_n)                         start.cu:227 */
int  pmc_lj_assign(pmc_lj_handle *h, const float *d_r, float *d_disk, int16_t *d_n);
/* cudaMemcpy(d_off) + subsweep_kernel<<<>>>(d_disk, d_n, d_off)   start.cu:242-245; off in {0,1}^3 */
int  pmc_lj_subsweep(pmc_lj_handle *h, float *d_disk, int16_t *d_n, const int off[3], uint64_t sweep);
/* shiftCells<<<>>>(d_disk, d_n, f, d)                    start.cu:255, f in {0,1,2}, d in (-w/2, w/2] */
int  pmc_lj_shift_cells(pmc_lj_handle *h, float *d_disk, int16_t *d_n, int f, float d);
/* FY_Shuffle + itoa + (f, d)                             start.cu:238,241,251-252 (ranges of kernel.cu:683-684) */
int  pmc_lj_schedule(const pmc_lj_handle *h, uint64_t sweep, int colour_order[8], int *f, float *d);
void pmc_lj_colour_to_off(int colour, int off[3]);
/* the loop body start.cu:237-260, n_sweeps times; trace_host (may be NULL): the sum of the accepted energy
 * changes of every sweep, i.e. energytrace[t+1] - energytrace[t] of kernel.cu:672-680 */
int  pmc_lj_sweep(pmc_lj_handle *h, float *d_disk, int16_t *d_n, uint64_t sweep0, int n_sweeps, double *trace_host);
/* total potential energy: calc_energy kernel.cu:452-470 (all pairs, minimum image, r <= w), on the device */
int  pmc_lj_energy(pmc_lj_handle *h, const float *d_disk, const int16_t *d_n, double *energy);
/* accept_counter kernel.cu:228,413 and the summed accepted energy change since the last reset */
int  pmc_lj_get_counters(pmc_lj_handle *h, uint64_t *trials, uint64_t *accepted, uint64_t *lost, uint32_t *status, double *dE);
int  pmc_lj_reset_counters(pmc_lj_handle *h);
/* disk_to_r kernel.cu:497-507 into HOST memory: r_host float[3][N], cells in order, slots in order */
int  pmc_lj_disk_to_r_host(pmc_lj_handle *h, const float *d_disk, const int16_t *d_n, float *r_host, int64_t *n_found);

#ifdef __cplusplus
}
#endif
#endif /* PMC_LJ_H */
