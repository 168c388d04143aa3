/*
 * pmc_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).  See pmc_oracle.h.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off matters: every float operation below is an individually rounded IEEE
 * binary32 operation (the only fused operation is the explicit fmaf), which is what the
 * CUDA kernels reproduce with __fadd_rn/__fmul_rn/__fmaf_rn, so that CPU and GPU
 * trajectories are bit-identical.
 *
 * Each function cites the reference lines it restates (paths relative to /root/reference).
 */
#include "pmc_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ geometry */

/* start.cu:14-27 (#define N_ATOMS, L, cellsPerSide, w, nmax, n_M, sigma) as runtime values.
 * L from (N, phi, sigma_d); cps = even number of cells of width >= cell_w. */
int oracle_make_geom(int64_t n_particles, float phi, float sigma_d, float cell_w,
                     int nmax, int n_M, float move_delta, uint64_t seed,
                     int cps_multiple, oracle_geom *g)
{
    if (n_particles <= 0 || !(phi > 0.0f) || !(sigma_d > 0.0f) || !(cell_w >= sigma_d) ||
        nmax < 1 || nmax > 8 || n_M < 1 || n_M > 64 || !(move_delta > 0.0f))
        return 1;
    if (cps_multiple < 2) cps_multiple = 2;
    if (cps_multiple & 1) return 1;
    double L_d = sqrt((double)n_particles * M_PI * (double)sigma_d * (double)sigma_d /
                      (4.0 * (double)phi));
    int64_t cps = (int64_t)floor(L_d / ((double)cps_multiple * (double)cell_w)) * cps_multiple;
    if (cps < 4 || cps > 46340) return 2;
    double w_d = L_d / (double)cps;
    /* COORDINATE GRID.  Every stored coordinate, the cell width w, every shift distance d and
     * every trial displacement is an integer multiple of q = 2^e, with e chosen so that all
     * multiples of q below 2^(e+24) > 2w are exactly representable in binary32.  Then x - d,
     * D +- w (shiftCells.h:62,97), px -+ w (apply_PBC subsweep.h:139-151) and every coordinate
     * difference are EXACT, so the grid shift is an exact translation, a pair's squared distance
     * fmaf(dx, dx, dy*dy) is the same number in every cell frame and at every later sweep, and
     * "no pair with d2 < sigma2" is an exact, bit-level invariant of the trajectory. */
    int ex;
    (void)frexp(2.0 * w_d, &ex);                 /* 2w = m * 2^ex, m in [0.5, 1) */
    double q = ldexp(1.0, ex - 24);
    double K = nearbyint(w_d / q);               /* cell width in grid units, < 2^23 */
    double M = floor((double)move_delta / q);    /* proposal half-width in grid units */
    /* the uniform proposal has 4096 levels (2k - 4095) * A * q per axis, A = floor(M / 4095) >= 1 */
    if (M < 4095.0 || M >= 4194304.0 || (double)move_delta > w_d) return 1;
    if (2.0 * floor(M / 4095.0) * q * ldexp(1.0, 137) >= ldexp(1.0, 127)) return 1;   /* the CUDA path's scale factor must be a finite float */
    memset(g, 0, sizeof(*g));
    g->n_particles = n_particles;
    g->cps = (int)cps;
    g->n_cells = cps * cps;
    g->nmax = nmax;
    g->n_M = n_M;
    g->w = (float)(K * q);
    g->K = (int)K;
    g->M = (int)M;
    g->A = (int)floor(M / 4095.0);
    g->L_box = (double)cps * (double)g->w;
    g->L = (float)g->L_box;
    g->half_L = g->L / 2.0f;
    g->sigma = sigma_d;
    g->sigma2 = sigma_d * sigma_d;
    g->delta = (float)(M * q);                        /* move_delta rounded down to the grid */
    g->dscale = (float)q;                             /* one grid step */
    g->seed = seed;
    return 0;
}

/* ------------------------------------------------------------------ Philox4x32-10 */

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* counter layout: {cell id | 0xFFFFFFFF for host draws, sweep lo, sweep hi, (tag<<16)|call} */
static void trial_rng(const oracle_geom *g, uint32_t cell, uint64_t sweep, uint32_t call,
                      uint32_t out[4])
{
    uint32_t ctr[4] = { cell, (uint32_t)sweep, (uint32_t)(sweep >> 32), call };
    uint32_t key[2] = { (uint32_t)g->seed, (uint32_t)(g->seed >> 32) };
    oracle_philox4x32_10(ctr, key, out);
}

/* ------------------------------------------------------------------ init_r */

/* start.cu:47-58: r[index] = L / 2.0 * (1.0 - float(2*ix+1) / N_cube); 2-D: N_side = sqrt(N).
 * The division is float/int -> float; everything outside it is double (L / 2.0). */
int oracle_init_r(const oracle_geom *g, float *r)
{
    int64_t N = g->n_particles;
    int64_t ns = (int64_t)floor(sqrt((double)N) + 0.5);
    if (ns * ns != N) return 1;
    for (int64_t iy = 0; iy < ns; iy++)
        for (int64_t ix = 0; ix < ns; ix++) {
            int64_t i = ix + iy * ns;
            float fx = (float)(2 * ix + 1) / (float)ns;
            float fy = (float)(2 * iy + 1) / (float)ns;
            r[i]     = (float)((double)g->L / 2.0 * (1.0 - (double)fx));
            r[i + N] = (float)((double)g->L / 2.0 * (1.0 - (double)fy));
        }
    return 0;
}

/* ------------------------------------------------------------------ assign */

/* start.cu:129-131: xlb = cellx*w - L/2.0f (float); xub = xlb + w.  Canonical form of the
 * membership rule: the unique c with xlb(c) < x <= xlb(c+1). */
static inline float xlb(const oracle_geom *g, int c)
{
    float cw = (float)c * g->w;
    return cw - g->half_L;
}

int oracle_cell_of(const oracle_geom *g, float x)
{
    if (!(x > xlb(g, 0)) || x > xlb(g, g->cps)) return -1;
    int c = (int)floorf((x + g->half_L) / g->w);
    if (c < 0) c = 0;
    if (c > g->cps - 1) c = g->cps - 1;
    while (c > 0 && !(x > xlb(g, c))) c--;
    while (c < g->cps - 1 && x > xlb(g, c + 1)) c++;
    return c;
}

/* global -> cell-local, snapped to the coordinate grid: k * q with k in [1, K] */
static inline float to_local(const oracle_geom *g, float x, int c)
{
    double origin = (double)c * (double)g->w - g->L_box * 0.5;
    double k = nearbyint(((double)x - origin) / (double)g->dscale);   /* exact scaling, ties to even */
    if (k > (double)g->K) k = (double)g->K;
    if (k < 1.0) k = 1.0;
    return (float)(k * (double)g->dscale);
}

static inline float to_global(const oracle_geom *g, float xl, int c)
{
    double origin = (double)c * (double)g->w - g->L_box * 0.5;
    return (float)(origin + (double)xl);
}

static void clear_cells(const oracle_geom *g, float *disk, int16_t *n)
{
    int nm = g->nmax;
    for (int64_t c = 0; c < g->n_cells; c++) {
        n[c] = 0;
        for (int s = 0; s < nm; s++) {
            disk[c * 2 * nm + s] = PMC_SENTINEL;
            disk[c * 2 * nm + nm + s] = 0.0f;
        }
    }
}

/* start.cu:125-145: every cell scans atoms 0..N-1 in order and appends members, i.e. slot
 * order = ascending atom index.  Same result in O(N): visit atoms in order, append. */
int64_t oracle_assign(const oracle_geom *g, const float *r, float *disk, int16_t *n)
{
    int64_t N = g->n_particles, lost = 0;
    int nm = g->nmax;
    clear_cells(g, disk, n);
    for (int64_t i = 0; i < N; i++) {
        float x = r[i], y = r[i + N];
        int cx = oracle_cell_of(g, x), cy = oracle_cell_of(g, y);
        if (cx < 0 || cy < 0) { lost++; continue; }
        int64_t c = cx + (int64_t)cy * g->cps;
        if (n[c] >= nm) { lost++; continue; }
        int s = n[c]++;
        disk[c * 2 * nm + s]      = to_local(g, x, cx);
        disk[c * 2 * nm + nm + s] = to_local(g, y, cy);
    }
    return lost;
}

/* kernel.cu:497-507 disk_to_r: cells in order, slots in order; global coordinates. */
int64_t oracle_disk_to_r(const oracle_geom *g, const float *disk, const int16_t *n, float *r)
{
    int64_t N = g->n_particles, k = 0;
    int nm = g->nmax;
    for (int64_t c = 0; c < g->n_cells; c++) {
        int cx = (int)(c % g->cps), cy = (int)(c / g->cps);
        for (int s = 0; s < n[c]; s++) {
            if (k < N) {
                r[k]     = to_global(g, disk[c * 2 * nm + s], cx);
                r[k + N] = to_global(g, disk[c * 2 * nm + nm + s], cy);
            }
            k++;
        }
    }
    return k;
}

/* ------------------------------------------------------------------ sub-sweep */

static inline int wrap(int c, int cps) { return c < 0 ? c + cps : (c >= cps ? c - cps : c); }

static oracle_trial_record *g_trace = NULL;
static int64_t g_trace_cap = 0, g_trace_n = 0;
void oracle_set_trace(oracle_trial_record *buf, int64_t cap) { g_trace = buf; g_trace_cap = cap; g_trace_n = 0; }
int64_t oracle_trace_count(void) { return g_trace_n; }

/* accept_move subsweep.h:194-217 for hard disks: the proposal is rejected when it leaves the
 * cell (out_of_bound :73-88) or when the new energy is +inf, i.e. some other disk of the own
 * cell (calculate_energy_in_cell :105-117, j != i) or of the 8 neighbour cells
 * (get_neighbors :119-137 with helper {0,-1,1}, calculate_energy_in_neighbors :153-172)
 * lies closer than sigma_d.  apply_PBC (:139-151) is implicit in cell-local coordinates: the
 * trial point is expressed in the neighbour's frame by subtracting helper * w (exact on the
 * coordinate grid). */
int oracle_trial(const oracle_geom *g, const float *disk, const int16_t *n,
                 int cx, int cy, int slot, float px, float py, float *min_d2)
{
    static const int helper[3] = { 0, -1, 1 };            /* subsweep.h:120 */
    const int nm = g->nmax, cps = g->cps;
    const float w = g->w;
    const int64_t cell = cx + (int64_t)cy * cps;          /* subsweep.h:14-16 */
    const float *X = disk + cell * 2 * nm, *Y = X + nm;
    const int cnt = n[cell];
    /* out_of_bound subsweep.h:73-88 (half-open like assign / shiftCells) */
    if (!(px > 0.0f && px <= w && py > 0.0f && py <= w)) return 1;
    float md2 = FLT_MAX;
    for (int j = 0; j < cnt; j++) {                       /* subsweep.h:105-117 */
        if (j == slot) continue;
        float dx = px - X[j], dy = py - Y[j];
        float d2 = fmaf(dx, dx, dy * dy);
        if (d2 < md2) md2 = d2;
    }
    for (int i = 0; i < 3; i++)                           /* subsweep.h:119-137 */
        for (int j = 0; j < 3; j++) {
            if (i == 0 && j == 0) continue;
            int nx = wrap(cx + helper[i], cps), ny = wrap(cy + helper[j], cps);
            int64_t nb = nx + (int64_t)ny * cps;
            const float *QX = disk + nb * 2 * nm, *QY = QX + nm;
            float pxs = px - (float)helper[i] * w;
            float pys = py - (float)helper[j] * w;
            for (int k = 0; k < n[nb]; k++) {             /* subsweep.h:153-172 */
                float dx = pxs - QX[k], dy = pys - QY[k];
                float d2 = fmaf(dx, dx, dy * dy);
                if (d2 < md2) md2 = d2;
            }
        }
    if (min_d2) *min_d2 = md2;
    return md2 < g->sigma2 ? 2 : 0;
}

/* one active cell: subsweep.h:250-298 */
static void subsweep_cell(const oracle_geom *g, float *disk, const int16_t *n,
                          int cx, int cy, uint64_t sweep,
                          uint64_t *trials, uint64_t *accepted)
{
    const int nm = g->nmax, cps = g->cps;
    int64_t cell = cx + (int64_t)cy * cps;                /* subsweep.h:14-16 */
    int cnt = n[cell];
    if (cnt == 0) return;                                 /* subsweep.h:252-254 */
    float *X = disk + cell * 2 * nm, *Y = X + nm;         /* cpy_to_Dsh subsweep.h:18-27 */
    /* all random words of this cell's sub-sweep.  Uniform proposal: ONE 32-bit word per trial
     * (shuffle: bits 24-31, dx: bits 12-23, dy: bits 0-11), one Philox call feeds four trials.
     * Gaussian proposal: two words per trial (23 + 1 bits per axis), one call feeds two trials. */
    uint32_t words[2 * 64 + 4];
    const int per_call = g->proposal == 1 ? 2 : 4;
    for (int c = 0; c < (g->n_M + per_call - 1) / per_call; c++)
        trial_rng(g, (uint32_t)cell, sweep, (uint32_t)c, words + 4 * c);
    /* random_shuffle subsweep.h:50-58 as intended: a physical Fisher-Yates shuffle of the
     * cell's slots, written back with the cell like the reference's D_sh.  Only the first
     * min(n_M, cnt) positions are ever visited by the trial loop, so the shuffle stops there
     * (partial Fisher-Yates: positions 0..k-1 hold an ordered sample without replacement).  Step s
     * takes its random bits from trial s's word(s): the top byte (uniform proposal: 8 bits, the
     * choice among m <= 8 remaining slots is uniform to 1 part in 256 / m) or 16 bits (Gaussian).
     * Which disk a trial moves never depends on the positions, so detailed balance is not affected. */
    int steps = g->n_M < cnt ? g->n_M : cnt;
    for (int s = 0; s < steps; s++) {
        int j;
        if (g->proposal == 1) {
            uint32_t ra = words[2 * s], rb = words[2 * s + 1];
            uint32_t b16 = ((ra & 0xFFu) << 8) | (rb & 0xFFu);
            j = s + (int)((b16 * (uint32_t)(cnt - s)) >> 16);
        } else {
            j = s + (int)(((words[s] >> 24) * (uint32_t)(cnt - s)) >> 8);
        }
        float t;
        t = X[s]; X[s] = X[j]; X[j] = t;
        t = Y[s]; Y[s] = Y[j]; Y[j] = t;
    }
    for (int s = 0; s < g->n_M; s++) {                    /* subsweep.h:279 */
        uint32_t ra = g->proposal == 1 ? words[2 * s] : words[s], rb = g->proposal == 1 ? words[2 * s + 1] : 0u;
        int slot = s % cnt;                               /* i = (i+1) mod atom_counts, subsweep.h:291-296 */
        /* make_move subsweep.h:60-71; proposal = uniform in the square on the
         * coordinate grid: 4096 equally spaced levels m * q per axis, m = (2k - 4095) * A with k a 12-bit
         * field of the word (x: bits 12-23, y: bits 0-11; the shuffle took bits 24-31) and A = floor(M / 4095):
         * half-width 4095 A q, at most 4095 q below move_delta.  P(m) == P(-m) exactly (k -> 4095 - k); a
         * symmetric proposal is all detailed balance needs.  x + m*q is exact.  (The GPU reads the field in
         * place as the denormal float k * 2^-137 and evaluates fmaf(k 2^-137, 2 A q 2^137, x - 4095 A q):
         * the same number, every step exact.) */
        int mx = (2 * (int)((ra >> 12) & 0xFFFu) - 4095) * g->A;
        int my = (2 * (int)(ra & 0xFFFu) - 4095) * g->A;
        if (g->proposal == 1) {
            /* the reference's proposal, make_move subsweep.h:60-71: x + curand_normal * sigma per axis.
             * Box-Muller on bits 8..30 of the two words (radius from ra, angle in the first quadrant from
             * rb), the signs from bit 31 of each word: P(m) == P(-m) exactly whatever libm does. */
            float u1 = ((float)((ra >> 8) & 0x7FFFFFu) + 0.5f) * 1.1920928955078125e-07f;   /* (0, 1) */
            float u2 = ((float)((rb >> 8) & 0x7FFFFFu) + 0.5f) * 5.9604644775390625e-08f;   /* (0, 1/2) */
            float rr = sqrtf(-2.0f * logf(u1)) * (float)g->M;
            float ang = 3.14159265358979323846f * u2;
            mx = (int)rintf(rr * cosf(ang));
            my = (int)rintf(rr * sinf(ang));
            if (ra >> 31) mx = -mx;
            if (rb >> 31) my = -my;
        }
        float px = fmaf((float)mx, g->dscale, X[slot]);
        float py = fmaf((float)my, g->dscale, Y[slot]);
        (*trials)++;
        int verdict = oracle_trial(g, disk, n, cx, cy, slot, px, py, NULL);
        if (g_trace && g_trace_n < g_trace_cap) {
            oracle_trial_record *t = g_trace + g_trace_n;
            t->sweep = sweep; t->cx = cx; t->cy = cy; t->slot = slot; t->cnt = cnt;
            t->verdict = verdict; t->trial = s; t->px = px; t->py = py;
            for (int k = 0; k < 8; k++) { t->own_x[k] = k < cnt ? X[k] : PMC_SENTINEL; t->own_y[k] = k < cnt ? Y[k] : 0.0f; }
        }
        if (g_trace) g_trace_n++;
        if (verdict == 0) {                               /* accept_move subsweep.h:194-217 */
            X[slot] = px; Y[slot] = py;                   /* cpy_proposed_to_D_sh :219-223 */
            (*accepted)++;
        }
    }
}

void oracle_subsweep(const oracle_geom *g, float *disk, const int16_t *n,
                     const int off[2], uint64_t sweep,
                     uint64_t *trials, uint64_t *accepted)
{
    /* subsweep.h:242-245: cell = 2*tid + offset */
    for (int cy = off[1]; cy < g->cps; cy += 2)
        for (int cx = off[0]; cx < g->cps; cx += 2)
            subsweep_cell(g, disk, n, cx, cy, sweep, trials, accepted);
}

/* ------------------------------------------------------------------ shiftCells */

/* V2 shiftCells.h:23-112 for one destination cell, cell-local coordinates.
 * src/dst are distinct buffers (the reference separates reads and writes with
 * __syncthreads in its single block, shiftCells.h:104). */
static int shift_cell(const oracle_geom *g, const float *src, const int16_t *nsrc,
                      float *dst, int16_t *ndst, int cx, int cy, int f, float d)
{
    const int nm = g->nmax, cps = g->cps;
    const float w = g->w;
    int dir = (d <= 0.0f) ? -1 : 1;                        /* shiftCells.h:38-44 */
    int64_t cell = cx + (int64_t)cy * cps;
    const float *S = src + cell * 2 * nm;
    float *D_sh = dst + cell * 2 * nm;
    int nNew = 0, lost = 0;
    for (int s = 0; s < nm; s++) { D_sh[s] = PMC_SENTINEL; D_sh[nm + s] = 0.0f; }
    for (int i = 0; i < nsrc[cell]; i++) {                 /* shiftCells.h:59-72 */
        float D = S[f * nm + i] - d;
        if (D > 0.0f && D <= w) {
            if (nNew < nm) {
                D_sh[f * nm + nNew] = D;
                D_sh[(1 - f) * nm + nNew] = S[(1 - f) * nm + i];
                nNew++;
            } else lost++;
        }
    }
    int nc[2] = { cx, cy };                                /* shiftCells.h:73-82 */
    nc[f] = wrap(nc[f] + dir, cps);
    int64_t nb = nc[0] + (int64_t)nc[1] * cps;
    const float *Q = src + nb * 2 * nm;
    float sshift = w * (float)dir;                         /* shiftCells.h:84-86 (float s[3]) */
    for (int i = 0; i < nsrc[nb]; i++) {                   /* shiftCells.h:91-102 */
        float D = Q[f * nm + i] - d;
        if (!(D > 0.0f && D <= w)) {
            if (nNew < nm) {
                D_sh[f * nm + nNew] = D + sshift;
                D_sh[(1 - f) * nm + nNew] = Q[(1 - f) * nm + i];
                nNew++;
            } else lost++;
        }
    }
    ndst[cell] = (int16_t)nNew;                            /* shiftCells.h:105 */
    return lost;
}

int64_t oracle_shift_cells(const oracle_geom *g, float *disk, int16_t *n, int f, float d)
{
    d = (float)(nearbyint((double)d / (double)g->dscale) * (double)g->dscale);   /* onto the coordinate grid */
    size_t db = (size_t)g->n_cells * 2 * g->nmax * sizeof(float);
    size_t nb = (size_t)g->n_cells * sizeof(int16_t);
    float *src = (float *)malloc(db);
    int16_t *nsrc = (int16_t *)malloc(nb);
    memcpy(src, disk, db); memcpy(nsrc, n, nb);
    int64_t lost = 0;
    for (int cy = 0; cy < g->cps; cy++)
        for (int cx = 0; cx < g->cps; cx++)
            lost += shift_cell(g, src, nsrc, disk, n, cx, cy, f, d);
    free(src); free(nsrc);
    return lost;
}

/* ------------------------------------------------------------------ host loop */

/* itoa start.cu:153-157: r[2] = n%2, r[1] = (n/2)%2, r[0] = (n/4)%2; 2-D: 2 bits. */
void oracle_colour_to_off(int colour, int off[2])
{
    off[1] = colour % 2;
    off[0] = (colour / 2) % 2;
}

/* start.cu:238 FY_Shuffle(cboard_index), :251-252 (f, d) with the range fix of
 * kernel.cu:683-684, drawn from Philox(seed, sweep) so every rank agrees. */
void oracle_schedule(const oracle_geom *g, uint64_t sweep, int order[4], int *f, float *d)
{
    uint32_t a[4], b[4];
    trial_rng(g, 0xFFFFFFFFu, sweep, (1u << 16) | 0u, a);
    trial_rng(g, 0xFFFFFFFFu, sweep, (1u << 16) | 1u, b);
    for (int i = 0; i < 4; i++) order[i] = i;
    for (int i = 3; i >= 1; i--) {
        int j = (int)(((uint64_t)a[3 - i] * (uint64_t)(i + 1)) >> 32);
        int t = order[i]; order[i] = order[j]; order[j] = t;
    }
    *f = (int)(a[3] >> 31);
    /* d uniform on the grid points of (-w/2, w/2] (kernel.cu:684 range): dk * q */
    int64_t dk = (int64_t)(((uint64_t)b[0] * (uint64_t)g->K) >> 32) + 1 - (g->K + 1) / 2;
    *d = (float)dk * g->dscale;
}

int64_t oracle_sweep(const oracle_geom *g, float *disk, int16_t *n,
                     uint64_t sweep0, int n_sweeps,
                     uint64_t *trials, uint64_t *accepted)
{
    int64_t lost = 0;
    for (int t = 0; t < n_sweeps; t++) {                   /* start.cu:237 */
        uint64_t sweep = sweep0 + (uint64_t)t;
        int order[4], f, off[2];
        float d;
        oracle_schedule(g, sweep, order, &f, &d);          /* start.cu:238 */
        for (int k = 0; k < 4; k++) {                      /* start.cu:239-250 */
            oracle_colour_to_off(order[k], off);
            oracle_subsweep(g, disk, n, off, sweep, trials, accepted);
        }
        lost += oracle_shift_cells(g, disk, n, f, d);      /* start.cu:255 */
    }
    return lost;
}

/* thread count of oracle_sweep_omp: torchrun exports OMP_NUM_THREADS=1 to its workers, the CPU arm of
 * bench.py sets the count from sched_getaffinity explicitly instead.  Returns the count now in force. */
int oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

int oracle_sweep_omp(const oracle_geom *g, float *disk, int16_t *n,
                     uint64_t sweep0, int n_sweeps,
                     uint64_t *trials, uint64_t *accepted, int64_t *lost_out)
{
    size_t db = (size_t)g->n_cells * 2 * g->nmax * sizeof(float);
    size_t nbytes = (size_t)g->n_cells * sizeof(int16_t);
    float *tmp = (float *)malloc(db);
    int16_t *ntmp = (int16_t *)malloc(nbytes);
    uint64_t tr = 0, ac = 0;
    int64_t lost = 0;
    int threads = 1;
#ifdef _OPENMP
    threads = omp_get_max_threads();
#endif
    const int half = g->cps / 2;
    for (int t = 0; t < n_sweeps; t++) {
        uint64_t sweep = sweep0 + (uint64_t)t;
        int order[4], f, off[2];
        float d;
        oracle_schedule(g, sweep, order, &f, &d);
        for (int k = 0; k < 4; k++) {
            oracle_colour_to_off(order[k], off);
#pragma omp parallel for schedule(static) reduction(+ : tr, ac)
            for (int64_t q = 0; q < (int64_t)half * half; q++) {
                int cx = 2 * (int)(q % half) + off[0], cy = 2 * (int)(q / half) + off[1];
                uint64_t t1 = 0, a1 = 0;
                subsweep_cell(g, disk, n, cx, cy, sweep, &t1, &a1);
                tr += t1; ac += a1;
            }
        }
        memcpy(tmp, disk, db); memcpy(ntmp, n, nbytes);
#pragma omp parallel for schedule(static) reduction(+ : lost)
        for (int64_t c = 0; c < g->n_cells; c++)
            lost += shift_cell(g, tmp, ntmp, disk, n, (int)(c % g->cps), (int)(c / g->cps), f, d);
    }
    free(tmp); free(ntmp);
    *trials += tr; *accepted += ac; *lost_out += lost;
    return threads;
}

/* ------------------------------------------------------------------ invariants / observables */

/* unordered pair enumeration: same cell i<j, plus half of the 8 neighbours */
static const int HALF_NB[4][2] = { { 1, 0 }, { -1, 1 }, { 0, 1 }, { 1, 1 } };

void oracle_check(const oracle_geom *g, const float *disk, const int16_t *n,
                  int64_t out[4], float *min_d2)
{
    const int nm = g->nmax, cps = g->cps;
    const float w = g->w;
    int64_t total = 0, oob = 0, ov = 0, badsent = 0;
    float md2 = FLT_MAX;
    for (int cy = 0; cy < cps; cy++)
        for (int cx = 0; cx < cps; cx++) {
            int64_t c = cx + (int64_t)cy * cps;
            const float *X = disk + c * 2 * nm, *Y = X + nm;
            int cnt = n[c];
            total += cnt;
            for (int s = 0; s < nm; s++) {
                if (s < cnt) {
                    if (!(X[s] > 0.0f && X[s] <= w)) oob++;
                    if (!(Y[s] > 0.0f && Y[s] <= w)) oob++;
                } else if (X[s] != PMC_SENTINEL) badsent++;
            }
            for (int i = 0; i < cnt; i++) {
                for (int j = i + 1; j < cnt; j++) {
                    float dx = X[i] - X[j], dy = Y[i] - Y[j];
                    float d2 = fmaf(dx, dx, dy * dy);
                    if (d2 < md2) md2 = d2;
                    if (d2 < g->sigma2) ov++;
                }
                for (int h = 0; h < 4; h++) {
                    int nx = wrap(cx + HALF_NB[h][0], cps), ny = wrap(cy + HALF_NB[h][1], cps);
                    int64_t nb = nx + (int64_t)ny * cps;
                    const float *QX = disk + nb * 2 * nm, *QY = QX + nm;
                    float pxs = X[i] - (float)HALF_NB[h][0] * w;
                    float pys = Y[i] - (float)HALF_NB[h][1] * w;
                    for (int k = 0; k < n[nb]; k++) {
                        float dx = pxs - QX[k], dy = pys - QY[k];
                        float d2 = fmaf(dx, dx, dy * dy);
                        if (d2 < md2) md2 = d2;
                        if (d2 < g->sigma2) ov++;
                    }
                }
            }
        }
    out[0] = total; out[1] = oob; out[2] = ov; out[3] = badsent;
    *min_d2 = md2;
}

void oracle_gr_hist(const oracle_geom *g, const float *disk, const int16_t *n,
                    float r_max, int nbins, uint64_t *hist)
{
    const int nm = g->nmax, cps = g->cps;
    const float w = g->w;
    const float inv_dr = (float)nbins / r_max;
    const float rmax2 = r_max * r_max;
    memset(hist, 0, (size_t)nbins * sizeof(uint64_t));
    for (int cy = 0; cy < cps; cy++)
        for (int cx = 0; cx < cps; cx++) {
            int64_t c = cx + (int64_t)cy * cps;
            const float *X = disk + c * 2 * nm, *Y = X + nm;
            int cnt = n[c];
            for (int i = 0; i < cnt; i++) {
                for (int j = i + 1; j < cnt; j++) {
                    float dx = X[i] - X[j], dy = Y[i] - Y[j];
                    float d2 = fmaf(dx, dx, dy * dy);
                    if (d2 < rmax2) {
                        int b = (int)(sqrtf(d2) * inv_dr);
                        if (b < nbins) hist[b]++;
                    }
                }
                for (int h = 0; h < 4; h++) {
                    int nx = wrap(cx + HALF_NB[h][0], cps), ny = wrap(cy + HALF_NB[h][1], cps);
                    int64_t nb = nx + (int64_t)ny * cps;
                    const float *QX = disk + nb * 2 * nm, *QY = QX + nm;
                    float pxs = X[i] - (float)HALF_NB[h][0] * w;
                    float pys = Y[i] - (float)HALF_NB[h][1] * w;
                    for (int k = 0; k < n[nb]; k++) {
                        float dx = pxs - QX[k], dy = pys - QY[k];
                        float d2 = fmaf(dx, dx, dy * dy);
                        if (d2 < rmax2) {
                            int b = (int)(sqrtf(d2) * inv_dr);
                            if (b < nbins) hist[b]++;
                        }
                    }
                }
            }
        }
}
