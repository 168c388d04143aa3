/*
 * pmc_oracle_lj.c -- CPU ORACLE of the 3-D Lennard-Jones mode (test infrastructure, NOT the product).
 *
 * The reference's ACTUAL physics: truncated LJ 4 (r^-12 - r^-6), cut at w
 * (calculate_pair_energy subsweep.h:90-103), 8 checkerboard colours, 26 neighbour cells
 * (get_neighbors subsweep.h:119-137 / make_nl kernel.cu:46-75), Metropolis acceptance at inverse temperature
 * beta (accept_move subsweep.h:194-217), the energy trace of V2 (kernel.cu:248,415,672-680) and the O(N^2)
 * host energy (calc_energy kernel.cu:452-470).  Arrays are the reference's own: r float[3][N],
 * disk float[cell][3][nmax] in GLOBAL coordinates, short n[cell], cell = cx + cy*cps + cz*cps^2
 * (start.cu:135-137, subsweep.h:14-16).
 *
 * So that the CUDA path can be compared BIT FOR BIT, every float operation is an individually rounded IEEE
 * binary32 operation in a fixed order (the order of the CUDA kernel: 32 lane-strided partial sums, then an
 * xor-butterfly), the pair energy uses 1 / r^2 instead of sqrtf + __powf, and the Metropolis test uses
 * pmc_exp_det, a polynomial exp built from fmaf only (1e-7 relative): no libm transcendental is on the
 * default path.  The reference's Gaussian proposal (curand_normal * sigma, subsweep.h:64) is the statistical
 * option `proposal = 1`; the default is uniform in the cube [-sigma, sigma]^3.
 * None of the reference's bugs listed in SURVEY H7 is reproduced.
 */
#include "pmc_oracle_lj.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed, uint32_t out[4])
{
    uint32_t ctr[4] = { c0, c1, c2, c3 }, key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    oracle_philox4x32_10(ctr, key, out);
}

int oracle_lj_make_geom(int64_t n_particles, float L, float beta, int cps, int nmax, int n_M, float sigma,
                        uint64_t seed, int proposal, oracle_lj_geom *g)
{
    if (n_particles <= 0 || !(L > 0.0f) || cps < 4 || (cps & 1) || cps > 1024 || nmax < 1 || nmax > 32 ||
        n_M < 1 || n_M > 64 || !(sigma > 0.0f) || !(beta >= 0.0f) || proposal < 0 || proposal > 1) return 1;
    memset(g, 0, sizeof(*g));
    g->n_particles = n_particles; g->L = L; g->half_L = L / 2.0f; g->beta = beta; g->cps = cps;
    g->n_cells = (int64_t)cps * cps * cps; g->nmax = nmax; g->n_M = n_M; g->sigma = sigma;
    g->w = L / (float)cps;                       /* start.cu:18: "L has to be a multiple of cell width" */
    g->rc2 = g->w * g->w;                        /* cut-off = w (subsweep.h:98) */
    g->dscale = sigma * 1.1920928955078125e-07f; /* 2^-23 */
    g->seed = seed; g->proposal = proposal;
    return 0;
}

/* start.cu:47-58: simple cubic lattice, r[dim][i] = L / 2.0 * (1.0 - float(2 i_dim + 1) / N_cube) */
int oracle_lj_init_r(const oracle_lj_geom *g, float *r)
{
    int64_t N = g->n_particles, nc = (int64_t)floor(cbrt((double)N) + 0.5);
    if (nc * nc * nc != N) return 1;
    for (int64_t iz = 0; iz < nc; iz++)
        for (int64_t iy = 0; iy < nc; iy++)
            for (int64_t ix = 0; ix < nc; ix++) {
                int64_t i = ix + iy * nc + iz * nc * nc;
                float f[3] = { (float)(2 * ix + 1) / (float)nc, (float)(2 * iy + 1) / (float)nc, (float)(2 * iz + 1) / (float)nc };
                for (int dim = 0; dim < 3; dim++) r[i + dim * N] = (float)((double)g->L / 2.0 * (1.0 - (double)f[dim]));
            }
    return 0;
}

static inline float xlb(const oracle_lj_geom *g, int c) { float cw = (float)c * g->w; return cw - g->half_L; }

/* start.cu:129-134: lb < x <= ub per axis; the unique such cell, -1 outside the box */
int oracle_lj_cell_of(const oracle_lj_geom *g, float x)
{
    if (!(x > xlb(g, 0)) || x > xlb(g, g->cps)) return -1;
    int c = (int)floorf((x + g->half_L) / g->w);
    if (c < 0) c = 0;
    if (c > g->cps - 1) c = g->cps - 1;
    while (c > 0 && !(x > xlb(g, c))) c--;
    while (c < g->cps - 1 && x > xlb(g, c + 1)) c++;
    return c;
}

/* assign start.cu:87-146: slot order = ascending atom index; returns particles lost (outside / overflow) */
int64_t oracle_lj_assign(const oracle_lj_geom *g, const float *r, float *disk, int16_t *n)
{
    const int nm = g->nmax;
    int64_t N = g->n_particles, lost = 0;
    memset(n, 0, (size_t)g->n_cells * sizeof(int16_t));
    memset(disk, 0, (size_t)g->n_cells * 3 * nm * sizeof(float));
    for (int64_t i = 0; i < N; i++) {
        int c[3];
        for (int dim = 0; dim < 3; dim++) c[dim] = oracle_lj_cell_of(g, r[i + dim * N]);
        if (c[0] < 0 || c[1] < 0 || c[2] < 0) { lost++; continue; }
        int64_t cell = c[0] + (int64_t)c[1] * g->cps + (int64_t)c[2] * g->cps * g->cps;
        if (n[cell] >= nm) { lost++; continue; }
        int s = n[cell]++;
        for (int dim = 0; dim < 3; dim++) disk[cell * 3 * nm + dim * nm + s] = r[i + dim * N];
    }
    return lost;
}

/* ------------------------------------------------------------------ energies */
float pmc_lj_pair(float dx, float dy, float dz, float rc2)
{
    float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    if (r2 > rc2) return 0.0f;                  /* subsweep.h:98-100: beyond the cut-off */
    float inv = 1.0f / r2;
    float i3 = (inv * inv) * inv;
    return ((i3 * i3) - i3) * 4.0f;             /* 4 (r^-12 - r^-6), subsweep.h:101-102 */
}

/* exp(x) for x <= 0 from individually rounded operations only (identical on CPU and GPU) */
float pmc_exp_det(float x)
{
    if (x < -87.0f) return 0.0f;
    float k = rintf(x * 1.44269504088896341f);
    float r = fmaf(k, -0.693145751953125f, x);          /* ln2 high part (exact product for |k| < 2^11) */
    r = fmaf(k, -1.42860682030941723e-06f, r);          /* ln2 low part */
    float p = 1.0f / 5040.0f;
    p = fmaf(p, r, 1.0f / 720.0f);
    p = fmaf(p, r, 1.0f / 120.0f);
    p = fmaf(p, r, 1.0f / 24.0f);
    p = fmaf(p, r, 1.0f / 6.0f);
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    return ldexpf(p, (int)k);
}

static inline int wrapc(int c, int cps) { return c < 0 ? c + cps : (c >= cps ? c - cps : c); }

/* The compact neighbourhood of one cell as V2 stages it (make_nl kernel.cu:46-75 order: entry 0 = self,
 * p = {0,-1,1}, x fastest; prefix-sum gather kernel.cu:256-278): own particles first, then the 26 neighbour
 * cells in that order, each shifted by its periodic image (apply_PBC subsweep.h:139-151). */
static int gather(const oracle_lj_geom *g, const float *disk, const int16_t *n, int cx, int cy, int cz,
                  float *lx, float *ly, float *lz)
{
    static const int p[3] = { 0, -1, 1 };
    const int nm = g->nmax, cps = g->cps;
    int tot = 0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) {
                int ux = cx + p[k], uy = cy + p[j], uz = cz + p[i];
                int nx = wrapc(ux, cps), ny = wrapc(uy, cps), nz = wrapc(uz, cps);
                float sx = ux < 0 ? -g->L : (ux >= cps ? g->L : 0.0f);
                float sy = uy < 0 ? -g->L : (uy >= cps ? g->L : 0.0f);
                float sz = uz < 0 ? -g->L : (uz >= cps ? g->L : 0.0f);
                int64_t c = nx + (int64_t)ny * cps + (int64_t)nz * cps * cps;
                const float *q = disk + c * 3 * nm;
                for (int s = 0; s < n[c]; s++) {
                    lx[tot] = q[s] + sx; ly[tot] = q[nm + s] + sy; lz[tot] = q[2 * nm + s] + sz;
                    tot++;
                }
            }
    return tot;
}

/* energy of a particle at (px, py, pz) with every staged particle except `skip`: 32 lane-strided partial
 * sums, then the xor-butterfly of the CUDA warp reduction (calculate_energy_in_cell + _in_neighbors
 * subsweep.h:105-117,153-172; V2 tree reduction kernel.cu:353-379) */
static float energy_at(const oracle_lj_geom *g, const float *lx, const float *ly, const float *lz, int tot, int skip,
                       float px, float py, float pz)
{
    float part[32];
    for (int lane = 0; lane < 32; lane++) {
        float e = 0.0f;
        for (int j = lane; j < tot; j += 32)
            if (j != skip) e = e + pmc_lj_pair(px - lx[j], py - ly[j], pz - lz[j], g->rc2);
        part[lane] = e;
    }
    for (int o = 16; o; o >>= 1) {
        float nxt[32];
        for (int lane = 0; lane < 32; lane++) nxt[lane] = part[lane] + part[lane ^ o];
        memcpy(part, nxt, sizeof(part));
    }
    return part[0];
}

static inline float uni_disp(uint32_t r, float dscale)
{
    float k = (float)(int)(r >> 9) - 4194304.0f;
    return fmaf(k, 2.0f, 1.0f) * dscale;         /* (2k + 1) * sigma * 2^-23: symmetric, one rounding */
}

/* one active cell of one colour: subsweep.h:250-298 (V2 kernel.cu:209-435) */
static void lj_cell(const oracle_lj_geom *g, float *disk, const int16_t *n, int cx, int cy, int cz, uint64_t sweep,
                    uint64_t *trials, uint64_t *accepted, double *dE)
{
    const int nm = g->nmax, cps = g->cps;
    const int64_t cell = cx + (int64_t)cy * cps + (int64_t)cz * cps * cps;
    const int cnt = n[cell];
    if (cnt == 0) return;                                   /* subsweep.h:252-254 */
    float lx[27 * 32], ly[27 * 32], lz[27 * 32];
    float *own = disk + cell * 3 * nm;
    /* random_shuffle subsweep.h:50-58 as intended: partial Fisher-Yates of the own slots (written back) */
    const int steps = g->n_M < cnt ? g->n_M : cnt;
    for (int s = 0; s < steps; s++) {
        uint32_t w4[4];         /* call (2 << 16) | (s / 8) feeds eight steps with 16 bits each */
        philox((uint32_t)cell, (uint32_t)sweep, (uint32_t)(sweep >> 32), (2u << 16) | (uint32_t)(s >> 3), g->seed, w4);
        uint32_t b16 = (w4[(s & 7) >> 1] >> (16 * (s & 1))) & 0xFFFFu;
        int j = s + (int)((b16 * (uint32_t)(cnt - s)) >> 16);
        for (int dim = 0; dim < 3; dim++) { float t = own[dim * nm + s]; own[dim * nm + s] = own[dim * nm + j]; own[dim * nm + j] = t; }
    }
    const int tot = gather(g, disk, n, cx, cy, cz, lx, ly, lz);
    const float lbx = xlb(g, cx), lby = xlb(g, cy), lbz = xlb(g, cz);
    const float ubx = xlb(g, cx + 1), uby = xlb(g, cy + 1), ubz = xlb(g, cz + 1);
    for (int s = 0; s < g->n_M; s++) {                      /* subsweep.h:279 */
        const int i = s % cnt;                              /* subsweep.h:291-296 */
        uint32_t w4[4];
        philox((uint32_t)cell, (uint32_t)sweep, (uint32_t)(sweep >> 32), (uint32_t)s, g->seed, w4);
        float ddx, ddy, ddz;
        if (g->proposal == 0) {
            ddx = uni_disp(w4[0], g->dscale); ddy = uni_disp(w4[1], g->dscale); ddz = uni_disp(w4[2], g->dscale);
        } else {                                            /* make_move subsweep.h:60-71: curand_normal * sigma */
            float u1 = ((float)((w4[0] >> 8) & 0x7FFFFFu) + 0.5f) * 1.1920928955078125e-07f;
            float u2 = ((float)(w4[1] >> 8) + 0.5f) * 5.9604644775390625e-08f;
            float u3 = ((float)((w4[2] >> 8) & 0x7FFFFFu) + 0.5f) * 1.1920928955078125e-07f;
            float u4 = ((float)(w4[3] >> 8) + 0.5f) * 5.9604644775390625e-08f;
            float ra = sqrtf(-2.0f * logf(u1)) * g->sigma, rb = sqrtf(-2.0f * logf(u3)) * g->sigma;
            ddx = ra * cosf(6.28318530717958648f * u2); ddy = ra * sinf(6.28318530717958648f * u2);
            ddz = rb * cosf(6.28318530717958648f * u4);
        }
        const float px = lx[i] + ddx, py = ly[i] + ddy, pz = lz[i] + ddz;
        (*trials)++;
        /* out_of_bound subsweep.h:73-88 (half-open like assign / shiftCells) */
        if (!(px > lbx && px <= ubx && py > lby && py <= uby && pz > lbz && pz <= ubz)) continue;
        const float e_old = energy_at(g, lx, ly, lz, tot, i, lx[i], ly[i], lz[i]);   /* subsweep.h:175-184 */
        const float e_new = energy_at(g, lx, ly, lz, tot, i, px, py, pz);            /* subsweep.h:186-191 */
        const float de = e_new - e_old;
        int acc = e_new < e_old;                            /* subsweep.h:209-211 */
        if (!acc) {
            /* Metropolis subsweep.h:212-216; the uniform is the trial's 4th word ((0, 1], 24 bits); the Gaussian
             * option spends that word on the z normal's angle and draws the uniform from one more call */
            uint32_t uw = w4[3];
            if (g->proposal == 1) { uint32_t x4[4]; philox((uint32_t)cell, (uint32_t)sweep, (uint32_t)(sweep >> 32), (3u << 16) | (uint32_t)s, g->seed, x4); uw = x4[0]; }
            const float u = (float)((uw >> 8) + 1u) * 5.9604644775390625e-08f;
            acc = u < pmc_exp_det(-(g->beta * de));
        }
        if (acc) {                                          /* cpy_proposed_to_D_sh subsweep.h:219-223 */
            lx[i] = px; ly[i] = py; lz[i] = pz;
            (*accepted)++;
            *dE += (double)de;                              /* d_Eblocks kernel.cu:248,415 */
        }
    }
    for (int s = 0; s < cnt; s++) { own[s] = lx[s]; own[nm + s] = ly[s]; own[2 * nm + s] = lz[s]; }   /* subsweep.h:29-36 */
}

void oracle_lj_subsweep(const oracle_lj_geom *g, float *disk, const int16_t *n, const int off[3], uint64_t sweep,
                        uint64_t *trials, uint64_t *accepted, double *dE)
{
    for (int cz = off[2]; cz < g->cps; cz += 2)
        for (int cy = off[1]; cy < g->cps; cy += 2)
            for (int cx = off[0]; cx < g->cps; cx += 2)
                lj_cell(g, disk, n, cx, cy, cz, sweep, trials, accepted, dE);
}

/* V2 shiftCells.h:23-112, global coordinates, operation order as there */
int64_t oracle_lj_shift_cells(const oracle_lj_geom *g, float *disk, int16_t *n, int f, float d)
{
    const int nm = g->nmax, cps = g->cps;
    const float w = g->w;
    size_t db = (size_t)g->n_cells * 3 * nm * sizeof(float), nb = (size_t)g->n_cells * sizeof(int16_t);
    float *src = (float *)malloc(db);
    int16_t *nsrc = (int16_t *)malloc(nb);
    memcpy(src, disk, db); memcpy(nsrc, n, nb);
    int64_t lost = 0;
    const int dir = (d <= 0.0f) ? -1 : 1;
    for (int64_t cell = 0; cell < g->n_cells; cell++) {
        int cid[3] = { (int)(cell % cps), (int)((cell / cps) % cps), (int)(cell / ((int64_t)cps * cps)) };
        const float offset = (float)cid[f] * w - g->half_L;                     /* shiftCells.h:50 */
        float *D_sh = disk + cell * 3 * nm;
        memset(D_sh, 0, (size_t)3 * nm * sizeof(float));
        const float *S = src + cell * 3 * nm;
        int nNew = 0;
        for (int i = 0; i < nsrc[cell]; i++) {                                  /* shiftCells.h:59-72 */
            float D = (S[f * nm + i] - offset) - d;
            if (D > 0.0f && D <= w) {
                for (int dim = 0; dim < 3; dim++) D_sh[dim * nm + nNew] = dim == f ? D + offset : S[dim * nm + i];
                nNew++;
            }
        }
        int nc[3] = { cid[0], cid[1], cid[2] };
        nc[f] = wrapc(nc[f] + dir, cps);                                        /* shiftCells.h:73-82 */
        const int64_t nbc = nc[0] + (int64_t)nc[1] * cps + (int64_t)nc[2] * cps * cps;
        const float off_nb = (float)nc[f] * w - g->half_L;
        const float sshift = w * (float)dir;                                    /* shiftCells.h:84-86 */
        const float *Q = src + nbc * 3 * nm;
        for (int i = 0; i < nsrc[nbc]; i++) {                                   /* shiftCells.h:91-102 */
            float D = (Q[f * nm + i] - off_nb) - d;
            if (!(D > 0.0f && D <= w)) {
                if (nNew < nm) {
                    for (int dim = 0; dim < 3; dim++) D_sh[dim * nm + nNew] = dim == f ? (D + offset) + sshift : Q[dim * nm + i];
                    nNew++;
                } else lost++;
            }
        }
        n[cell] = (int16_t)nNew;
    }
    free(src); free(nsrc);
    return lost;
}

/* host randomness start.cu:238,251-252 (ranges of kernel.cu:683-684) from (seed, sweep) */
void oracle_lj_schedule(const oracle_lj_geom *g, uint64_t sweep, int order[8], int *f, float *d)
{
    uint32_t a[4], b[4];
    philox(0xFFFFFFFFu, (uint32_t)sweep, (uint32_t)(sweep >> 32), (4u << 16) | 0u, g->seed, a);
    philox(0xFFFFFFFFu, (uint32_t)sweep, (uint32_t)(sweep >> 32), (4u << 16) | 1u, g->seed, b);
    for (int i = 0; i < 8; i++) order[i] = i;
    uint32_t pool[8] = { a[0], a[1], a[2], a[3], b[0], b[1], b[2], b[3] };
    for (int i = 7; i >= 1; i--) {                          /* FY_Shuffle start.cu:34-44, unbiased */
        int j = (int)(((uint64_t)pool[7 - i] * (uint64_t)(i + 1)) >> 32);
        int t = order[i]; order[i] = order[j]; order[j] = t;
    }
    *f = (int)(((uint64_t)b[3] * 3u) >> 32);                /* kernel.cu:683: 0, 1, 2 */
    float u = (float)((pool[7] >> 8) + 1u) * 5.9604644775390625e-08f;      /* (0, 1] */
    *d = (u - 0.5f) * g->w;                                 /* kernel.cu:684: (-w/2, w/2] */
}

void oracle_lj_colour_to_off(int colour, int off[3])        /* itoa start.cu:153-157 */
{
    off[2] = colour % 2; off[1] = (colour / 2) % 2; off[0] = (colour / 4) % 2;
}

/* start.cu:237-260 with the V2 energy trace: trace[t] = sum of accepted dE of sweep t (kernel.cu:672-680) */
int64_t oracle_lj_sweep(const oracle_lj_geom *g, float *disk, int16_t *n, uint64_t sweep0, int n_sweeps,
                        uint64_t *trials, uint64_t *accepted, double *trace)
{
    int64_t lost = 0;
    for (int t = 0; t < n_sweeps; t++) {
        uint64_t sweep = sweep0 + (uint64_t)t;
        int order[8], f, off[3];
        float d;
        double dE = 0.0;
        oracle_lj_schedule(g, sweep, order, &f, &d);
        for (int k = 0; k < 8; k++) {
            oracle_lj_colour_to_off(order[k], off);
            oracle_lj_subsweep(g, disk, n, off, sweep, trials, accepted, &dE);
        }
        if (trace) trace[t] = dE;
        lost += oracle_lj_shift_cells(g, disk, n, f, d);
    }
    return lost;
}

/* calc_energy kernel.cu:452-470: all pairs, minimum image, dist <= rc; double accumulation */
double oracle_lj_energy(const oracle_lj_geom *g, const float *disk, const int16_t *n)
{
    const int nm = g->nmax;
    int64_t tot = 0;
    for (int64_t c = 0; c < g->n_cells; c++) tot += n[c];
    float *x = (float *)malloc((size_t)tot * 3 * sizeof(float));
    int64_t k = 0;
    for (int64_t c = 0; c < g->n_cells; c++)
        for (int s = 0; s < n[c]; s++, k++)
            for (int dim = 0; dim < 3; dim++) x[k * 3 + dim] = disk[c * 3 * nm + dim * nm + s];
    double e = 0.0;
    for (int64_t i = 0; i < tot; i++)
        for (int64_t j = i + 1; j < tot; j++) {
            float dd[3];
            for (int dim = 0; dim < 3; dim++) {
                float del = fabsf(x[i * 3 + dim] - x[j * 3 + dim]);
                if (del > g->half_L) del -= g->L;
                dd[dim] = del;
            }
            e += (double)pmc_lj_pair(dd[0], dd[1], dd[2], g->rc2);
        }
    free(x);
    return e;
}

void oracle_lj_probe(const oracle_lj_geom *g, const float *disk, const int16_t *n, int cx, int cy, int cz, int slot,
                     float px, float py, float pz, int *oob, float *e_cell, float *e_nbrs)
{
    float lx[27 * 32], ly[27 * 32], lz[27 * 32];
    const int64_t cell = cx + (int64_t)cy * g->cps + (int64_t)cz * g->cps * g->cps;
    const int cnt = n[cell], tot = gather(g, disk, n, cx, cy, cz, lx, ly, lz);
    *oob = !(px > xlb(g, cx) && px <= xlb(g, cx + 1) && py > xlb(g, cy) && py <= xlb(g, cy + 1) &&
             pz > xlb(g, cz) && pz <= xlb(g, cz + 1));
    float ec = 0.0f, en = 0.0f;
    for (int j = 0; j < cnt; j++) if (j != slot) ec += pmc_lj_pair(px - lx[j], py - ly[j], pz - lz[j], g->rc2);
    for (int j = cnt; j < tot; j++) en += pmc_lj_pair(px - lx[j], py - ly[j], pz - lz[j], g->rc2);
    *e_cell = ec; *e_nbrs = en;
}
