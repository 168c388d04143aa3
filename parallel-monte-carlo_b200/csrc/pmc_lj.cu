// pmc_lj.cu -- the 3-D Lennard-Jones mode (include/pmc_lj.h): the reference's actual physics for sm_100a.
// One warp per active cell (V2's block-per-cell subSweep kernel.cu:209-435 at warp granularity): the 27 cells
// are gathered compactly into shared memory in make_nl order (kernel.cu:46-75, 256-278), the lanes split
// the pair energies of a trial (calculate_energy_in_cell / _in_neighbors subsweep.h:105-117,153-172), reduce
// them with an xor-butterfly of shuffles (V2: shared-memory tree kernel.cu:353-379) and take the Metropolis
// decision (accept_move subsweep.h:194-217) redundantly, so nothing is broadcast.
// Every float operation on the default path is an individually rounded intrinsic in the order of
// oracle/pmc_oracle_lj.c: results are bit-identical to the CPU oracle.
#include "pmc_internal.cuh"
#include "../../include/pmc_lj.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

struct LjGeom {
    long long n_particles, n_cells;
    int cps, nmax, n_M, proposal;
    float L, half_L, w, rc2, beta, sigma, dscale;
    unsigned seed_lo, seed_hi;
};

struct LjCounters {
    unsigned long long trials, accepted, lost;
    unsigned status, pad;
    double dE;
};

__device__ __forceinline__ float lj_xlb(int c, const LjGeom &g) { return __fadd_rn(__fmul_rn((float)c, g.w), -g.half_L); }

__device__ __forceinline__ int lj_cell_of(float x, const LjGeom &g)
{
    if (!(x > lj_xlb(0, g)) || x > lj_xlb(g.cps, g)) return -1;
    int c = (int)floorf(__fdiv_rn(__fadd_rn(x, g.half_L), g.w));
    c = c < 0 ? 0 : (c > g.cps - 1 ? g.cps - 1 : c);
    while (c > 0 && !(x > lj_xlb(c, g))) c--;
    while (c < g.cps - 1 && x > lj_xlb(c + 1, g)) c++;
    return c;
}

// subsweep.h:90-103 with 1 / r^2 in place of sqrtf + __powf (oracle pmc_lj_pair)
__device__ __forceinline__ float lj_pair(float dx, float dy, float dz, float rc2)
{
    const float r2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
    if (r2 > rc2) return 0.0f;
    const float inv = __fdiv_rn(1.0f, r2);
    const float i3 = __fmul_rn(__fmul_rn(inv, inv), inv);
    return __fmul_rn(__fadd_rn(__fmul_rn(i3, i3), -i3), 4.0f);
}

// exp(x), x <= 0, from individually rounded operations only (oracle pmc_exp_det)
__device__ __forceinline__ float exp_det(float x)
{
    if (x < -87.0f) return 0.0f;
    const float k = rintf(__fmul_rn(x, 1.44269504088896341f));
    float r = __fmaf_rn(k, -0.693145751953125f, x);
    r = __fmaf_rn(k, -1.42860682030941723e-06f, r);
    float p = 1.0f / 5040.0f;
    p = __fmaf_rn(p, r, 1.0f / 720.0f);
    p = __fmaf_rn(p, r, 1.0f / 120.0f);
    p = __fmaf_rn(p, r, 1.0f / 24.0f);
    p = __fmaf_rn(p, r, 1.0f / 6.0f);
    p = __fmaf_rn(p, r, 0.5f);
    p = __fmaf_rn(p, r, 1.0f);
    p = __fmaf_rn(p, r, 1.0f);
    return ldexpf(p, (int)k);
}

__device__ __forceinline__ float uni_disp(uint32_t r, float dscale)
{
    const float k = __fadd_rn((float)(int)(r >> 9), -4194304.0f);
    return __fmul_rn(__fmaf_rn(k, 2.0f, 1.0f), dscale);
}

// ------------------------------------------------------------------ init_r (start.cu:47-58)
__global__ void lj_init_r_kernel(float *__restrict__ r, long long N, int nc, float L)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int iz = (int)(i / ((long long)nc * nc)), iy = (int)((i / nc) % nc), ix = (int)(i % nc);
    const int idx[3] = { ix, iy, iz };
    const double hl = __ddiv_rn((double)L, 2.0);
#pragma unroll
    for (int dim = 0; dim < 3; dim++) {
        const float f = __fdiv_rn((float)(2 * idx[dim] + 1), (float)nc);
        r[i + dim * N] = (float)__dmul_rn(hl, __dsub_rn(1.0, (double)f));
    }
}

// ------------------------------------------------------------------ assign (start.cu:87-146)
__global__ void lj_assign_rank_kernel(const float *__restrict__ r, LjGeom g, unsigned *__restrict__ cnt32,
                                      unsigned *__restrict__ idx_tmp, uint2 *__restrict__ ovf, unsigned *ovf_count, LjCounters *ctr)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= g.n_particles) return;
    int c[3];
#pragma unroll
    for (int dim = 0; dim < 3; dim++) c[dim] = lj_cell_of(__ldg(r + i + dim * g.n_particles), g);
    if (c[0] < 0 || c[1] < 0 || c[2] < 0) {
        atomicAdd(&ctr->lost, 1ull);
        atomicOr(&ctr->status, PMC_STATUS_LOST);
        return;
    }
    const long long cell = c[0] + (long long)c[1] * g.cps + (long long)c[2] * g.cps * g.cps;
    const unsigned s = atomicAdd(cnt32 + cell, 1u);
    if (s < (unsigned)g.nmax) idx_tmp[cell * g.nmax + s] = (unsigned)i;
    else ovf[atomicAdd(ovf_count, 1u)] = make_uint2((unsigned)cell, (unsigned)i);
}

// one thread per cell: arrivals ordered by particle index (= the reference's slot order), gathered
__global__ void lj_assign_fill_kernel(const float *__restrict__ r, LjGeom g, const unsigned *__restrict__ cnt32,
                                      const unsigned *__restrict__ idx_tmp, const uint2 *__restrict__ ovf,
                                      const unsigned *__restrict__ ovf_count, float *__restrict__ disk,
                                      int16_t *__restrict__ n, LjCounters *ctr)
{
    const long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (cell >= g.n_cells) return;
    const int nm = g.nmax;
    const unsigned c32 = cnt32[cell];
    const int cnt = c32 > (unsigned)nm ? nm : (int)c32;
    unsigned v[32];
    for (int k = 0; k < cnt; k++) v[k] = idx_tmp[cell * nm + k];
    for (int a = 1; a < cnt; a++) {                     // insertion sort, <= 32 keys
        const unsigned key = v[a];
        int b = a - 1;
        while (b >= 0 && v[b] > key) { v[b + 1] = v[b]; b--; }
        v[b + 1] = key;
    }
    if (c32 > (unsigned)nm) {
        atomicOr(&ctr->status, PMC_STATUS_OVERFLOW);
        atomicAdd(&ctr->lost, (unsigned long long)(c32 - nm));
        const unsigned m = *ovf_count;                  // rare: keep the nmax LOWEST particle indices
        for (unsigned k = 0; k < m; k++) {
            const uint2 e = ovf[k];
            if (e.x != (unsigned)cell || e.y >= v[cnt - 1]) continue;
            int b = cnt - 2;
            while (b >= 0 && v[b] > e.y) { v[b + 1] = v[b]; b--; }
            v[b + 1] = e.y;
        }
    }
    float *p = disk + cell * 3 * nm;
    for (int k = 0; k < nm; k++)
#pragma unroll
        for (int dim = 0; dim < 3; dim++) p[dim * nm + k] = k < cnt ? __ldg(r + v[k] + dim * g.n_particles) : 0.0f;
    n[cell] = (int16_t)cnt;
}

// ------------------------------------------------------------------ sub-sweep (subsweep.h:240-300)
constexpr int kLjWarps = 4;

__device__ __forceinline__ int wrapc(int c, int cps) { return c < 0 ? c + cps : (c >= cps ? c - cps : c); }

// compact gather of the 27 cells around (cx, cy, cz) in make_nl order (entry 0 = self), periodic images applied;
// returns the number of staged particles, *cnt0 = particles of the own cell (the first entries)
__device__ __forceinline__ int lj_gather(const float *__restrict__ disk, const int16_t *__restrict__ n, const LjGeom &g,
                                         int cx, int cy, int cz, float *lx, float *ly, float *lz, int lane, int *cnt0)
{
    const int p3[3] = { 0, -1, 1 };
    const int cps = g.cps, nm = g.nmax;
    // lane k < 27 owns neighbour entry k = i*9 + j*3 + kk (z, y, x) like make_nl
    int mycnt = 0;
    long long mycell = 0;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    if (lane < 27) {
        const int i = lane / 9, j = (lane / 3) % 3, kk = lane % 3;
        const int ux = cx + p3[kk], uy = cy + p3[j], uz = cz + p3[i];
        mycell = wrapc(ux, cps) + (long long)wrapc(uy, cps) * cps + (long long)wrapc(uz, cps) * cps * cps;
        sx = ux < 0 ? -g.L : (ux >= cps ? g.L : 0.0f);
        sy = uy < 0 ? -g.L : (uy >= cps ? g.L : 0.0f);
        sz = uz < 0 ? -g.L : (uz >= cps ? g.L : 0.0f);
        const int c = __ldg(n + mycell);
        mycnt = c < 0 ? 0 : (c > nm ? nm : c);
    }
    int inc = mycnt;                                    // inclusive prefix sum over the 27 entries (kernel.cu:153-158)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    const int tot = __shfl_sync(0xffffffffu, inc, 31);
    *cnt0 = __shfl_sync(0xffffffffu, mycnt, 0);
    for (int k = 0; k < 27; k++) {
        const int ck = __shfl_sync(0xffffffffu, mycnt, k), off = __shfl_sync(0xffffffffu, inc, k) - ck;
        const long long cell = __shfl_sync(0xffffffffu, mycell, k);
        const float kx = __shfl_sync(0xffffffffu, sx, k), ky = __shfl_sync(0xffffffffu, sy, k), kz = __shfl_sync(0xffffffffu, sz, k);
        const float *q = disk + cell * 3 * nm;
        for (int s = lane; s < ck; s += 32) {
            lx[off + s] = __fadd_rn(q[s], kx);
            ly[off + s] = __fadd_rn(q[nm + s], ky);
            lz[off + s] = __fadd_rn(q[2 * nm + s], kz);
        }
    }
    __syncwarp();
    return tot;
}

__device__ __forceinline__ float lj_energy_at(const float *lx, const float *ly, const float *lz, int tot, int skip,
                                              float px, float py, float pz, float rc2, int lane)
{
    float e = 0.0f;
    for (int j = lane; j < tot; j += 32)
        if (j != skip) e = __fadd_rn(e, lj_pair(__fadd_rn(px, -lx[j]), __fadd_rn(py, -ly[j]), __fadd_rn(pz, -lz[j]), rc2));
#pragma unroll
    for (int o = 16; o; o >>= 1) e = __fadd_rn(e, __shfl_xor_sync(0xffffffffu, e, o));
    return e;
}

__global__ void __launch_bounds__(32 * kLjWarps)
lj_subsweep_kernel(float *__restrict__ disk, const int16_t *__restrict__ n, const LjGeom g, int ox, int oy, int oz,
                   unsigned sweep_lo, unsigned sweep_hi, LjCounters *ctr)
{
    extern __shared__ float lj_sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int half = g.cps / 2;
    const long long wg = (long long)blockIdx.x * kLjWarps + wid;
    if (wg >= (long long)half * half * half) return;
    const int cx = 2 * (int)(wg % half) + ox, cy = 2 * (int)((wg / half) % half) + oy, cz = 2 * (int)(wg / ((long long)half * half)) + oz;
    const long long cell = cx + (long long)cy * g.cps + (long long)cz * g.cps * g.cps;      // subsweep.h:242-245
    const int nm = g.nmax, cap = 27 * nm;
    float *lx = lj_sm + (size_t)wid * 3 * cap, *ly = lx + cap, *lz = ly + cap;
    int cnt;
    const int tot = lj_gather(disk, n, g, cx, cy, cz, lx, ly, lz, lane, &cnt);
    if (cnt == 0) return;                                                                   // subsweep.h:252-254
    // random_shuffle subsweep.h:50-58 as intended: partial Fisher-Yates of the own slots (the first entries)
    const int steps = g.n_M < cnt ? g.n_M : cnt;
    if (lane == 0) {
        uint32_t w4[4] = { 0, 0, 0, 0 };
        for (int s = 0; s < steps; s++) {
            if ((s & 7) == 0)
                philox4x32_10((uint32_t)cell, sweep_lo, sweep_hi, (2u << 16) | (uint32_t)(s >> 3), g.seed_lo, g.seed_hi, w4[0], w4[1], w4[2], w4[3]);
            const uint32_t b16 = (w4[(s & 7) >> 1] >> (16 * (s & 1))) & 0xFFFFu;
            const int j = s + (int)((b16 * (uint32_t)(cnt - s)) >> 16);
            float t;
            t = lx[s]; lx[s] = lx[j]; lx[j] = t;
            t = ly[s]; ly[s] = ly[j]; ly[j] = t;
            t = lz[s]; lz[s] = lz[j]; lz[j] = t;
        }
    }
    __syncwarp();
    const float lbx = lj_xlb(cx, g), lby = lj_xlb(cy, g), lbz = lj_xlb(cz, g);
    const float ubx = lj_xlb(cx + 1, g), uby = lj_xlb(cy + 1, g), ubz = lj_xlb(cz + 1, g);
    unsigned n_acc = 0;
    double dE = 0.0;
    for (int s = 0; s < g.n_M; s++) {                                                       // subsweep.h:279
        const int i = s % cnt;                                                              // subsweep.h:291-296
        uint32_t w0, w1, w2, w3;
        philox4x32_10((uint32_t)cell, sweep_lo, sweep_hi, (uint32_t)s, g.seed_lo, g.seed_hi, w0, w1, w2, w3);
        float ddx, ddy, ddz;
        if (g.proposal == PMC_PROPOSAL_UNIFORM) {
            ddx = uni_disp(w0, g.dscale); ddy = uni_disp(w1, g.dscale); ddz = uni_disp(w2, g.dscale);
        } else {                                                                            // make_move subsweep.h:60-71
            const float u1 = ((float)((w0 >> 8) & 0x7FFFFFu) + 0.5f) * 1.1920928955078125e-07f;
            const float u2 = ((float)(w1 >> 8) + 0.5f) * 5.9604644775390625e-08f;
            const float u3 = ((float)((w2 >> 8) & 0x7FFFFFu) + 0.5f) * 1.1920928955078125e-07f;
            const float u4 = ((float)(w3 >> 8) + 0.5f) * 5.9604644775390625e-08f;
            const float ra = sqrtf(-2.0f * logf(u1)) * g.sigma, rb = sqrtf(-2.0f * logf(u3)) * g.sigma;
            float sn, cs, sn2, cs2;
            sincospif(2.0f * u2, &sn, &cs);
            sincospif(2.0f * u4, &sn2, &cs2);
            ddx = ra * cs; ddy = ra * sn; ddz = rb * cs2;
        }
        const float xi = lx[i], yi = ly[i], zi = lz[i];
        const float px = __fadd_rn(xi, ddx), py = __fadd_rn(yi, ddy), pz = __fadd_rn(zi, ddz);
        // out_of_bound subsweep.h:73-88 (half-open like assign / shiftCells); warp-uniform
        if (!(px > lbx && px <= ubx && py > lby && py <= uby && pz > lbz && pz <= ubz)) continue;
        const float e_old = lj_energy_at(lx, ly, lz, tot, i, xi, yi, zi, g.rc2, lane);     // subsweep.h:175-184
        const float e_new = lj_energy_at(lx, ly, lz, tot, i, px, py, pz, g.rc2, lane);     // subsweep.h:186-191
        const float de = __fadd_rn(e_new, -e_old);
        bool acc = e_new < e_old;                                                           // subsweep.h:209-211
        if (!acc) {                                                                         // Metropolis subsweep.h:212-216
            uint32_t uw = w3;
            if (g.proposal != PMC_PROPOSAL_UNIFORM) {
                uint32_t x1, x2, x3;
                philox4x32_10((uint32_t)cell, sweep_lo, sweep_hi, (3u << 16) | (uint32_t)s, g.seed_lo, g.seed_hi, uw, x1, x2, x3);
            }
            const float u = __fmul_rn((float)((uw >> 8) + 1u), 5.9604644775390625e-08f);
            acc = u < exp_det(-__fmul_rn(g.beta, de));
        }
        if (acc) {                                                                          // cpy_proposed_to_D_sh subsweep.h:219-223
            __syncwarp();
            if (lane == 0) { lx[i] = px; ly[i] = py; lz[i] = pz; }
            __syncwarp();
            n_acc++;
            dE += (double)de;                                                               // d_Eblocks kernel.cu:248,415
        }
    }
    for (int s = lane; s < cnt; s += 32) {                                                  // cpy_D_sh_to_Disk subsweep.h:29-36
        float *p = disk + cell * 3 * nm;
        p[s] = lx[s]; p[nm + s] = ly[s]; p[2 * nm + s] = lz[s];
    }
    if (lane == 0) {
        atomicAdd(&ctr->trials, (unsigned long long)g.n_M);
        atomicAdd(&ctr->accepted, (unsigned long long)n_acc);
        if (n_acc) atomicAdd(&ctr->dE, dE);
    }
}

// ------------------------------------------------------------------ shiftCells (V2 shiftCells.h:23-112), out of place
__global__ void lj_shift_kernel(const float *__restrict__ src, const int16_t *__restrict__ nsrc, float *__restrict__ dst,
                                int16_t *__restrict__ ndst, LjGeom g, int f, float d, LjCounters *ctr)
{
    const long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (cell >= g.n_cells) return;
    const int nm = g.nmax, cps = g.cps;
    const float w = g.w;
    int cid[3] = { (int)(cell % cps), (int)((cell / cps) % cps), (int)(cell / ((long long)cps * cps)) };
    const int dir = (d <= 0.0f) ? -1 : 1;                                                   // shiftCells.h:38-44
    const float offset = __fadd_rn(__fmul_rn((float)cid[f], w), -g.half_L);                 // shiftCells.h:50
    float *D = dst + cell * 3 * nm;
    for (int k = 0; k < 3 * nm; k++) D[k] = 0.0f;
    const float *S = src + cell * 3 * nm;
    int cnt = nsrc[cell];
    cnt = cnt < 0 ? 0 : (cnt > nm ? nm : cnt);
    int nNew = 0, lost = 0;
    for (int i = 0; i < cnt; i++) {                                                         // shiftCells.h:59-72
        const float Dl = __fadd_rn(__fadd_rn(S[f * nm + i], -offset), -d);
        if (Dl > 0.0f && Dl <= w) {
            for (int dim = 0; dim < 3; dim++) D[dim * nm + nNew] = dim == f ? __fadd_rn(Dl, offset) : S[dim * nm + i];
            nNew++;
        }
    }
    int nc[3] = { cid[0], cid[1], cid[2] };
    nc[f] = wrapc(nc[f] + dir, cps);                                                        // shiftCells.h:73-82
    const long long nb = nc[0] + (long long)nc[1] * cps + (long long)nc[2] * cps * cps;
    const float off_nb = __fadd_rn(__fmul_rn((float)nc[f], w), -g.half_L);
    const float sshift = __fmul_rn(w, (float)dir);                                          // shiftCells.h:84-86
    const float *Q = src + nb * 3 * nm;
    int cq = nsrc[nb];
    cq = cq < 0 ? 0 : (cq > nm ? nm : cq);
    for (int i = 0; i < cq; i++) {                                                          // shiftCells.h:91-102
        const float Dl = __fadd_rn(__fadd_rn(Q[f * nm + i], -off_nb), -d);
        if (!(Dl > 0.0f && Dl <= w)) {
            if (nNew < nm) {
                for (int dim = 0; dim < 3; dim++) D[dim * nm + nNew] = dim == f ? __fadd_rn(__fadd_rn(Dl, offset), sshift) : Q[dim * nm + i];
                nNew++;
            } else lost++;
        }
    }
    ndst[cell] = (int16_t)nNew;
    if (lost) { atomicOr(&ctr->status, PMC_STATUS_OVERFLOW); atomicAdd(&ctr->lost, (unsigned long long)lost); }
}

// ------------------------------------------------------------------ total energy (calc_energy kernel.cu:452-470)
// one warp per cell: pairs inside the cell plus the 13 "forward" neighbour cells, every pair once
__global__ void __launch_bounds__(32 * kLjWarps)
lj_energy_kernel(const float *__restrict__ disk, const int16_t *__restrict__ n, const LjGeom g, double *out)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long cell = (long long)blockIdx.x * kLjWarps + wid;
    if (cell >= g.n_cells) return;
    const int nm = g.nmax, cps = g.cps;
    const int cx = (int)(cell % cps), cy = (int)((cell / cps) % cps), cz = (int)(cell / ((long long)cps * cps));
    int cnt = n[cell];
    cnt = cnt < 0 ? 0 : (cnt > nm ? nm : cnt);
    const float *P = disk + cell * 3 * nm;
    double e = 0.0;
    for (int t = lane; t < cnt * cnt; t += 32) {
        const int i = t / cnt, j = t % cnt;
        if (j > i) e += (double)lj_pair(__fadd_rn(P[i], -P[j]), __fadd_rn(P[nm + i], -P[nm + j]), __fadd_rn(P[2 * nm + i], -P[2 * nm + j]), g.rc2);
    }
    for (int dz = 0; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                if (!(dz > 0 || (dz == 0 && (dy > 0 || (dy == 0 && dx > 0))))) continue;
                const int ux = cx + dx, uy = cy + dy, uz = cz + dz;
                const long long nb = wrapc(ux, cps) + (long long)wrapc(uy, cps) * cps + (long long)wrapc(uz, cps) * cps * cps;
                const float sx = ux < 0 ? -g.L : (ux >= cps ? g.L : 0.0f), sy = uy < 0 ? -g.L : (uy >= cps ? g.L : 0.0f);
                const float sz = uz >= cps ? g.L : 0.0f;
                int cq = n[nb];
                cq = cq < 0 ? 0 : (cq > nm ? nm : cq);
                const float *Q = disk + nb * 3 * nm;
                for (int t = lane; t < cnt * cq; t += 32) {
                    const int i = t / cq, j = t % cq;
                    e += (double)lj_pair(__fadd_rn(P[i], -__fadd_rn(Q[j], sx)), __fadd_rn(P[nm + i], -__fadd_rn(Q[nm + j], sy)),
                                         __fadd_rn(P[2 * nm + i], -__fadd_rn(Q[2 * nm + j], sz)), g.rc2);
                }
            }
    for (int o = 16; o; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if (lane == 0 && e != 0.0) atomicAdd(out, e);
}

}  // namespace

// ------------------------------------------------------------------ C-ABI
struct pmc_lj_handle {
    pmc_lj_params p;
    LjGeom g;
    int device;
    cudaStream_t stream;
    bool own_stream;
    LjCounters *d_ctr, *h_ctr;
    float *scratch_disk;
    int16_t *scratch_n;
    double *d_energy;
    unsigned status_sticky;
};

#define CKL(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)
struct LjGuard {
    int prev = -1, dev;
    explicit LjGuard(int d) : dev(d) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; if (prev != dev) cudaSetDevice(dev); }
    ~LjGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};

// blocking like the reference (a sync after every launch); overflow / lost particles become the return code
static int lj_finish(pmc_lj_handle *h)
{
    CKL(cudaMemcpyAsync(h->h_ctr, h->d_ctr, sizeof(LjCounters), cudaMemcpyDeviceToHost, h->stream));
    CKL(cudaStreamSynchronize(h->stream));
    const unsigned st = h->h_ctr->status;
    if (!st) return 0;
    CKL(cudaMemsetAsync(&h->d_ctr->status, 0, sizeof(unsigned), h->stream));
    h->status_sticky |= st;
    return (st & PMC_STATUS_OVERFLOW) ? PMC_E_OVERFLOW : PMC_E_LOST;
}

static void lj_host_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed, uint32_t out[4])
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

extern "C" {

int pmc_lj_create(const pmc_lj_params *pp, pmc_lj_handle **out)
{
    if (!pp || !out) return PMC_E_INVALID;
    const pmc_lj_params &p = *pp;
    if (p.n_particles <= 0 || !(p.L > 0.0f) || p.cells_per_side < 4 || (p.cells_per_side & 1) || p.cells_per_side > 1024 ||
        p.n_M < 1 || p.n_M > 64 || !(p.sigma > 0.0f) || !(p.beta >= 0.0f) ||
        (p.proposal != PMC_PROPOSAL_UNIFORM && p.proposal != PMC_PROPOSAL_GAUSSIAN)) return PMC_E_INVALID;
    if (p.nmax < 1 || p.nmax > 32) return PMC_E_UNSUPPORTED;
    pmc_lj_handle *h = (pmc_lj_handle *)calloc(1, sizeof(pmc_lj_handle));
    if (!h) return PMC_E_INVALID;
    h->p = p;
    LjGeom &g = h->g;
    g.n_particles = p.n_particles; g.cps = p.cells_per_side; g.n_cells = (long long)g.cps * g.cps * g.cps;
    g.nmax = p.nmax; g.n_M = p.n_M; g.proposal = p.proposal;
    g.L = p.L; g.half_L = p.L / 2.0f; g.w = p.L / (float)g.cps; g.rc2 = g.w * g.w; g.beta = p.beta; g.sigma = p.sigma;
    g.dscale = p.sigma * 1.1920928955078125e-07f;
    g.seed_lo = (unsigned)p.seed; g.seed_hi = (unsigned)(p.seed >> 32);
    int caller = -1;
    cudaGetDevice(&caller);
    if (p.device >= 0 && cudaSetDevice(p.device) != cudaSuccess) { free(h); return PMC_E_INVALID; }
    cudaError_t e = cudaGetDevice(&h->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    h->own_stream = e == cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc(&h->d_ctr, sizeof(LjCounters));
    if (e == cudaSuccess) e = cudaMemset(h->d_ctr, 0, sizeof(LjCounters));
    if (e == cudaSuccess) e = cudaMallocHost(&h->h_ctr, sizeof(LjCounters));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_energy, sizeof(double));
    if (caller >= 0) cudaSetDevice(caller);
    if (e != cudaSuccess) { pmc_lj_destroy(h); return (int)e; }
    *out = h;
    return 0;
}

int pmc_lj_destroy(pmc_lj_handle *h)
{
    if (!h) return PMC_E_INVALID;
    LjGuard guard(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_ctr); cudaFree(h->scratch_disk); cudaFree(h->scratch_n); cudaFree(h->d_energy);
    if (h->h_ctr) cudaFreeHost(h->h_ctr);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    free(h);
    return 0;
}

size_t pmc_lj_r_bytes(const pmc_lj_handle *h) { return h ? (size_t)h->p.n_particles * 3 * sizeof(float) : 0; }
size_t pmc_lj_disk_bytes(const pmc_lj_handle *h) { return h ? (size_t)h->g.n_cells * 3 * h->g.nmax * sizeof(float) : 0; }
size_t pmc_lj_n_bytes(const pmc_lj_handle *h) { return h ? (size_t)h->g.n_cells * sizeof(int16_t) : 0; }

int pmc_lj_set_stream(pmc_lj_handle *h, void *cuda_stream)
{
    if (!h) return PMC_E_INVALID;
    LjGuard guard(h->device);
    CKL(cudaStreamSynchronize(h->stream));
    if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
    h->stream = (cudaStream_t)cuda_stream;
    return 0;
}

int pmc_lj_init_r(pmc_lj_handle *h, float *d_r)
{
    if (!h || !d_r) return PMC_E_INVALID;
    LjGuard guard(h->device);
    const long long N = h->p.n_particles, nc = (long long)floor(cbrt((double)N) + 0.5);
    if (nc * nc * nc != N) return PMC_E_NOT_SQUARE;
    lj_init_r_kernel<<<(unsigned)((N + 255) / 256), 256, 0, h->stream>>>(d_r, N, (int)nc, h->g.L);
    CKL(cudaGetLastError());
    return lj_finish(h);
}

int pmc_lj_assign(pmc_lj_handle *h, const float *d_r, float *d_disk, int16_t *d_n)
{
    if (!h || !d_r || !d_disk || !d_n) return PMC_E_INVALID;
    LjGuard guard(h->device);
    const LjGeom &g = h->g;
    unsigned *cnt32 = nullptr, *idx_tmp = nullptr;
    uint2 *ovf = nullptr;
    cudaStream_t st = h->stream;
    CKL(cudaMallocAsync(&cnt32, (size_t)(g.n_cells + 1) * sizeof(unsigned), st));
    cudaError_t e = cudaMallocAsync(&idx_tmp, (size_t)g.n_cells * g.nmax * sizeof(unsigned), st);
    if (e == cudaSuccess) e = cudaMallocAsync(&ovf, (size_t)g.n_particles * sizeof(uint2), st);
    if (e == cudaSuccess) {
        cudaMemsetAsync(cnt32, 0, (size_t)(g.n_cells + 1) * sizeof(unsigned), st);
        lj_assign_rank_kernel<<<(unsigned)((g.n_particles + 255) / 256), 256, 0, st>>>(d_r, g, cnt32, idx_tmp, ovf, cnt32 + g.n_cells, h->d_ctr);
        lj_assign_fill_kernel<<<(unsigned)((g.n_cells + 127) / 128), 128, 0, st>>>(d_r, g, cnt32, idx_tmp, ovf, cnt32 + g.n_cells, d_disk, d_n, h->d_ctr);
        e = cudaGetLastError();
    }
    cudaFreeAsync(cnt32, st);
    if (idx_tmp) cudaFreeAsync(idx_tmp, st);
    if (ovf) cudaFreeAsync(ovf, st);
    CKL(e);
    return lj_finish(h);
}

void pmc_lj_colour_to_off(int colour, int off[3])          // itoa start.cu:153-157
{
    off[2] = colour % 2; off[1] = (colour / 2) % 2; off[0] = (colour / 4) % 2;
}

int pmc_lj_schedule(const pmc_lj_handle *h, uint64_t sweep, int order[8], int *f, float *d)
{
    if (!h || !order || !f || !d) return PMC_E_INVALID;
    uint32_t a[4], b[4];
    lj_host_philox(0xFFFFFFFFu, (uint32_t)sweep, (uint32_t)(sweep >> 32), (4u << 16) | 0u, h->p.seed, a);
    lj_host_philox(0xFFFFFFFFu, (uint32_t)sweep, (uint32_t)(sweep >> 32), (4u << 16) | 1u, h->p.seed, b);
    const uint32_t pool[8] = { a[0], a[1], a[2], a[3], b[0], b[1], b[2], b[3] };
    for (int i = 0; i < 8; i++) order[i] = i;
    for (int i = 7; i >= 1; i--) {                          // FY_Shuffle start.cu:34-44, unbiased
        const int j = (int)(((uint64_t)pool[7 - i] * (uint64_t)(i + 1)) >> 32);
        const int t = order[i]; order[i] = order[j]; order[j] = t;
    }
    *f = (int)(((uint64_t)b[3] * 3u) >> 32);                // kernel.cu:683
    volatile float u = (float)((pool[7] >> 8) + 1u) * 5.9604644775390625e-08f;
    volatile float um = u - 0.5f;
    *d = um * h->g.w;                                       // kernel.cu:684: (-w/2, w/2]
    return 0;
}

static int lj_launch_subsweep(pmc_lj_handle *h, float *d_disk, int16_t *d_n, const int off[3], uint64_t sweep)
{
    const LjGeom &g = h->g;
    const long long half = g.cps / 2, active = half * half * half;
    const size_t smem = (size_t)kLjWarps * 3 * 27 * g.nmax * sizeof(float);
    if (smem > 48 * 1024) CKL(cudaFuncSetAttribute(lj_subsweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lj_subsweep_kernel<<<(unsigned)((active + kLjWarps - 1) / kLjWarps), 32 * kLjWarps, smem, h->stream>>>(
        d_disk, d_n, g, off[0], off[1], off[2], (unsigned)sweep, (unsigned)(sweep >> 32), h->d_ctr);
    return (int)cudaGetLastError();
}

int pmc_lj_subsweep(pmc_lj_handle *h, float *d_disk, int16_t *d_n, const int off[3], uint64_t sweep)
{
    if (!h || !d_disk || !d_n || !off || ((off[0] | off[1] | off[2]) & ~1)) return PMC_E_INVALID;
    LjGuard guard(h->device);
    int rc = lj_launch_subsweep(h, d_disk, d_n, off, sweep);
    if (rc) return rc;
    return lj_finish(h);
}

static int lj_launch_shift(pmc_lj_handle *h, float *d_disk, int16_t *d_n, int f, float d)
{
    if (!h->scratch_disk) CKL(cudaMalloc(&h->scratch_disk, pmc_lj_disk_bytes(h)));
    if (!h->scratch_n) CKL(cudaMalloc(&h->scratch_n, pmc_lj_n_bytes(h)));
    CKL(cudaMemcpyAsync(h->scratch_disk, d_disk, pmc_lj_disk_bytes(h), cudaMemcpyDeviceToDevice, h->stream));
    CKL(cudaMemcpyAsync(h->scratch_n, d_n, pmc_lj_n_bytes(h), cudaMemcpyDeviceToDevice, h->stream));
    lj_shift_kernel<<<(unsigned)((h->g.n_cells + 127) / 128), 128, 0, h->stream>>>(h->scratch_disk, h->scratch_n, d_disk, d_n, h->g, f, d, h->d_ctr);
    return (int)cudaGetLastError();
}

int pmc_lj_shift_cells(pmc_lj_handle *h, float *d_disk, int16_t *d_n, int f, float d)
{
    if (!h || !d_disk || !d_n || f < 0 || f > 2) return PMC_E_INVALID;
    if (!(fabsf(d) <= 0.5f * h->g.w * 1.0001f)) return PMC_E_INVALID;      // shiftCells.h:7 contract
    LjGuard guard(h->device);
    int rc = lj_launch_shift(h, d_disk, d_n, f, d);
    if (rc) return rc;
    return lj_finish(h);
}

int pmc_lj_sweep(pmc_lj_handle *h, float *d_disk, int16_t *d_n, uint64_t sweep0, int n_sweeps, double *trace_host)
{
    if (!h || !d_disk || !d_n || n_sweeps < 0) return PMC_E_INVALID;
    LjGuard guard(h->device);
    std::vector<double> marks;
    double *d_marks = nullptr;
    if (trace_host && n_sweeps > 0) CKL(cudaMallocAsync(&d_marks, (size_t)(n_sweeps + 1) * sizeof(double), h->stream));
    if (d_marks) CKL(cudaMemcpyAsync(d_marks, &h->d_ctr->dE, sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    for (int t = 0; t < n_sweeps; t++) {                    // start.cu:237
        const uint64_t sweep = sweep0 + (uint64_t)t;
        int order[8], f, off[3];
        float d;
        pmc_lj_schedule(h, sweep, order, &f, &d);           // :238, :251-252
        for (int k = 0; k < 8; k++) {                       // :239
            pmc_lj_colour_to_off(order[k], off);            // :241
            int rc = lj_launch_subsweep(h, d_disk, d_n, off, sweep);        // :242-245
            if (rc) return rc;
        }
        // energytrace[t + 1] - energytrace[t] = the accepted energy changes of this sweep (kernel.cu:672-680)
        if (d_marks) CKL(cudaMemcpyAsync(d_marks + t + 1, &h->d_ctr->dE, sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        int rc = lj_launch_shift(h, d_disk, d_n, f, d);     // :255
        if (rc) return rc;
    }
    if (d_marks) {
        marks.resize((size_t)n_sweeps + 1);
        cudaError_t e = cudaMemcpyAsync(marks.data(), d_marks, marks.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        cudaFreeAsync(d_marks, h->stream);
        CKL(e);
        for (int t = 0; t < n_sweeps; t++) trace_host[t] = marks[(size_t)t + 1] - marks[(size_t)t];
    }
    return lj_finish(h);
}

int pmc_lj_energy(pmc_lj_handle *h, const float *d_disk, const int16_t *d_n, double *energy)
{
    if (!h || !d_disk || !d_n || !energy) return PMC_E_INVALID;
    LjGuard guard(h->device);
    CKL(cudaMemsetAsync(h->d_energy, 0, sizeof(double), h->stream));
    lj_energy_kernel<<<(unsigned)((h->g.n_cells + kLjWarps - 1) / kLjWarps), 32 * kLjWarps, 0, h->stream>>>(d_disk, d_n, h->g, h->d_energy);
    CKL(cudaGetLastError());
    CKL(cudaMemcpyAsync(energy, h->d_energy, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CKL(cudaStreamSynchronize(h->stream));
    return 0;
}

int pmc_lj_get_counters(pmc_lj_handle *h, uint64_t *trials, uint64_t *accepted, uint64_t *lost, uint32_t *status, double *dE)
{
    if (!h) return PMC_E_INVALID;
    LjGuard guard(h->device);
    CKL(cudaMemcpyAsync(h->h_ctr, h->d_ctr, sizeof(LjCounters), cudaMemcpyDeviceToHost, h->stream));
    CKL(cudaStreamSynchronize(h->stream));
    if (trials) *trials = h->h_ctr->trials;
    if (accepted) *accepted = h->h_ctr->accepted;
    if (lost) *lost = h->h_ctr->lost;
    if (status) *status = h->h_ctr->status | h->status_sticky;
    if (dE) *dE = h->h_ctr->dE;
    return 0;
}

int pmc_lj_reset_counters(pmc_lj_handle *h)
{
    if (!h) return PMC_E_INVALID;
    LjGuard guard(h->device);
    h->status_sticky = 0;
    CKL(cudaMemsetAsync(h->d_ctr, 0, sizeof(LjCounters), h->stream));
    CKL(cudaStreamSynchronize(h->stream));
    return 0;
}

int pmc_lj_disk_to_r_host(pmc_lj_handle *h, const float *d_disk, const int16_t *d_n, float *r_host, int64_t *n_found)
{
    if (!h || !d_disk || !d_n || !r_host) return PMC_E_INVALID;
    LjGuard guard(h->device);
    const LjGeom &g = h->g;
    std::vector<float> disk(pmc_lj_disk_bytes(h) / sizeof(float));
    std::vector<int16_t> n((size_t)g.n_cells);
    CKL(cudaMemcpyAsync(disk.data(), d_disk, pmc_lj_disk_bytes(h), cudaMemcpyDeviceToHost, h->stream));
    CKL(cudaMemcpyAsync(n.data(), d_n, pmc_lj_n_bytes(h), cudaMemcpyDeviceToHost, h->stream));
    CKL(cudaStreamSynchronize(h->stream));
    const long long N = g.n_particles;
    long long k = 0;
    for (long long c = 0; c < g.n_cells; c++)               // disk_to_r kernel.cu:497-507
        for (int s = 0; s < n[(size_t)c]; s++, k++)
            if (k < N)
                for (int dim = 0; dim < 3; dim++) r_host[k + dim * N] = disk[(size_t)(c * 3 * g.nmax + dim * g.nmax + s)];
    if (n_found) *n_found = k;
    return 0;
}

}  // extern "C"
