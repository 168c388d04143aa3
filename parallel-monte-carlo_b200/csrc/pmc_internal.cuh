// pmc_internal.cuh -- shared between the translation units of libpmc_b200.so.
// B200 (sm_100a) only.  Every float operation that takes part in a parity contract is an
// explicit round-to-nearest intrinsic (__fadd_rn / __fmul_rn / __fmaf_rn / __f*2_rn), so nvcc
// can neither contract nor reassociate it and results are bit-identical to oracle/pmc_oracle.c.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pmc.h"

#define PMC_NMAX 8          // slots per cell in this build: one cell = 4 x float4 = 64 B

// geometry + slab description, passed by value to every kernel
struct DevGeom {
    int cps;                // cellsPerSide (whole box)
    int row0;               // first owned cell row (global)
    int rows;               // owned rows
    int ghost;              // ghost rows on each side (0 when single rank)
    int local_rows;         // rows + 2*ghost
    int wrap_y;             // 1: single rank, rows wrap modulo cps
    int n_M;
    float w;
    float sigma;
    float sigma2;
    float dscale;           // q: the coordinate grid quantum (a power of two; pmc.h "coordinate grid")
    int K;                  // w / q: cell width in grid units (< 2^23)
    int M;                  // move_delta / q: trial displacements are m * q, m in [-M, M]
    int A;                  // uniform proposal: displacements are (2k - 4095) * A * q, k a 12-bit field; A = M / 4095
    float dstep, doff;      // 2 A q * 2^137 and 4095 A q: p = fmaf(grid_disp, dstep, x - doff) (grid_disp_x / _y below)
    int proposal;           // PMC_PROPOSAL_UNIFORM / PMC_PROPOSAL_GAUSSIAN
    float L;
    float half_L;
    double L_box;
    unsigned seed_lo, seed_hi;
    long long n_particles;
};

struct Counters {           // device-resident, one per handle
    unsigned long long trials;
    unsigned long long accepted;
    unsigned long long lost;
    unsigned int status;
    unsigned int pad;
};

struct SweepArgs {
    int offx[4], offy[4];   // colour offsets in execution order (itoa start.cu:153-157)
    unsigned offmask;       // the same, packed: bit 2k = offx[k], bit 2k+1 = offy[k]
    unsigned sweep_lo, sweep_hi;
    unsigned ph_e1, ph_e2, ph_e3;   // fused fast path: what rounds 0-2 of the cell's Philox call owe to (seed, sweep) alone (pmc4_philox_prepare)
    int sanitize_in;        // input comes from the caller: unused slots may hold garbage
    int shift_on;           // apply a pending shiftCells(f, d) while staging the tile
    int shift_f;
    float shift_d;
    int by_off;             // fused fast path: first tile row of this launch (boundary / interior split of slab runs)
    int by_n1, by_off2;     // grid rows >= by_n1 map to tile rows by_off2 + (row - by_n1) (second band)
    int prefetch_ahead;     // fused fast path: L2-prefetch the tile of block id + this (0 = off)
    // fused fast path: crowded-cell flags (one word per 2 x 2 block of internal cells, stamped with the
    // epoch of the sweep that stored a cell with 7 or 8 disks there; never cleared)
    const unsigned *flag_in;    // flags of the state being read, valid where == epoch_in
    unsigned *flag_out;         // flags of the state being written, stamped with epoch_out
    unsigned epoch_in, epoch_out;
    // fused fast path, filled by pmc4_plan_sweep from the colour order and the shift: owned extent of a
    // tile, halo per axis, and how far from the edge of the staged region colour k still has to be
    // computed (4 bits per colour, >= 1)
    int tx, ty, hx, hy;
    unsigned lo_x, lo_y;
    int sh_nseg, sh_inv, sh_inv2;   // shift_store_pass: kNT / tx and the multiply-shift reciprocals of tx and of kNT / tx
    // per colour, everything about "which cells are active" that depends on the tile only through the parity of its
    // first owned column / row (tile extents may be odd): bits 0-3 i0, 4-7 j0 (first active column / row of the
    // region), 8-11 / 12-15 lo_x / lo_y, 16-23 / 24-31 chunk offset of the first active cell / of its left neighbour
    // in the staged box (lane part excluded).  Index = (row0 & 1) * 8 + (col0 & 1) * 4 + k.
    unsigned colour_word[16];
    // result-invariant path selection (pmc_set_tuning; tests use it to drive the rare paths): 8 treat every tile
    // as crowded, 16 never use the 4-slot instantiation, 64 full halo for every colour order.
    // Builds with -DPMC_DEBUG additionally honour the phase-isolation bits (env PMC_DBG_SKIP) that DO change
    // results: 1 skip sub-sweeps, 2 skip shift, 4 skip store, 32 skip the crowded-cell flag lookup.  They
    // do not exist in the shipped library.
    int dbg_skip;
};
#ifdef PMC_DEBUG
#define PMC_DBG_BIT(a, bit) ((a).dbg_skip & (bit))
#else
#define PMC_DBG_BIT(a, bit) 0
#endif
constexpr int kTuneForceBits = 8 | 16 | 64;

// ---- launchers implemented in the .cu files (all asynchronous on `st`)
cudaError_t pmc_launch_init_r(const DevGeom &g, float *d_r, cudaStream_t st);
cudaError_t pmc_launch_assign(const DevGeom &g, const float *d_r, float4 *disk, int16_t *n,
                              Counters *ctr, cudaStream_t st);
cudaError_t pmc_launch_shift(const DevGeom &g, const float4 *src, const int16_t *nsrc,
                             float4 *dst, int16_t *ndst, int f, float d, Counters *ctr,
                             cudaStream_t st);
cudaError_t pmc_launch_subsweep(const DevGeom &g, float4 *disk, const int16_t *n,
                                const SweepArgs &a, Counters *ctr, cudaStream_t st);
cudaError_t pmc_launch_fused_sweep(const DevGeom &g, const float4 *din, const int16_t *nin,
                                   float4 *dout, int16_t *nout, const SweepArgs &a,
                                   Counters *ctr, cudaStream_t st);
cudaError_t pmc_launch_check(const DevGeom &g, const float4 *disk, const int16_t *n,
                             long long *out4, unsigned *min_d2_bits, cudaStream_t st);
cudaError_t pmc_launch_gr_hist(const DevGeom &g, const float4 *disk, const int16_t *n,
                               float r_max, int nbins, unsigned long long *hist, cudaStream_t st);
cudaError_t pmc_launch_disk_to_r(const DevGeom &g, const float4 *disk, const int16_t *n, float *d_r, long long r_stride,
                                 long long r_cap, unsigned long long **scratch, int *nblocks, cudaStream_t st);
int pmc_fused_launch_count();   // kernels per fused sweep (for bench gpu_launches)

// ---- fast fused sweep on the handle-owned internal layout (pmc_sweep4.cu)
constexpr int kMX = 6, kMY = 5; // margin columns (even) / rows holding periodic images or slab ghosts
struct Geom4 {
    int cps, row0, rows, wrap_y;
    int CH;                     // float4 chunks per (row, plane, parity) run
    int ROWS;                   // allocated rows
    int FW, FH;                 // crowded-cell flag grid: words per row, rows
    float w, hw, sigma2, dscale;    // dscale = q, the coordinate grid quantum
    float dstep, doff;              // trial proposal p = fmaf(grid_disp, dstep, x - doff) (DevGeom)
    unsigned seed_lo, seed_hi;
    int try_ns4;                // mean occupancy is low: worth scanning for tiles whose cells all hold <= 4 disks
    unsigned pk0[10], pk1[10];  // Philox round keys seed + r * (0x9E3779B9, 0xBB67AE85): constant-bank operands
};
void pmc4_plan_sweep(SweepArgs &a, int full_halo);
void pmc4_philox_prepare(SweepArgs &a, const Geom4 &g);     // after sweep_lo / sweep_hi are set
int pmc4_tile_rows(const Geom4 &g, const SweepArgs &a);
void pmc4_alloc_shape(int cps, int rows, int *CH, int *ROWS, int *FW, int *FH);
int pmc4_make_tensor_map(void *tmap_out128, const float4 *base, const Geom4 &g, int half);
cudaError_t pmc4_launch_import(const Geom4 &g, int ghost, const float4 *disk, const int16_t *n, float4 *out,
                               unsigned *flags, unsigned epoch, cudaStream_t st);
cudaError_t pmc4_launch_export(const Geom4 &g, int ghost, const float4 *in, float4 *disk, int16_t *n, cudaStream_t st);
// slab ring: add the stamps (== epoch) of the received flag rows to flag rows row_lo.. / row_hi.. of `flags`
cudaError_t pmc4_launch_flag_merge(unsigned *flags, const unsigned *recv_lo, int row_lo, const unsigned *recv_hi, int row_hi,
                                   int nrows, int FW, unsigned epoch, cudaStream_t st);
// tile rows [by0, by0 + nby) of one planned sweep (nby <= 0: all rows), optional second band
// [by1, by1 + nby1) in the same launch; fast = 1: the 3-plane / 4-CTA kernel (the crowded-cell flags under the box must
// be valid: slab ghost rows get theirs through the ring)
cudaError_t pmc4_launch_sweep(const Geom4 &g, const void *tmap_in, const void *tmap_half, float4 *dout, const SweepArgs &a,
                              Counters *ctr, cudaStream_t st, int fast, int by0 = 0, int nby = 0, int by1 = 0, int nby1 = 0);

#ifdef __CUDACC__
// ------------------------------------------------------------------ Philox4x32-10
// Salmon et al., SC'11.  Counter = {cell id, sweep lo, sweep hi, call}, key = seed.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1,
                                              uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

// the same with the ten round keys precomputed on the host (kernel-parameter constants)
__device__ __forceinline__ void philox4x32_10_keys(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                   const unsigned (&k0)[10], const unsigned (&k1)[10],
                                                   uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0[r];
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1[r];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

// philox4x32_10_keys(cell, sweep_lo, sweep_hi, 0, ...) with the work that does not depend on the cell done once per
// sweep on the host: only counter word 0 differs between the cells of a sweep, so in round 0 one of the two products
// and one of the two xors are the same for every cell, in round 1 one product, and the words they produce enter rounds
// 1 and 2 as constants folded into the round keys (e1, e2, e3: pmc4_philox_prepare).  37 instead of 40 instructions,
// the same four words bit for bit.
__device__ __forceinline__ void philox_cell(uint32_t cell, unsigned e1, unsigned e2, unsigned e3,
                                            const unsigned (&k0)[10], const unsigned (&k1)[10],
                                            uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3)
{
    unsigned long long p0 = (unsigned long long)0xD2511F53u * cell;            // round 0: c = {cell, lo, hi, 0}
    uint32_t c2 = (uint32_t)(p0 >> 32) ^ k1[0], c3 = (uint32_t)p0;
    unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;              // round 1: c0, c1 are sweep constants
    uint32_t c0 = (uint32_t)(p1 >> 32) ^ e1, c1 = (uint32_t)p1;
    c2 = c3 ^ e2;
    p0 = (unsigned long long)0xD2511F53u * c0;                                  // round 2: c3 is a sweep constant
    p1 = (unsigned long long)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0[2], n2 = (uint32_t)(p0 >> 32) ^ e3;
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
#pragma unroll
    for (int r = 3; r < 10; r++) {
        p0 = (unsigned long long)0xD2511F53u * c0;
        p1 = (unsigned long long)0xCD9E8D57u * c2;
        n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0[r];
        n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1[r];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

// Trial displacement (make_move subsweep.h:60-71, proposal uniform in the square on the coordinate grid):
// 4096 equally spaced levels m * q per axis, m = (2k - 4095) * A, k a 12-bit field of the trial's random word
// (x: bits 12-23, y: bits 0-11, shuffle: bits 24-31), A = floor(M / 4095); k -> 4095 - k maps m -> -m, so the
// proposal is exactly symmetric.  No integer -> float conversion: a 12-bit field sitting in bits 12-23 of a word
// whose other bits are zero IS the float k * 2^-137 (bits 0-22 are a denormal's mantissa, bit 23 continues it
// linearly), FMA units take denormals at full speed, and
//     p = fmaf(k * 2^-137, dstep, x - doff),   dstep = 2 A q * 2^137,   doff = 4095 A q
// is x + m q with every step exact (x - doff is on the grid and below 2^24 q; the product is exact inside the FMA).
// x: one LOP3; y: one shift + one LOP3.  Identical to oracle/pmc_oracle.c subsweep_cell.
__device__ __forceinline__ float grid_disp_x(uint32_t r) { return __uint_as_float(r & 0x00FFF000u); }
__device__ __forceinline__ float grid_disp_y(uint32_t r) { return __uint_as_float((r << 12) & 0x00FFF000u); }

// The reference's Gaussian proposal (make_move subsweep.h:60-71: x + curand_normal * sigma per axis) in grid
// units: Box-Muller on bits 8..30 of the two words, signs from bit 31 (oracle subsweep_cell, proposal == 1).
__device__ __forceinline__ void gauss_disp(uint32_t ra, uint32_t rb, int M, float &mx, float &my)
{
    const float u1 = __fmul_rn(__fadd_rn((float)((ra >> 8) & 0x7FFFFFu), 0.5f), 1.1920928955078125e-07f);   // (0, 1)
    const float u2 = __fmul_rn(__fadd_rn((float)((rb >> 8) & 0x7FFFFFu), 0.5f), 5.9604644775390625e-08f);   // (0, 1/2)
    const float rr = __fmul_rn(__fsqrt_rn(__fmul_rn(-2.0f, logf(u1))), (float)M);
    float sn, cs;
    sincospif(u2, &sn, &cs);
    mx = rintf(__fmul_rn(rr, cs));
    my = rintf(__fmul_rn(rr, sn));
    if (ra >> 31) mx = -mx;
    if (rb >> 31) my = -my;
}

// ------------------------------------------------------------------ small helpers
__device__ __forceinline__ int wrap_mod(int v, int m)
{
    v %= m;
    return v < 0 ? v + m : v;
}

// one cell in registers: x[8], y[8] as four float4 (global layout [cell][dim][slot])
struct CellRegs {
    float4 x03, x47, y03, y47;
    int cnt;
};

__device__ __forceinline__ float f4get(const float4 &lo, const float4 &hi, int i)
{
    // static index after unrolling
    switch (i) {
    case 0: return lo.x; case 1: return lo.y; case 2: return lo.z; case 3: return lo.w;
    case 4: return hi.x; case 5: return hi.y; case 6: return hi.z; default: return hi.w;
    }
}

// V2 shiftCells.h:23-112 for one destination cell in cell-local coordinates: stayers of
// `own` in slot order, then immigrants from `up` (the cell at +dir along axis F) in slot
// order.  put(slot, f_coord, other_coord) stores one particle; returns the new count
// (clamped to PMC_NMAX) and the number of particles that did not fit in *dropped.
template <int F, typename Put>
__device__ __forceinline__ int shift_one_cell(const CellRegs &own, const CellRegs &up,
                                              float d, float w, float sshift, Put put, int *dropped)
{
    int nNew = 0, drop = 0;
#pragma unroll
    for (int i = 0; i < PMC_NMAX; i++) {
        float fc = F == 0 ? f4get(own.x03, own.x47, i) : f4get(own.y03, own.y47, i);
        float oc = F == 0 ? f4get(own.y03, own.y47, i) : f4get(own.x03, own.x47, i);
        float D = __fadd_rn(fc, -d);
        if (i < own.cnt && D > 0.0f && D <= w) {           // shiftCells.h:62
            put(nNew, D, oc);                              // nNew < 8 always holds here
            nNew++;
        }
    }
#pragma unroll
    for (int i = 0; i < PMC_NMAX; i++) {
        float fc = F == 0 ? f4get(up.x03, up.x47, i) : f4get(up.y03, up.y47, i);
        float oc = F == 0 ? f4get(up.y03, up.y47, i) : f4get(up.x03, up.x47, i);
        float D = __fadd_rn(fc, -d);
        if (i < up.cnt && !(D > 0.0f && D <= w)) {         // shiftCells.h:94
            if (nNew < PMC_NMAX) {
                put(nNew, __fadd_rn(D, sshift), oc);       // shiftCells.h:97
                nNew++;
            } else drop++;
        }
    }
    *dropped = drop;
    return nNew;
}
#endif  // __CUDACC__
