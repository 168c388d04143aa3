// pmc_sweep4.cu -- the fused MC sweep (start.cu:237-260 loop body: 4 x subsweep_kernel
// subsweep.h:240-300, then shiftCells shiftCells.h:23-112) as ONE sm_100a kernel per sweep.
// This is the throughput path behind pmc_sweep(); pmc_sweep.cu keeps the generic kernel
// (any n_M, w < 2 sigma, tiny boxes, the single-colour call site).
//
// Differences from the generic kernel, chosen to cut issued instructions (ncu: the ALU pipe and
// the issue slots are the tightest resources of this path) and shared-memory wavefronts:
//   * INTERNAL STATE LAYOUT (handle-owned, never seen by the caller): one cell = 4 float4
//     chunks   P0 = x0..x3   P1 = y0..y3   P2 = x4 x5 y4 y5   P3 = x6 x7 y6 y7
//         chunk(X, Y, P) = ((Y*4 + P)*2 + (X & 1)) * CH + (X >> 1)
//     Even and odd columns are split so that the same-colour cells a warp works on are
//     contiguous (conflict-free LDS.128) and ONE 4-D TMA box per plane lands a tile in
//     shared memory in exactly the layout the sub-sweeps read: no index arithmetic at all.
//     The box is surrounded by margins (kMX columns, kMY rows) holding periodic images, written
//     by the CTA that produces the original cell (single GPU) or by the NCCL ring (slab rows),
//     so a tile never wraps.
//   * cells hold <= 6 disks in all but ~1e-5 of the cases at phi = 0.70, w = 2 sigma.  Every
//     store records, per block of 2 x 2 internal cells, whether it wrote a cell with 7 or 8
//     disks (an epoch-stamped flag word, so nothing is ever cleared).  A tile whose staged box
//     touches no flagged block takes the FAST path: plane P3 is not even staged (3 TMA boxes,
//     56 KB of shared memory, 64 registers: FOUR CTAs per SM instead of three) and the NS = 6
//     instantiation runs (3 instead of 4 chunks per neighbour cell, 24 instead of 32 pair
//     tests per trial); if all staged cells hold <= 4 (dilute systems) NS = 4 runs and P2 is
//     not touched either.  A flagged tile is processed by the same CTA in two or three
//     half-height pieces with all four planes (they fit the same 56 KB), NS chosen from a scan.
//   * the staged box is fixed (36 x 33 cells); how much of it is owned is planned per sweep
//     from the colour order (pmc4_plan_sweep): the halo an order needs is 2-4 cells per axis,
//     a shallower halo leaves room for a larger tile (24..30 x 24..28 owned cells).
//   * the cell count lives in-band: unused slots have x = sentinel; a cell with fewer than 8
//     (6) disks carries its count in the bits of y7 (y5).  No count array on the hot path.
//   * the grid shift of THIS sweep is applied while the tile leaves shared memory (the tile
//     carries one extra upstream row / column), so only owned cells are re-binned.
//   * neighbour cells needed by a trial: with w >= 2 sigma a proposal in the left half of its
//     cell can only touch the left column of neighbours, etc.: one compare per axis.
// Every random number is a pure function of (seed, sweep, global cell id, trial), so the
// redundantly recomputed halo cells get the same bits in every CTA and on every GPU.
#include "pmc_internal.cuh"
#include <cuda.h>      // CUtensorMap type only; the encoder comes from cudaGetDriverEntryPoint
#include <stdlib.h>
#include <atomic>

namespace {

constexpr float kSent = PMC_SENTINEL;
constexpr float kSentTest = 1.0e17f;      // x < kSentTest <=> slot in use
constexpr int kNT = 256;                  // threads per CTA of every kernel in this file
constexpr int kFB = 1;                    // log2 of the flag block edge (2 x 2 internal cells per flag word)

// The staged box: 36 columns (18 per parity) x SYB rows of cells per plane, the same for every
// tiling.  How much of it is owned is decided per sweep (SweepArgs::tx, ty, hx, hy): the halo a sweep
// needs depends on its colour order, and a shallower halo leaves room for a larger tile.
template <int SYB_>
struct Box4 {
    static constexpr int HB = 18;                                // staged chunks per parity row
    static constexpr int PITCH = 2 * HB;                         // chunks per staged row
    static constexpr int SYB = SYB_;                             // staged rows
    static constexpr int PLB = PITCH * SYB;                      // chunks the TMA box brings per plane
    static constexpr int PLC = (PLB + 7) / 8 * 8;                // plane stride (128-byte aligned)
    // active cells per colour: at most 16 per row (half a warp, two conflict-free quarter-warps)
    // in at most 16 rows: one thread each
    static constexpr int NAX = 16;
    static constexpr size_t PLANE_BYTES = (size_t)PLC * 16;
    static_assert(4 * HB <= 256 && SYB <= 256, "TMA box extents");
    static_assert((SYB - 1) / 2 <= kNT / NAX, "one thread per active cell of a colour");
    static_assert((PLC - PLB) * 16 >= 32, "padding behind the planes: mbarrier (plane 0), dump word (plane 1)");
};
constexpr int kSYB = 33, kSYBH = 21;      // rows of the full box; of the half-height box of a crowded tile
using BoxF = Box4<kSYB>;
using BoxH = Box4<kSYBH>;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned phase)
{
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    } while (!ok);
}
// one 4-D box {4*HB floats, 2 parities, 1 plane, SYB rows} -> dense [SYB][2][HB] float4 in smem
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}

// the same box, HBM -> L2 only: issued for the tile that a CTA slot of this wave will stage next
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap *tm, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// forced 16-byte shared-memory load (ptxas otherwise splits a float4 whose components are
// consumed one by one into four LDS.32, each a 4-way bank conflict at a 16-byte lane stride)
template <int BYTE_OFF>
__device__ __forceinline__ float4 lds128(unsigned saddr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr), "n"(BYTE_OFF));
    return v;
}

template <int BYTE_OFF>
__device__ __forceinline__ void sts128(unsigned saddr, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0+%1], {%2, %3, %4, %5};"
                 ::"r"(saddr), "n"(BYTE_OFF), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// in-band count: y7 when fewer than 8 disks (P3 = x6 x7 y6 y7), y5 when fewer than 6 (P2 = x4 x5 y4 y5)
__device__ __forceinline__ int decode_cnt8(const float4 &p3) { return p3.y < kSentTest ? 8 : __float_as_int(p3.w); }
__device__ __forceinline__ int decode_cnt6(const float4 &p2) { return p2.y < kSentTest ? 6 : __float_as_int(p2.w); }

// two slots per instruction (Blackwell packed FP32): d2 = (q.x + npx)^2 + (q.y + npy)^2, the
// oracle's fmaf(dx, dx, dy*dy) with dx = pxs - qx (sign is irrelevant after squaring)
__device__ __forceinline__ float2 pair2(float qx0, float qx1, float qy0, float qy1, float2 npx, float2 npy)
{
    const float2 dx = __fadd2_rn(make_float2(qx0, qx1), npx);
    const float2 dy = __fadd2_rn(make_float2(qy0, qy1), npy);
    return __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
}

// smallest squared distance between the trial point (negated, in this cell's frame) and the
// first NS slots of one staged cell; unused slots hold the sentinel
template <int NS, int PLC>
__device__ __forceinline__ float cell_min_d2(const char *cp, float npx, float npy, float macc)
{
    const float4 p0 = *reinterpret_cast<const float4 *>(cp), p1 = *reinterpret_cast<const float4 *>(cp + PLC * 16);
    // sign bit of y3 = "this cell holds 5 or more disks" (cell-local coordinates are positive):
    // only then are P2 (and P3) fetched, so quarter-warps whose 8 neighbour cells all hold <= 4
    // disks (61 % of them at phi = 0.70) spend no shared-memory wavefront on those planes
    const bool more = NS >= 6 && __float_as_int(p1.w) < 0;
    const float2 nx = make_float2(npx, npx), ny = make_float2(npy, npy);
    const float2 a = pair2(p0.x, p0.y, p1.x, p1.y, nx, ny);
    const float2 b = pair2(p0.z, p0.w, p1.z, fabsf(p1.w), nx, ny);
    // running minimum `macc` threaded through, every step a 3-input min (FMNMX3): two new values per instruction
    float m = fminf(fminf(a.x, a.y), b.x);
    m = fminf(fminf(m, b.y), macc);
    if (NS >= 6) {
        if (more) {
            const float4 p2 = *reinterpret_cast<const float4 *>(cp + 2 * PLC * 16);
            const float2 c = pair2(p2.x, p2.y, p2.z, p2.w, nx, ny);
            m = fminf(fminf(m, c.x), c.y);
            if (NS == 8) {
                const float4 p3 = *reinterpret_cast<const float4 *>(cp + 3 * PLC * 16);
                const float2 e = pair2(p3.x, p3.y, p3.z, p3.w, nx, ny);
                m = fminf(fminf(m, e.x), e.y);
            }
        }
    }
    return m;
}

// V2 shiftCells.h:23-112.  The new content of a destination cell - the stayers of the cell itself in slot
// order, then the immigrants from the cell upstream in slot order - is scattered into the cell's own staged
// chunks in PAIR order: chunk c (plane c) = (x_2c, y_2c, x_2c+1, y_2c+1), so that one 8-byte store places a
// disk (st.shared.v2.f32) and the slot -> address walk is two adds (+8, then +plane stride - 8, alternating); the way out to HBM
// converts to P0..P3 in registers.
// One source cell, one of the two roles.  STAY: its disks that remain (shiftCells.h:62: 0 < D <= w with D = c - d
// along axis F) go to the walk position (sa, step) of their own cell.  !STAY: its disks that leave (shiftCells.h:94)
// go, re-based by sshift (shiftCells.h:97), to the walk position of the cell downstream; that walk may run past
// the last staged plane (7 or more disks want to be in the cell: rare): nothing is stored at or beyond `lim`, but
// the walk goes on, so that the final position counts every disk.  A slot is in use iff its x is below `thr`
// (kSentTest; a lane is switched off with thr <= 0: cell-local coordinates are positive); a disk stays iff
// 0 < D <= wl (a lane is switched off with wl <= 0).
// One candidate = one straight-line block of predicated instructions (written in PTX: ptxas otherwise turns
// every candidate into a branch with its own convergence barrier).
template <int NS, int F, bool STAY>
__device__ __forceinline__ void walk_scatter(const CellRegs &c, float d, float wl, float thr, float sshift,
                                             unsigned &sa, unsigned &step, unsigned PS, unsigned lim)
{
#pragma unroll
    for (int i = 0; i < NS; i++) {
        const float xi = f4get(c.x03, c.x47, i), yi = f4get(c.y03, c.y47, i);
        const float fc = F == 0 ? xi : yi;
        if (STAY) {
            // D = fc - d (one rounding, == __fadd_rn(fc, -d)); p = 0 < D <= wl [and slot in use]; store (D, y) / (x, D)
            if (F == 0)
                asm volatile("{\n .reg .pred p;\n .reg .f32 D;\n"
                             " sub.rn.f32 D, %2, %4;\n setp.gt.f32 p, D, 0f00000000;\n setp.le.and.f32 p, D, %5, p;\n"
                             " @p st.shared.v2.f32 [%0], {D, %3};\n @p add.u32 %0, %0, %1;\n @p sub.u32 %1, %6, %1;\n}"
                             : "+r"(sa), "+r"(step) : "f"(xi), "f"(yi), "f"(d), "f"(wl), "r"(PS) : "memory");
            else
                asm volatile("{\n .reg .pred p;\n .reg .f32 D;\n"
                             " sub.rn.f32 D, %2, %4;\n setp.gt.f32 p, D, 0f00000000;\n setp.le.and.f32 p, D, %5, p;\n"
                             " setp.lt.and.f32 p, %3, %7, p;\n"
                             " @p st.shared.v2.f32 [%0], {%3, D};\n @p add.u32 %0, %0, %1;\n @p sub.u32 %1, %6, %1;\n}"
                             : "+r"(sa), "+r"(step) : "f"(yi), "f"(xi), "f"(d), "f"(wl), "r"(PS), "f"(thr) : "memory");
        } else {
            // p = slot in use and not (0 < D <= wl); Ds = D + sshift; stored only below lim, walked in any case
            if (F == 0)
                asm volatile("{\n .reg .pred p, q;\n .reg .f32 D;\n"
                             " sub.rn.f32 D, %2, %4;\n setp.gt.f32 q, D, 0f00000000;\n setp.le.and.f32 q, D, %5, q;\n"
                             " setp.lt.and.f32 p, %2, %7, !q;\n add.rn.f32 D, D, %8;\n setp.lt.and.u32 q, %0, %9, p;\n"
                             " @q st.shared.v2.f32 [%0], {D, %3};\n @p add.u32 %0, %0, %1;\n @p sub.u32 %1, %6, %1;\n}"
                             : "+r"(sa), "+r"(step) : "f"(xi), "f"(yi), "f"(d), "f"(wl), "r"(PS), "f"(thr), "f"(sshift), "r"(lim)
                             : "memory");
            else
                asm volatile("{\n .reg .pred p, q;\n .reg .f32 D;\n"
                             " sub.rn.f32 D, %2, %4;\n setp.gt.f32 q, D, 0f00000000;\n setp.le.and.f32 q, D, %5, q;\n"
                             " setp.lt.and.f32 p, %3, %7, !q;\n add.rn.f32 D, D, %8;\n setp.lt.and.u32 q, %0, %9, p;\n"
                             " @q st.shared.v2.f32 [%0], {%3, D};\n @p add.u32 %0, %0, %1;\n @p sub.u32 %1, %6, %1;\n}"
                             : "+r"(sa), "+r"(step) : "f"(yi), "f"(xi), "f"(d), "f"(wl), "r"(PS), "f"(thr), "f"(sshift), "r"(lim)
                             : "memory");
        }
        (void)fc;
    }
}
// walk offset (bytes from the cell's first chunk) -> number of disks walked over
__device__ __forceinline__ int walk_count(unsigned off, unsigned PS) { return 2 * (int)(off / PS) + ((off & 8u) ? 1 : 0); }

// a cell with 7 or 8 disks was produced (rare): stamp the flag word of its 2 x 2 block - and of
// the blocks of its periodic images - so that the next sweep stages P3 for the tiles that see it;
// on the fast path (P3 not in shared memory) also write its P3 chunk, images included
__device__ __noinline__ void crowded_cell_out(float4 *dout, unsigned *flag_out, unsigned epoch, int cps, int rows,
                                              int wrap_y, int CH, int FW, int X, int Y, int write_p3, float4 p3)
{
    const int ux = X - kMX, uy = Y - kMY;
    const int ximg = ux < kMX ? cps : (ux >= cps - kMX ? -cps : 0);
    const int yimg = wrap_y ? (uy < kMY ? rows : (uy >= rows - kMY ? -rows : 0)) : 0;
    for (int q = 0; q < 4; q++) {
        if (((q & 1) && !ximg) || ((q & 2) && !yimg)) continue;
        const int XX = X + ((q & 1) ? ximg : 0), YY = Y + ((q & 2) ? yimg : 0);
        flag_out[(YY >> kFB) * FW + (XX >> kFB)] = epoch;
        if (write_p3) dout[((long long)(YY * 4 + 3) * 2 + (XX & 1)) * CH + (XX >> 1)] = p3;
    }
}

// fast path, rare: a cell ends up with 7 or 8 disks (or more: dropped), but slots 6 and 7 exist only
// in HBM.  Finds the immigrants that did not fit (the `first` earlier ones are placed), writes the
// cell's P3 chunk to HBM (the caller then skips its own P3 store) and stamps the crowded-cell flags.
// Deliberately not inlined: nothing of it may cost the common path a register.
__device__ __noinline__ int shift_overflow3(int F, float4 ux03, float2 ux45, float4 uy03, float2 uy45, int ucnt,
                                            float d, float w, float sshift, int first, int n_total,
                                            int owned, float4 *dout, const Geom4 *gp, const SweepArgs *ap,
                                            int X, int Y)
{
    const float ux[6] = { ux03.x, ux03.y, ux03.z, ux03.w, ux45.x, ux45.y };
    const float uy[6] = { uy03.x, uy03.y, uy03.z, uy03.w, uy45.x, uy45.y };
    float4 p3 = make_float4(kSent, kSent, 0.f, 0.f);
    int k = 0;
    for (int i = 0; i < 6; i++) {
        const float fc = F == 0 ? ux[i] : uy[i], oc = F == 0 ? uy[i] : ux[i];
        const float D = __fadd_rn(fc, -d);
        if (i < ucnt && !(D > 0.0f && D <= w)) {
            const float Ds = __fadd_rn(D, sshift);
            const float vx = F == 0 ? Ds : oc, vy = F == 0 ? oc : Ds;
            if (k == first) { p3.x = vx; p3.z = vy; }
            if (k == first + 1) { p3.y = vx; p3.w = vy; }
            k++;
        }
    }
    const int n = n_total < PMC_NMAX ? n_total : PMC_NMAX;
    if (n < PMC_NMAX) p3.w = __int_as_float(n);
    if (owned)
        crowded_cell_out(dout, ap->flag_out, ap->epoch_out, gp->cps, gp->rows, gp->wrap_y, gp->CH, gp->FW, X, Y, 1, p3);
    return n_total - n;                                 // dropped
}

// everything one CTA knows about its tile
struct TileCtx {
    int RX, RY;             // region the sub-sweeps work on (owned + halo + the extra upstream strip)
    int rx0, ry0;           // unwrapped global column / owned-relative row of region (0, 0)
    int xs;                 // region column i is staged column i + xs
    int ox0, oy0;           // region coordinates of the owned tile's corner
    int nox, noy;           // owned extent after clipping at the box / slab edge
    int X0, Y0;             // internal array coordinates of region (0, 0)
    int tx, ty;             // nominal owned extent of this tiling
    int wraps;              // the region reaches beyond the box: global cell coordinates need the periodic wrap
    int cwsel;              // which set of SweepArgs::colour_word: (row0 & 1) * 8 + (col0 & 1) * 4
    int edge;               // owned cells of this tile have periodic images in the margins
};

// ---- one colour: one thread per active cell (subsweep.h:242-245), own cell in registers
template <int NS, typename TL>
__device__ __forceinline__ void colour_pass(float4 *sm, const TileCtx &t, const Geom4 &g, const SweepArgs &a,
                                            int k, int aq, int bq, unsigned &my_cnt)
{
    constexpr int PITCH = TL::PITCH, PLC = TL::PLC;
    const int cps = g.cps;
    const float w = g.w, hw = g.hw, sigma2 = g.sigma2, dstep = g.dstep, doff = g.doff;
    // which cells this colour works on (SweepArgs::colour_word, planned on the host per parity of the tile origin).
    // Cells closer than lo to the region edge are stale or irrelevant; (i0, j0) = first active cell.
    const unsigned cw = a.colour_word[t.cwsel + k];
    const int i0 = (int)(cw & 15u), j0 = (int)((cw >> 4) & 15u), lox = (int)((cw >> 8) & 15u), loy = (int)((cw >> 12) & 15u);
    const int i = i0 + 2 * aq, j = j0 + 2 * bq;
    if (!(i < t.RX - lox && j < t.RY - loy)) return;
    // staged chunk of a cell = a tile-independent part (in cw) + the lane part
    char *const cbase = reinterpret_cast<char *>(sm) + (2 * bq * PITCH + aq) * 16;
    char *const cown = cbase + ((cw >> 16) & 255u) * 16;
    const char *const cL = cbase + (cw >> 24) * 16;             // left neighbour; right = cL + 16
    const unsigned sown = smem_u32(cown);
    const float4 p0 = lds128<0>(sown), p1 = lds128<PLC * 16>(sown), p2 = lds128<2 * PLC * 16>(sown);
    float4 p3 = make_float4(kSent, kSent, 0.f, 0.f);
    if (NS == 8) p3 = lds128<3 * PLC * 16>(sown);
    const int cnt = NS == 8 ? decode_cnt8(p3) : decode_cnt6(p2);
    if (cnt == 0) return;                       // subsweep.h:252-254
    const bool owned = (unsigned)(i - t.ox0) < (unsigned)t.nox && (unsigned)(j - t.oy0) < (unsigned)t.noy;
    int gx = t.rx0 + i, gy = g.row0 + t.ry0 + j;
    if (t.wraps) {                              // CTA-uniform: only tiles at the edge of the box hold periodic images
        gx += gx < 0 ? cps : 0; gx -= gx >= cps ? cps : 0;
        gy += gy < 0 ? cps : 0; gy -= gy >= cps ? cps : 0;
    }
    const uint32_t cell_id = (uint32_t)gy * (uint32_t)cps + (uint32_t)gx;

    // neighbour part of one trial: smallest d2 against the 3 neighbour cells that can hold a
    // disk closer than sigma (w >= 2 sigma); inb = the proposal stays inside the cell
    // (out_of_bound subsweep.h:73-88)
    auto neighbours_min_d2 = [&](const float px, const float py, bool &inb) -> float {
        inb = px > 0.0f && px <= w && py > 0.0f && py <= w;
        const bool goL = px <= hw, goD = py <= hw;
        const float npxs = -__fadd_rn(px, goL ? w : -w);     // -(px - helper*w), subsweep.h:139-151
        const float npys = -__fadd_rn(py, goD ? w : -w);
        const char *cH = cL + (goL ? 0 : 16);
        const int dV = goD ? -PITCH * 16 : PITCH * 16;
        float m = cell_min_d2<NS, PLC>(cH, npxs, -py, 3.0e38f);
        m = cell_min_d2<NS, PLC>(cown + dV, -px, npys, m);
        return cell_min_d2<NS, PLC>(cH + dV, npxs, npys, m);
    };

    float ox[8] = { p0.x, p0.y, p0.z, p0.w, p2.x, p2.y, p3.x, p3.y };
    float oy[8] = { p1.x, p1.y, p1.z, fabsf(p1.w), p2.z, p2.w, p3.z, p3.w };     // y3 carries the "5 or more" flag in its sign
    // ONE Philox call per cell and sub-sweep: word s feeds trial s (shuffle: bits 24-31, dx: bits 12-23, dy: bits 0-11)
    uint32_t rw[4];
    philox_cell(cell_id, a.ph_e1, a.ph_e2, a.ph_e3, g.pk0, g.pk1, rw[0], rw[1], rw[2], rw[3]);
    // random_shuffle subsweep.h:50-58: physical partial Fisher-Yates, steps 0..3
#pragma unroll
    for (int s = 0; s < 4; s++) {
        // jd = jj - s = ((r >> 24) * (cnt - s)) >> 8 as the high half of one product.  s >= cnt (no slot left): cnt - s
        // wraps to 2^32 - k, the product's high half is 0 (top byte 0) or >= 2^24 - 1: never one of the 1 .. 5 tested below
        const int jd = (int)__umulhi(rw[s] & 0xFF000000u, (uint32_t)(cnt - s));
        const float tx = ox[s], ty = oy[s];
        float nx = tx, ny = ty;
#pragma unroll
        for (int q = s + 1; q < NS; q++) {
            const bool p = (jd == q - s);
            nx = p ? ox[q] : nx; ny = p ? oy[q] : ny;
            ox[q] = p ? tx : ox[q]; oy[q] = p ? ty : oy[q];
        }
        ox[s] = nx; oy[s] = ny;
    }
    // trials 0..3 move slot s mod cnt (subsweep.h:279-297), all register indices static
    unsigned n_acc = 0;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const bool cA = cnt > s;                            // slot == s
        const bool cB = (s == 3) && (cnt == 2);             // slot == 1 (only s = 3, cnt = 2)
        const float x = cA ? ox[s] : (cB ? ox[1] : ox[0]);
        const float y = cA ? oy[s] : (cB ? oy[1] : oy[0]);
        const float px = __fmaf_rn(grid_disp_x(rw[s]), dstep, __fadd_rn(x, -doff));      // make_move subsweep.h:60-71
        const float py = __fmaf_rn(grid_disp_y(rw[s]), dstep, __fadd_rn(y, -doff));
        bool inb;
        float m = neighbours_min_d2(px, py, inb);
        // own cell (calculate_energy_in_cell subsweep.h:105-117), j != slot
        const float2 npx = make_float2(-px, -px), npy = make_float2(-py, -py);
        float2 d01 = pair2(ox[0], ox[1], oy[0], oy[1], npx, npy);
        float2 d23 = pair2(ox[2], ox[3], oy[2], oy[3], npx, npy);
        // j != slot: of slots 0..3 the moved one is left out (one select per pair it can be in, not one per slot)
        if (s == 0) m = fminf(fminf(m, d01.y), fminf(d23.x, d23.y));
        if (s == 1) m = fminf(fminf(m, cA ? d01.x : d01.y), fminf(d23.x, d23.y));                       // slot 1, or 0 (cnt == 1)
        if (s == 2) m = fminf(fminf(m, cA ? d01.x : d23.x), fminf(d01.y, d23.y));                       // slot 2, or 0 (cnt <= 2)
        if (s == 3) {                                                                                   // slot 3, 1 (cnt == 2) or 0 (cnt == 1, 3)
            const float u = (cA | cB) ? d01.x : d01.y, v = cA ? d01.y : d23.y;
            m = fminf(fminf(m, u), fminf(v, d23.x));
        }
        if (NS >= 6) {
            const float2 d45 = pair2(ox[4], ox[5], oy[4], oy[5], npx, npy);
            m = fminf(fminf(m, d45.x), d45.y);
        }
        if (NS == 8) {
            const float2 d67 = pair2(ox[6], ox[7], oy[6], oy[7], npx, npy);
            m = fminf(fminf(m, d67.x), d67.y);
        }
        // accept_move subsweep.h:194-217 (hard disks: accept iff in bounds and no overlap)
        const bool acc = inb && !(m < sigma2);      // out of the cell: rejected whatever the neighbours say
        if (acc) n_acc += 1u;
        if (s == 0) { ox[0] = acc ? px : ox[0]; oy[0] = acc ? py : oy[0]; }
        else {
            const bool w0 = acc & !cA & !cB, w1 = acc & cB, ws = acc & cA;
            ox[s] = ws ? px : ox[s]; oy[s] = ws ? py : oy[s];
            ox[0] = w0 ? px : ox[0]; oy[0] = w0 ? py : oy[0];
            if (s == 3) { ox[1] = w1 ? px : ox[1]; oy[1] = w1 ? py : oy[1]; }
        }
    }
    if (owned) my_cnt += (4u << 16) + n_acc;     // trials in the high half, accepted in the low half (<= 64 per thread)
    // cpy_D_sh_to_Disk subsweep.h:29-36 (shuffled order is written back, like the reference)
    float4 *pown = reinterpret_cast<float4 *>(cown);
    pown[0] = make_float4(ox[0], ox[1], ox[2], ox[3]);
    pown[PLC] = make_float4(oy[0], oy[1], oy[2], cnt >= 5 ? -oy[3] : oy[3]);
    if (NS >= 6) pown[2 * PLC] = make_float4(ox[4], ox[5], oy[4], oy[5]);
    if (NS == 8) pown[3 * PLC] = make_float4(ox[6], ox[7], oy[6], oy[7]);
}

// ---- shiftCells(f, d) of this sweep for the owned cells and their way to HBM, in ONE pass (v11/v12; v4-v10
// had a shift pass that left the tile in shared memory and a separate store pass).  One thread per destination
// cell: the new content is scattered into the cell's own staged chunks in pair order (walk_scatter), the same
// thread reads the chunks back, turns them into P0..P2 in registers (+ the in-band count, the "5 or more" flag,
// the constant P3 of the fast path; slots beyond the new count are blanked in registers, nothing is pre-cleared)
// and stores them.  No second pass over the tile, no count / flag round trip through shared memory.
// Lanes run along x, permuted so that every quarter-warp touches 8 consecutive chunks of one column parity:
// conflict-free LDS.128 / STS and 128-byte store segments.
//   F = 0 (shift along x), PUSH: warp = K rows, lane = one cell of the row (tx owned + the extra upstream column).
//     A lane keeps ONE cell in registers: it scatters its stayers into its own cell, gets the walk position of the
//     cell downstream from that cell's lane by shuffle and appends its leavers there.  All hazards are inside the
//     warp: __syncwarp() after the loads and before the read-back, no CTA barrier.
//   F = 1 (shift along y), PULL: thread = (segment of K rows, column); it walks its segment from the upstream end
//     and keeps the raw cell u+1 in registers as `up` of cell u.  The raw cell beyond the segment belongs to
//     another warp: it is read before the one CTA barrier of this pass.
template <int NS, typename TL, int NPL, int F>
__device__ __forceinline__ void shift_store_pass(float4 *sm, const TileCtx &t, const Geom4 &g, const SweepArgs &a,
                                                 float4 *__restrict__ dout, int sdir, float d, int tid, Counters *ctr)
{
    constexpr int HB = TL::HB, PITCH = TL::PITCH, PLC = TL::PLC;
    constexpr unsigned PS = PLC * 16;
    constexpr int NSL = 2 * NPL;                                    // slots that exist in shared memory
    const int TX = t.tx, TY = t.ty;
    const float w = g.w;
    const float sshift = sdir > 0 ? w : -w;                         // shiftCells.h:84-86
    const unsigned ps = 2u * (unsigned)g.CH;                        // float4 chunks between planes (rows: 4 * ps)

    auto cell_addr = [&](int i, int j) -> unsigned {
        const int is = i + t.xs;
        return smem_u32(sm + j * PITCH + (is & 1) * HB + (is >> 1));
    };
    // chunk index of plane 0 of a region cell in the internal array (fits 32 bits: < 2^31 chunks at N = 2^28)
    auto chunk_index = [&](int i, int j) -> unsigned {
        const unsigned X = (unsigned)(t.X0 + i), Y = (unsigned)(t.Y0 + j);
        return (Y * 8u + (X & 1u)) * (unsigned)g.CH + (X >> 1);
    };
    auto load_cell = [&](unsigned sp, CellRegs &c) {                // P0..P3 -> plain x / y registers
        const float4 p0 = lds128<0>(sp), p1 = lds128<PS>(sp);
        c.x03 = p0; c.y03 = make_float4(p1.x, p1.y, p1.z, fabsf(p1.w));
        c.x47 = make_float4(kSent, kSent, kSent, kSent);
        c.y47 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (NS >= 6) {                                              // higher slots: known to be unused in this tile
            const float4 p2 = lds128<2 * PS>(sp);
            c.x47.x = p2.x; c.x47.y = p2.y; c.y47.x = p2.z; c.y47.y = p2.w;
        }
        if (NS == 8) {
            const float4 p3 = lds128<(NPL == 4 ? 3 : 2) * PS>(sp);
            c.x47.z = p3.x; c.x47.w = p3.y; c.y47.z = p3.z; c.y47.w = p3.w;
        }
    };
    auto clear_cell = [&](unsigned sp) {                            // two unused slots in pair order per chunk
        const float4 empty2 = make_float4(kSent, 0.f, kSent, 0.f);
        sts128<0>(sp, empty2); sts128<PS>(sp, empty2); sts128<2 * PS>(sp, empty2);
        if (NPL == 4) sts128<3 * PS>(sp, empty2);
    };
    // the scattered cell -> HBM.  n_total = disks that want to be in the cell (more than NSL: rare path);
    // gidx = chunk_index of the cell, (i, j) only for the rare paths
    auto store_cell = [&](unsigned sp, unsigned gidx, bool owned, int i, int j, int n_total) {
        int nNew = n_total;
        bool p3_done = false;
        if (n_total > NSL) {
            p3_done = NPL == 3;                         // the pusher / puller wrote P3 of this cell to HBM
            nNew = min(n_total, PMC_NMAX);
            if (n_total > PMC_NMAX) {
                atomicOr(&ctr->status, PMC_STATUS_OVERFLOW);
                if (owned) atomicAdd(&ctr->lost, (unsigned long long)(n_total - PMC_NMAX));
            }
        }
        if (NPL == 4 && nNew >= 7 && owned)
            crowded_cell_out(dout, a.flag_out, a.epoch_out, g.cps, g.rows, g.wrap_y, g.CH, g.FW, t.X0 + i, t.Y0 + j, 0,
                             make_float4(0.f, 0.f, 0.f, 0.f));
        // pair order -> P0..P3, the in-band counts, the "5 or more" flag (sign of y3)
        const float4 c0 = lds128<0>(sp), c1 = lds128<PS>(sp), c2 = lds128<2 * PS>(sp);
        const float4 q0 = make_float4(c0.x, c0.z, c1.x, c1.z);
        const float4 q1 = make_float4(c0.y, c0.w, c1.y, nNew >= 5 ? -c1.w : c1.w);
        const float4 q2 = make_float4(c2.x, c2.z, c2.y, nNew < 6 ? __int_as_float(nNew) : c2.w);
        float4 q3 = make_float4(kSent, kSent, 0.f, __int_as_float(nNew));
        if (NPL == 4) {
            const float4 c3 = lds128<3 * PS>(sp);
            q3 = make_float4(c3.x, c3.z, c3.y, nNew < PMC_NMAX ? __int_as_float(nNew) : c3.w);
        }
        if (!owned) return;
        dout[gidx] = q0; dout[gidx + ps] = q1; dout[gidx + 2 * ps] = q2;
        if (!p3_done) dout[gidx + 3 * ps] = q3;
        if (t.edge) {                                   // CTA-uniform: periodic images into the margins
            const int ux = t.X0 + i - kMX, uy = t.Y0 + j - kMY, cps = g.cps;
            const int ximg = ux < kMX ? cps / 2 : (ux >= cps - kMX ? -(cps / 2) : 0);   // cps is even: parity is kept
            const int yimg = g.wrap_y ? (uy < kMY ? g.rows : (uy >= g.rows - kMY ? -g.rows : 0)) : 0;
#pragma unroll 1
            for (int q = 1; q < 4; q++) {
                if (((q & 1) && !ximg) || ((q & 2) && !yimg)) continue;
                float4 *di = dout + gidx + ((q & 1) ? ximg : 0) + ((q & 2) ? (long long)yimg * 4 * ps : 0);
                di[0] = q0; di[ps] = q1; di[2 * ps] = q2;
                if (!p3_done) di[3 * ps] = q3;
            }
        }
    };
    // fast path, rare: more disks than staged slots.  The immigrants that did not fit go straight to the P3 chunk in HBM.
    auto overflow3 = [&](const CellRegs &up, int n_stay, int n_total, int i, int j) {
        int ucnt = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) ucnt += f4get(up.x03, up.x47, k) < kSentTest ? 1 : 0;
        const bool owned = (unsigned)(i - t.ox0) < (unsigned)t.nox && (unsigned)(j - t.oy0) < (unsigned)t.noy;
        shift_overflow3(F, up.x03, make_float2(up.x47.x, up.x47.y), up.y03, make_float2(up.y47.x, up.y47.y), ucnt,
                        d, w, sshift, NSL - n_stay, n_total, owned, dout, &g, &a, t.X0 + i, t.Y0 + j);
    };

    if (F == 1) {
        // column strips: nseg = kNT / TX segments of K rows; (seg, c) by an exact multiply-shift (tid < 256);
        // lanes take the even columns first, then the odd ones
        const int seg = (tid * a.sh_inv) >> 16, c = tid - seg * TX, hx2 = TX >> 1;
        const int col = c < hx2 ? 2 * c : 2 * (c - hx2) + 1;
        const int K = ((TY + a.sh_nseg - 1) * a.sh_inv2) >> 16, k0 = seg * K;      // ceil(TY / nseg)
        const int len = seg < a.sh_nseg ? min(K, TY - k0) : 0;
        const int i = t.ox0 + col;
        const int jb = t.oy0 + (sdir > 0 ? k0 : TY - 1 - k0);       // cell u of the strip = row jb + u * sdir
        const int rs = sdir * PITCH * 16;                           // bytes from cell u to cell u + 1
        const bool xown = (unsigned)col < (unsigned)t.nox;
        unsigned sp = cell_addr(i, jb) + (len - 1) * rs;            // cell len - 1 (upstream end)
        unsigned gidx = chunk_index(i, jb + (len - 1) * sdir);
        const unsigned grs = (unsigned)(sdir * 4) * ps;             // chunks from cell u to cell u + 1
        CellRegs A, B;
        if (len > 0) {
            load_cell(sp, A);
            load_cell(sp + rs, B);
        }
        __syncthreads();
        auto emit = [&](int u, const CellRegs &own, const CellRegs &up) {
            const int j = jb + u * sdir;
            clear_cell(sp);
            unsigned sa = sp, step = 8;
            walk_scatter<NS, 1, true>(own, d, w, kSentTest, sshift, sa, step, PS, 0u);
            const int n_stay = walk_count(sa - sp, PS);
            walk_scatter<NS, 1, false>(up, d, w, kSentTest, sshift, sa, step, PS, sp + NPL * PS);
            const int n_total = walk_count(sa - sp, PS);
            if (NPL == 3 && n_total > NSL) overflow3(up, n_stay, n_total, i, j);
            store_cell(sp, gidx, xown && (unsigned)(j - t.oy0) < (unsigned)t.noy, i, j, n_total);
        };
        int u = len - 1;
#pragma unroll 1
        for (int it = 0; it < 2; it++) {                // K <= 4: two ping-pong steps per iteration, no register copies
            if (u >= 0) {
                emit(u, A, B);
                sp -= rs; gidx -= grs;
                if (u > 0) load_cell(sp, B);
            }
            u--;
            if (u >= 0) {
                emit(u, B, A);
                sp -= rs; gidx -= grs;
                if (u > 0) load_cell(sp, A);
            }
            u--;
        }
    } else {
        // cell u of a row: u = 0 the most downstream owned column ... TX - 1, u = TX the extra upstream column.
        // lane -> u: quarter-warps hold 8 cells of one parity (0 2 .. 14 | 1 3 .. 15 | 16 .. 30 | 17 .. 31)
        const int lane = tid & 31, seg = tid >> 5;
        const int u = ((lane >> 4) << 4) + ((lane & 7) << 1) + ((lane >> 3) & 1);
        auto lane_of = [](int uu) { return ((uu >> 4) << 4) + ((uu & 1) << 3) + ((uu & 15) >> 1); };
        const int lane_dn = lane_of(max(u - 1, 0)), lane_up = lane_of(min(u + 1, 31));
        const bool act = u <= TX, dest = u < TX, push = act && u >= 1;
        const int K = (TY + kNT / 32 - 1) / (kNT / 32), k0 = seg * K;
        const int len = min(K, TY - k0);                // warp-uniform
        const int i = sdir > 0 ? t.ox0 + u : t.ox0 + TX - 1 - u;
        const bool xown = dest && (unsigned)(i - t.ox0) < (unsigned)t.nox;
        const float w_stay = dest ? w : -1.0f, thr_push = push ? kSentTest : -1.0f;    // lane switches of walk_scatter
        unsigned sp = cell_addr(i, t.oy0 + k0), sp_dn = cell_addr(i - sdir, t.oy0 + k0);
        unsigned gidx = chunk_index(i, t.oy0 + k0);
#pragma unroll 1
        for (int v = 0; v < len; v++, sp += PITCH * 16, sp_dn += PITCH * 16, gidx += 4 * ps) {
            const int j = t.oy0 + k0 + v;
            CellRegs c;
            if (act) load_cell(sp, c);
            __syncwarp();
            if (dest) clear_cell(sp);
            unsigned sa = sp, step = 8;
            walk_scatter<NS, 0, true>(c, d, w_stay, kSentTest, sshift, sa, step, PS, 0u);
            __syncwarp();                               // the cleared chunks before the neighbour lane's leavers
            const unsigned off_dn = __shfl_sync(0xffffffffu, sa - sp, lane_dn);     // where the cell downstream stands
            unsigned sb = sp_dn + off_dn, stepb = (off_dn & 8u) ? PS - 8 : 8;
            walk_scatter<NS, 0, false>(c, d, w, thr_push, sshift, sb, stepb, PS, sp_dn + NPL * PS);
            const int n_dn = walk_count(sb - sp_dn, PS);
            if (NPL == 3 && push && n_dn > NSL) overflow3(c, walk_count(off_dn, PS), n_dn, i - sdir, j);
            const int n_total = __shfl_sync(0xffffffffu, n_dn, lane_up);
            __syncwarp();
            if (dest) store_cell(sp, gidx, xown && (unsigned)(j - t.oy0) < (unsigned)t.noy, i, j, n_total);
        }
    }
}

// ---- one tile of one sweep: stage, 4 colours, shiftCells, store.  (col0, row0) = first owned column /
// owned-relative row, tx x ty = nominal owned extent (a.tx x a.ty, or half the rows for a crowded tile);
// `phase` = parity of the mbarrier phase this staging completes.
// NPL = 3 is the fast path: it first looks up the crowded-cell flags of the blocks its box
// touches and returns true, having done nothing, when one is set.
template <typename TL, int NPL>
__device__ __forceinline__ bool process_tile(const CUtensorMap *tmap_p, float4 *__restrict__ dout, const Geom4 &g,
                                             const SweepArgs &a, Counters *ctr, int col0, int row0, int tx, int ty,
                                             float4 *sm, uint64_t *mbar, unsigned phase, bool prefetch,
                                             unsigned &my_cnt)
{
    constexpr int PLC = TL::PLC, NAX = TL::NAX;
    const int tid = threadIdx.x;
    const int cps = g.cps;
    const int HX = a.hx, HY = a.hy;                 // halo this sweep's colour order needs, per axis (2..4)

    // this sweep's grid shift: the tile carries one extra row / column on the upstream side
    const bool do_shift = a.shift_on && !PMC_DBG_BIT(a, 2);
    const int sdir = (a.shift_d <= 0.0f) ? -1 : 1;                       // shiftCells.h:38-44
    const int exl = (do_shift && a.shift_f == 0 && sdir < 0), exh = (do_shift && a.shift_f == 0 && sdir > 0);
    const int eyl = (do_shift && a.shift_f == 1 && sdir < 0), eyh = (do_shift && a.shift_f == 1 && sdir > 0);
    TileCtx t;
    t.tx = tx; t.ty = ty;
    t.RX = tx + 2 * HX + exl + exh; t.RY = ty + 2 * HY + eyl + eyh;
    t.rx0 = col0 - HX - exl;
    t.ry0 = row0 - HY - eyl;
    t.X0 = t.rx0 + kMX; t.Y0 = t.ry0 + kMY;         // region (0, 0) in internal array coordinates (>= 0)
    t.xs = t.X0 & 1;
    t.cwsel = ((row0 & 1) << 3) | ((col0 & 1) << 2);
    t.ox0 = HX + exl; t.oy0 = HY + eyl;
    t.nox = min(tx, cps - col0); t.noy = min(ty, g.rows - row0);
    const int Xb0 = t.X0 - t.xs;                    // first column of the staged box
    t.wraps = t.rx0 < 0 || t.rx0 + t.RX > cps || g.row0 + t.ry0 < 0 || g.row0 + t.ry0 + t.RY > cps;
    t.edge = col0 < kMX || col0 + tx > cps - kMX || (g.wrap_y && (row0 < kMY || row0 + ty > g.rows - kMY));

    // ------------------------------------------------------------ stage the tile: one TMA box per plane
    if (tid == 0) {
        mbar_expect_tx(mbar, (unsigned)(NPL * TL::PLB * 16));
#pragma unroll
        for (int p = 0; p < NPL; p++) tma_load_4d(sm + p * PLC, tmap_p, 4 * (Xb0 >> 1), 0, p, t.Y0, mbar);
        if (prefetch && a.prefetch_ahead > 0) {
            // warm L2 for the tile a CTA slot freed by this wave will stage (blocks are issued in order)
            const int nb = blockIdx.y * gridDim.x + blockIdx.x + a.prefetch_ahead;
            if (nb < (int)(gridDim.x * gridDim.y)) {
                const int gr2 = nb / gridDim.x, bx2 = nb - gr2 * gridDim.x;
                const int by2 = gr2 < a.by_n1 ? gr2 + a.by_off : gr2 - a.by_n1 + a.by_off2;
                const int X2 = bx2 * a.tx - HX - exl + kMX, Y2 = by2 * a.ty - HY - eyl + kMY;
#pragma unroll
                for (int p = 0; p < NPL; p++) tma_prefetch_4d(tmap_p, 4 * ((X2 - (X2 & 1)) >> 1), 0, p, Y2);
            }
        }
    }
    int crowded = 0;
    if (NPL == 3) {
        // flag words under the region (every cell of it is read with the 6-slot assumption): at most
        // 19 per row, a lane each, the warps stride over the rows; all while the TMA is in flight
        const int bx0 = t.X0 >> kFB, nbx = ((t.X0 + t.RX - 1) >> kFB) - bx0 + 1;
        const int by0 = t.Y0 >> kFB, nby = ((t.Y0 + t.RY - 1) >> kFB) - by0 + 1;
        static_assert(((TL::PITCH - 1) >> kFB) + 2 <= 32, "flag lanes");
        if ((tid & 31) < nbx && !PMC_DBG_BIT(a, 32))
            for (int by = tid >> 5; by < nby; by += kNT / 32)
                crowded |= __ldg(a.flag_in + (by0 + by) * g.FW + bx0 + (tid & 31)) == a.epoch_in;
        if (a.dbg_skip & 8) crowded = 1;
    }
    if (tid == 0) mbar_wait(mbar, phase);           // one poller; the others observe the completed phase once
    crowded = __syncthreads_or(crowded);
    mbar_wait(mbar, phase);
    if (NPL == 3 && crowded) return true;

    // which slots are in use anywhere in the staged box?  x6 (7 or 8 disks: NS = 8, only possible
    // with P3 staged), x4 (5 or 6: NS = 6), else NS = 4: bit 1 / bit 0 of `big`
    int big = 0;
    if (NPL == 4) {
        const float *x4 = reinterpret_cast<const float *>(sm + 2 * PLC);
        const float *x6 = reinterpret_cast<const float *>(sm + 3 * PLC);
        if (g.try_ns4) {
#pragma unroll 1
            for (int c = tid; c < TL::PLB; c += kNT)
                big |= (x6[c * 4] < kSentTest ? 2 : 0) | (x4[c * 4] < kSentTest ? 1 : 0);
        } else {            // dense system: no tile will qualify for NS = 4, scan one plane only
#pragma unroll 1
            for (int c = tid; c < TL::PLB; c += kNT) big |= x6[c * 4] < kSentTest ? 2 : 1;
        }
        big = __syncthreads_or(big & 2) ? 2 : (__syncthreads_or(big & 1) ? 1 : 0);
    } else {
        big = 1;
        if (g.try_ns4) {
            const float *x4 = reinterpret_cast<const float *>(sm + 2 * PLC);
            big = 0;
#pragma unroll 1
            for (int c = tid; c < TL::PLB; c += kNT) big |= x4[c * 4] < kSentTest ? 1 : 0;
            big = __syncthreads_or(big);
        }
    }
    if ((a.dbg_skip & 16) && big == 0) big = 1;
    const bool ns8 = NPL == 4 && big == 2, ns4 = big == 0;

    // ------------------------------------------------------------ the four sub-sweeps
    const int bq = tid / NAX, aq = tid - bq * NAX;  // fixed thread -> (column, row) of the active lattice
    if (!PMC_DBG_BIT(a, 1)) {
        if (ns4) {
#pragma unroll 1
            for (int k = 0; k < 4; k++) {
                colour_pass<4, TL>(sm, t, g, a, k, aq, bq, my_cnt);
                __syncthreads();
            }
        } else if (!ns8) {
#pragma unroll 1
            for (int k = 0; k < 4; k++) {
                colour_pass<6, TL>(sm, t, g, a, k, aq, bq, my_cnt);
                __syncthreads();
            }
        } else if (NPL == 4) {
#pragma unroll 1
            for (int k = 0; k < 4; k++) {
                colour_pass<8, TL>(sm, t, g, a, k, aq, bq, my_cnt);
                __syncthreads();
            }
        }
    }

    // ------------------------------------------------------------ this sweep's shiftCells + owned tile -> HBM
    // (no shift = a shift by d = 0: every disk stays, nothing immigrates)
    {
        const float d = do_shift ? a.shift_d : 0.0f;
        if (a.shift_f == 0 || !do_shift) {
            if (ns4) shift_store_pass<4, TL, NPL, 0>(sm, t, g, a, dout, sdir, d, tid, ctr);
            else if (!ns8) shift_store_pass<6, TL, NPL, 0>(sm, t, g, a, dout, sdir, d, tid, ctr);
            else if (NPL == 4) shift_store_pass<8, TL, NPL, 0>(sm, t, g, a, dout, sdir, d, tid, ctr);
        } else {
            if (ns4) shift_store_pass<4, TL, NPL, 1>(sm, t, g, a, dout, sdir, d, tid, ctr);
            else if (!ns8) shift_store_pass<6, TL, NPL, 1>(sm, t, g, a, dout, sdir, d, tid, ctr);
            else if (NPL == 4) shift_store_pass<8, TL, NPL, 1>(sm, t, g, a, dout, sdir, d, tid, ctr);
        }
    }
    return false;
}

// acceptance counts reduced warp-level, one atomic per warp (kernel.cu:228,413 accept_counter)
__device__ __forceinline__ void flush_counters(Counters *ctr, unsigned my_cnt)
{
    const unsigned my_trials = __reduce_add_sync(0xffffffffu, my_cnt >> 16);
    const unsigned my_acc = __reduce_add_sync(0xffffffffu, my_cnt & 0xFFFFu);
    if ((threadIdx.x & 31) == 0 && my_trials) {
        atomicAdd(&ctr->trials, (unsigned long long)my_trials);
        atomicAdd(&ctr->accepted, (unsigned long long)my_acc);
    }
}

// ---- crowded tile (a staged cell holds 7 or 8 disks): all four planes, half the rows at a time, in the
// shared memory of the fast tile.  Rare, and deliberately NOT inlined: the fast path keeps its own
// register allocation (64 registers without spills).  Returns (trials << 16) | accepted of this thread.
__device__ __noinline__ unsigned crowded_tile(const CUtensorMap *tmap_half, float4 *dout, const Geom4 *gp,
                                                        const SweepArgs *ap, Counters *ctr, int col0, int row0,
                                                        float4 *sm, uint64_t *mbar_fast)
{
    uint64_t *mbar2 = reinterpret_cast<uint64_t *>(sm + BoxH::PLB);
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(mbar_fast)) : "memory");
        mbar_init(mbar2, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    unsigned my_cnt = 0;
    // rows a half-height box can own: its kSYBH rows hold the halo on both sides and the upstream row
    const int tyh = (kSYBH - 1 - 2 * ap->hy) & ~1;
    unsigned phase = 0;
#pragma unroll 1
    for (int r = 0; r < ap->ty; r += tyh) {
        const int ty = min(tyh, ap->ty - r);
        if (row0 + r < gp->rows) {
            process_tile<BoxH, 4>(tmap_half, dout, *gp, *ap, ctr, col0, row0 + r, ap->tx, ty, sm, mbar2, phase, false,
                                  my_cnt);
            phase ^= 1u;
        }
        __syncthreads();                        // every thread is done with the tile before the next box lands on it
        if (threadIdx.x == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    return my_cnt;
}

// ---- one launch = one sweep.  FAST: 3 planes staged, 4 CTAs per SM, crowded tiles re-done in two or
// three 4-plane pieces in the same shared memory.  !FAST: 4 planes, 3 CTAs per SM, no flag
// lookup (pmc_set_tuning "four_plane": the A/B partner of the fast path; round 1 ran the slab boundary rows on it).
template <int MINB, bool FAST>
__global__ void __launch_bounds__(kNT, MINB)
sweep4_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_half,
              float4 *__restrict__ dout, const __grid_constant__ Geom4 g, const __grid_constant__ SweepArgs a, Counters *ctr)
{
    static_assert(4 * BoxH::PLANE_BYTES <= 3 * BoxF::PLANE_BYTES, "the half-height boxes reuse the fast tile's shared memory");
    extern __shared__ __align__(128) float4 sm[];
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sm + BoxF::PLB);
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // tile row (a launch may cover one or two bands of rows)
    const int tby = (int)blockIdx.y < a.by_n1 ? (int)blockIdx.y + a.by_off : (int)blockIdx.y - a.by_n1 + a.by_off2;
    const int col0 = (int)blockIdx.x * a.tx, row0 = tby * a.ty;
    unsigned my_cnt = 0;
    if (!FAST) {
        process_tile<BoxF, 4>(&tmap, dout, g, a, ctr, col0, row0, a.tx, a.ty, sm, mbar, 0u, true, my_cnt);
    } else if (process_tile<BoxF, 3>(&tmap, dout, g, a, ctr, col0, row0, a.tx, a.ty, sm, mbar, 0u, true, my_cnt)) {
        my_cnt += crowded_tile(&tmap_half, dout, &g, &a, ctr, col0, row0, sm, mbar);
    }
    flush_counters(ctr, my_cnt);
}

// ------------------------------------------------------------------ caller layout <-> internal layout
// caller: disk float[cell][2][8] + int16 n[cell] (include/pmc.h).  One thread per (internal
// cell, plane); margins are filled with the periodic images, everything beyond with empty cells.
__global__ void import4_kernel(const float *__restrict__ disk, const int16_t *__restrict__ n,
                               float4 *__restrict__ out, Geom4 g, int ghost, unsigned *__restrict__ flags, unsigned epoch)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int cols = 2 * g.CH;
    const long long cell = t >> 2;
    const int pl = (int)(t & 3);
    if (cell >= (long long)cols * g.ROWS) return;
    const int Y = (int)(cell / cols), X = (int)(cell - (long long)Y * cols);
    int gx = X - kMX, ly = Y - kMY;
    bool have = gx >= -kMX && gx < g.cps + kMX && ly >= -kMY && ly < g.rows + kMY;
    gx += gx < 0 ? g.cps : 0; gx -= gx >= g.cps ? g.cps : 0;
    long long src;
    if (g.wrap_y) {
        ly += ly < 0 ? g.rows : 0; ly -= ly >= g.rows ? g.rows : 0;
        src = (long long)ly * g.cps + gx;
    } else {
        src = (long long)(ly + ghost) * g.cps + gx;     // the caller's array carries `ghost` rows on each side
        have = have && (ly + ghost >= 0) && (ly + ghost < g.rows + 2 * ghost);
    }
    // first slot of this plane's x pair / quad, number of slots per coordinate
    const int s0 = pl < 2 ? 0 : (pl == 2 ? 4 : 6), ns = pl < 2 ? 4 : 2;
    float v[4];
    if (pl == 0) { v[0] = v[1] = v[2] = v[3] = kSent; }
    else if (pl == 1) { v[0] = v[1] = v[2] = v[3] = 0.f; }
    else { v[0] = v[1] = kSent; v[2] = v[3] = 0.f; }
    if (have) {
        int cnt = (int)__ldg(n + src);
        cnt = cnt < 0 ? 0 : (cnt > PMC_NMAX ? PMC_NMAX : cnt);
        const float *c = disk + src * 16;               // x0..x7, y0..y7
        if (pl == 0) {
#pragma unroll
            for (int q = 0; q < 4; q++) if (q < cnt) v[q] = __ldg(c + q);
        } else if (pl == 1) {
#pragma unroll
            for (int q = 0; q < 4; q++) if (q < cnt) v[q] = __ldg(c + 8 + q);
            if (cnt >= 5) v[3] = -v[3];                 // "5 or more" flag: sign of y3
        } else {
#pragma unroll
            for (int q = 0; q < 2; q++)
                if (q < ns && s0 + q < cnt) { v[q] = __ldg(c + s0 + q); v[2 + q] = __ldg(c + 8 + s0 + q); }
            // in-band count (the caller's unused slots may hold garbage: never copied)
            if (pl == 2 && cnt < 6) v[3] = __int_as_float(cnt);
            if (pl == 3 && cnt < PMC_NMAX) v[3] = __int_as_float(cnt);
            if (pl == 3 && cnt >= 7) flags[(Y >> kFB) * g.FW + (X >> kFB)] = epoch;      // crowded cell: P3 must be staged
        }
    }
    out[((long long)(Y * 4 + pl) * 2 + (X & 1)) * g.CH + (X >> 1)] = make_float4(v[0], v[1], v[2], v[3]);
}

// internal -> caller layout: every cell of the caller's array (slab: ghost rows included);
// one thread per (cell, caller chunk): x03, x47, y03, y47
__global__ void export4_kernel(const float4 *__restrict__ in, float4 *__restrict__ disk,
                               int16_t *__restrict__ n, Geom4 g, int ghost)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long cell = t >> 2;
    const int ch = (int)(t & 3);
    if (cell >= (long long)(g.rows + 2 * ghost) * g.cps) return;
    const int lr = (int)(cell / g.cps), gx = (int)(cell - (long long)lr * g.cps);
    const int X = gx + kMX, Y = lr - ghost + kMY;
    const float4 *cp = in + ((long long)(Y * 4) * 2 + (X & 1)) * g.CH + (X >> 1);
    const long long ps = (long long)2 * g.CH;           // plane stride
    const float4 p3 = __ldg(cp + 3 * ps);
    const int cnt = decode_cnt8(p3);
    float4 v;
    int s0;
    if (ch == 0) { v = __ldg(cp); s0 = 0; }
    else if (ch == 2) { v = __ldg(cp + ps); v.w = fabsf(v.w); s0 = 0; }
    else {
        const float4 p2 = __ldg(cp + 2 * ps);
        v = ch == 1 ? make_float4(p2.x, p2.y, p3.x, p3.y) : make_float4(p2.z, p2.w, p3.z, p3.w);
        s0 = 4;
    }
    // pmc.h: unused slots hold x = sentinel, y = 0 (the in-band counts are internal)
    const float fill = ch < 2 ? kSent : 0.f;
    v.x = s0 + 0 < cnt ? v.x : fill; v.y = s0 + 1 < cnt ? v.y : fill;
    v.z = s0 + 2 < cnt ? v.z : fill; v.w = s0 + 3 < cnt ? v.w : fill;
    disk[cell * 4 + ch] = v;
    if (ch == 0) n[cell] = (int16_t)cnt;
}

__global__ void flag_merge_kernel(unsigned *__restrict__ flags, const unsigned *__restrict__ recv_lo, int row_lo,
                                  const unsigned *__restrict__ recv_hi, int row_hi, int nwords, int FW, unsigned epoch)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    if (recv_lo[i] == epoch) flags[(size_t)row_lo * FW + i] = epoch;
    if (recv_hi[i] == epoch) flags[(size_t)row_hi * FW + i] = epoch;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

template <int MINB, bool FAST>
cudaError_t launch_cfg(const Geom4 &g, const void *tmap_in, const void *tmap_half, float4 *dout, const SweepArgs &a_in,
                       Counters *ctr, cudaStream_t st, int by0, int nby, int by1, int nby1)
{
    constexpr int SMEM = (int)((FAST ? 3 : 4) * BoxF::PLANE_BYTES);
    auto kern = sweep4_kernel<MINB, FAST>;
    static std::atomic<unsigned long long> attr_set{ 0ull };   // function attributes are per device; setting twice is harmless
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !((attr_set.load(std::memory_order_acquire) >> dev) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set.fetch_or(1ull << dev, std::memory_order_release);
    }
    if (a_in.tx < 2 || a_in.ty < 2) return cudaErrorInvalidValue;      // pmc4_plan_sweep was not called
    const int gy = (g.rows + a_in.ty - 1) / a_in.ty;
    SweepArgs a = a_in;
    a.by_off = by0;
    if (nby <= 0) { a.by_off = 0; nby = gy; }
    if (a.by_off + nby > gy) nby = gy - a.by_off;
    if (nby <= 0) return cudaSuccess;
    a.by_n1 = nby; a.by_off2 = by1;
    if (nby1 < 0 || by1 + nby1 > gy) nby1 = 0;
    dim3 grid((g.cps + a.tx - 1) / a.tx, nby + nby1);
    kern<<<grid, kNT, SMEM, st>>>(*(const CUtensorMap *)tmap_in, *(const CUtensorMap *)tmap_half, dout, g, a, ctr);
    return cudaGetLastError();
}

}  // namespace

// The tiling of one sweep, from its colour order and shift (a.offx / offy / shift_* must be set).
// A halo cell of colour k at distance d from the owned cells matters only if d later colours
// alternate the parity along that axis (each step towards the owned cells crosses one column /
// row), so per axis the halo is the longest parity-alternating subsequence of the colour order
// (2, 3 or 4) and colour k is computed 1 + las(0) - las(k) cells in from the region edge.
// The staged box is fixed (36 x 33 cells, at most 16 x 16 active cells per colour): a shallower
// halo leaves room for more owned columns / rows.
void pmc4_plan_sweep(SweepArgs &a, int full_halo)
{
    int lasx[4], lasy[4];
    for (int k = 3; k >= 0; k--) {
        lasx[k] = lasy[k] = 1;
        for (int m = k + 1; m < 4; m++) {
            if (a.offx[m] != a.offx[k] && lasx[m] + 1 > lasx[k]) lasx[k] = lasx[m] + 1;
            if (a.offy[m] != a.offy[k] && lasy[m] + 1 > lasy[k]) lasy[k] = lasy[m] + 1;
        }
    }
    if (full_halo) for (int k = 0; k < 4; k++) lasx[k] = lasy[k] = 4 - k;
    a.hx = lasx[0]; a.hy = lasy[0];
    a.lo_x = a.lo_y = 0;
    for (int k = 0; k < 4; k++) {
        a.lo_x |= (unsigned)(1 + lasx[0] - lasx[k]) << (4 * k);
        a.lo_y |= (unsigned)(1 + lasy[0] - lasy[k]) << (4 * k);
    }
    const int ex = (a.shift_on && a.shift_f == 0) ? 1 : 0, ey = (a.shift_on && a.shift_f == 1) ? 1 : 0;
    {   // region (0, 0) sits at column col0 - hx - exl, row row0 - hy - eyl (the slab origin is even); tile extents
        // may be odd, so the parities of col0 / row0 alternate from tile to tile: one set of words per parity pair
        const int sdir = (a.shift_d <= 0.0f) ? -1 : 1;
        const int exl = ex && sdir < 0, eyl = ey && sdir < 0;
        for (int py = 0; py < 2; py++)
            for (int px = 0; px < 2; px++) {
                const int xs = (a.hx + exl + px) & 1;                       // parity of the region's first internal column (kMX even)
                for (int k = 0; k < 4; k++) {
                    const int lox = (int)((a.lo_x >> (4 * k)) & 15u), loy = (int)((a.lo_y >> (4 * k)) & 15u);
                    const int pi = (a.offx[k] + a.hx + exl + px) & 1, pj = (a.offy[k] + a.hy + eyl + py) & 1;
                    const int i0 = lox + ((pi - lox) & 1), j0 = loy + ((pj - loy) & 1);
                    const int isu = i0 + xs, par = isu & 1;
                    const int off_own = j0 * BoxF::PITCH + par * BoxF::HB + (isu >> 1);
                    const int off_left = j0 * BoxF::PITCH + (1 - par) * BoxF::HB + ((isu - 1) >> 1);
                    a.colour_word[py * 8 + px * 4 + k] = (unsigned)i0 | ((unsigned)j0 << 4) | ((unsigned)lox << 8) | ((unsigned)loy << 12) |
                                                         ((unsigned)off_own << 16) | ((unsigned)off_left << 24);
                }
            }
    }
    // columns: region tx + 2 hx + ex, of which all but the two edge columns can be active in colour 0: <= 32;
    // rows: region ty + 2 hy + ey <= kSYB
    // (both extents are odd in sweeps that shift along x: 34 - 2 hx - 1 columns, 33 - 2 hy rows)
    a.tx = 34 - 2 * a.hx - ex;
    a.ty = kSYB - 2 * a.hy - ey;
    // shift_store_pass, shift along y: thread -> (segment, column) = (tid / tx, tid % tx) and rows per segment
    // = ceil(ty / nseg) by exact multiply-shifts (tid < 256, ty + nseg <= 64)
    a.sh_nseg = kNT / a.tx;
    a.sh_inv = 65536 / a.tx + 1;
    a.sh_inv2 = 65536 / a.sh_nseg + 1;
}

// The sweep-only part of the cells' Philox calls (philox_cell): counter = {cell, sweep_lo, sweep_hi, 0}.
void pmc4_philox_prepare(SweepArgs &a, const Geom4 &g)
{
    const unsigned long long p1 = 0xCD9E8D57ull * a.sweep_hi;                   // round 0, the product of counter word 2
    const unsigned a0 = (unsigned)(p1 >> 32) ^ a.sweep_lo ^ g.pk0[0];           // -> counter word 0 of round 1
    const unsigned a1 = (unsigned)p1;                                           // -> counter word 1 of round 1
    const unsigned long long q0 = 0xD2511F53ull * a0;                           // round 1, the product of counter word 0
    a.ph_e1 = a1 ^ g.pk0[1];
    a.ph_e2 = (unsigned)(q0 >> 32) ^ g.pk1[1];
    a.ph_e3 = (unsigned)q0 ^ g.pk1[2];
}

int pmc4_tile_rows(const Geom4 &g, const SweepArgs &a) { return (g.rows + a.ty - 1) / a.ty; }

// rows / chunk columns the internal array needs so that every staged box is in bounds, and the
// extent of the crowded-cell flag grid over it
void pmc4_alloc_shape(int cps, int rows, int *CH, int *ROWS, int *FW, int *FH)
{
    // a box starts at most 5 cells before its first owned column / row, the last tile starts before cps / rows
    const int cols = kMX + cps + BoxF::PITCH + 2, rws = kMY + rows + kSYB + 2;
    *CH = (cols + 1) / 2;
    *ROWS = rws;
    *FW = ((2 * *CH) >> kFB) + 1;
    *FH = (*ROWS >> kFB) + 1;
}

// half = 0: the box of a full tile; 1: the box of a half-height tile (crowded tiles)
int pmc4_make_tensor_map(void *tmap_out, const float4 *base, const Geom4 &g, int half)
{
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return (int)(e != cudaSuccess ? e : cudaErrorNotSupported);
        encode = (EncodeTiledFn)fn;
    }
    const int hb = BoxF::HB, syb = half ? kSYBH : kSYB;
    const cuuint64_t dims[4] = { (cuuint64_t)4 * g.CH, 2, 4, (cuuint64_t)g.ROWS };
    const cuuint64_t strides[3] = { (cuuint64_t)g.CH * 16, (cuuint64_t)g.CH * 32, (cuuint64_t)g.CH * 128 };
    const cuuint32_t box[4] = { (cuuint32_t)(4 * hb), 2, 1, (cuuint32_t)syb };
    const cuuint32_t estr[4] = { 1, 1, 1, 1 };
    CUresult r = encode((CUtensorMap *)tmap_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

cudaError_t pmc4_launch_flag_merge(unsigned *flags, const unsigned *recv_lo, int row_lo, const unsigned *recv_hi, int row_hi,
                                   int nrows, int FW, unsigned epoch, cudaStream_t st)
{
    const int nwords = nrows * FW;
    flag_merge_kernel<<<(nwords + 255) / 256, 256, 0, st>>>(flags, recv_lo, row_lo, recv_hi, row_hi, nwords, FW, epoch);
    return cudaGetLastError();
}

cudaError_t pmc4_launch_import(const Geom4 &g, int ghost, const float4 *disk, const int16_t *n, float4 *out,
                               unsigned *flags, unsigned epoch, cudaStream_t st)
{
    const long long threads = (long long)2 * g.CH * g.ROWS * 4;
    import4_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>((const float *)disk, n, out, g, ghost, flags, epoch);
    return cudaGetLastError();
}

cudaError_t pmc4_launch_export(const Geom4 &g, int ghost, const float4 *in, float4 *disk, int16_t *n, cudaStream_t st)
{
    const long long threads = (long long)(g.rows + 2 * ghost) * g.cps * 4;
    export4_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(in, disk, n, g, ghost);
    return cudaGetLastError();
}

// fast = 1: the 3-plane kernel with the crowded-cell flag lookup (tile rows whose boxes hold no
// ghost rows); 0: the 4-plane kernel.  a must have been planned (pmc4_plan_sweep).
cudaError_t pmc4_launch_sweep(const Geom4 &g, const void *tmap_in, const void *tmap_half, float4 *dout, const SweepArgs &a,
                              Counters *ctr, cudaStream_t st, int fast, int by0, int nby, int by1, int nby1)
{
    if (fast) return launch_cfg<4, true>(g, tmap_in, tmap_half, dout, a, ctr, st, by0, nby, by1, nby1);
    return launch_cfg<3, false>(g, tmap_in, tmap_half, dout, a, ctr, st, by0, nby, by1, nby1);
}
