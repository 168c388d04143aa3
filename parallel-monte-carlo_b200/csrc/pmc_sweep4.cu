// pmc_sweep4.cu -- the fused MC sweep (start.cu:237-260 loop body: 4 x subsweep_kernel
// subsweep.h:240-300, then shiftCells shiftCells.h:23-112) as ONE sm_100a kernel per sweep.
// This is the throughput path behind pmc_sweep(); pmc_sweep.cu keeps the generic kernel
// (any n_M, w < 2 sigma, tiny boxes, the single-colour call site).
//
// Differences from the generic kernel, chosen to cut issued instructions AND shared-memory
// wavefronts per trial (ncu: the LSU data pipe is the tightest resource of this path):
//   * INTERNAL STATE LAYOUT (handle-owned, never seen by the caller): one cell = 4 float4
//     chunks   P0 = x0..x3   P1 = y0..y3   P2 = x4 x5 y4 y5   P3 = x6 x7 y6 y7
//         chunk(X, Y, P) = ((Y*4 + P)*2 + (X & 1)) * CH + (X >> 1)
//     Even and odd columns are split so that the same-colour cells a warp works on are
//     contiguous (conflict-free LDS.128) and ONE 4-D TMA box per plane lands a tile in
//     shared memory in exactly the layout the sub-sweeps read: no index arithmetic at all.
//     The box is surrounded by margins (kMX columns, kMY rows) holding periodic images, written
//     by the CTA that produces the original cell (single GPU) or by the NCCL ring (slab rows),
//     so a tile never wraps.
//   * cells hold <= 6 disks in all but ~1e-5 of the cases at phi = 0.70, w = 2 sigma: a tile
//     whose staged cells all hold <= 6 runs the NS = 6 instantiation, which never touches P3
//     (3 instead of 4 chunks per neighbour cell, 24 instead of 32 pair tests per trial); a
//     tile whose cells all hold <= 4 (dilute systems) runs NS = 4 and touches neither P3 nor
//     P2; NS = 8 covers the rest.  The choice is per tile, from a scan of the staged box.
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
#include <mutex>

namespace {

constexpr float kSent = PMC_SENTINEL;
constexpr float kSentTest = 1.0e17f;      // x < kSentTest <=> slot in use

template <int TX, int TY>
struct Tile4 {
    static constexpr int H = 4;                                  // halo: one cell per colour
    static constexpr int HB = ((TX + 10) / 2 + 1) / 2 * 2;       // staged chunks per parity row (even)
    static constexpr int PITCH = 2 * HB;                         // chunks per staged row
    static constexpr int SYB = TY + 2 * H + 1;                   // staged rows
    static constexpr int PLB = PITCH * SYB;                      // chunks the TMA box brings per plane
    static constexpr int PLC = (PLB + 7) / 8 * 8;                // plane stride (128-byte aligned)
    // active cells per colour and row: at most 16 (half a warp, two conflict-free quarter-warps).
    // TX = 24 always fits; TX = 26 fits when the tile carries no extra column (shift along y).
    static constexpr int NAX = 16, NAY = (TY + 2 * H) / 2;
    static constexpr int THREADS = NAX * NAY;
    static constexpr size_t SMEM = (size_t)4 * PLC * 16 + 16;
    static_assert(TX % 2 == 0 && TY % 2 == 0 && TX + 2 * H <= 2 * NAX + 2, "tile shape");
    static_assert(4 * HB <= 256 && SYB <= 256, "TMA box extents");
};

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
__device__ __forceinline__ float cell_min_d2(const float4 *cp, float npx, float npy)
{
    const float4 p0 = cp[0], p1 = cp[PLC];
    // sign bit of y3 = "this cell holds 5 or more disks" (cell-local coordinates are positive):
    // only then are P2 (and P3) fetched, so quarter-warps whose 8 neighbour cells all hold <= 4
    // disks (61 % of them at phi = 0.70) spend no shared-memory wavefront on those planes
    const bool more = NS >= 6 && __float_as_int(p1.w) < 0;
    const float2 nx = make_float2(npx, npx), ny = make_float2(npy, npy);
    const float2 a = pair2(p0.x, p0.y, p1.x, p1.y, nx, ny);
    const float2 b = pair2(p0.z, p0.w, p1.z, fabsf(p1.w), nx, ny);
    float m = fminf(fminf(a.x, a.y), fminf(b.x, b.y));
    if (NS >= 6) {
        if (more) {
            const float4 p2 = cp[2 * PLC];
            const float2 c = pair2(p2.x, p2.y, p2.z, p2.w, nx, ny);
            m = fminf(m, fminf(c.x, c.y));
            if (NS == 8) {
                const float4 p3 = cp[3 * PLC];
                const float2 e = pair2(p3.x, p3.y, p3.z, p3.w, nx, ny);
                m = fminf(m, fminf(e.x, e.y));
            }
        }
    }
    return m;
}

// +-(odd integer < 2^24) as a float without I2F: 0x4B800000 | m23 is the float 2^24 + 2*m23
__device__ __forceinline__ float signed_odd24(uint32_t r)
{
    const float a = __uint_as_float(((r >> 8) & 0x7FFFFFu) | 0x4B800000u);
    const float mag = __fadd_rn(a, -16777215.0f);
    return __uint_as_float(__float_as_uint(mag) | (r & 0x80000000u));
}

// V2 shiftCells.h:23-112 for one destination cell.  The result is written in place into the
// staged tile in the PLAIN plane order (x0-3 | x4-7 | y0-3 | y4-7: slot -> address is one
// shift and one multiply-add); the store phase converts to P0..P3 on the way out.
// fx points at the first float of the destination cell's first chunk.
template <int NS, int F, int PLC>
__device__ __forceinline__ int shift_into_tile(const CellRegs &own, const CellRegs &up, float d, float w,
                                               float sshift, float *fx, int *dropped)
{
    constexpr int PF = PLC * 4;                         // floats between consecutive planes
    float *pf = F == 0 ? fx : fx + 2 * PF;              // f-coordinate plane (slots 0-3)
    constexpr int OFF = F == 0 ? 2 * PF : -2 * PF;      // to the other coordinate
    int n = 0, drop = 0;
#pragma unroll
    for (int i = 0; i < NS; i++) {
        const float fc = F == 0 ? f4get(own.x03, own.x47, i) : f4get(own.y03, own.y47, i);
        const float oc = F == 0 ? f4get(own.y03, own.y47, i) : f4get(own.x03, own.x47, i);
        const float D = __fadd_rn(fc, -d);
        // unused x slots hold the sentinel and fail on their own; unused y slots hold 0 / the count
        if ((F == 0 || i < own.cnt) && D > 0.0f && D <= w) {    // shiftCells.h:62
            float *p = pf + n + (n >> 2) * (PF - 4);
            p[0] = D; p[OFF] = oc;
            n++;
        }
    }
#pragma unroll
    for (int i = 0; i < NS; i++) {
        const float fc = F == 0 ? f4get(up.x03, up.x47, i) : f4get(up.y03, up.y47, i);
        const float oc = F == 0 ? f4get(up.y03, up.y47, i) : f4get(up.x03, up.x47, i);
        const float D = __fadd_rn(fc, -d);
        if (i < up.cnt && !(D > 0.0f && D <= w)) {      // shiftCells.h:94
            if (n < PMC_NMAX) {
                float *p = pf + n + (n >> 2) * (PF - 4);
                p[0] = __fadd_rn(D, sshift); p[OFF] = oc;   // shiftCells.h:97
                n++;
            } else drop++;
        }
    }
    *dropped = drop;
    return n;
}

// everything one CTA knows about its tile
struct TileCtx {
    int RX, RY;             // region the sub-sweeps work on (owned + halo + the extra upstream strip)
    int rx0, ry0;           // unwrapped global column / owned-relative row of region (0, 0)
    int xs;                 // region column i is staged column i + xs
    int ox0, oy0;           // region coordinates of the owned tile's corner
    int nox, noy;           // owned extent after clipping at the box / slab edge
};

// ---- one colour: one thread per active cell (subsweep.h:242-245), own cell in registers
template <int NS, int TX, int TY>
__device__ __forceinline__ void colour_pass(float4 *sm, const TileCtx &t, const Geom4 &g, const SweepArgs &a,
                                            int k, int aq, int bq, unsigned &my_trials, unsigned &my_acc)
{
    using TL = Tile4<TX, TY>;
    constexpr int HB = TL::HB, PITCH = TL::PITCH, PLC = TL::PLC;
    const int cps = g.cps;
    const float w = g.w, hw = g.hw, sigma2 = g.sigma2, dscale = g.dscale;
    const int lo = k + 1;                       // cells closer than lo to the region edge are stale
    const int pi = ((int)((a.offmask >> (2 * k)) & 1u) - t.rx0) & 1;       // region-column parity of the active colour
    const int pj = ((int)((a.offmask >> (2 * k + 1)) & 1u) - (g.row0 + t.ry0)) & 1;
    const int i = lo + ((pi - lo) & 1) + 2 * aq, j = lo + ((pj - lo) & 1) + 2 * bq;
    if (!(i < t.RX - lo && j < t.RY - lo)) return;
    const int is = i + t.xs, par = is & 1;
    float4 *pown = sm + j * PITCH + par * HB + (is >> 1);
    const float4 *pL = sm + j * PITCH + (1 - par) * HB + ((is - 1) >> 1);      // left neighbour; right = pL + 1
    const unsigned sown = smem_u32(pown);
    const float4 p0 = lds128<0>(sown), p1 = lds128<PLC * 16>(sown), p2 = lds128<2 * PLC * 16>(sown);
    float4 p3 = make_float4(kSent, kSent, 0.f, 0.f);
    if (NS == 8) p3 = lds128<3 * PLC * 16>(sown);
    const int cnt = NS == 8 ? decode_cnt8(p3) : decode_cnt6(p2);
    if (cnt == 0) return;                       // subsweep.h:252-254
    const bool owned = (unsigned)(i - t.ox0) < (unsigned)t.nox && (unsigned)(j - t.oy0) < (unsigned)t.noy;
    int gx = t.rx0 + i, gy = g.row0 + t.ry0 + j;
    gx += gx < 0 ? cps : 0; gx -= gx >= cps ? cps : 0;
    gy += gy < 0 ? cps : 0; gy -= gy >= cps ? cps : 0;
    const uint32_t cell_id = (uint32_t)gy * (uint32_t)cps + (uint32_t)gx;

    // neighbour part of one trial: smallest d2 against the 3 neighbour cells that can hold a
    // disk closer than sigma (w >= 2 sigma), or -1 when the proposal leaves the cell
    // (out_of_bound subsweep.h:73-88)
    auto neighbours_min_d2 = [&](const float px, const float py) -> float {
        const bool inb = px > 0.0f && px <= w && py > 0.0f && py <= w;
        const bool goL = px <= hw, goD = py <= hw;
        const float npxs = -__fadd_rn(px, goL ? w : -w);     // -(px - helper*w), subsweep.h:139-151
        const float npys = -__fadd_rn(py, goD ? w : -w);
        const float4 *pH = pL + (goL ? 0 : 1);
        const int dV = goD ? -PITCH : PITCH;
        float m = cell_min_d2<NS, PLC>(pH, npxs, -py);
        m = fminf(m, cell_min_d2<NS, PLC>(pown + dV, -px, npys));
        m = fminf(m, cell_min_d2<NS, PLC>(pH + dV, npxs, npys));
        return inb ? m : -1.0f;                 // out of the cell: rejected whatever the neighbours say
    };

    float ox[8] = { p0.x, p0.y, p0.z, p0.w, p2.x, p2.y, p3.x, p3.y };
    float oy[8] = { p1.x, p1.y, p1.z, fabsf(p1.w), p2.z, p2.w, p3.z, p3.w };     // y3 carries the "5 or more" flag in its sign
    uint32_t rw[8];
    philox4x32_10_keys(cell_id, a.sweep_lo, a.sweep_hi, 0u, g.pk0, g.pk1, rw[0], rw[1], rw[2], rw[3]);
    philox4x32_10_keys(cell_id, a.sweep_lo, a.sweep_hi, 1u, g.pk0, g.pk1, rw[4], rw[5], rw[6], rw[7]);
    // random_shuffle subsweep.h:50-58: physical partial Fisher-Yates, steps 0..3
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const uint32_t b16 = ((rw[2 * s] & 0xFFu) << 8) | (rw[2 * s + 1] & 0xFFu);
        const int mrem = cnt > s ? cnt - s : 1;             // s >= cnt: no-op (jj == s)
        const int jj = s + (int)((b16 * (uint32_t)mrem) >> 16);
        const float tx = ox[s], ty = oy[s];
        float nx = tx, ny = ty;
#pragma unroll
        for (int q = s + 1; q < NS; q++) {
            const bool p = (jj == q);
            nx = p ? ox[q] : nx; ny = p ? oy[q] : ny;
            ox[q] = p ? tx : ox[q]; oy[q] = p ? ty : oy[q];
        }
        ox[s] = nx; oy[s] = ny;
    }
    // trials 0..3 move slot s mod cnt (subsweep.h:279-297), all register indices static
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const bool cA = cnt > s;                            // slot == s
        const bool cB = (s == 3) && (cnt == 2);             // slot == 1 (only s = 3, cnt = 2)
        const float x = cA ? ox[s] : (cB ? ox[1] : ox[0]);
        const float y = cA ? oy[s] : (cB ? oy[1] : oy[0]);
        const float px = __fmaf_rn(signed_odd24(rw[2 * s]), dscale, x);     // make_move subsweep.h:60-71
        const float py = __fmaf_rn(signed_odd24(rw[2 * s + 1]), dscale, y);
        my_trials += owned ? 1u : 0u;
        float m = neighbours_min_d2(px, py);
        // own cell (calculate_energy_in_cell subsweep.h:105-117), j != slot
        const float2 npx = make_float2(-px, -px), npy = make_float2(-py, -py);
        float2 d01 = pair2(ox[0], ox[1], oy[0], oy[1], npx, npy);
        float2 d23 = pair2(ox[2], ox[3], oy[2], oy[3], npx, npy);
        const float big = 3.0e38f;
        if (s == 0) d01.x = big;
        if (s == 1) { d01.y = cA ? big : d01.y; d01.x = cA ? d01.x : big; }
        if (s == 2) { d23.x = cA ? big : d23.x; d01.x = cA ? d01.x : big; }
        if (s == 3) { d23.y = cA ? big : d23.y; d01.y = cB ? big : d01.y; d01.x = (cA | cB) ? d01.x : big; }
        m = fminf(m, fminf(fminf(d01.x, d01.y), fminf(d23.x, d23.y)));
        if (NS >= 6) {
            const float2 d45 = pair2(ox[4], ox[5], oy[4], oy[5], npx, npy);
            m = fminf(m, fminf(d45.x, d45.y));
        }
        if (NS == 8) {
            const float2 d67 = pair2(ox[6], ox[7], oy[6], oy[7], npx, npy);
            m = fminf(m, fminf(d67.x, d67.y));
        }
        // accept_move subsweep.h:194-217 (hard disks: accept iff in bounds and no overlap)
        const bool acc = !(m < sigma2);
        my_acc += (acc && owned) ? 1u : 0u;
        if (s == 0) { ox[0] = acc ? px : ox[0]; oy[0] = acc ? py : oy[0]; }
        else {
            const bool w0 = acc & !cA & !cB, w1 = acc & cB, ws = acc & cA;
            ox[s] = ws ? px : ox[s]; oy[s] = ws ? py : oy[s];
            ox[0] = w0 ? px : ox[0]; oy[0] = w0 ? py : oy[0];
            if (s == 3) { ox[1] = w1 ? px : ox[1]; oy[1] = w1 ? py : oy[1]; }
        }
    }
    // cpy_D_sh_to_Disk subsweep.h:29-36 (shuffled order is written back, like the reference)
    pown[0] = make_float4(ox[0], ox[1], ox[2], ox[3]);
    pown[PLC] = make_float4(oy[0], oy[1], oy[2], cnt >= 5 ? -oy[3] : oy[3]);
    if (NS >= 6) pown[2 * PLC] = make_float4(ox[4], ox[5], oy[4], oy[5]);
    if (NS == 8) pown[3 * PLC] = make_float4(ox[6], ox[7], oy[6], oy[7]);
}

// ---- shiftCells(f, d) of this sweep for the owned cells, in place (plain plane order out)
template <int NS, int TX, int TY, bool UNROLL>
__device__ __forceinline__ void shift_pass(float4 *sm, const TileCtx &t, const SweepArgs &a, float w, int sdir,
                                           int tid, Counters *ctr)
{
    using TL = Tile4<TX, TY>;
    constexpr int HB = TL::HB, PITCH = TL::PITCH, PLC = TL::PLC, THREADS = TL::THREADS;
    const float d = a.shift_d;
    const float sshift = __fmul_rn(w, (float)sdir);                 // shiftCells.h:84-86
    // One thread owns a strip of K consecutive owned cells along the shift axis and walks it
    // from the downstream end to the upstream end: cell u is rewritten only after raw cell
    // u+1 has been read.  The raw cell after the strip (next strip, or the extra upstream
    // row / column) is read before the barrier.
    constexpr int SEG1 = THREADS / TX, K1 = (TY + SEG1 - 1) / SEG1;     // f = 1: column strips of K1 rows
    constexpr int SEG0 = THREADS / TY < 8 ? THREADS / TY : 8, K0 = (TX + SEG0 - 1) / SEG0;   // f = 0: row strips of K0 columns
    int i0, j0, len, di, dj;
    if (a.shift_f == 1) {
        const int seg = tid / TX, col = tid - seg * TX, k0 = seg * K1;
        len = (seg < SEG1 && k0 < TY) ? min(K1, TY - k0) : 0;
        i0 = t.ox0 + col; j0 = t.oy0 + (sdir > 0 ? k0 : TY - 1 - k0);
        di = 0; dj = sdir;
    } else {
        const int row = tid / SEG0, seg = tid - row * SEG0, k0 = seg * K0;
        len = (row < TY && k0 < TX) ? min(K0, TX - k0) : 0;
        j0 = t.oy0 + row; i0 = t.ox0 + (sdir > 0 ? k0 : TX - 1 - k0);
        di = sdir; dj = 0;
    }
    auto cell_ptr = [&](int i, int j) -> float4 * {
        const int is = i + t.xs;
        return sm + j * PITCH + (is & 1) * HB + (is >> 1);
    };
    auto load_cell = [&](int i, int j, CellRegs &c) {           // P0..P3 -> plain x / y registers
        const float4 *p = cell_ptr(i, j);
        const float4 p0 = p[0], p1 = p[PLC], p2 = p[2 * PLC];
        c.x03 = p0; c.y03 = make_float4(p1.x, p1.y, p1.z, fabsf(p1.w));
        if (NS == 8) {
            const float4 p3 = p[3 * PLC];
            c.x47 = make_float4(p2.x, p2.y, p3.x, p3.y); c.y47 = make_float4(p2.z, p2.w, p3.z, p3.w);
            c.cnt = decode_cnt8(p3);
        } else {
            // NS = 6 / 4: the higher slots are known to be unused in this tile
            c.x47 = NS == 6 ? make_float4(p2.x, p2.y, kSent, kSent) : make_float4(kSent, kSent, kSent, kSent);
            c.y47 = NS == 6 ? make_float4(p2.z, p2.w, 0.f, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
            c.cnt = decode_cnt6(p2);
        }
    };
    CellRegs cur, edge;
    if (len > 0) {
        load_cell(i0, j0, cur);
        load_cell(i0 + len * di, j0 + len * dj, edge);
    }
    __syncthreads();
    constexpr int KMAX = K0 > K1 ? K0 : K1;
    // UNROLL: the walk is unrolled (no register moves between iterations); the persistent kernel,
    // which carries more live state, keeps it rolled to stay spill-free at 80 registers
#pragma unroll(UNROLL ? KMAX : 1)
    for (int u = 0; u < KMAX; u++) {
        if (u < len) {
            const int i = i0 + u * di, j = j0 + u * dj;
            CellRegs up = edge;
            if (u + 1 < len) load_cell(i + di, j + dj, up);
            float4 *p = cell_ptr(i, j);
            p[0] = make_float4(kSent, kSent, kSent, kSent);
            p[PLC] = make_float4(kSent, kSent, kSent, kSent);
            p[2 * PLC] = make_float4(0.f, 0.f, 0.f, 0.f);
            p[3 * PLC] = make_float4(0.f, 0.f, 0.f, 0.f);
            float *fx = reinterpret_cast<float *>(p);
            int dropped, nNew;
            if (a.shift_f == 0) nNew = shift_into_tile<NS, 0, PLC>(cur, up, d, w, sshift, fx, &dropped);
            else nNew = shift_into_tile<NS, 1, PLC>(cur, up, d, w, sshift, fx, &dropped);
            // in-band counts (plain order: y5 = plane 3 word 1, y7 = plane 3 word 3)
            if (nNew < PMC_NMAX) fx[3 * PLC * 4 + 3] = __int_as_float(nNew);
            if (nNew < 6) fx[3 * PLC * 4 + 1] = __int_as_float(nNew);
            if (nNew >= 5) fx[2 * PLC * 4 + 3] = -fx[2 * PLC * 4 + 3];      // "5 or more" flag: sign of y3
            if (dropped) {
                atomicOr(&ctr->status, PMC_STATUS_OVERFLOW);
                if ((unsigned)(i - t.ox0) < (unsigned)t.nox && (unsigned)(j - t.oy0) < (unsigned)t.noy)
                    atomicAdd(&ctr->lost, (unsigned long long)dropped);
            }
            cur = up;
        }
    }
    __syncthreads();
}

// ---- one tile of one sweep: stage, 4 colours, shiftCells, store.  tbx / tby = tile column / row;
// `phase` = parity of the mbarrier phase this staging completes (the barrier is reused by
// the persistent kernel); pf_* describe the launch grid for the L2 prefetch (pf_n = 0: none).
template <int TX, int TY, bool UNROLL_SHIFT>
__device__ __forceinline__ void process_tile(const CUtensorMap *tmap_p, float4 *__restrict__ dout, const Geom4 &g,
                                             const SweepArgs &a, Counters *ctr, int tbx, int tby,
                                             float4 *sm, uint64_t *mbar, unsigned phase,
                                             unsigned &my_trials, unsigned &my_acc)
{
    using TL = Tile4<TX, TY>;
    constexpr int H = TL::H, HB = TL::HB, PITCH = TL::PITCH, PLC = TL::PLC, NAX = TL::NAX, THREADS = TL::THREADS;
    const int tid = threadIdx.x;
    const int cps = g.cps;

    // this sweep's grid shift: the tile carries one extra row / column on the upstream side
    const bool do_shift = a.shift_on && !(a.dbg_skip & 2);
    const int sdir = (a.shift_d <= 0.0f) ? -1 : 1;                       // shiftCells.h:38-44
    const int exl = (do_shift && a.shift_f == 0 && sdir < 0), exh = (do_shift && a.shift_f == 0 && sdir > 0);
    const int eyl = (do_shift && a.shift_f == 1 && sdir < 0), eyh = (do_shift && a.shift_f == 1 && sdir > 0);
    TileCtx t;
    t.RX = TX + 2 * H + exl + exh; t.RY = TY + 2 * H + eyl + eyh;
    t.rx0 = tbx * TX - H - exl;
    t.ry0 = tby * TY - H - eyl;
    const int X0 = t.rx0 + kMX, Y0 = t.ry0 + kMY;   // region (0, 0) in internal array coordinates (>= 0)
    t.xs = X0 & 1;
    t.ox0 = H + exl; t.oy0 = H + eyl;
    t.nox = min(TX, cps - tbx * TX); t.noy = min(TY, g.rows - tby * TY);

    // ------------------------------------------------------------ stage the tile: 4 TMA boxes
    if (tid == 0) {
        mbar_expect_tx(mbar, (unsigned)(4 * TL::PLB * 16));
#pragma unroll
        for (int p = 0; p < 4; p++) tma_load_4d(sm + p * PLC, tmap_p, 4 * ((X0 - t.xs) >> 1), 0, p, Y0, mbar);
        if (a.prefetch_ahead > 0) {
            // warm L2 for the tile a CTA slot freed by this wave will stage (blocks are issued in order)
            const int nb = blockIdx.y * gridDim.x + blockIdx.x + a.prefetch_ahead;
            if (nb < (int)(gridDim.x * gridDim.y)) {
                const int gr2 = nb / gridDim.x, bx2 = nb - gr2 * gridDim.x;
                const int by2 = gr2 < a.by_n1 ? gr2 + a.by_off : gr2 - a.by_n1 + a.by_off2;
                const int X2 = bx2 * TX - H - exl + kMX, Y2 = by2 * TY - H - eyl + kMY;
#pragma unroll
                for (int p = 0; p < 4; p++) tma_prefetch_4d(tmap_p, 4 * ((X2 - (X2 & 1)) >> 1), 0, p, Y2);
            }
        }
        mbar_wait(mbar, phase);     // one poller; the others observe the completed phase once
    }
    __syncthreads();
    mbar_wait(mbar, phase);

    // does any staged cell hold 7 or 8 disks (x6 in use), or 5 or 6 (x4 in use)?  Otherwise P3
    // (P3 and P2) are never needed: bit 1 / bit 0 of `big`
    int big = 0;
    {
        const float *x4 = reinterpret_cast<const float *>(sm + 2 * PLC);
        const float *x6 = reinterpret_cast<const float *>(sm + 3 * PLC);
        if (g.try_ns4) {
#pragma unroll 1
            for (int c = tid; c < TL::PLB; c += THREADS)
                big |= (x6[c * 4] < kSentTest ? 2 : 0) | (x4[c * 4] < kSentTest ? 1 : 0);
        } else {            // dense system: no tile will qualify for NS = 4, scan one plane only
#pragma unroll 1
            for (int c = tid; c < TL::PLB; c += THREADS) big |= x6[c * 4] < kSentTest ? 2 : 1;
        }
    }
    big = __syncthreads_or(big & 2) ? 2 : (__syncthreads_or(big & 1) ? 1 : 0);
    if (a.dbg_skip & 8) big = 2;
    if ((a.dbg_skip & 16) && big == 0) big = 1;
    const bool ns8 = big == 2, ns4 = big == 0;

    // ------------------------------------------------------------ the four sub-sweeps
    const int bq = tid / NAX, aq = tid - bq * NAX;  // fixed thread -> (column, row) of the active lattice
    if (!(a.dbg_skip & 1)) {
        if (ns4) {
#pragma unroll 1
            for (int k = 0; k < 4; k++) {
                colour_pass<4, TX, TY>(sm, t, g, a, k, aq, bq, my_trials, my_acc);
                __syncthreads();
            }
        } else if (!ns8) {
#pragma unroll 1
            for (int k = 0; k < 4; k++) {
                colour_pass<6, TX, TY>(sm, t, g, a, k, aq, bq, my_trials, my_acc);
                __syncthreads();
            }
        } else {
#pragma unroll 1
            for (int k = 0; k < 4; k++) {
                colour_pass<8, TX, TY>(sm, t, g, a, k, aq, bq, my_trials, my_acc);
                __syncthreads();
            }
        }
    }

    // ------------------------------------------------------------ this sweep's shiftCells, owned cells only
    if (do_shift) {
        if (ns4) shift_pass<4, TX, TY, UNROLL_SHIFT>(sm, t, a, g.w, sdir, tid, ctr);
        else if (!ns8) shift_pass<6, TX, TY, UNROLL_SHIFT>(sm, t, a, g.w, sdir, tid, ctr);
        else shift_pass<8, TX, TY, UNROLL_SHIFT>(sm, t, a, g.w, sdir, tid, ctr);
    }

    // ------------------------------------------------------------ owned tile -> HBM (+ periodic images into the margins)
    if (!(a.dbg_skip & 4)) {
        // thread -> fixed (chunk column h, parity, plane), rows strided: a warp stores runs of
        // TX/2 consecutive float4.  After the shift the staged cells are in plain plane order
        // (x0-3 | x4-7 | y0-3 | y4-7) and are converted to P0..P3 here.
        constexpr int HX = TX / 2, RSTEP = THREADS / (8 * HX);   // threads beyond RSTEP * 8 * HX do not store
        const int h = tid % HX, pr = (tid / HX) & 1, pl = (tid / (2 * HX)) & 3, rg = tid / (8 * HX);
        const int ox = 2 * h + pr;                                  // owned column (parity == internal column parity)
        const int ux = tbx * TX + ox, uy0 = tby * TY;
        if (ox < t.nox && rg < RSTEP) {
            const int is = t.ox0 + ox + t.xs;
            const float4 *cell0 = sm + (is & 1) * HB + (is >> 1) + (t.oy0 + rg) * PITCH;
            // source of output plane pl: P0 <- x03, P1 <- y03, P2 <- x47.xy y47.xy, P3 <- x47.zw y47.zw
            const float4 *src = cell0 + (do_shift ? (pl == 0 ? 0 : (pl == 1 ? 2 : 1)) : pl) * PLC;
            const int half = (pl == 3) ? 2 : 0;                     // float offset of the pair inside x47 / y47
            const bool split = do_shift && pl >= 2;
            auto fetch = [&](const float4 *s) -> float4 {
                if (!split) return *s;
                const float2 xa = *reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(s) + half);
                const float2 ya = *reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(s + 2 * PLC) + half);
                return make_float4(xa.x, xa.y, ya.x, ya.y);
            };
            const long long rstride = (long long)8 * g.CH;          // float4 chunks per internal row
            float4 *dst = dout + ((long long)(kMY + uy0 + rg) * 4 + pl) * 2 * g.CH + (long long)pr * g.CH + ((kMX + ux) >> 1);
            // periodic image of this column inside the margins (cps is even: parity is kept)
            const int ximg = ux < kMX ? cps / 2 : (ux >= cps - kMX ? -(cps / 2) : 0);
            const bool yedge = g.wrap_y && (uy0 < kMY || uy0 + t.noy > g.rows - kMY);
            if (!ximg && !yedge) {
#pragma unroll 2
                for (int oyy = rg; oyy < t.noy; oyy += RSTEP) {
                    *dst = fetch(src);
                    src += RSTEP * PITCH; dst += RSTEP * rstride;
                }
            } else {
#pragma unroll 1
                for (int oyy = rg; oyy < t.noy; oyy += RSTEP) {
                    const float4 v = fetch(src);
                    const int uy = uy0 + oyy;
                    dst[0] = v;
                    if (ximg) dst[ximg] = v;
                    if (g.wrap_y) {
                        const int yimg = uy < kMY ? g.rows : (uy >= g.rows - kMY ? -g.rows : 0);
                        if (yimg) {
                            float4 *di = dst + (long long)yimg * rstride;
                            di[0] = v;
                            if (ximg) di[ximg] = v;
                        }
                    }
                    src += RSTEP * PITCH; dst += RSTEP * rstride;
                }
            }
        }
    }

}

// acceptance counts reduced warp-level, one atomic per warp (kernel.cu:228,413 accept_counter)
__device__ __forceinline__ void flush_counters(Counters *ctr, unsigned my_trials, unsigned my_acc)
{
    my_trials = __reduce_add_sync(0xffffffffu, my_trials);
    my_acc = __reduce_add_sync(0xffffffffu, my_acc);
    if ((threadIdx.x & 31) == 0 && my_trials) {
        atomicAdd(&ctr->trials, (unsigned long long)my_trials);
        atomicAdd(&ctr->accepted, (unsigned long long)my_acc);
    }
}

// ---- one launch = one sweep (slab runs, and the fallback of the persistent kernel)
template <int TX, int TY, int MINB>
__global__ void __launch_bounds__(Tile4<TX, TY>::THREADS, MINB)
sweep4_kernel(const __grid_constant__ CUtensorMap tmap, float4 *__restrict__ dout, const Geom4 g,
              const SweepArgs a, Counters *ctr)
{
    using TL = Tile4<TX, TY>;
    extern __shared__ __align__(128) float4 sm[];
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sm + 4 * TL::PLC);
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // tile row (a launch may cover one or two bands of rows)
    const int tby = (int)blockIdx.y < a.by_n1 ? (int)blockIdx.y + a.by_off : (int)blockIdx.y - a.by_n1 + a.by_off2;
    unsigned my_trials = 0, my_acc = 0;
    process_tile<TX, TY, true>(&tmap, dout, g, a, ctr, (int)blockIdx.x, tby, sm, mbar, 0u, my_trials, my_acc);
    flush_counters(ctr, my_trials, my_acc);
}

// ---- persistent kernel: n_steps sweeps in ONE cooperative launch (single GPU).  CTA c owns
// tiles c, c + G, c + 2G, ... of every sweep; there is no grid-wide barrier between sweeps:
// a tile of sweep s starts as soon as the (up to 4 x 4, normally 3 x 3) tiles of sweep s-1
// that produced its staged box - and that were the last readers of the cells it is about to
// overwrite in the other ping-pong buffer - have published done[tile] >= s.  This removes the
// per-sweep launch and the idle tail of the last partial wave (~5 % at N = 2^24).
struct StepArgs { unsigned offmask, sweep_lo, sweep_hi; int shift_f; float shift_d; };
constexpr int kStepCap = 1024;
// per-sweep arguments of the batch in flight: constant memory, indexed by the (warp-uniform) sweep
// counter, so they live in the uniform datapath instead of in vector registers
__constant__ StepArgs c_steps[kStepCap];

__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int TX, int TY, int MINB>
__global__ void __launch_bounds__(Tile4<TX, TY>::THREADS, MINB)
sweep4_persistent(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                  float4 *__restrict__ buf0, float4 *__restrict__ buf1, const Geom4 g,
                  int n_steps, int src0, int *done, Counters *ctr, int dbg)
{
    using TL = Tile4<TX, TY>;
    extern __shared__ __align__(128) float4 sm[];
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sm + 4 * TL::PLC);
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int gx = (g.cps + TX - 1) / TX, gy = (g.rows + TY - 1) / TY, ntiles = gx * gy;
    unsigned phase = 0;
#pragma unroll 1
    for (int s = 0; s < n_steps; s++) {
        SweepArgs sa;
        sa.offmask = c_steps[s].offmask; sa.sweep_lo = c_steps[s].sweep_lo; sa.sweep_hi = c_steps[s].sweep_hi;
        sa.shift_on = 1; sa.shift_f = c_steps[s].shift_f; sa.shift_d = c_steps[s].shift_d;
        sa.sanitize_in = 0; sa.by_off = 0; sa.by_n1 = 0x7fffffff; sa.by_off2 = 0; sa.prefetch_ahead = 0; sa.dbg_skip = dbg;
        const int src = (src0 + s) & 1;
#pragma unroll 1
        for (int it = blockIdx.x; it < ntiles; it += gridDim.x) {
            // every other sweep starts half a grid away: the tiles a CTA runs first never depend on
            // the tiles the previous sweep finished last (first and last tile rows are periodic
            // neighbours), so there is no idle tail between sweeps
            int tile = it + ((s & 1) ? ntiles / 2 : 0);
            tile -= tile >= ntiles ? ntiles : 0;
            const int tby = tile / gx, tbx = tile - tby * gx;
            if (s > 0 && tid < 16) {
                // producers (sweep s-1) of every cell this tile stages or overwrites: the tiles under
                // the four corner-ish columns / rows of the staged box, wrapped like the images
                const int x0 = tbx * TX, y0 = tby * TY;
                const int xl = min(x0 + TX, g.cps) - 1, yl = min(y0 + TY, g.rows) - 1;
                const int sx = tid & 3, sy = tid >> 2;
                int cx = sx == 0 ? x0 - kMX : (sx == 1 ? x0 : (sx == 2 ? xl : xl + kMX));
                int cy = sy == 0 ? y0 - kMY : (sy == 1 ? y0 : (sy == 2 ? yl : yl + kMY));
                cx += cx < 0 ? g.cps : 0; cx -= cx >= g.cps ? g.cps : 0;
                cy += cy < 0 ? g.rows : 0; cy -= cy >= g.rows ? g.rows : 0;
                const int *flag = done + (cy / TY) * gx + cx / TX;
                while (ld_acquire(flag) < s) __nanosleep(64);
            }
            __syncthreads();
            if (tid == 0) asm volatile("fence.proxy.async;" ::: "memory");   // the TMA reads what the producers stored
            unsigned my_trials = 0, my_acc = 0;
            process_tile<TX, TY, false>(src ? &tm1 : &tm0, src ? buf0 : buf1, g, sa, ctr, tbx, tby, sm, mbar, phase, my_trials, my_acc);
            flush_counters(ctr, my_trials, my_acc);
            phase ^= 1u;
            // bar.sync orders every thread's stores before thread 0's release (cumulative at gpu scope);
            // it also frees the shared memory for the next box
            __syncthreads();
            if (tid == 0) st_release(done + tile, s + 1);
        }
    }
}

// ------------------------------------------------------------------ caller layout <-> internal layout
// caller: disk float[cell][2][8] + int16 n[cell] (include/pmc.h).  One thread per (internal
// cell, plane); margins are filled with the periodic images, everything beyond with empty cells.
__global__ void import4_kernel(const float *__restrict__ disk, const int16_t *__restrict__ n,
                               float4 *__restrict__ out, Geom4 g, int ghost)
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

// tile configurations (PMC_TILE4 selects; 1 is the tuned default)
struct TileCfg { int tx, ty, hb, syb, h; };
template <int TX, int TY> constexpr TileCfg cfg_of() { return { TX, TY, Tile4<TX, TY>::HB, Tile4<TX, TY>::SYB, Tile4<TX, TY>::H }; }
constexpr TileCfg kCfgs[] = { cfg_of<24, 40>(), cfg_of<24, 24>(), cfg_of<24, 32>(), cfg_of<24, 48>() };
// default (index 0): sweeps that shift along y use 26-column tiles (16, 15, 14, 13 active cells
// per half-warp instead of 15, 14, 13, 12); the staged box is the same, so is the tensor map
constexpr TileCfg kCfgWideY = cfg_of<26, 40>();
static_assert(kCfgWideY.hb == kCfgs[0].hb && kCfgWideY.syb == kCfgs[0].syb, "one TMA box for both default tilings");

int tile_index()
{
    static const int idx = [] {
        const char *e = getenv("PMC_TILE4");
        const int i = e ? atoi(e) : 1;      // tuned default: 24 / 26 x 24 tiles, 3 CTAs per SM
        return (i >= 0 && i < (int)(sizeof(kCfgs) / sizeof(kCfgs[0]))) ? i : 1;
    }();
    return idx;
}

template <int TX, int TY, int MINB>
cudaError_t launch_cfg(const Geom4 &g, const void *tmap_in, float4 *dout, const SweepArgs &a_in, Counters *ctr,
                       cudaStream_t st, int by0, int nby, int by1, int nby1)
{
    using TL = Tile4<TX, TY>;
    auto kern = sweep4_kernel<TX, TY, MINB>;
    static bool attr_set[64] = { false };           // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL::SMEM);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const int gy = (g.rows + TY - 1) / TY;
    SweepArgs a = a_in;
    a.by_off = by0;
    if (nby <= 0) { a.by_off = 0; nby = gy; }
    if (a.by_off + nby > gy) nby = gy - a.by_off;
    if (nby <= 0) return cudaSuccess;
    a.by_n1 = nby; a.by_off2 = by1;
    if (nby1 < 0 || by1 + nby1 > gy) nby1 = 0;
    dim3 grid((g.cps + TX - 1) / TX, nby + nby1);
    kern<<<grid, TL::THREADS, TL::SMEM, st>>>(*(const CUtensorMap *)tmap_in, dout, g, a, ctr);
    return cudaGetLastError();
}

}  // namespace

// persistent multi-sweep launch (single GPU, default tiling): returns cudaErrorNotSupported when
// the geometry does not qualify (the caller then launches one kernel per sweep)
int pmc4_step_capacity() { return kStepCap; }

// steps_host: n_steps Pmc4Step records (pageable is fine).  The constant bank is shared by every
// handle of this process on this device: a launch waits for the previous batch to finish.
cudaError_t pmc4_launch_persistent(const Geom4 &g, const void *tmap0, const void *tmap1, float4 *buf0, float4 *buf1,
                                   const void *steps_host, int n_steps, int src0, int *done_dev, Counters *ctr,
                                   int dbg, cudaStream_t st)
{
    constexpr int TX = 24, TY = 24, MINB = 3;
    using TL = Tile4<TX, TY>;
    if (tile_index() != 1 || !g.wrap_y) return cudaErrorNotSupported;
    // a clipped last tile narrower than the margin would make the image producers span two tiles
    const int remx = g.cps % TX, remy = g.rows % TY;
    if ((remx && remx < kMX) || (remy && remy < kMY) || g.cps < 2 * TX || g.rows < 2 * TY) return cudaErrorNotSupported;
    auto kern = sweep4_persistent<TX, TY, MINB>;
    static int grid_max = 0;
    if (!grid_max) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL::SMEM);
        if (e != cudaSuccess) return e;
        int dev = 0, sms = 0, per_sm = 0, coop = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TL::THREADS, TL::SMEM);
        if (e != cudaSuccess) return e;
        if (!coop || per_sm < 1) return cudaErrorNotSupported;
        grid_max = sms * per_sm;
    }
    if (n_steps > kStepCap) return cudaErrorInvalidValue;
    const int gx = (g.cps + TX - 1) / TX, gy = (g.rows + TY - 1) / TY;
    const int grid = gx * gy < grid_max ? gx * gy : grid_max;
    static std::mutex mu;
    static cudaEvent_t batch_done = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    cudaError_t e;
    if (!batch_done) { e = cudaEventCreateWithFlags(&batch_done, cudaEventDisableTiming); if (e != cudaSuccess) return e; }
    else { e = cudaEventSynchronize(batch_done); if (e != cudaSuccess) return e; }     // c_steps is free again
    e = cudaMemcpyToSymbolAsync(c_steps, steps_host, (size_t)n_steps * sizeof(StepArgs), 0, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);                  // steps_host may be pageable and reused by the caller
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(done_dev, 0, (size_t)gx * gy * sizeof(int), st);
    if (e != cudaSuccess) return e;
    CUtensorMap t0 = *(const CUtensorMap *)tmap0, t1 = *(const CUtensorMap *)tmap1;
    Geom4 gg = g;
    void *args[] = { &t0, &t1, &buf0, &buf1, &gg, &n_steps, &src0, &done_dev, &ctr, &dbg };
    e = cudaLaunchCooperativeKernel((void *)kern, dim3(grid), dim3(TL::THREADS), args, TL::SMEM, st);
    if (e != cudaSuccess) return e;
    return cudaEventRecord(batch_done, st);
}

int pmc4_tile_count(const Geom4 &g)
{
    const TileCfg c = kCfgs[tile_index()];
    return ((g.cps + c.tx - 1) / c.tx) * ((g.rows + c.ty - 1) / c.ty);
}

int pmc4_tile_rows(const Geom4 &g) { const int ty = kCfgs[tile_index()].ty; return (g.rows + ty - 1) / ty; }

int pmc4_tile_x() { return kCfgs[tile_index()].tx; }
int pmc4_tile_y() { return kCfgs[tile_index()].ty; }

// rows / chunk columns the internal array needs so that every staged box is in bounds
void pmc4_alloc_shape(int cps, int rows, int *CH, int *ROWS)
{
    const TileCfg c = kCfgs[tile_index()];
    const int gx = (cps + c.tx - 1) / c.tx, gy = (rows + c.ty - 1) / c.ty;
    // last staged column: kMX + (gx-1)*TX - H - 1 (rounded down to even) + 2*HB - 1
    int cols = kMX + (gx - 1) * c.tx - c.h + 2 * c.hb + 2;
    if (tile_index() <= 1) {
        const int gxw = (cps + kCfgWideY.tx - 1) / kCfgWideY.tx;
        const int colsw = kMX + (gxw - 1) * kCfgWideY.tx - kCfgWideY.h + 2 * kCfgWideY.hb + 2;
        cols = colsw > cols ? colsw : cols;
    }
    const int cols_img = cps + 2 * kMX;
    const int cc = cols > cols_img ? cols : cols_img;
    *CH = (cc + 1) / 2;
    const int r = kMY + (gy - 1) * c.ty - c.h + c.syb + 1;
    const int r_img = rows + 2 * kMY;
    *ROWS = r > r_img ? r : r_img;
}

int pmc4_make_tensor_map(void *tmap_out, const float4 *base, const Geom4 &g)
{
    const TileCfg c = kCfgs[tile_index()];
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return (int)(e != cudaSuccess ? e : cudaErrorNotSupported);
        encode = (EncodeTiledFn)fn;
    }
    const cuuint64_t dims[4] = { (cuuint64_t)4 * g.CH, 2, 4, (cuuint64_t)g.ROWS };
    const cuuint64_t strides[3] = { (cuuint64_t)g.CH * 16, (cuuint64_t)g.CH * 32, (cuuint64_t)g.CH * 128 };
    const cuuint32_t box[4] = { (cuuint32_t)(4 * c.hb), 2, 1, (cuuint32_t)c.syb };
    const cuuint32_t estr[4] = { 1, 1, 1, 1 };
    CUresult r = encode((CUtensorMap *)tmap_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

cudaError_t pmc4_launch_import(const Geom4 &g, int ghost, const float4 *disk, const int16_t *n, float4 *out, cudaStream_t st)
{
    const long long threads = (long long)2 * g.CH * g.ROWS * 4;
    import4_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>((const float *)disk, n, out, g, ghost);
    return cudaGetLastError();
}

cudaError_t pmc4_launch_export(const Geom4 &g, int ghost, const float4 *in, float4 *disk, int16_t *n, cudaStream_t st)
{
    const long long threads = (long long)(g.rows + 2 * ghost) * g.cps * 4;
    export4_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(in, disk, n, g, ghost);
    return cudaGetLastError();
}

cudaError_t pmc4_launch_sweep(const Geom4 &g, const void *tmap_in, float4 *dout, const SweepArgs &a,
                              Counters *ctr, cudaStream_t st, int by0, int nby, int by1, int nby1)
{
    switch (tile_index()) {
    case 1:
        if (a.shift_on && a.shift_f == 0) return launch_cfg<24, 24, 3>(g, tmap_in, dout, a, ctr, st, by0, nby, by1, nby1);
        return launch_cfg<26, 24, 3>(g, tmap_in, dout, a, ctr, st, by0, nby, by1, nby1);
    case 2: return launch_cfg<24, 32, 2>(g, tmap_in, dout, a, ctr, st, by0, nby, by1, nby1);
    case 3: return launch_cfg<24, 48, 1>(g, tmap_in, dout, a, ctr, st, by0, nby, by1, nby1);
    default:
        if (a.shift_on && a.shift_f == 0) return launch_cfg<24, 40, 2>(g, tmap_in, dout, a, ctr, st, by0, nby, by1, nby1);
        return launch_cfg<26, 40, 2>(g, tmap_in, dout, a, ctr, st, by0, nby, by1, nby1);
    }
}
