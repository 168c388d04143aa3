// pmc_sweep.cu -- the checkerboard sub-sweep (reference subsweep.h:240-300) for sm_100a.
//
// One kernel template, two instantiations:
//   NCOL = 1  one colour, in place  -> pmc_subsweep   (call site start.cu:242-245)
//   NCOL = 4  a whole MC sweep: pending shiftCells applied while staging, then the four
//             colours back to back on a tile held in shared memory (temporal blocking with
//             a halo of 4 cells recomputed redundantly), out of place -> pmc_sweep.
// The redundant halo work is legal because every random number is a pure function of
// (seed, sweep, global cell id, trial index): two CTAs recomputing the same cell get the
// same bits, so the result is independent of the tiling and of the number of GPUs.
//
// Phases of one CTA:
//   0a  raw tile (+1 upstream row or column) HBM -> shared memory with 16-byte cp.async,
//       four lanes per cell so every warp instruction moves 512 contiguous bytes
//   0b  pending shiftCells(f, d) in place in shared memory, one thread per cell, batches
//       ordered downstream -> upstream so no cell is overwritten before it has been read
//   1-4 sub-sweeps: one thread per active cell (the reference's V1 mapping,
//       subsweep.h:242-245), own cell in registers, neighbour cells read with LDS.128
//   5   owned tile shared memory -> HBM, again four lanes per cell (coalesced STG.128)
// Shared-memory tile layout: four float4 planes (x0-3, x4-7, y0-3, y4-7), each plane stored
// row by row (pitch a multiple of 4 chunks so rows j-1 / j+1 share bank groups) with even
// and odd columns split so that the same-colour cells a warp works on are contiguous.
// Overlap tests use the Blackwell packed-FP32 instructions (FADD2 / FMUL2 / FFMA2).
#include "pmc_internal.cuh"
#include <stdlib.h>

namespace {

constexpr float kSent = PMC_SENTINEL;

template <int NCOL, int TX, int TY>
struct Tile {
    static constexpr int H = NCOL;                  // halo cells on each side
    static constexpr int RX = TX + 2 * H, RY = TY + 2 * H;   // region the sub-sweeps work on
    static constexpr int SX = RX + 1, SY = RY + 1;  // staged area: + the upstream row / column of a pending shift
    static constexpr int PITCH = (SX + 3) / 4 * 4;  // float4 chunks per staged row
    static constexpr int HALF = PITCH / 2;
    static constexpr int PL = PITCH * SY + 1;       // plane stride (odd: the 4 planes of a cell fall in 4 bank groups)
    // most active cells per row / column in a sub-sweep; NAX a multiple of 8 keeps every
    // quarter-warp inside one row of same-colour cells (conflict-free LDS.128)
    static constexpr int NAX = (RX - 2) / 2, NAY = (RY - 2) / 2;
    static constexpr size_t SMEM = (size_t)PL * 64 + ((PITCH * SY + 15) / 16) * 16;
    static_assert(RX % 2 == 0 && RY % 2 == 0, "region edges must be even");
    static_assert((SX + 1) / 2 <= HALF, "even columns must fit in half a row");
};

// staged-area coordinates (i, j) -> chunk index inside a plane
template <int PITCH, int HALF>
__device__ __forceinline__ int sidx(int i, int j) { return j * PITCH + (i & 1) * HALF + (i >> 1); }

// v in [-m, 2m) unless the box is smaller than a tile region (then a real modulo)
__device__ __forceinline__ int wrap_fast(int v, int m, bool small)
{
    if (small) return wrap_mod(v, m);
    return v < 0 ? v + m : (v >= m ? v - m : v);
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;\n" ::: "memory");
}

// Slots >= cnt are overwritten with `fill` (pmc.h: the caller's unused slots may hold
// garbage, as in the reference).
__device__ __forceinline__ void sanitize_x(CellRegs &c, float fill)
{
    const int n = c.cnt;
    c.x03.x = n > 0 ? c.x03.x : fill; c.x03.y = n > 1 ? c.x03.y : fill;
    c.x03.z = n > 2 ? c.x03.z : fill; c.x03.w = n > 3 ? c.x03.w : fill;
    c.x47.x = n > 4 ? c.x47.x : fill; c.x47.y = n > 5 ? c.x47.y : fill;
    c.x47.z = n > 6 ? c.x47.z : fill; c.x47.w = n > 7 ? c.x47.w : fill;
}

__device__ __forceinline__ void sanitize_y(CellRegs &c, float fill)
{
    const int n = c.cnt;
    c.y03.x = n > 0 ? c.y03.x : fill; c.y03.y = n > 1 ? c.y03.y : fill;
    c.y03.z = n > 2 ? c.y03.z : fill; c.y03.w = n > 3 ? c.y03.w : fill;
    c.y47.x = n > 4 ? c.y47.x : fill; c.y47.y = n > 5 ? c.y47.y : fill;
    c.y47.z = n > 6 ? c.y47.z : fill; c.y47.w = n > 7 ? c.y47.w : fill;
}

// pending shiftCells(f, d) of the previous sweep (V2 shiftCells.h:23-112) for one destination
// cell, written straight into the staged tile.  `own`/`up` must be sanitised so that unused
// slots can never pass the respective test (own: sentinel -> never a stayer; up: a value
// that stays -> never an immigrant).  pf points at slot 0 of the f-coordinate plane of the
// destination cell, OFF is the float offset from there to the other coordinate's plane.
template <int F, int PL>
__device__ __forceinline__ int shift_into_tile(const CellRegs &own, const CellRegs &up, float d,
                                               float w, float sshift, float *pf, int *dropped)
{
    constexpr int PLANE = PL * 4;                       // floats between the 0-3 and 4-7 planes
    constexpr int OFF = (F == 0) ? 2 * PLANE : -2 * PLANE;
    int n = 0, drop = 0;
#pragma unroll
    for (int i = 0; i < PMC_NMAX; i++) {
        const float fc = F == 0 ? f4get(own.x03, own.x47, i) : f4get(own.y03, own.y47, i);
        const float oc = F == 0 ? f4get(own.y03, own.y47, i) : f4get(own.x03, own.x47, i);
        const float D = __fadd_rn(fc, -d);
        if (D > 0.0f && D <= w) {                       // shiftCells.h:62 (sentinel slots fail D <= w)
            float *p = pf + n + (n >> 2) * (PLANE - 4);
            p[0] = D; p[OFF] = oc;
            n++;
        }
    }
#pragma unroll
    for (int i = 0; i < PMC_NMAX; i++) {
        const float fc = F == 0 ? f4get(up.x03, up.x47, i) : f4get(up.y03, up.y47, i);
        const float oc = F == 0 ? f4get(up.y03, up.y47, i) : f4get(up.x03, up.x47, i);
        const float D = __fadd_rn(fc, -d);
        if (!(D > 0.0f && D <= w)) {                    // shiftCells.h:94
            if (n < PMC_NMAX) {
                float *p = pf + n + (n >> 2) * (PLANE - 4);
                p[0] = __fadd_rn(D, sshift); p[OFF] = oc;   // shiftCells.h:97
                n++;
            } else drop++;
        }
    }
    *dropped = drop;
    return n;
}

// two slots per instruction: d2 = (q.x + npx)^2 + (q.y + npy)^2
__device__ __forceinline__ float2 pair2_d2(float qx0, float qx1, float qy0, float qy1,
                                           float2 npx, float2 npy)
{
    const float2 dx = __fadd2_rn(make_float2(qx0, qx1), npx);
    const float2 dy = __fadd2_rn(make_float2(qy0, qy1), npy);
    const float2 t = __fmul2_rn(dy, dy);
    return __ffma2_rn(dx, dx, t);
}

// smallest squared distance between the trial point (in this cell's frame, negated:
// npx = -pxs) and the 8 slots of one staged cell.  Unused slots hold the sentinel (1e36).
template <int PL>
__device__ __forceinline__ float cell_min_d2(const float4 *cellp, float npx, float npy)
{
    const float4 x03 = cellp[0], x47 = cellp[PL], y03 = cellp[2 * PL], y47 = cellp[3 * PL];
    const float2 nx = make_float2(npx, npx), ny = make_float2(npy, npy);
    const float2 a = pair2_d2(x03.x, x03.y, y03.x, y03.y, nx, ny);
    const float2 b = pair2_d2(x03.z, x03.w, y03.z, y03.w, nx, ny);
    const float2 c = pair2_d2(x47.x, x47.y, y47.x, y47.y, nx, ny);
    const float2 e = pair2_d2(x47.z, x47.w, y47.z, y47.w, nx, ny);
    return fminf(fminf(fminf(a.x, a.y), fminf(b.x, b.y)), fminf(fminf(c.x, c.y), fminf(e.x, e.y)));
}

// d2 of the trial point against two slots held in registers
__device__ __forceinline__ float2 reg2_d2(float qx0, float qx1, float qy0, float qy1, float2 npx, float2 npy)
{
    return pair2_d2(qx0, qx1, qy0, qy1, npx, npy);
}

// NM: compile-time n_M of the register fast path (4), or 0 for the generic path (any n_M)
template <int NCOL, int TX, int TY, int THREADS, int MINB, int NM>
__global__ void __launch_bounds__(THREADS, MINB)
sweep_tile_kernel(const float4 *din, const int16_t *nin, float4 *dout, int16_t *nout,
                  const DevGeom g, const SweepArgs a, Counters *ctr)
{
    using TL = Tile<NCOL, TX, TY>;
    constexpr int H = TL::H, RX = TL::RX, RY = TL::RY, PITCH = TL::PITCH, HALF = TL::HALF, PL = TL::PL;
    constexpr int NAX = TL::NAX, NAY = TL::NAY;
    static_assert(NAX * NAY <= THREADS, "one thread per active cell");
    extern __shared__ float4 sm[];
    unsigned char *scnt = reinterpret_cast<unsigned char *>(sm + 4 * PL);
    auto SID = [](int i, int j) { return sidx<PITCH, HALF>(i, j); };

    const int tid = threadIdx.x;
    const int cps = g.cps;
    const float w = g.w;
    const bool small = cps < (RX > RY ? RX : RY) + 2;   // region may wrap more than once

    // pending shiftCells(f, d) of the previous sweep
    const bool do_shift = (NCOL != 1) && a.shift_on && !PMC_DBG_BIT(a, 2);
    const int sdir = (a.shift_d <= 0.0f) ? -1 : 1;                       // shiftCells.h:38-44
    const int sdx = (do_shift && a.shift_f == 0) ? sdir : 0, sdy = (do_shift && a.shift_f == 1) ? sdir : 0;
    // staged area = region + one upstream row / column; (xoff, yoff) = staged coords of region (0, 0)
    const int xoff = sdx < 0 ? 1 : 0, yoff = sdy < 0 ? 1 : 0;
    const int nsx = RX + (sdx != 0), nsy = RY + (sdy != 0);
    // unwrapped global column / owned-relative row of staged (0, 0)
    const int ux0 = blockIdx.x * TX - H - xoff;
    const int uy0 = blockIdx.y * TY - H - yoff;

    // ------------------------------------------------------------ phase 0a: raw tile -> shared memory
    {
        constexpr int SX = TL::SX, SY = TL::SY, SXY = SX * SY;
        // (i, j) staged coordinates -> global cell index, or -1 (do not touch) / -2 (empty cell)
        auto gcell = [&](int i, int j) -> int {
            if (i >= nsx || j >= nsy) return -1;
            const int ux = ux0 + i, uy = uy0 + j;
            if (NCOL == 1) {
                // in-place mode: ring cells of the active colour belong to other CTAs, may be
                // written concurrently and are never read by this CTA -> do not touch them
                const bool ring = (i < H) | (i >= H + TX) | (j < H) | (j >= H + TY);
                if (ring && ((ux & 1) == a.offx[0]) && (((g.row0 + uy) & 1) == a.offy[0])) return -1;
            }
            const int gx = wrap_fast(ux, cps, small);
            int lr;
            if (g.wrap_y) lr = wrap_fast(uy, cps, small);
            else {
                lr = uy + g.ghost;
                // rows outside this slab's storage read as empty cells (they can only influence
                // cells this CTA does not own)
                if (lr < 0 || lr >= g.local_rows) return -2;
            }
            return lr * cps + gx;          // < 2^31 (cps <= 46340)
        };
        // particle data: 16-byte cp.async, lane -> (cell, plane): 512 contiguous bytes per warp
        // instruction.  A thread keeps its plane and walks the staged cells with a fixed stride,
        // so (i, j) and the global row are updated incrementally (no divisions in the loop).
        {
            static_assert(THREADS % 4 == 0, "four lanes per cell");
            constexpr int CSTEP = THREADS / 4, DJ = CSTEP / SX, DI = CSTEP % SX;
            const int plane = tid & 3;
            int c = tid >> 2;
            int j = c / SX, i = c - j * SX;
            float4 *splane = sm + plane * PL;
            const float4 fillv = plane < 2 ? make_float4(kSent, kSent, kSent, kSent) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
            for (; c < SXY; c += CSTEP) {
                const int cell = gcell(i, j);
                float4 *dst = splane + SID(i, j);
                if (cell >= 0) cp_async16(dst, din + (long long)cell * 4 + plane);
                else if (cell == -2) *dst = fillv;
                i += DI; j += DJ;
                if (i >= SX) { i -= SX; j++; }
            }
        }
        // counts: all loads of a thread in flight together
        constexpr int CPT = (SXY + THREADS - 1) / THREADS;
        int cn[CPT];
#pragma unroll
        for (int u = 0; u < CPT; u++) {
            const int c = tid + u * THREADS;
            const int j = c / SX, i = c - j * SX;
            const int cell = (c < SXY) ? gcell(i, j) : -1;
            cn[u] = cell >= 0 ? (int)__ldg(nin + cell) : cell;
        }
#pragma unroll
        for (int u = 0; u < CPT; u++) {
            const int c = tid + u * THREADS;
            const int j = c / SX, i = c - j * SX;
            if (cn[u] == -1) continue;
            const int n = cn[u];
            scnt[SID(i, j)] = (unsigned char)(n < 0 ? 0 : (n > PMC_NMAX ? PMC_NMAX : n));
        }
        cp_async_wait_all();
        __syncthreads();
    }

    // ------------------------------------------------------------ phase 0b: sanitise / pending shift, in place
    if (!do_shift) {
        if (a.sanitize_in) {
#pragma unroll 1
            for (int c = tid; c < RX * RY; c += THREADS) {
                const int j = c / RX, i = c - j * RX;
                if (NCOL == 1) {
                    const bool ring = (i < H) | (i >= H + TX) | (j < H) | (j >= H + TY);
                    if (ring && (((ux0 + i) & 1) == a.offx[0]) && (((g.row0 + uy0 + j) & 1) == a.offy[0])) continue;
                }
                const int sid = SID(i, j);
                CellRegs r;
                r.x03 = sm[sid]; r.x47 = sm[sid + PL]; r.cnt = scnt[sid];
                sanitize_x(r, kSent);
                sm[sid] = r.x03; sm[sid + PL] = r.x47;
            }
            __syncthreads();
        }
    } else {
        const float sshift = __fmul_rn(w, (float)sdir);                     // shiftCells.h:84-86
        const float stay_fill = __fadd_rn(a.shift_d, __fmul_rn(0.5f, w));   // D = w/2: never an immigrant
        // One thread owns a strip of consecutive cells along the shift axis and walks it from
        // the downstream end to the upstream end: cell k is rewritten only after raw cell k+1
        // has been read.  The one raw cell a strip needs from the next strip (owned by another
        // thread) is read before the barrier, so a single __syncthreads makes the in-place
        // update race-free.
        constexpr int SEGY = THREADS / RX, KY = (RY + SEGY - 1) / SEGY;     // f = 1: strips of KY rows in one column
        constexpr int SEGX = THREADS / RY, KX = (RX + SEGX - 1) / SEGX;     // f = 0: strips of KX columns in one row
        int i0, j0, len, di, dj;        // strip start (region coords), length, step = upstream direction
        bool act;
        if (a.shift_f == 1) {
            const int seg = tid / RX;
            i0 = tid - seg * RX;
            const int k0 = seg * KY;
            act = (seg < SEGY) && (k0 < RY);
            len = act ? (RY - k0 < KY ? RY - k0 : KY) : 0;
            j0 = sdir > 0 ? k0 : RY - 1 - k0;
            di = 0; dj = sdir;
        } else {
            const int row = tid / SEGX, seg = tid - row * SEGX;
            j0 = row;
            const int k0 = seg * KX;
            act = (row < RY) && (k0 < RX);
            len = act ? (RX - k0 < KX ? RX - k0 : KX) : 0;
            i0 = sdir > 0 ? k0 : RX - 1 - k0;
            di = sdir; dj = 0;
        }
        auto load_staged = [&](int i, int j, CellRegs &c) {     // region coords (may be the extra upstream row / column)
            const int sd = SID(i + xoff, j + yoff);
            c.x03 = sm[sd]; c.x47 = sm[sd + PL]; c.y03 = sm[sd + 2 * PL]; c.y47 = sm[sd + 3 * PL];
            c.cnt = scnt[sd];
        };
        CellRegs cur, edge;
        if (act) {
            load_staged(i0, j0, cur);
            load_staged(i0 + len * di, j0 + len * dj, edge);    // first raw cell of the next strip
        }
        __syncthreads();
#pragma unroll 1
        for (int k = 0; k < len; k++) {
            const int i = i0 + k * di, j = j0 + k * dj;
            CellRegs up;
            if (k + 1 < len) load_staged(i + di, j + dj, up);
            else up = edge;
            const CellRegs nxt = up;
            const int sid = SID(i + xoff, j + yoff);
            sm[sid] = make_float4(kSent, kSent, kSent, kSent);
            sm[sid + PL] = make_float4(kSent, kSent, kSent, kSent);
            sm[sid + 2 * PL] = make_float4(0.f, 0.f, 0.f, 0.f);
            sm[sid + 3 * PL] = make_float4(0.f, 0.f, 0.f, 0.f);
            float *base = reinterpret_cast<float *>(sm + sid);
            int dropped, nNew;
            if (a.shift_f == 0) {
                sanitize_x(cur, kSent); sanitize_x(up, stay_fill);
                nNew = shift_into_tile<0, PL>(cur, up, a.shift_d, w, sshift, base, &dropped);
            } else {
                sanitize_y(cur, kSent); sanitize_y(up, stay_fill);
                nNew = shift_into_tile<1, PL>(cur, up, a.shift_d, w, sshift, base + 2 * PL * 4, &dropped);
            }
            scnt[sid] = (unsigned char)nNew;
            if (dropped) {
                atomicOr(&ctr->status, PMC_STATUS_OVERFLOW);
                const int ux = ux0 + xoff + i, uy = uy0 + yoff + j;
                const bool owned = (i >= H) & (i < H + TX) & (j >= H) & (j < H + TY) & (ux < cps) & (uy < g.rows);
                if (owned) atomicAdd(&ctr->lost, (unsigned long long)dropped);
            }
            cur = nxt;
        }
        __syncthreads();
    }
    // from here on everything is in region coordinates; staged (i, j) = region (i, j) + (xoff, yoff)
    const int rx0 = ux0 + xoff, ry0 = uy0 + yoff;   // unwrapped global column / owned-relative row of region (0, 0)

    // ------------------------------------------------------------ the sub-sweeps
    unsigned my_trials = 0, my_acc = 0;
    const float sigma = g.sigma, sigma2 = g.sigma2;
    // Gaussian: m * q + x (m from gauss_disp, in grid units); uniform: fmaf(grid_disp, dstep, x - doff)
    const bool gauss_p = g.proposal == PMC_PROPOSAL_GAUSSIAN;
    const float dscale = gauss_p ? g.dscale : g.dstep, doff = gauss_p ? 0.0f : g.doff;
    const int bq = tid / NAX, aq = tid - bq * NAX;  // fixed thread -> (column, row) of the active lattice

#pragma unroll 1
    for (int k = 0; k < (PMC_DBG_BIT(a, 1) ? 0 : NCOL); k++) {
        const int lo = (NCOL == 1) ? H : k + 1;     // cells closer than lo to the region edge are stale
        const int pi = ((int)((a.offmask >> (2 * k)) & 1u) - rx0) & 1;       // region-column parity of the active colour
        const int pj = ((int)((a.offmask >> (2 * k + 1)) & 1u) - (g.row0 + ry0)) & 1;
        const int i_first = lo + ((pi - lo) & 1), j_first = lo + ((pj - lo) & 1);
        const int na = (RX - lo - i_first + 1) >> 1, nb = (RY - lo - j_first + 1) >> 1;
        const int i = i_first + 2 * aq, j = j_first + 2 * bq;
        const int sid = SID(i + xoff, j + yoff);
        const int cnt = (aq < na && bq < nb) ? (int)scnt[sid] : 0;
        if (cnt != 0) {                             // subsweep.h:252-254
            const int ux = rx0 + i, uy = ry0 + j;
            const bool owned = (i >= H) & (i < H + TX) & (j >= H) & (j < H + TY) & (ux < cps) & (uy < g.rows);
            const uint32_t cell_id = (uint32_t)wrap_fast(g.row0 + uy, cps, small) * (uint32_t)cps +
                                     (uint32_t)wrap_fast(ux, cps, small);
            const int sidL = SID(i + xoff - 1, j + yoff);
            const int sidR = SID(i + xoff + 1, j + yoff);
            float4 *pown = sm + sid;

            // neighbour part of one trial: smallest d2 against the needed neighbour cells, or
            // a negative value when the proposal leaves the cell (out_of_bound subsweep.h:73-88)
            auto neighbours_min_d2 = [&](const float px, const float py) -> float {
                if (!(px > 0.0f && px <= w && py > 0.0f && py <= w)) return -1.0f;
                // which neighbour columns / rows can hold a disk closer than sigma?  Exact
                // conservative tests (monotonicity of IEEE rounding): a skipped cell could
                // not have produced d2 < sigma2 in the oracle's arithmetic.
                const float pxl = __fadd_rn(px, w), pxr = __fadd_rn(px, -w);
                const float pyd = __fadd_rn(py, w), pyu = __fadd_rn(py, -w);
                const bool needL = !(__fadd_rn(pxl, -w) >= sigma), needR = !(pxr <= -sigma);
                const bool needD = !(__fadd_rn(pyd, -w) >= sigma), needU = !(pyu <= -sigma);
                float m;
                if (!((needL & needR) | (needD & needU))) {
                    // fast path (always taken when w >= 2 sigma): at most 3 neighbour cells
                    const float npxH = needL ? -pxl : (needR ? -pxr : kSent);
                    const float npyV = needD ? -pyd : (needU ? -pyu : kSent);
                    const float4 *pH = sm + (needL ? sidL : sidR);
                    const int dV = needD ? -PITCH : PITCH;
                    m = cell_min_d2<PL>(pH, npxH, -py);
                    m = fminf(m, cell_min_d2<PL>(pown + dV, -px, npyV));
                    m = fminf(m, cell_min_d2<PL>(pH + dV, npxH, npyV));
                } else {
                    // generic path (w < 2 sigma): every needed neighbour of the 3x3 block
                    m = 3.0e38f;
#pragma unroll 1
                    for (int dj = -1; dj <= 1; dj++) {
                        if ((dj < 0 && !needD) || (dj > 0 && !needU)) continue;
                        const float npy = dj < 0 ? -pyd : (dj > 0 ? -pyu : -py);
#pragma unroll 1
                        for (int di = -1; di <= 1; di++) {
                            if ((di < 0 && !needL) || (di > 0 && !needR) || (di == 0 && dj == 0)) continue;
                            const float npx = di < 0 ? -pxl : (di > 0 ? -pxr : -px);
                            const int s2 = (di < 0 ? sidL : (di > 0 ? sidR : sid)) + dj * PITCH;
                            m = fminf(m, cell_min_d2<PL>(sm + s2, npx, npy));
                        }
                    }
                }
                return m;
            };

            if (NM == 4) {
                // ---------------- register fast path, n_M == 4 ----------------
                // own cell lives in registers for the whole sub-sweep: no shared-memory
                // traffic for the moving disk, its own-cell test or the commit
                const float4 x03 = pown[0], x47 = pown[PL], y03 = pown[2 * PL], y47 = pown[3 * PL];
                float ox[8] = { x03.x, x03.y, x03.z, x03.w, x47.x, x47.y, x47.z, x47.w };
                float oy[8] = { y03.x, y03.y, y03.z, y03.w, y47.x, y47.y, y47.z, y47.w };
                // uniform proposal: one word per trial (one Philox call); Gaussian: two words per trial (two calls)
                const bool gauss = g.proposal == PMC_PROPOSAL_GAUSSIAN;
                uint32_t rw[8];
                philox4x32_10(cell_id, a.sweep_lo, a.sweep_hi, 0u, g.seed_lo, g.seed_hi, rw[0], rw[1], rw[2], rw[3]);
                if (gauss) philox4x32_10(cell_id, a.sweep_lo, a.sweep_hi, 1u, g.seed_lo, g.seed_hi, rw[4], rw[5], rw[6], rw[7]);
                // random_shuffle subsweep.h:50-58: physical partial Fisher-Yates, steps 0..3
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    const uint32_t b16 = ((rw[2 * s] & 0xFFu) << 8) | (rw[2 * s + 1] & 0xFFu);
                    const int mrem = cnt > s ? cnt - s : 1;             // s >= cnt: no-op (jj == s)
                    const int jj = s + (gauss ? (int)((b16 * (uint32_t)mrem) >> 16)
                                              : (int)(((rw[s] >> 24) * (uint32_t)mrem) >> 8));
                    const float tx = ox[s], ty = oy[s];
                    float nx = tx, ny = ty;
#pragma unroll
                    for (int q = s + 1; q < 8; q++) {
                        const bool p = (jj == q);
                        nx = p ? ox[q] : nx; ny = p ? oy[q] : ny;
                        ox[q] = p ? tx : ox[q]; oy[q] = p ? ty : oy[q];
                    }
                    ox[s] = nx; oy[s] = ny;
                }
                // trials 0..3 move slot s mod cnt (subsweep.h:279-297), all indices static
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    // slot = s mod cnt: s < cnt -> s, else one of the earlier slots
                    const bool cA = cnt > s;                            // slot == s
                    const bool cB = (s == 3) && (cnt == 2);             // slot == 1 (only s = 3, cnt = 2)
                    const float x = cA ? ox[s] : (cB ? ox[1] : ox[0]);
                    const float y = cA ? oy[s] : (cB ? oy[1] : oy[0]);
                    float mx = grid_disp_x(rw[s]), my = grid_disp_y(rw[s]);
                    if (gauss) gauss_disp(rw[2 * s], rw[2 * s + 1], g.M, mx, my);
                    const float px = __fmaf_rn(mx, dscale, __fadd_rn(x, -doff));     // make_move subsweep.h:60-71
                    const float py = __fmaf_rn(my, dscale, __fadd_rn(y, -doff));
                    my_trials += owned ? 1u : 0u;
                    float m = neighbours_min_d2(px, py);
                    // own cell (calculate_energy_in_cell subsweep.h:105-117), j != slot
                    const float2 npx = make_float2(-px, -px), npy = make_float2(-py, -py);
                    float2 d01 = reg2_d2(ox[0], ox[1], oy[0], oy[1], npx, npy);
                    float2 d23 = reg2_d2(ox[2], ox[3], oy[2], oy[3], npx, npy);
                    const float2 d45 = reg2_d2(ox[4], ox[5], oy[4], oy[5], npx, npy);
                    const float2 d67 = reg2_d2(ox[6], ox[7], oy[6], oy[7], npx, npy);
                    const float big = 3.0e38f;
                    if (s == 0) d01.x = big;
                    if (s == 1) { d01.y = cA ? big : d01.y; d01.x = cA ? d01.x : big; }
                    if (s == 2) { d23.x = cA ? big : d23.x; d01.x = cA ? d01.x : big; }
                    if (s == 3) { d23.y = cA ? big : d23.y; d01.y = cB ? big : d01.y; d01.x = (cA | cB) ? d01.x : big; }
                    m = fminf(m, fminf(fminf(fminf(d01.x, d01.y), fminf(d23.x, d23.y)),
                                       fminf(fminf(d45.x, d45.y), fminf(d67.x, d67.y))));
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
                pown[PL] = make_float4(ox[4], ox[5], ox[6], ox[7]);
                pown[2 * PL] = make_float4(oy[0], oy[1], oy[2], oy[3]);
                pown[3 * PL] = make_float4(oy[4], oy[5], oy[6], oy[7]);
            } else {
                // ---------------- generic path, any n_M: own cell stays in shared memory ----------------
                const int n_M = g.n_M;
                float *fown = reinterpret_cast<float *>(pown);
                auto slot_ptr = [&](int slot) { return fown + (slot >> 2) * (PL * 4) + (slot & 3); };
                const int steps = n_M < cnt ? n_M : cnt;
                const bool gauss = g.proposal == PMC_PROPOSAL_GAUSSIAN;
                const int per_call = gauss ? 2 : 4;         // trials fed by one Philox call (oracle subsweep_cell)
#pragma unroll 1
                for (int s0 = 0; s0 < steps; s0 += per_call) {     // shuffle first (needs every call's low bytes)
                    uint32_t r[4];
                    philox4x32_10(cell_id, a.sweep_lo, a.sweep_hi, (uint32_t)(s0 / per_call), g.seed_lo, g.seed_hi, r[0], r[1], r[2], r[3]);
#pragma unroll 1
                    for (int h = 0; h < per_call && s0 + h < steps; h++) {
                        const int s = s0 + h;
                        const uint32_t ra = gauss ? (h ? r[2] : r[0]) : (h == 0 ? r[0] : h == 1 ? r[1] : h == 2 ? r[2] : r[3]);
                        const uint32_t rb = h ? r[3] : r[1];
                        const uint32_t b16 = ((ra & 0xFFu) << 8) | (rb & 0xFFu);
                        const int jj = s + (gauss ? (int)((b16 * (uint32_t)(cnt - s)) >> 16)
                                                  : (int)(((ra >> 24) * (uint32_t)(cnt - s)) >> 8));
                        float *ps = slot_ptr(s), *pj = slot_ptr(jj);
                        const float xs = ps[0], ys = ps[2 * PL * 4], xj = pj[0], yj = pj[2 * PL * 4];
                        ps[0] = xj; ps[2 * PL * 4] = yj; pj[0] = xs; pj[2 * PL * 4] = ys;
                    }
                }
                int it = 0;                                 // i of subsweep.h:278,291-296
#pragma unroll 1
                for (int s = 0; s < n_M; s++) {             // subsweep.h:279
                    uint32_t r0, r1, r2, r3;
                    philox4x32_10(cell_id, a.sweep_lo, a.sweep_hi, (uint32_t)(s / per_call), g.seed_lo, g.seed_hi, r0, r1, r2, r3);
                    const int h = s % per_call;
                    const uint32_t ra = gauss ? (h ? r2 : r0) : (h == 0 ? r0 : h == 1 ? r1 : h == 2 ? r2 : r3);
                    const uint32_t rb = h ? r3 : r1;
                    float *fx = slot_ptr(it), *fy = fx + 2 * PL * 4;
                    it = (it + 1 >= cnt) ? 0 : it + 1;
                    const float x = *fx, y = *fy;
                    float mx = grid_disp_x(ra), my = grid_disp_y(ra);
                    if (gauss) gauss_disp(ra, rb, g.M, mx, my);
                    const float px = __fmaf_rn(mx, dscale, __fadd_rn(x, -doff));
                    const float py = __fmaf_rn(my, dscale, __fadd_rn(y, -doff));
                    my_trials += owned ? 1u : 0u;
                    float m = neighbours_min_d2(px, py);
                    if (m >= 0.0f) {
                        *fx = kSent;                        // hide the moving disk from its own cell test
                        m = fminf(m, cell_min_d2<PL>(pown, -px, -py));
                    }
                    const bool acc = !(m < sigma2);
                    *fx = acc ? px : x;
                    if (acc) { *fy = py; my_acc += owned ? 1u : 0u; }
                }
            }
        }
        __syncthreads();
    }

    // ------------------------------------------------------------ phase 5: owned tile -> HBM
    if (!PMC_DBG_BIT(a, 4)) {
        constexpr int CSTEP = THREADS / 4, DJ = CSTEP / TX, DI = CSTEP % TX;
        const int plane = tid & 3;
        int c = tid >> 2;
        int jj = c / TX, ii = c - jj * TX;
        const float4 *splane = sm + plane * PL;
#pragma unroll 1
        for (; c < TX * TY; c += CSTEP) {
            const int i = H + ii, j = H + jj;
            const int ux = rx0 + i, uy = ry0 + j;
            const bool skip = (ux >= cps) | (uy >= g.rows) |
                              ((NCOL == 1) && (((ux & 1) != a.offx[0]) | (((g.row0 + uy) & 1) != a.offy[0])));
            if (!skip) {
                const int sid = SID(i + xoff, j + yoff);
                const int cell = (uy + g.ghost) * cps + ux;
                dout[(long long)cell * 4 + plane] = splane[sid];
                if (NCOL != 1 && plane == 0) nout[cell] = (int16_t)scnt[sid];
            }
            ii += DI; jj += DJ;
            if (ii >= TX) { ii -= TX; jj++; }
        }
    }

    // acceptance counts reduced warp-level, one atomic per warp (kernel.cu:228,413 accept_counter)
    my_trials = __reduce_add_sync(0xffffffffu, my_trials);
    my_acc = __reduce_add_sync(0xffffffffu, my_acc);
    if ((tid & 31) == 0 && my_trials) {
        atomicAdd(&ctr->trials, (unsigned long long)my_trials);
        atomicAdd(&ctr->accepted, (unsigned long long)my_acc);
    }
}

// single colour (in place): 32 x 32 tile, 16 x 16 active cells
constexpr int kT1 = 32, kThreads1 = 256, kMinB1 = 2;
// fused sweep: 26 x 32 tile -> 34 x 40 staged region (88 KB, 2 CTAs / SM), 16 x 19 active cells
constexpr int kTX4 = 26, kTY4 = 32, kThreads4 = 320, kMinB4 = 2;

template <typename K>
cudaError_t set_smem(K kernel, size_t bytes)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace

int pmc_fused_launch_count() { return 1; }

cudaError_t pmc_launch_subsweep(const DevGeom &g, float4 *disk, const int16_t *n,
                                const SweepArgs &a, Counters *ctr, cudaStream_t st)
{
    constexpr size_t smem = Tile<1, kT1, kT1>::SMEM;
    dim3 grid((g.cps + kT1 - 1) / kT1, (g.rows + kT1 - 1) / kT1);
    cudaError_t e;
    if (g.n_M == 4) {
        auto kern = sweep_tile_kernel<1, kT1, kT1, kThreads1, kMinB1, 4>;
        if ((e = set_smem(kern, smem)) != cudaSuccess) return e;
        kern<<<grid, kThreads1, smem, st>>>(disk, n, disk, nullptr, g, a, ctr);
    } else {
        auto kern = sweep_tile_kernel<1, kT1, kT1, kThreads1, kMinB1, 0>;
        if ((e = set_smem(kern, smem)) != cudaSuccess) return e;
        kern<<<grid, kThreads1, smem, st>>>(disk, n, disk, nullptr, g, a, ctr);
    }
    return cudaGetLastError();
}

template <int TX, int TY, int THREADS, int MINB>
cudaError_t launch_fused_cfg(const DevGeom &g, const float4 *din, const int16_t *nin,
                             float4 *dout, int16_t *nout, const SweepArgs &a,
                             Counters *ctr, cudaStream_t st)
{
    constexpr size_t smem = Tile<4, TX, TY>::SMEM;
    dim3 grid((g.cps + TX - 1) / TX, (g.rows + TY - 1) / TY);
    cudaError_t e;
    if (g.n_M == 4) {
        auto kern = sweep_tile_kernel<4, TX, TY, THREADS, MINB, 4>;
        if ((e = set_smem(kern, smem)) != cudaSuccess) return e;
        kern<<<grid, THREADS, smem, st>>>(din, nin, dout, nout, g, a, ctr);
    } else {
        auto kern = sweep_tile_kernel<4, TX, TY, THREADS, MINB, 0>;
        if ((e = set_smem(kern, smem)) != cudaSuccess) return e;
        kern<<<grid, THREADS, smem, st>>>(din, nin, dout, nout, g, a, ctr);
    }
    return cudaGetLastError();
}

cudaError_t pmc_launch_fused_sweep(const DevGeom &g, const float4 *din, const int16_t *nin,
                                   float4 *dout, int16_t *nout, const SweepArgs &a,
                                   Counters *ctr, cudaStream_t st)
{
    // 26 x 32 tile, 2 CTAs/SM (the 26 x 20 / 3 CTAs and 26 x 38 / 2 CTAs variants measured slower; round 1)
    return launch_fused_cfg<kTX4, kTY4, kThreads4, kMinB4>(g, din, nin, dout, nout, a, ctr, st);
}
