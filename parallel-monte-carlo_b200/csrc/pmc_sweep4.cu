// pmc_sweep4.cu -- the fused MC sweep (start.cu:237-260 loop body: 4 x subsweep_kernel
// subsweep.h:240-300, then shiftCells shiftCells.h:23-112) as ONE sm_100a kernel per sweep.
// This is the throughput path behind pmc_sweep(); pmc_sweep.cu keeps the generic kernel
// (any n_M, w < 2 sigma, tiny boxes, the single-colour call site).
//
// Differences from the generic kernel, all chosen to cut issued instructions per trial:
//   * INTERNAL STATE LAYOUT (handle-owned, never seen by the caller): float4 chunks
//         chunk(X, Y, plane) = ((Y*4 + plane)*2 + (X & 1)) * CH + (X >> 1)
//     planes = x slots 0-3, x slots 4-7, y slots 0-3, y slots 4-7; even and odd columns are
//     split so that the same-colour cells a warp works on are contiguous (conflict-free
//     LDS.128) and so that ONE 4-D TMA box per plane lands a tile in shared memory in
//     exactly the layout the sub-sweeps read: no per-cell index arithmetic at all.
//     The box is surrounded by margins (kMX columns, kMY rows) holding periodic images, written
//     by the CTA that produces the original cell (single GPU) or by the NCCL ring (slab rows),
//     so a tile never wraps.
//   * the cell count lives in-band: unused slots have x = sentinel; when a cell holds fewer
//     than 8 particles the bits of its y[7] are the count.  No count array on the hot path.
//   * the grid shift of THIS sweep is applied while the tile leaves shared memory (the tile
//     carries one extra upstream row / column), so only owned cells are re-binned (the
//     generic kernel re-bins its whole halo while staging).
//   * neighbour cells needed by a trial: with w >= 2 sigma a proposal in the left half of its
//     cell can only touch the left column of neighbours, etc.: one compare per axis.
// Every random number is a pure function of (seed, sweep, global cell id, trial), so the
// redundantly recomputed halo cells get the same bits in every CTA and on every GPU.
#include "pmc_internal.cuh"
#include <cuda.h>      // CUtensorMap type only; the encoder comes from cudaGetDriverEntryPoint
#include <stdlib.h>

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
    static constexpr int NAX = (TX + 2 * H) / 2, NAY = (TY + 2 * H) / 2;   // active cells per colour
    static constexpr int THREADS = NAX * NAY;
    static constexpr size_t SMEM = (size_t)4 * PLC * 16 + 16;
    static_assert(TX % 4 == 0 && TY % 2 == 0, "owned runs must be 32-byte aligned in HBM");
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

__device__ __forceinline__ int decode_cnt(float x7, float y7) { return x7 < kSentTest ? 8 : __float_as_int(y7); }

// two slots per instruction (Blackwell packed FP32): d2 = (q.x + npx)^2 + (q.y + npy)^2, the
// oracle's fmaf(dx, dx, dy*dy) with dx = pxs - qx (sign is irrelevant after squaring)
__device__ __forceinline__ float2 pair2(float qx0, float qx1, float qy0, float qy1, float2 npx, float2 npy)
{
    const float2 dx = __fadd2_rn(make_float2(qx0, qx1), npx);
    const float2 dy = __fadd2_rn(make_float2(qy0, qy1), npy);
    return __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
}

template <int PLC>
__device__ __forceinline__ float cell_min_d2(const float4 *cp, float npx, float npy)
{
    const float4 x03 = cp[0], x47 = cp[PLC], y03 = cp[2 * PLC], y47 = cp[3 * PLC];
    const float2 nx = make_float2(npx, npx), ny = make_float2(npy, npy);
    const float2 a = pair2(x03.x, x03.y, y03.x, y03.y, nx, ny);
    const float2 b = pair2(x03.z, x03.w, y03.z, y03.w, nx, ny);
    const float2 c = pair2(x47.x, x47.y, y47.x, y47.y, nx, ny);
    const float2 e = pair2(x47.z, x47.w, y47.z, y47.w, nx, ny);
    return fminf(fminf(fminf(a.x, a.y), fminf(b.x, b.y)), fminf(fminf(c.x, c.y), fminf(e.x, e.y)));
}

// +-(odd integer < 2^24) as a float without I2F: 0x4B800000 | m23 is the float 2^24 + 2*m23
__device__ __forceinline__ float signed_odd24(uint32_t r)
{
    const float a = __uint_as_float(((r >> 8) & 0x7FFFFFu) | 0x4B800000u);
    const float mag = __fadd_rn(a, -16777215.0f);
    return __uint_as_float(__float_as_uint(mag) | (r & 0x80000000u));
}

// V2 shiftCells.h:23-112 for one destination cell, written into the staged tile in place.
// fx points at x slot 0 of the destination cell.
template <int F, int PLC>
__device__ __forceinline__ int shift_into_tile(const CellRegs &own, const CellRegs &up, float d, float w,
                                               float sshift, float *fx, int *dropped)
{
    constexpr int PF = PLC * 4;                         // floats between consecutive planes
    float *pf = F == 0 ? fx : fx + 2 * PF;              // f-coordinate plane (slots 0-3)
    constexpr int OFF = F == 0 ? 2 * PF : -2 * PF;      // to the other coordinate
    int n = 0, drop = 0;
#pragma unroll
    for (int i = 0; i < PMC_NMAX; i++) {
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
    for (int i = 0; i < PMC_NMAX; i++) {
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

template <int TX, int TY, int MINB>
__global__ void __launch_bounds__(Tile4<TX, TY>::THREADS, MINB)
sweep4_kernel(const __grid_constant__ CUtensorMap tmap, float4 *__restrict__ dout, const Geom4 g,
              const SweepArgs a, Counters *ctr)
{
    using TL = Tile4<TX, TY>;
    constexpr int H = TL::H, HB = TL::HB, PITCH = TL::PITCH, PLC = TL::PLC, NAX = TL::NAX, THREADS = TL::THREADS;
    extern __shared__ __align__(128) float4 sm[];
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sm + 4 * PLC);

    const int tid = threadIdx.x;
    const int cps = g.cps;
    const float w = g.w;

    // this sweep's grid shift: the tile carries one extra row / column on the upstream side
    const bool do_shift = a.shift_on && !(a.dbg_skip & 2);
    const int sdir = (a.shift_d <= 0.0f) ? -1 : 1;                       // shiftCells.h:38-44
    const int exl = (do_shift && a.shift_f == 0 && sdir < 0), exh = (do_shift && a.shift_f == 0 && sdir > 0);
    const int eyl = (do_shift && a.shift_f == 1 && sdir < 0), eyh = (do_shift && a.shift_f == 1 && sdir > 0);
    const int RX = TX + 2 * H + exl + exh, RY = TY + 2 * H + eyl + eyh;  // region the sub-sweeps work on
    const int rx0 = blockIdx.x * TX - H - exl;      // unwrapped global column of region column 0
    const int ry0 = blockIdx.y * TY - H - eyl;      // owned-relative row of region row 0
    const int X0 = rx0 + kMX, Y0 = ry0 + kMY;       // the same in internal array coordinates (>= 0)
    const int xs = X0 & 1;                          // region column i is staged column i + xs

    // ------------------------------------------------------------ stage the tile: 4 TMA boxes
    if (tid == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(mbar, (unsigned)(4 * TL::PLB * 16));
#pragma unroll
        for (int p = 0; p < 4; p++) tma_load_4d(sm + p * PLC, &tmap, 4 * ((X0 - xs) >> 1), 0, p, Y0, mbar);
    }
    if (tid == 0) mbar_wait(mbar, 0);   // one poller; the others observe the completed phase once
    __syncthreads();
    mbar_wait(mbar, 0);

    // ------------------------------------------------------------ the four sub-sweeps
    unsigned my_trials = 0, my_acc = 0;
    const float hw = g.hw, sigma2 = g.sigma2, dscale = g.dscale;
    const int bq = tid / NAX, aq = tid - bq * NAX;  // fixed thread -> (column, row) of the active lattice
    const int ox0 = H + exl, oy0 = H + eyl;         // region coordinates of the owned tile's corner
    const int nox = min(TX, cps - (int)blockIdx.x * TX), noy = min(TY, g.rows - (int)blockIdx.y * TY);

#pragma unroll 1
    for (int k = 0; k < ((a.dbg_skip & 1) ? 0 : 4); k++) {
        const int lo = k + 1;                       // cells closer than lo to the region edge are stale
        const int pi = ((int)((a.offmask >> (2 * k)) & 1u) - rx0) & 1;       // region-column parity of the active colour
        const int pj = ((int)((a.offmask >> (2 * k + 1)) & 1u) - (g.row0 + ry0)) & 1;
        const int i = lo + ((pi - lo) & 1) + 2 * aq, j = lo + ((pj - lo) & 1) + 2 * bq;
        if (i < RX - lo && j < RY - lo) {
            const int is = i + xs, par = is & 1;
            float4 *pown = sm + j * PITCH + par * HB + (is >> 1);
            const float4 *pL = sm + j * PITCH + (1 - par) * HB + ((is - 1) >> 1);      // left neighbour; right = pL + 1
            const float4 x03 = pown[0], x47 = pown[PLC], y03 = pown[2 * PLC], y47 = pown[3 * PLC];
            const int cnt = decode_cnt(x47.w, y47.w);
            if (cnt != 0) {                         // subsweep.h:252-254
                const bool owned = (unsigned)(i - ox0) < (unsigned)nox && (unsigned)(j - oy0) < (unsigned)noy;
                int gx = rx0 + i, gy = g.row0 + ry0 + j;
                gx += gx < 0 ? cps : 0; gx -= gx >= cps ? cps : 0;
                gy += gy < 0 ? cps : 0; gy -= gy >= cps ? cps : 0;
                const uint32_t cell_id = (uint32_t)gy * (uint32_t)cps + (uint32_t)gx;

                // neighbour part of one trial: smallest d2 against the 3 neighbour cells that can
                // hold a disk closer than sigma (w >= 2 sigma), or -1 when the proposal leaves the
                // cell (out_of_bound subsweep.h:73-88)
                auto neighbours_min_d2 = [&](const float px, const float py) -> float {
                    const bool inb = px > 0.0f && px <= w && py > 0.0f && py <= w;
                    const bool goL = px <= hw, goD = py <= hw;
                    const float npxs = -__fadd_rn(px, goL ? w : -w);     // -(px - helper*w), subsweep.h:139-151
                    const float npys = -__fadd_rn(py, goD ? w : -w);
                    const float4 *pH = pL + (goL ? 0 : 1);
                    const int dV = goD ? -PITCH : PITCH;
                    float m = cell_min_d2<PLC>(pH, npxs, -py);
                    m = fminf(m, cell_min_d2<PLC>(pown + dV, -px, npys));
                    m = fminf(m, cell_min_d2<PLC>(pH + dV, npxs, npys));
                    return inb ? m : -1.0f;      // out of the cell: rejected whatever the neighbours say
                };

                float ox[8] = { x03.x, x03.y, x03.z, x03.w, x47.x, x47.y, x47.z, x47.w };
                float oy[8] = { y03.x, y03.y, y03.z, y03.w, y47.x, y47.y, y47.z, y47.w };
                uint32_t rw[8];
                philox4x32_10(cell_id, a.sweep_lo, a.sweep_hi, 0u, g.seed_lo, g.seed_hi, rw[0], rw[1], rw[2], rw[3]);
                philox4x32_10(cell_id, a.sweep_lo, a.sweep_hi, 1u, g.seed_lo, g.seed_hi, rw[4], rw[5], rw[6], rw[7]);
                // random_shuffle subsweep.h:50-58: physical partial Fisher-Yates, steps 0..3
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    const uint32_t b16 = ((rw[2 * s] & 0xFFu) << 8) | (rw[2 * s + 1] & 0xFFu);
                    const int mrem = cnt > s ? cnt - s : 1;             // s >= cnt: no-op (jj == s)
                    const int jj = s + (int)((b16 * (uint32_t)mrem) >> 16);
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
                    const float2 d45 = pair2(ox[4], ox[5], oy[4], oy[5], npx, npy);
                    const float2 d67 = pair2(ox[6], ox[7], oy[6], oy[7], npx, npy);
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
                pown[PLC] = make_float4(ox[4], ox[5], ox[6], ox[7]);
                pown[2 * PLC] = make_float4(oy[0], oy[1], oy[2], oy[3]);
                pown[3 * PLC] = make_float4(oy[4], oy[5], oy[6], oy[7]);
            }
        }
        __syncthreads();
    }

    // ------------------------------------------------------------ shiftCells(f, d) of this sweep, owned cells, in place
    if (do_shift) {
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
            i0 = ox0 + col; j0 = oy0 + (sdir > 0 ? k0 : TY - 1 - k0);
            di = 0; dj = sdir;
        } else {
            const int row = tid / SEG0, seg = tid - row * SEG0, k0 = seg * K0;
            len = (row < TY && k0 < TX) ? min(K0, TX - k0) : 0;
            j0 = oy0 + row; i0 = ox0 + (sdir > 0 ? k0 : TX - 1 - k0);
            di = sdir; dj = 0;
        }
        auto cell_ptr = [&](int i, int j) -> float4 * {
            const int is = i + xs;
            return sm + j * PITCH + (is & 1) * HB + (is >> 1);
        };
        auto load_cell = [&](int i, int j, CellRegs &c) {
            const float4 *p = cell_ptr(i, j);
            c.x03 = p[0]; c.x47 = p[PLC]; c.y03 = p[2 * PLC]; c.y47 = p[3 * PLC];
            c.cnt = decode_cnt(c.x47.w, c.y47.w);
        };
        CellRegs cur, edge;
        if (len > 0) {
            load_cell(i0, j0, cur);
            load_cell(i0 + len * di, j0 + len * dj, edge);
        }
        __syncthreads();
        constexpr int KMAX = K0 > K1 ? K0 : K1;
#pragma unroll
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
                if (a.shift_f == 0) nNew = shift_into_tile<0, PLC>(cur, up, d, w, sshift, fx, &dropped);
                else nNew = shift_into_tile<1, PLC>(cur, up, d, w, sshift, fx, &dropped);
                if (nNew < PMC_NMAX) fx[3 * PLC * 4 + 3] = __int_as_float(nNew);
                if (dropped) {
                    atomicOr(&ctr->status, PMC_STATUS_OVERFLOW);
                    if ((unsigned)(i - ox0) < (unsigned)nox && (unsigned)(j - oy0) < (unsigned)noy)
                        atomicAdd(&ctr->lost, (unsigned long long)dropped);
                }
                cur = up;
            }
        }
        __syncthreads();
    }

    // ------------------------------------------------------------ owned tile -> HBM (+ periodic images into the margins)
    if (!(a.dbg_skip & 4)) {
        // thread -> fixed (chunk column h, parity, plane), rows strided: a warp stores runs of
        // TX/2 consecutive float4
        constexpr int HX = TX / 2, RSTEP = THREADS / (8 * HX);
        static_assert(THREADS % (8 * HX) == 0, "store mapping");
        const int h = tid % HX, pr = (tid / HX) & 1, pl = (tid / (2 * HX)) & 3, rg = tid / (8 * HX);
        const int ox = 2 * h + pr;                                  // owned column (parity == internal column parity)
        const int ux = blockIdx.x * TX + ox, uy0 = blockIdx.y * TY;
        if (ox < nox) {
            const int is = ox0 + ox + xs;
            const float4 *src = sm + pl * PLC + (is & 1) * HB + (is >> 1) + (oy0 + rg) * PITCH;
            const long long rstride = (long long)8 * g.CH;          // float4 chunks per internal row
            float4 *dst = dout + ((long long)(kMY + uy0 + rg) * 4 + pl) * 2 * g.CH + (long long)pr * g.CH + ((kMX + ux) >> 1);
            // periodic image of this column inside the margins (cps is even: parity is kept)
            const int ximg = ux < kMX ? cps / 2 : (ux >= cps - kMX ? -(cps / 2) : 0);
            const bool yedge = g.wrap_y && (uy0 < kMY || uy0 + noy > g.rows - kMY);
            if (!ximg && !yedge) {
#pragma unroll 2
                for (int oyy = rg; oyy < noy; oyy += RSTEP) {
                    *dst = *src;
                    src += RSTEP * PITCH; dst += RSTEP * rstride;
                }
            } else {
#pragma unroll 1
                for (int oyy = rg; oyy < noy; oyy += RSTEP) {
                    const float4 v = *src;
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

    // acceptance counts reduced warp-level, one atomic per warp (kernel.cu:228,413 accept_counter)
    my_trials = __reduce_add_sync(0xffffffffu, my_trials);
    my_acc = __reduce_add_sync(0xffffffffu, my_acc);
    if ((tid & 31) == 0 && my_trials) {
        atomicAdd(&ctr->trials, (unsigned long long)my_trials);
        atomicAdd(&ctr->accepted, (unsigned long long)my_acc);
    }
}

// ------------------------------------------------------------------ caller layout <-> internal layout
// caller: disk float[cell][2][8] (= 4 float4 per cell: x03, x47, y03, y47) + int16 n[cell]
// (include/pmc.h).  One thread per (internal cell, plane); margins are filled with the
// periodic images, everything beyond with empty cells.
__global__ void import4_kernel(const float4 *__restrict__ disk, const int16_t *__restrict__ n,
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
    float4 v = pl < 2 ? make_float4(kSent, kSent, kSent, kSent) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (have) {
        int cnt = (int)__ldg(n + src);
        cnt = cnt < 0 ? 0 : (cnt > PMC_NMAX ? PMC_NMAX : cnt);
        const float4 q = __ldg(disk + src * 4 + pl);
        const int s0 = (pl & 1) * 4;
        const float fill = pl < 2 ? kSent : 0.0f;       // the caller's unused slots may hold garbage
        v.x = s0 + 0 < cnt ? q.x : fill; v.y = s0 + 1 < cnt ? q.y : fill;
        v.z = s0 + 2 < cnt ? q.z : fill; v.w = s0 + 3 < cnt ? q.w : fill;
        if (pl == 3 && cnt < PMC_NMAX) v.w = __int_as_float(cnt);
    }
    out[((long long)(Y * 4 + pl) * 2 + (X & 1)) * g.CH + (X >> 1)] = v;
}

// internal -> caller layout: every cell of the caller's array (slab: ghost rows included)
__global__ void export4_kernel(const float4 *__restrict__ in, float4 *__restrict__ disk,
                               int16_t *__restrict__ n, Geom4 g, int ghost)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long cell = t >> 2;
    const int pl = (int)(t & 3);
    const long long ncell = (long long)(g.rows + 2 * ghost) * g.cps;
    const bool live = cell < ncell;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
        const int lr = (int)(cell / g.cps), gx = (int)(cell - (long long)lr * g.cps);
        const int X = gx + kMX, Y = lr - ghost + kMY;
        v = __ldg(in + ((long long)(Y * 4 + pl) * 2 + (X & 1)) * g.CH + (X >> 1));
    }
    // the four lanes of a cell are adjacent: lane 1 holds x[7], lane 3 holds y[7]
    const int base = (threadIdx.x & 31) & ~3;
    const float x7 = __shfl_sync(0xffffffffu, v.w, base + 1);
    const float y7 = __shfl_sync(0xffffffffu, v.w, base + 3);
    if (!live) return;
    const int cnt = decode_cnt(x7, y7);
    if (pl == 3 && cnt < PMC_NMAX) v.w = 0.0f;          // pmc.h: unused slots hold x = sentinel, y = 0
    disk[cell * 4 + pl] = v;
    if (pl == 0) n[cell] = (int16_t)cnt;
}

constexpr int kTX = 24, kTY = 40, kMinB = 2;

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

}  // namespace

int pmc4_tile_x() { return kTX; }
int pmc4_tile_y() { return kTY; }

// rows / chunk columns the internal array needs so that every staged box is in bounds
void pmc4_alloc_shape(int cps, int rows, int *CH, int *ROWS)
{
    using TL = Tile4<kTX, kTY>;
    const int gx = (cps + kTX - 1) / kTX, gy = (rows + kTY - 1) / kTY;
    // last staged column: kMX + (gx-1)*TX - H - 1 (rounded down to even) + 2*HB - 1
    const int cols = kMX + (gx - 1) * kTX - TL::H + 2 * TL::HB + 2;
    const int cols_img = cps + 2 * kMX;
    const int c = cols > cols_img ? cols : cols_img;
    *CH = (c + 1) / 2;
    const int r = kMY + (gy - 1) * kTY - TL::H + TL::SYB + 1;
    const int r_img = rows + 2 * kMY;
    *ROWS = r > r_img ? r : r_img;
}

int pmc4_make_tensor_map(void *tmap_out, const float4 *base, const Geom4 &g)
{
    using TL = Tile4<kTX, kTY>;
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
    const cuuint32_t box[4] = { 4 * TL::HB, 2, 1, TL::SYB };
    const cuuint32_t estr[4] = { 1, 1, 1, 1 };
    CUresult r = encode((CUtensorMap *)tmap_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

cudaError_t pmc4_launch_import(const Geom4 &g, int ghost, const float4 *disk, const int16_t *n, float4 *out, cudaStream_t st)
{
    const long long threads = (long long)2 * g.CH * g.ROWS * 4;
    import4_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(disk, n, out, g, ghost);
    return cudaGetLastError();
}

cudaError_t pmc4_launch_export(const Geom4 &g, int ghost, const float4 *in, float4 *disk, int16_t *n, cudaStream_t st)
{
    const long long threads = (long long)(g.rows + 2 * ghost) * g.cps * 4;
    export4_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(in, disk, n, g, ghost);
    return cudaGetLastError();
}

cudaError_t pmc4_launch_sweep(const Geom4 &g, const void *tmap_in, float4 *dout, const SweepArgs &a,
                              Counters *ctr, cudaStream_t st)
{
    using TL = Tile4<kTX, kTY>;
    auto kern = sweep4_kernel<kTX, kTY, kMinB>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL::SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    dim3 grid((g.cps + kTX - 1) / kTX, (g.rows + kTY - 1) / kTY);
    kern<<<grid, TL::THREADS, TL::SMEM, st>>>(*(const CUtensorMap *)tmap_in, dout, g, a, ctr);
    return cudaGetLastError();
}
