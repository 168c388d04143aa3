// pmc_sweep.cu -- the checkerboard sub-sweep (reference subsweep.h:240-300) for sm_100a.
//
// One kernel template, two instantiations:
//   NCOL = 1  one colour, in place  -> pmc_subsweep   (call site start.cu:242-245)
//   NCOL = 4  a whole MC sweep: pending shiftCells applied while loading, then the four
//             colours back to back on a tile held in shared memory (temporal blocking with
//             a halo of 4 cells recomputed redundantly), out of place -> pmc_sweep.
// The redundant halo work is legal because every random number is a pure function of
// (seed, sweep, global cell id, trial index): two CTAs recomputing the same cell get the
// same bits, so the result is independent of the tiling and of the number of GPUs.
//
// Work decomposition: one thread per active cell (the reference's V1 mapping,
// subsweep.h:242-245), trials strictly sequential inside a cell (subsweep.h:279-297).
// Shared-memory tile layout: four float4 planes (x0-3, x4-7, y0-3, y4-7), each plane stored
// row by row with even and odd columns split so that the same-colour cells a warp works on
// are contiguous (conflict-free LDS.128).  Overlap tests use the Blackwell packed-FP32
// instructions (FADD2 / FMUL2 / FFMA2): two slots per instruction, IEEE-RN per component.
#include "pmc_internal.cuh"

namespace {

constexpr float kSent = PMC_SENTINEL;

template <int NCOL, int T>
struct Tile {
    static constexpr int H = NCOL;          // halo cells on each side
    static constexpr int R = T + 2 * H;     // region edge (cells)
    static constexpr int RR = R * R;
    static constexpr int HALF = R / 2;
    static constexpr size_t SMEM = (size_t)RR * 64 + ((RR + 15) / 16) * 16;
    static_assert(R % 2 == 0, "region edge must be even");
};

// ---- global -> registers, with pmc.h's "unused slots are garbage-tolerant" rule
__device__ __forceinline__ void sanitize(CellRegs &c)
{
    int n = c.cnt;
    n = n < 0 ? 0 : (n > PMC_NMAX ? PMC_NMAX : n);
    c.cnt = n;
    c.x03.x = n > 0 ? c.x03.x : kSent; c.x03.y = n > 1 ? c.x03.y : kSent;
    c.x03.z = n > 2 ? c.x03.z : kSent; c.x03.w = n > 3 ? c.x03.w : kSent;
    c.x47.x = n > 4 ? c.x47.x : kSent; c.x47.y = n > 5 ? c.x47.y : kSent;
    c.x47.z = n > 6 ? c.x47.z : kSent; c.x47.w = n > 7 ? c.x47.w : kSent;
}

// (ux, uy): unwrapped column / owned-relative row.  Rows outside the slab's storage read
// as empty cells (they can only influence cells this CTA does not own).
__device__ __forceinline__ void load_cell(const float4 *__restrict__ din,
                                          const int16_t *__restrict__ nin,
                                          const DevGeom &g, int ux, int uy, CellRegs &c)
{
    int gx = wrap_mod(ux, g.cps);
    int lr;
    bool valid = true;
    if (g.wrap_y) lr = wrap_mod(uy, g.cps);
    else { lr = uy + g.ghost; valid = (lr >= 0) && (lr < g.local_rows); }
    if (valid) {
        long long cell = (long long)lr * g.cps + gx;
        const float4 *p = din + cell * 4;
        c.x03 = __ldg(p); c.x47 = __ldg(p + 1); c.y03 = __ldg(p + 2); c.y47 = __ldg(p + 3);
        c.cnt = __ldg(nin + cell);
    } else {
        c.x03 = c.x47 = c.y03 = c.y47 = make_float4(0.f, 0.f, 0.f, 0.f);
        c.cnt = 0;
    }
    sanitize(c);
}

// two slots per instruction: d2 = (q.x + npx)^2 + (q.y + npy)^2 < sigma2 ?
__device__ __forceinline__ bool pair2_hit(float qx0, float qx1, float qy0, float qy1,
                                          float2 npx, float2 npy, float sigma2)
{
    float2 dx = __fadd2_rn(make_float2(qx0, qx1), npx);
    float2 dy = __fadd2_rn(make_float2(qy0, qy1), npy);
    float2 t = __fmul2_rn(dy, dy);
    float2 d2 = __ffma2_rn(dx, dx, t);
    return (d2.x < sigma2) | (d2.y < sigma2);
}

// all 8 slots of one staged cell against the trial point (already in that cell's frame,
// negated: npx = -pxs).  Unused slots hold the sentinel and can never hit.
template <int RR>
__device__ __forceinline__ bool cell_hit(const float4 *cellp, float npx, float npy, float sigma2)
{
    const float4 x03 = cellp[0], x47 = cellp[RR], y03 = cellp[2 * RR], y47 = cellp[3 * RR];
    const float2 nx = make_float2(npx, npx), ny = make_float2(npy, npy);
    bool h = pair2_hit(x03.x, x03.y, y03.x, y03.y, nx, ny, sigma2);
    h |= pair2_hit(x03.z, x03.w, y03.z, y03.w, nx, ny, sigma2);
    h |= pair2_hit(x47.x, x47.y, y47.x, y47.y, nx, ny, sigma2);
    h |= pair2_hit(x47.z, x47.w, y47.z, y47.w, nx, ny, sigma2);
    return h;
}

template <int NCOL, int T, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
sweep_tile_kernel(const float4 *din, const int16_t *nin, float4 *dout, int16_t *nout,
                  const DevGeom g, const SweepArgs a, Counters *ctr)
{
    using TL = Tile<NCOL, T>;
    constexpr int H = TL::H, R = TL::R, RR = TL::RR, HALF = TL::HALF;
    extern __shared__ float4 sm[];
    unsigned char *scnt = reinterpret_cast<unsigned char *>(sm + 4 * RR);

    const int tid = threadIdx.x;
    const int cps = g.cps;
    const int ux0 = blockIdx.x * T - H;     // unwrapped global column of region column 0
    const int uy0 = blockIdx.y * T - H;     // owned-relative row of region row 0
    const float w = g.w;

    // ------------------------------------------------------------ phase 0: stage the tile
    for (int idx = tid; idx < RR; idx += THREADS) {
        const int j = idx / R, i = idx - j * R;
        const int ux = ux0 + i, uy = uy0 + j;
        if (NCOL == 1) {
            // in-place mode: ring cells of the active colour belong to other CTAs, may be
            // written concurrently and are never read by this CTA -> do not touch them
            const bool ring = (i < H) | (i >= H + T) | (j < H) | (j >= H + T);
            if (ring && ((ux & 1) == a.offx[0]) && (((g.row0 + uy) & 1) == a.offy[0])) continue;
        }
        const int sid = j * R + (i & 1) * HALF + (i >> 1);
        CellRegs own;
        load_cell(din, nin, g, ux, uy, own);
        if (NCOL == 1 || !a.shift_on) {
            sm[sid] = own.x03; sm[sid + RR] = own.x47;
            sm[sid + 2 * RR] = own.y03; sm[sid + 3 * RR] = own.y47;
            scnt[sid] = (unsigned char)own.cnt;
        } else {
            // pending shiftCells(f, d) of the previous sweep (shiftCells.h:23-112), applied
            // as a gather while loading: destination cell pulls from itself and from the
            // one upstream neighbour at +dir along f
            const int dir = (a.shift_d <= 0.0f) ? -1 : 1;
            CellRegs up;
            load_cell(din, nin, g, ux + (a.shift_f == 0 ? dir : 0), uy + (a.shift_f == 1 ? dir : 0), up);
            sm[sid] = make_float4(kSent, kSent, kSent, kSent);
            sm[sid + RR] = make_float4(kSent, kSent, kSent, kSent);
            sm[sid + 2 * RR] = make_float4(0.f, 0.f, 0.f, 0.f);
            sm[sid + 3 * RR] = make_float4(0.f, 0.f, 0.f, 0.f);
            float *base = reinterpret_cast<float *>(sm + sid);
            int dropped, nNew;
            const float sshift = __fmul_rn(w, (float)dir);
            if (a.shift_f == 0) {
                auto put = [&](int slot, float fc, float oc) {
                    float *p = base + (slot >> 2) * (RR * 4) + (slot & 3);
                    p[0] = fc; p[2 * RR * 4] = oc;
                };
                nNew = shift_one_cell<0>(own, up, a.shift_d, w, sshift, put, &dropped);
            } else {
                auto put = [&](int slot, float fc, float oc) {
                    float *p = base + (slot >> 2) * (RR * 4) + (slot & 3);
                    p[0] = oc; p[2 * RR * 4] = fc;
                };
                nNew = shift_one_cell<1>(own, up, a.shift_d, w, sshift, put, &dropped);
            }
            scnt[sid] = (unsigned char)nNew;
            if (dropped) {
                atomicOr(&ctr->status, PMC_STATUS_OVERFLOW);
                const bool owned = (i >= H) & (i < H + T) & (j >= H) & (j < H + T) & (ux < cps) & (uy < g.rows);
                if (owned) atomicAdd(&ctr->lost, (unsigned long long)dropped);
            }
        }
    }
    __syncthreads();

    // ------------------------------------------------------------ the sub-sweeps
    unsigned my_trials = 0, my_acc = 0;
    const float sigma = g.sigma, sigma2 = g.sigma2, dscale = g.dscale;
    const int n_M = g.n_M;

#pragma unroll 1
    for (int k = 0; k < NCOL; k++) {
        const int lo = (NCOL == 1) ? H : k + 1;     // cells closer than lo to the region edge are stale
        const int pi = (a.offx[k] - ux0) & 1;       // region-column parity of the active colour
        const int pj = (a.offy[k] - (g.row0 + uy0)) & 1;
        const int i_first = lo + ((pi - lo) & 1), j_first = lo + ((pj - lo) & 1);
        const int na = (R - lo - i_first + 1) >> 1, nb = (R - lo - j_first + 1) >> 1;

#pragma unroll 1
        for (int q = tid; q < na * nb; q += THREADS) {
            const int bq = q / na, aq = q - bq * na;
            const int i = i_first + 2 * aq, j = j_first + 2 * bq;
            const int sid = j * R + (i & 1) * HALF + (i >> 1);
            const int cnt = scnt[sid];
            if (cnt == 0) continue;                 // subsweep.h:252-254
            const int ux = ux0 + i, uy = uy0 + j;
            const bool owned = (i >= H) & (i < H + T) & (j >= H) & (j < H + T) & (ux < cps) & (uy < g.rows);
            const uint32_t cell_id = (uint32_t)wrap_mod(g.row0 + uy, cps) * (uint32_t)cps + (uint32_t)wrap_mod(ux, cps);
            const int sidL = j * R + ((i - 1) & 1) * HALF + ((i - 1) >> 1);
            const int sidR = j * R + ((i + 1) & 1) * HALF + ((i + 1) >> 1);
            const float4 *pown = sm + sid;
            float *fown = reinterpret_cast<float *>(sm + sid);

            uint32_t perm = 0x76543210u;            // lazily shuffled trial order (random_shuffle subsweep.h:50-58)
            int it = 0;                             // i of subsweep.h:278,291-296
            uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
#pragma unroll 1
            for (int s = 0; s < n_M; s++) {         // subsweep.h:279
                uint32_t ra, rb;
                if ((s & 1) == 0) {
                    philox4x32_10(cell_id, a.sweep_lo, a.sweep_hi, (uint32_t)(s >> 1), g.seed_lo, g.seed_hi, r0, r1, r2, r3);
                    ra = r0; rb = r1;
                } else { ra = r2; rb = r3; }
                if (s < cnt) {                      // Fisher-Yates step s
                    const uint32_t b16 = ((ra & 0xFFu) << 8) | (rb & 0xFFu);
                    const int jj = s + (int)((b16 * (uint32_t)(cnt - s)) >> 16);
                    const uint32_t x = ((perm >> (4 * s)) ^ (perm >> (4 * jj))) & 15u;
                    perm ^= (x << (4 * s)) | (x << (4 * jj));
                }
                const int slot = (perm >> (4 * it)) & 15;
                it = (it + 1 >= cnt) ? 0 : it + 1;
                float *fx = fown + (slot >> 2) * (RR * 4) + (slot & 3);
                float *fy = fx + 2 * RR * 4;
                const float x = *fx, y = *fy;
                // make_move subsweep.h:60-71 (uniform square, exactly symmetric set)
                const int mx = (int)(((ra >> 8) << 1) | 1u) - (1 << 24);
                const int my = (int)(((rb >> 8) << 1) | 1u) - (1 << 24);
                const float px = __fadd_rn(x, __fmul_rn((float)mx, dscale));
                const float py = __fadd_rn(y, __fmul_rn((float)my, dscale));
                my_trials += owned ? 1u : 0u;
                // out_of_bound subsweep.h:73-88
                if (!(px > 0.0f && px <= w && py > 0.0f && py <= w)) continue;
                // which neighbour columns / rows can hold a disk closer than sigma?  Exact
                // conservative tests (monotonicity of IEEE rounding): a skipped cell could
                // not have produced d2 < sigma2 in the oracle's arithmetic.
                const float pxl = __fadd_rn(px, w), pxr = __fadd_rn(px, -w);
                const float pyd = __fadd_rn(py, w), pyu = __fadd_rn(py, -w);
                const bool needL = !(__fadd_rn(pxl, -w) >= sigma), needR = !(pxr <= -sigma);
                const bool needD = !(__fadd_rn(pyd, -w) >= sigma), needU = !(pyu <= -sigma);
                *fx = kSent;                        // hide the moving disk from its own cell test (j != i, subsweep.h:109)
                bool hit;
                if (!((needL & needR) | (needD & needU))) {
                    // fast path (always taken when w >= 2 sigma): own + at most 3 cells
                    const float npx = -px, npy = -py;
                    const float npxH = needL ? -pxl : (needR ? -pxr : kSent);
                    const float npyV = needD ? -pyd : (needU ? -pyu : kSent);
                    const float4 *pH = sm + (needL ? sidL : sidR);
                    const int dV = needD ? -R : R;
                    hit = cell_hit<RR>(pown, npx, npy, sigma2);
                    hit |= cell_hit<RR>(pH, npxH, npy, sigma2);
                    hit |= cell_hit<RR>(pown + dV, npx, npyV, sigma2);
                    hit |= cell_hit<RR>(pH + dV, npxH, npyV, sigma2);
                } else {
                    // generic path (w < 2 sigma): every needed cell of the 3x3 block
                    hit = false;
#pragma unroll 1
                    for (int dj = -1; dj <= 1; dj++) {
                        if ((dj < 0 && !needD) || (dj > 0 && !needU)) continue;
                        const float npy = dj < 0 ? -pyd : (dj > 0 ? -pyu : -py);
#pragma unroll 1
                        for (int di = -1; di <= 1; di++) {
                            if ((di < 0 && !needL) || (di > 0 && !needR)) continue;
                            const float npx = di < 0 ? -pxl : (di > 0 ? -pxr : -px);
                            const int s2 = (di < 0 ? sidL : (di > 0 ? sidR : sid)) + dj * R;
                            hit |= cell_hit<RR>(sm + s2, npx, npy, sigma2);
                        }
                    }
                }
                // accept_move subsweep.h:194-217 (hard disks: accept iff no overlap)
                *fx = hit ? x : px;
                if (!hit) { *fy = py; my_acc += owned ? 1u : 0u; }
            }
        }
        __syncthreads();
    }

    // ------------------------------------------------------------ write back the owned tile
    for (int idx = tid; idx < T * T; idx += THREADS) {
        const int jj = idx / T, ii = idx - jj * T;
        const int i = H + ii, j = H + jj;
        const int ux = ux0 + i, uy = uy0 + j;
        if (ux >= cps || uy >= g.rows) continue;
        if (NCOL == 1 && (((ux & 1) != a.offx[0]) || (((g.row0 + uy) & 1) != a.offy[0]))) continue;
        const int sid = j * R + (i & 1) * HALF + (i >> 1);
        const long long cell = (long long)(uy + g.ghost) * cps + ux;
        float4 *p = dout + cell * 4;
        p[0] = sm[sid]; p[1] = sm[sid + RR]; p[2] = sm[sid + 2 * RR]; p[3] = sm[sid + 3 * RR];
        if (NCOL != 1) nout[cell] = (int16_t)scnt[sid];
    }

    // acceptance counts reduced warp-level, one atomic per warp (kernel.cu:228,413 accept_counter)
    my_trials = __reduce_add_sync(0xffffffffu, my_trials);
    my_acc = __reduce_add_sync(0xffffffffu, my_acc);
    if ((tid & 31) == 0 && my_trials) {
        atomicAdd(&ctr->trials, (unsigned long long)my_trials);
        atomicAdd(&ctr->accepted, (unsigned long long)my_acc);
    }
}

constexpr int kT1 = 32, kThreads1 = 256, kMinB1 = 2;   // single colour
constexpr int kT4 = 32, kThreads4 = 384, kMinB4 = 2;   // fused sweep

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
    auto kern = sweep_tile_kernel<1, kT1, kThreads1, kMinB1>;
    constexpr size_t smem = Tile<1, kT1>::SMEM;
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    dim3 grid((g.cps + kT1 - 1) / kT1, (g.rows + kT1 - 1) / kT1);
    kern<<<grid, kThreads1, smem, st>>>(disk, n, disk, nullptr, g, a, ctr);
    return cudaGetLastError();
}

cudaError_t pmc_launch_fused_sweep(const DevGeom &g, const float4 *din, const int16_t *nin,
                                   float4 *dout, int16_t *nout, const SweepArgs &a,
                                   Counters *ctr, cudaStream_t st)
{
    auto kern = sweep_tile_kernel<4, kT4, kThreads4, kMinB4>;
    constexpr size_t smem = Tile<4, kT4>::SMEM;
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    dim3 grid((g.cps + kT4 - 1) / kT4, (g.rows + kT4 - 1) / kT4);
    kern<<<grid, kThreads4, smem, st>>>(din, nin, dout, nout, g, a, ctr);
    return cudaGetLastError();
}
