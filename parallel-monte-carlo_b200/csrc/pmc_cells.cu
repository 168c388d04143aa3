// pmc_cells.cu -- cell-list build (assign, start.cu:87-146), lattice init (init_r,
// start.cu:47-58), stand-alone shiftCells (V2 shiftCells.h:23-112) and the observables
// (invariant check, g(r) histogram) for sm_100a.
#include "pmc_internal.cuh"
#include <float.h>

namespace {

constexpr float kSent = PMC_SENTINEL;

// ------------------------------------------------------------------ init_r
// start.cu:54-56: r[index] = L / 2.0 * (1.0 - float(2*ix+1) / N_cube): the division is
// float / int -> float, the rest double.  2-D: N_side = sqrt(N).
__global__ void init_r_kernel(float *__restrict__ r, long long N, int ns, float L)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= N) return;
    int iy = (int)(i / ns), ix = (int)(i - (long long)iy * ns);
    float fx = __fdiv_rn((float)(2 * ix + 1), (float)ns);
    float fy = __fdiv_rn((float)(2 * iy + 1), (float)ns);
    double hl = __ddiv_rn((double)L, 2.0);
    r[i]     = (float)__dmul_rn(hl, __dsub_rn(1.0, (double)fx));
    r[i + N] = (float)__dmul_rn(hl, __dsub_rn(1.0, (double)fy));
}

// ------------------------------------------------------------------ assign
// start.cu:129-131: xlb = cellx*w - L/2.0f (float).  Canonical membership:
// the unique c with xlb(c) < x <= xlb(c+1); -1 when x is outside the box.
__device__ __forceinline__ float xlb(int c, float w, float half_L)
{
    return __fadd_rn(__fmul_rn((float)c, w), -half_L);
}

__device__ __forceinline__ int cell_of(float x, const DevGeom &g)
{
    if (!(x > xlb(0, g.w, g.half_L)) || x > xlb(g.cps, g.w, g.half_L)) return -1;
    int c = (int)floorf(__fdiv_rn(__fadd_rn(x, g.half_L), g.w));
    c = c < 0 ? 0 : (c > g.cps - 1 ? g.cps - 1 : c);
    while (c > 0 && !(x > xlb(c, g.w, g.half_L))) c--;
    while (c < g.cps - 1 && x > xlb(c + 1, g.w, g.half_L)) c++;
    return c;
}

// global -> cell-local, snapped to the coordinate grid: k * q with k in [1, K] (oracle to_local)
__device__ __forceinline__ float to_local(float x, int c, const DevGeom &g)
{
    const double origin = __dsub_rn(__dmul_rn((double)c, (double)g.w), __dmul_rn(g.L_box, 0.5));
    const double q = (double)g.dscale;
    double k = rint(__dmul_rn(__dsub_rn((double)x, origin), 1.0 / q));    // q = 2^e: exact scaling (no division), ties to even
    k = fmin(k, (double)g.K);
    k = fmax(k, 1.0);
    return (float)__dmul_rn(k, q);
}

// local storage row of global row gy, or -1 if this rank does not store it
__device__ __forceinline__ int local_row(int gy, const DevGeom &g)
{
    if (g.wrap_y) return gy;
    int lr = wrap_mod(gy - (g.row0 - g.ghost), g.cps);
    return lr < g.local_rows ? lr : -1;
}

// pass 1: one thread per particle: cell id, atomic arrival rank, remember the particle index.
// Arrivals beyond the 8th go to a global list with room for every particle, so that an overflowing cell keeps
// its 8 LOWEST particle indices (what a scan over atoms 0..N-1, start.cu:133-140, keeps when it stops
// writing at nmax) whatever order the atomics resolve in and however many particles overflow.
// (Measured and rejected: warp-aggregated ranks through __match_any_sync, 236 us instead of 120 us at
// N = 2^24 - the match costs more than the one atomic per lane it saves; profiles/r2.)
__global__ void assign_rank_kernel(const float *__restrict__ r, DevGeom g,
                                   unsigned *__restrict__ cnt32, unsigned *__restrict__ idx_tmp,
                                   uint2 *__restrict__ ovf, unsigned *__restrict__ ovf_count, Counters *ctr)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= g.n_particles) return;
    float x = __ldg(r + i), y = __ldg(r + i + g.n_particles);
    int cx = cell_of(x, g), cy = cell_of(y, g);
    if (cx < 0 || cy < 0) {
        if (g.row0 == 0 || g.wrap_y) atomicAdd(&ctr->lost, 1ull);   // counted once (rank 0)
        atomicOr(&ctr->status, PMC_STATUS_LOST);
        return;
    }
    int lr = local_row(cy, g);
    if (lr < 0) return;
    long long cell = (long long)lr * g.cps + cx;
    unsigned s = atomicAdd(cnt32 + cell, 1u);
    if (s < PMC_NMAX) idx_tmp[cell * PMC_NMAX + s] = (unsigned)i;
    else ovf[atomicAdd(ovf_count, 1u)] = make_uint2((unsigned)cell, (unsigned)i);
}

__device__ __forceinline__ void cswap(unsigned &a, unsigned &b)
{
    unsigned lo = min(a, b), hi = max(a, b);
    a = lo; b = hi;
}

// pass 2: one thread per cell: order the arrivals by particle index (= the reference's slot
// order, start.cu:133-140 scans atoms 0..N-1), gather, convert to cell-local, write the cell.
__global__ void assign_fill_kernel(const float *__restrict__ r, DevGeom g,
                                   const unsigned *__restrict__ cnt32,
                                   const unsigned *__restrict__ idx_tmp,
                                   const uint2 *__restrict__ ovf, const unsigned *__restrict__ ovf_count,
                                   float4 *__restrict__ disk, int16_t *__restrict__ n, Counters *ctr)
{
    long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long local_cells = (long long)g.local_rows * g.cps;
    if (cell >= local_cells) return;
    unsigned c32 = cnt32[cell];
    int cnt = c32 > PMC_NMAX ? PMC_NMAX : (int)c32;
    if (c32 > PMC_NMAX) {
        atomicOr(&ctr->status, PMC_STATUS_OVERFLOW);
        int lr = (int)(cell / g.cps);
        if (lr >= g.ghost && lr < g.ghost + g.rows) atomicAdd(&ctr->lost, (unsigned long long)(c32 - PMC_NMAX));
    }
    const uint4 *ip = reinterpret_cast<const uint4 *>(idx_tmp + cell * PMC_NMAX);
    uint4 a = ip[0], b = ip[1];
    unsigned v[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
#pragma unroll
    for (int k = 0; k < 8; k++) if (k >= cnt) v[k] = 0xFFFFFFFFu;
    // 19-comparator sorting network for 8 keys
    cswap(v[0], v[1]); cswap(v[2], v[3]); cswap(v[4], v[5]); cswap(v[6], v[7]);
    cswap(v[0], v[2]); cswap(v[1], v[3]); cswap(v[4], v[6]); cswap(v[5], v[7]);
    cswap(v[1], v[2]); cswap(v[5], v[6]); cswap(v[0], v[4]); cswap(v[3], v[7]);
    cswap(v[1], v[5]); cswap(v[2], v[6]);
    cswap(v[1], v[4]); cswap(v[3], v[6]);
    cswap(v[2], v[4]); cswap(v[3], v[5]);
    cswap(v[3], v[4]);
    if (c32 > PMC_NMAX) {
        // rare: merge the late arrivals, keep the 8 smallest indices (v stays sorted ascending)
        const unsigned m = *ovf_count;
        for (unsigned k = 0; k < m; k++) {
            const uint2 e = ovf[k];
            if (e.x != (unsigned)cell || e.y >= v[7]) continue;
            v[7] = e.y;
            cswap(v[6], v[7]); cswap(v[5], v[6]); cswap(v[4], v[5]); cswap(v[3], v[4]);
            cswap(v[2], v[3]); cswap(v[1], v[2]); cswap(v[0], v[1]);
        }
    }
    int lr = (int)(cell / g.cps), cx = (int)(cell - (long long)lr * g.cps);
    int cy = g.wrap_y ? lr : wrap_mod(g.row0 - g.ghost + lr, g.cps);
    float xs[8], ys[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        xs[k] = kSent; ys[k] = 0.0f;
        if (k < cnt) {
            xs[k] = to_local(__ldg(r + v[k]), cx, g);
            ys[k] = to_local(__ldg(r + v[k] + g.n_particles), cy, g);
        }
    }
    float4 *p = disk + cell * 4;
    p[0] = make_float4(xs[0], xs[1], xs[2], xs[3]);
    p[1] = make_float4(xs[4], xs[5], xs[6], xs[7]);
    p[2] = make_float4(ys[0], ys[1], ys[2], ys[3]);
    p[3] = make_float4(ys[4], ys[5], ys[6], ys[7]);
    n[cell] = (int16_t)cnt;
}

// ------------------------------------------------------------------ stand-alone shiftCells
__device__ __forceinline__ void load_local_cell(const float4 *__restrict__ src,
                                                const int16_t *__restrict__ nsrc,
                                                long long cell, CellRegs &c)
{
    const float4 *p = src + cell * 4;
    c.x03 = __ldg(p); c.x47 = __ldg(p + 1); c.y03 = __ldg(p + 2); c.y47 = __ldg(p + 3);
    int n = __ldg(nsrc + cell);
    c.cnt = n < 0 ? 0 : (n > PMC_NMAX ? PMC_NMAX : n);
}

// one thread per stored cell, out of place.  Compaction goes through a conflict-free
// shared-memory transpose buffer [16 words][block] so the global stores stay 128-bit.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
shift_kernel(const float4 *__restrict__ src, const int16_t *__restrict__ nsrc,
             float4 *__restrict__ dst, int16_t *__restrict__ ndst, DevGeom g, int f, float d,
             Counters *ctr)
{
    __shared__ float buf[16][THREADS];
    long long cell = blockIdx.x * (long long)THREADS + threadIdx.x;
    long long local_cells = (long long)g.local_rows * g.cps;
    if (cell >= local_cells) return;
    const int t = threadIdx.x;
    int lr = (int)(cell / g.cps), cx = (int)(cell - (long long)lr * g.cps);
    const int dir = (d <= 0.0f) ? -1 : 1;
    // upstream neighbour at +dir along f (shiftCells.h:73-82)
    int ulr = lr, ucx = cx;
    bool have_up = true;
    if (f == 0) ucx = wrap_mod(cx + dir, g.cps);
    else if (g.wrap_y) ulr = wrap_mod(lr + dir, g.cps);
    else { ulr = lr + dir; have_up = (ulr >= 0) && (ulr < g.local_rows); }
    CellRegs own, up;
    load_local_cell(src, nsrc, cell, own);
    if (have_up) load_local_cell(src, nsrc, (long long)ulr * g.cps + ucx, up);
    else { up.x03 = up.x47 = up.y03 = up.y47 = make_float4(0, 0, 0, 0); up.cnt = 0; }
#pragma unroll
    for (int k = 0; k < 8; k++) { buf[k][t] = kSent; buf[8 + k][t] = 0.0f; }
    int dropped, nNew;
    const float sshift = __fmul_rn(g.w, (float)dir);
    if (f == 0) {
        auto put = [&](int slot, float fc, float oc) { buf[slot][t] = fc; buf[8 + slot][t] = oc; };
        nNew = shift_one_cell<0>(own, up, d, g.w, sshift, put, &dropped);
    } else {
        auto put = [&](int slot, float fc, float oc) { buf[slot][t] = oc; buf[8 + slot][t] = fc; };
        nNew = shift_one_cell<1>(own, up, d, g.w, sshift, put, &dropped);
    }
    if (dropped) {
        atomicOr(&ctr->status, PMC_STATUS_OVERFLOW);
        if (lr >= g.ghost && lr < g.ghost + g.rows) atomicAdd(&ctr->lost, (unsigned long long)dropped);
    }
    float4 *p = dst + cell * 4;
    p[0] = make_float4(buf[0][t], buf[1][t], buf[2][t], buf[3][t]);
    p[1] = make_float4(buf[4][t], buf[5][t], buf[6][t], buf[7][t]);
    p[2] = make_float4(buf[8][t], buf[9][t], buf[10][t], buf[11][t]);
    p[3] = make_float4(buf[12][t], buf[13][t], buf[14][t], buf[15][t]);
    ndst[cell] = (int16_t)nNew;
}

// ------------------------------------------------------------------ observables
struct PairVisitorCheck {
    float sigma2;
    float min_d2;
    long long overlaps;
    __device__ __forceinline__ void operator()(float d2)
    {
        min_d2 = fminf(min_d2, d2);
        overlaps += d2 < sigma2 ? 1 : 0;
    }
};

// every unordered pair once: same cell i<j, plus the 4 "forward" neighbours
// (+1,0) (-1,+1) (0,+1) (+1,+1); arithmetic identical to oracle_check / oracle_gr_hist.
template <typename V>
__device__ __forceinline__ void visit_pairs(const float4 *__restrict__ disk,
                                            const int16_t *__restrict__ n, const DevGeom &g,
                                            int lr, int cx, const CellRegs &own, V &visit)
{
    const int hx[4] = { 1, -1, 0, 1 }, hy[4] = { 0, 1, 1, 1 };
    float X[8] = { own.x03.x, own.x03.y, own.x03.z, own.x03.w, own.x47.x, own.x47.y, own.x47.z, own.x47.w };
    float Y[8] = { own.y03.x, own.y03.y, own.y03.z, own.y03.w, own.y47.x, own.y47.y, own.y47.z, own.y47.w };
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = i + 1; j < 8; j++)
            if (j < own.cnt) {
                float dx = __fadd_rn(X[i], -X[j]), dy = __fadd_rn(Y[i], -Y[j]);
                visit(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
            }
#pragma unroll
    for (int h = 0; h < 4; h++) {
        int ncx = wrap_mod(cx + hx[h], g.cps);
        int nlr = g.wrap_y ? wrap_mod(lr + hy[h], g.cps) : lr + hy[h];
        if (nlr < 0 || nlr >= g.local_rows) continue;
        CellRegs q;
        load_local_cell(disk, n, (long long)nlr * g.cps + ncx, q);
        float QX[8] = { q.x03.x, q.x03.y, q.x03.z, q.x03.w, q.x47.x, q.x47.y, q.x47.z, q.x47.w };
        float QY[8] = { q.y03.x, q.y03.y, q.y03.z, q.y03.w, q.y47.x, q.y47.y, q.y47.z, q.y47.w };
        const float sx = __fmul_rn((float)hx[h], g.w), sy = __fmul_rn((float)hy[h], g.w);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (i >= own.cnt) continue;
            const float pxs = __fadd_rn(X[i], -sx), pys = __fadd_rn(Y[i], -sy);
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (k < q.cnt) {
                    float dx = __fadd_rn(pxs, -QX[k]), dy = __fadd_rn(pys, -QY[k]);
                    visit(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
                }
        }
    }
}

__global__ void check_kernel(const float4 *__restrict__ disk, const int16_t *__restrict__ n,
                             DevGeom g, long long *out4, unsigned *min_d2_bits)
{
    long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long owned_cells = (long long)g.rows * g.cps;
    long long total = 0, oob = 0, badsent = 0;
    PairVisitorCheck v{ g.sigma2, FLT_MAX, 0 };
    if (q < owned_cells) {
        int lr = (int)(q / g.cps) + g.ghost, cx = (int)(q % g.cps);
        long long cell = (long long)lr * g.cps + cx;
        CellRegs own;
        load_local_cell(disk, n, cell, own);
        total = own.cnt;
        float X[8] = { own.x03.x, own.x03.y, own.x03.z, own.x03.w, own.x47.x, own.x47.y, own.x47.z, own.x47.w };
        float Y[8] = { own.y03.x, own.y03.y, own.y03.z, own.y03.w, own.y47.x, own.y47.y, own.y47.z, own.y47.w };
#pragma unroll
        for (int s = 0; s < 8; s++) {
            if (s < own.cnt) {
                oob += !(X[s] > 0.0f && X[s] <= g.w);
                oob += !(Y[s] > 0.0f && Y[s] <= g.w);
            } else badsent += (X[s] != kSent);
        }
        visit_pairs(disk, n, g, lr, cx, own, v);
    }
    // warp reduce, then one atomic per warp
    for (int o = 16; o; o >>= 1) {
        total += __shfl_xor_sync(0xffffffffu, total, o);
        oob += __shfl_xor_sync(0xffffffffu, oob, o);
        badsent += __shfl_xor_sync(0xffffffffu, badsent, o);
        v.overlaps += __shfl_xor_sync(0xffffffffu, v.overlaps, o);
        v.min_d2 = fminf(v.min_d2, __shfl_xor_sync(0xffffffffu, v.min_d2, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd((unsigned long long *)&out4[0], (unsigned long long)total);
        atomicAdd((unsigned long long *)&out4[1], (unsigned long long)oob);
        atomicAdd((unsigned long long *)&out4[2], (unsigned long long)v.overlaps);
        atomicAdd((unsigned long long *)&out4[3], (unsigned long long)badsent);
        atomicMin(min_d2_bits, __float_as_uint(v.min_d2));   // d2 >= 0: uint order == float order
    }
}

struct PairVisitorHist {
    unsigned *sh;
    float rmax2, inv_dr;
    int nbins;
    __device__ __forceinline__ void operator()(float d2)
    {
        if (d2 < rmax2) {
            int b = (int)__fmul_rn(__fsqrt_rn(d2), inv_dr);
            if (b < nbins) atomicAdd(sh + b, 1u);
        }
    }
};

// shared-memory histogram per CTA, flushed with one 64-bit atomic per non-empty bin
__global__ void gr_hist_kernel(const float4 *__restrict__ disk, const int16_t *__restrict__ n,
                               DevGeom g, float rmax2, float inv_dr, int nbins,
                               unsigned long long *hist)
{
    extern __shared__ unsigned sh[];
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) sh[b] = 0;
    __syncthreads();
    long long owned_cells = (long long)g.rows * g.cps;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < owned_cells;
         q += (long long)gridDim.x * blockDim.x) {
        int lr = (int)(q / g.cps) + g.ghost, cx = (int)(q % g.cps);
        CellRegs own;
        load_local_cell(disk, n, (long long)lr * g.cps + cx, own);
        PairVisitorHist v{ sh, rmax2, inv_dr, nbins };
        visit_pairs(disk, n, g, lr, cx, own, v);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += blockDim.x)
        if (sh[b]) atomicAdd(hist + b, (unsigned long long)sh[b]);
}

// ------------------------------------------------------------------ disk -> r (disk_to_r kernel.cu:497-507)
// Particles in cell order, slots in order: particle k of the output is the k-th particle met when the owned
// cells are walked row by row.  Three small kernels: per-block particle counts, an exclusive scan of the block
// sums (one CTA), then every block scans its own cells again and writes its particles (global coordinates,
// the same double-precision conversion as oracle_disk_to_r).
constexpr int kD2RCells = 1024;     // cells per block (4 per thread)

__device__ __forceinline__ int block_exclusive_scan(int v, int *total)
{
    __shared__ int warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int ws = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, ws, o); if (lane >= o) ws += t; }
        warp_sums[lane] = ws;
    }
    __syncthreads();
    const int base = wid ? warp_sums[wid - 1] : 0;
    if (total) *total = warp_sums[(blockDim.x >> 5) - 1];
    __syncthreads();
    return base + inc - v;
}

__device__ __forceinline__ int owned_count(const int16_t *__restrict__ n, const DevGeom &g, long long q)
{
    if (q >= (long long)g.rows * g.cps) return 0;
    const int c = __ldg(n + q + (long long)g.ghost * g.cps);
    return c < 0 ? 0 : (c > PMC_NMAX ? PMC_NMAX : c);
}

__global__ void __launch_bounds__(256) d2r_count_kernel(const int16_t *__restrict__ n, DevGeom g, unsigned long long *block_sums)
{
    const long long q0 = (long long)blockIdx.x * kD2RCells + threadIdx.x * 4;
    int v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) v += owned_count(n, g, q0 + k);
    int total;
    block_exclusive_scan(v, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = (unsigned long long)total;
}

// one CTA: exclusive scan of the block sums in place; block_sums[nblocks] = grand total
__global__ void __launch_bounds__(1024) d2r_scan_kernel(unsigned long long *block_sums, int nblocks)
{
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0ull;
    __syncthreads();
    for (int b0 = 0; b0 < nblocks; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const unsigned long long v = i < nblocks ? block_sums[i] : 0ull;
        // per-block sums are < 2^13: scan them as int, carry the 64-bit running total separately
        int total;
        const int ex = block_exclusive_scan((int)v, &total);
        if (i < nblocks) block_sums[i] = carry + (unsigned long long)ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += (unsigned long long)total;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[nblocks] = carry;
}

__global__ void __launch_bounds__(256) d2r_write_kernel(const float4 *__restrict__ disk, const int16_t *__restrict__ n, DevGeom g,
                                                       const unsigned long long *__restrict__ block_sums, float *__restrict__ r,
                                                       long long r_stride, long long r_cap)
{
    const long long q0 = (long long)blockIdx.x * kD2RCells + threadIdx.x * 4;
    int cnt[4], v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { cnt[k] = owned_count(n, g, q0 + k); v += cnt[k]; }
    long long pos = (long long)block_sums[blockIdx.x] + block_exclusive_scan(v, nullptr);
    const double half = __dmul_rn(g.L_box, 0.5);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (!cnt[k]) continue;
        const long long q = q0 + k;
        const int row = (int)(q / g.cps), cx = (int)(q - (long long)row * g.cps), cy = g.row0 + row;
        const float4 *p = disk + (q + (long long)g.ghost * g.cps) * 4;
        const float4 x03 = __ldg(p), x47 = __ldg(p + 1), y03 = __ldg(p + 2), y47 = __ldg(p + 3);
        const float X[8] = { x03.x, x03.y, x03.z, x03.w, x47.x, x47.y, x47.z, x47.w };
        const float Y[8] = { y03.x, y03.y, y03.z, y03.w, y47.x, y47.y, y47.z, y47.w };
        const double ox = __dsub_rn(__dmul_rn((double)cx, (double)g.w), half), oy = __dsub_rn(__dmul_rn((double)cy, (double)g.w), half);
#pragma unroll
        for (int s = 0; s < 8; s++)
            if (s < cnt[k] && pos + s < r_cap) {
                r[pos + s] = (float)__dadd_rn(ox, (double)X[s]);
                r[pos + s + r_stride] = (float)__dadd_rn(oy, (double)Y[s]);
            }
        pos += cnt[k];
    }
}

}  // namespace

// d_r: SoA [2][r_stride] global coordinates of this rank's owned particles (at most r_cap of them are written).
// *scratch (stream-ordered allocation, the caller frees it with cudaFreeAsync) holds the exclusive block offsets;
// (*scratch)[*nblocks] = how many particles there are.
cudaError_t pmc_launch_disk_to_r(const DevGeom &g, const float4 *disk, const int16_t *n, float *d_r, long long r_stride,
                                 long long r_cap, unsigned long long **scratch, int *nblocks, cudaStream_t st)
{
    const long long owned = (long long)g.rows * g.cps;
    const int blocks = (int)((owned + kD2RCells - 1) / kD2RCells);
    unsigned long long *sums = nullptr;
    cudaError_t e = cudaMallocAsync(&sums, (size_t)(blocks + 1) * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    d2r_count_kernel<<<blocks, 256, 0, st>>>(n, g, sums);
    d2r_scan_kernel<<<1, 1024, 0, st>>>(sums, blocks);
    d2r_write_kernel<<<blocks, 256, 0, st>>>(disk, n, g, sums, d_r, r_stride, r_cap);
    e = cudaGetLastError();
    *scratch = sums;
    *nblocks = blocks;
    return e;
}

cudaError_t pmc_launch_init_r(const DevGeom &g, float *d_r, cudaStream_t st)
{
    long long N = g.n_particles;
    int ns = (int)(sqrt((double)N) + 0.5);
    int threads = 256;
    long long blocks = (N + threads - 1) / threads;
    init_r_kernel<<<(unsigned)blocks, threads, 0, st>>>(d_r, N, ns, g.L);
    return cudaGetLastError();
}

cudaError_t pmc_launch_assign(const DevGeom &g, const float *d_r, float4 *disk, int16_t *n,
                              Counters *ctr, cudaStream_t st)
{
    long long local_cells = (long long)g.local_rows * g.cps;
    unsigned *cnt32 = nullptr, *idx_tmp = nullptr;
    uint2 *ovf = nullptr;
    // cnt32[local_cells] is followed by the overflow-list counter (zeroed by the same memset)
    cudaError_t e = cudaMallocAsync(&cnt32, (local_cells + 1) * sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    // room for every particle: touched only by arrivals beyond the 8th of a cell, costs address space, not bandwidth
    e = cudaMallocAsync(&ovf, (size_t)g.n_particles * sizeof(uint2), st);
    if (e != cudaSuccess) { cudaFreeAsync(cnt32, st); return e; }
    e = cudaMallocAsync(&idx_tmp, local_cells * PMC_NMAX * sizeof(unsigned), st);
    if (e != cudaSuccess) { cudaFreeAsync(cnt32, st); cudaFreeAsync(ovf, st); return e; }
    cudaMemsetAsync(cnt32, 0, (local_cells + 1) * sizeof(unsigned), st);
    unsigned *ovf_count = cnt32 + local_cells;
    int threads = 256;
    long long b1 = (g.n_particles + threads - 1) / threads;
    assign_rank_kernel<<<(unsigned)b1, threads, 0, st>>>(d_r, g, cnt32, idx_tmp, ovf, ovf_count, ctr);
    long long b2 = (local_cells + threads - 1) / threads;
    assign_fill_kernel<<<(unsigned)b2, threads, 0, st>>>(d_r, g, cnt32, idx_tmp, ovf, ovf_count, disk, n, ctr);
    e = cudaGetLastError();
    cudaFreeAsync(cnt32, st);
    cudaFreeAsync(idx_tmp, st);
    cudaFreeAsync(ovf, st);
    return e;
}

cudaError_t pmc_launch_shift(const DevGeom &g, const float4 *src, const int16_t *nsrc,
                             float4 *dst, int16_t *ndst, int f, float d, Counters *ctr,
                             cudaStream_t st)
{
    constexpr int threads = 128;
    long long local_cells = (long long)g.local_rows * g.cps;
    long long blocks = (local_cells + threads - 1) / threads;
    shift_kernel<threads><<<(unsigned)blocks, threads, 0, st>>>(src, nsrc, dst, ndst, g, f, d, ctr);
    return cudaGetLastError();
}

cudaError_t pmc_launch_check(const DevGeom &g, const float4 *disk, const int16_t *n,
                             long long *out4, unsigned *min_d2_bits, cudaStream_t st)
{
    cudaMemsetAsync(out4, 0, 4 * sizeof(long long), st);
    cudaMemsetAsync(min_d2_bits, 0x7f, sizeof(unsigned), st);   // 0x7f7f7f7f = 3.39e38
    int threads = 128;
    long long owned = (long long)g.rows * g.cps;
    long long blocks = (owned + threads - 1) / threads;
    check_kernel<<<(unsigned)blocks, threads, 0, st>>>(disk, n, g, out4, min_d2_bits);
    return cudaGetLastError();
}

cudaError_t pmc_launch_gr_hist(const DevGeom &g, const float4 *disk, const int16_t *n,
                               float r_max, int nbins, unsigned long long *hist, cudaStream_t st)
{
    cudaMemsetAsync(hist, 0, (size_t)nbins * sizeof(unsigned long long), st);
    int threads = 128;
    long long owned = (long long)g.rows * g.cps;
    long long blocks = (owned + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    float inv_dr = (float)nbins / r_max;     // same single IEEE division as oracle_gr_hist
    float rmax2 = r_max * r_max;
    gr_hist_kernel<<<(unsigned)blocks, threads, (size_t)nbins * sizeof(unsigned), st>>>(
        disk, n, g, rmax2, inv_dr, nbins, hist);
    return cudaGetLastError();
}
