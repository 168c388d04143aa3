// ref_harness.cu -- runs the REFERENCE's own kernels (assign kernel.cu:92-150, V2 shiftCells
// shiftCells.h:23-112) on fixed inputs and dumps their outputs as JSON, so that the CPU oracle
// can be pinned against what the reference itself computes.  TEST INFRASTRUCTURE ONLY.
//
// The reference sources are compiled from where they lie (REF_KERNEL_CU is an absolute path
// under /root/reference given by oracle/Makefile); nothing is copied into this repository.
// kernel.cu carries its own main(); it is renamed out of the way.  The reference's
// compile-time geometry applies unchanged: N_ATOMS 800, L 10, cellsPerSide 4, w 2.5, nmax 30.
//
// Inputs are a 2-D configuration embedded in the bottom z-layer of the reference's 3-D box
// (cz = 0), on a dyadic grid (multiples of 2^-10) so that global <-> cell-local conversions
// are exact and the comparison with the 2-D cell-local oracle is bit-for-bit.  Particles
// n_real .. 799 are parked at x = 1000, outside every cell: the reference's membership test
// (kernel.cu:134) drops them.
//
// usage: ref_harness <seed> <n_real> [n_crowd]   (JSON on stdout; GPU required)
// n_crowd > 0: the first n_crowd particles (after the six face probes) alternate between cells (1,1) and
// (2,1), so that cells hold 7-8 particles and shiftCells moves several particles between two crowded
// cells (our nmax is 8, the reference's 30: the tests compare the first 8 slots and the overflow count).
#define main pmc_ref_kernel_cu_main
#include REF_KERNEL_CU
#undef main

#include <cstdint>
#include <vector>

static uint64_t lcg_state;
static uint32_t lcg()
{
    lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
    return (uint32_t)(lcg_state >> 33);
}
// dyadic value in (-5, 5]: k * 2^-10, k in (-5120, 5120]
static float dyadic_coord() { return (float)((int)(lcg() % 10240u) - 5119) / 1024.0f; }

static void dump_state(const char *op, int f, float d, const float *disk, const short *n, bool last)
{
    printf("  {\"op\": \"%s\", \"f\": %d, \"d\": %.9g, \"n\": [", op, f, d);
    for (int c = 0; c < CPS3; c++) printf("%d%s", (int)n[c], c + 1 < CPS3 ? ", " : "");
    printf("],\n   \"cells\": {");
    bool first = true;
    for (int c = 0; c < CPS3; c++) {
        if (n[c] <= 0) continue;
        printf("%s\"%d\": [", first ? "" : ", ", c);
        first = false;
        for (int dim = 0; dim < 3; dim++) {
            printf("[");
            for (int s = 0; s < n[c]; s++)
                printf("%.9g%s", disk[c * nmax * 3 + dim * nmax + s], s + 1 < n[c] ? ", " : "");
            printf("]%s", dim < 2 ? ", " : "");
        }
        printf("]");
    }
    printf("}}%s\n", last ? "" : ",");
}

int main(int argc, char **argv)
{
    const uint64_t seed = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
    const int n_real = argc > 2 ? atoi(argv[2]) : 40;
    const int n_crowd = argc > 3 ? atoi(argv[3]) : 0;
    lcg_state = seed * 2654435761ull + 12345ull;

    std::vector<float> r(3 * N_ATOMS);
    for (int i = 0; i < N_ATOMS; i++) {
        if (i < n_real) {
            r[i] = dyadic_coord();
            r[i + N_ATOMS] = dyadic_coord();
            r[i + 2 * N_ATOMS] = -3.75f;          // layer cz = 0: (-5, -2.5]
        } else {
            r[i] = 1000.0f; r[i + N_ATOMS] = 1000.0f; r[i + 2 * N_ATOMS] = 1000.0f;
        }
    }
    for (int k = 0; k < n_crowd && 6 + k < n_real; k++) {
        const int i = 6 + k;
        const float x0 = (k & 1) ? 0.0f : -2.5f;                    // lower face of cell column 2 / 1
        r[i] = x0 + (float)(1 + (int)(lcg() % 2560u)) / 1024.0f;     // (x0, x0 + 2.5]
        r[i + N_ATOMS] = -2.5f + (float)(1 + (int)(lcg() % 2560u)) / 1024.0f;   // cell row 1
    }
    // edge cases of the half-open rule lb < x <= ub (kernel.cu:134): exactly on cell faces,
    // on the upper box face (kept) and on the lower box face (dropped)
    if (n_real >= 8) {
        r[0] = -2.5f;  r[0 + N_ATOMS] = 0.0f;
        r[1] = 5.0f;   r[1 + N_ATOMS] = 2.5f;
        r[2] = -5.0f;  r[2 + N_ATOMS] = 1.0f;      // dropped: x == lower face
        r[3] = 0.0f;   r[3 + N_ATOMS] = 5.0f;
        r[4] = 2.5f;   r[4 + N_ATOMS] = -2.5f;
        r[5] = -4.9990234375f; r[5 + N_ATOMS] = -4.9990234375f;
    }

    float *d_r, *d_disk;
    short *d_n;
    const size_t rsize = sizeof(float) * 3 * N_ATOMS, dsize = sizeof(float) * 3 * nmax * CPS3, nsize = sizeof(short) * CPS3;
    cudaMalloc(&d_r, rsize); cudaMalloc(&d_disk, dsize); cudaMalloc(&d_n, nsize);
    cudaMemset(d_disk, 0, dsize);
    cudaMemcpy(d_r, r.data(), rsize, cudaMemcpyHostToDevice);
    std::vector<float> disk(3 * nmax * CPS3);
    std::vector<short> n(CPS3);

    printf("{\"source\": \"reference kernels assign (kernel.cu:92-150) and shiftCells (V2 shiftCells.h:23-112), run unmodified on a B200\",\n");
    printf(" \"params\": {\"N_ATOMS\": %d, \"L\": %.9g, \"cellsPerSide\": %d, \"w\": %.9g, \"nmax\": %d, \"seed\": %llu, \"n_real\": %d, \"n_crowd\": %d},\n",
           N_ATOMS, (float)L, cellsPerSide, (float)w, nmax, (unsigned long long)seed, n_real, n_crowd);
    printf(" \"r\": [");
    for (int dim = 0; dim < 3; dim++) {
        printf("[");
        for (int i = 0; i < n_real; i++) printf("%.9g%s", r[i + dim * N_ATOMS], i + 1 < n_real ? ", " : "");
        printf("]%s", dim < 2 ? ", " : "");
    }
    printf("],\n \"steps\": [\n");

    // launch exactly as the reference's main does (kernel.cu:645 and :687)
    assign<<<int(ceil(float(CPS3) / BLOCK_SIZE)), BLOCK_SIZE>>>(d_r, d_disk, d_n);
    cudaError_t st = cudaDeviceSynchronize();
    if (st != cudaSuccess) { fprintf(stderr, "assign failed: %s\n", cudaGetErrorString(st)); return 2; }
    cudaMemcpy(disk.data(), d_disk, dsize, cudaMemcpyDeviceToHost);
    cudaMemcpy(n.data(), d_n, nsize, cudaMemcpyDeviceToHost);
    dump_state("assign", -1, 0.0f, disk.data(), n.data(), false);

    // (f, d): d in (-w/2, w/2] (shiftCells.h:7); includes the probe f=0, d=w/4 (start.cu:253),
    // the closed end d = w/2, d = 0 (dir = -1, nothing moves), one grid step, both signs
    const int fs[] = { 0, 1, 0, 1, 0, 1, 0 };
    const float ds[] = { 0.625f, -1.0f, 1.25f, 0.0009765625f, 0.0f, -1.2490234375f, -0.3330078125f };
    const int nshift = 7;
    dim3 bs_shiftCells(cellsPerSide, cellsPerSide, cellsPerSide);
    for (int k = 0; k < nshift; k++) {
        shiftCells<<<1, bs_shiftCells>>>(d_disk, d_n, fs[k], ds[k]);
        st = cudaDeviceSynchronize();
        if (st != cudaSuccess) { fprintf(stderr, "shiftCells failed: %s\n", cudaGetErrorString(st)); return 3; }
        cudaMemcpy(disk.data(), d_disk, dsize, cudaMemcpyDeviceToHost);
        cudaMemcpy(n.data(), d_n, nsize, cudaMemcpyDeviceToHost);
        dump_state("shiftCells", fs[k], ds[k], disk.data(), n.data(), k + 1 == nshift);
    }
    printf(" ]}\n");
    cudaFree(d_r); cudaFree(d_disk); cudaFree(d_n);
    return 0;
}
