// ref_harness_lj.cu -- a SHORT 3-D Lennard-Jones run made of the REFERENCE's own device functions and kernels:
//   make_move :60-71, out_of_bound :73-88, calculate_pair_energy :90-103, calculate_energy_in_cell :105-117,
//   get_neighbors :119-137, apply_PBC :139-151, calculate_energy_in_neighbors :153-172, accept_move :194-217,
//   cpy_to_Dsh / cpy_D_sh_to_Disk / cpy_proposed_to_D_sh (subsweep.h, V1), init_r and assign (kernel.cu, V2 ==
//   start.cu) and shiftCells (V2 shiftCells.h:23-112),
// all compiled unmodified from where they lie under /root/reference.  TEST INFRASTRUCTURE ONLY.
//
// What is NOT taken from the reference is exactly its list of bugs: the sub-sweep loop below is
// subsweep_kernel (subsweep.h:240-300) restated with (i) a cuRAND state that differs per launch (the reference
// re-seeds the same stream every launch, :259), (ii) random_shuffle (:50-58) with a working random_int (the
// reference's casts a (0, 1] uniform to int, :38-40, which makes the shuffle a fixed rotation; without a real
// shuffle the slots that get the 9th and 10th trial are always the first "stayers" of shiftCells, i.e. interior
// particles, and the acceptance ratio comes out 1 % high - measured), and the host loop is start.cu:237-260
// with an unbiased colour shuffle and f in {0, 1, 2} (start.cu:251 yields -1; fixed in kernel.cu:683).
// Output: acceptance ratio and potential energy per particle (calc_energy kernel.cu:452-470, restated on the
// host) after a burn-in, for several seeds -> tests/golden/ref_lj_stats.json, against which the statistical
// parity of PMC_PROPOSAL_GAUSSIAN runs of the LJ mode is tested (3 sigma over seeds).
//
// The box holds 512 particles (8^3 lattice of init_r); the V2 arrays are dimensioned for its N_ATOMS = 800, the other
// 288 atoms are parked outside every cell (the membership test kernel.cu:134 drops them), as in ref_harness.cu.
// usage: ref_harness_lj <n_seeds> <burn_sweeps> <sample_sweeps>        (GPU required)
#include <ctime>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "cuda_runtime.h"
#include "math.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <device_launch_parameters.h>
#include <device_functions.h>
#include <curand.h>
#include <curand_kernel.h>

// V2 translation unit, unmodified: its #define block (kernel.cu:17-30: N_ATOMS 800, L 10, beta 0.3, cellsPerSide 4,
// w 2.5, nmax 30, n_M 15, sigma 0.5) is the one in force for everything below; init_r, assign and (through its
// #include "shiftCells.h") the V2 shiftCells kernel come from here.  Its main() is renamed out of the way.
namespace v2 {
#define main pmc_ref_kernel_cu_main_unused
#include REF_KERNEL_CU
#undef main
}
// V1 device functions, unmodified, under the same macros (subsweep.h takes every parameter from the including
// translation unit, start.cu:14-30)
namespace v1 {
const int CPS2 = cellsPerSide * cellsPerSide;
const int CPS3 = CPS2 * cellsPerSide;
#include REF_SUBSWEEP_H
}
using v1::CPS3;

static const int kReal = 512;           // particles really in the box (8^3 lattice); the other N_ATOMS - 512 are parked outside
static const int kTrials = 10;          // n_M of start.cu:21 (the V2 block says 15; the loop below is ours, the count is a choice)

__device__ unsigned long long g_trials, g_accepts;

// subsweep_kernel (subsweep.h:240-300) with a per-launch random stream and without the degenerate shuffle
__global__ void subsweep_fixed(float *disk, short int *n, int *offset, unsigned long long seed, unsigned long long launch)
{
    using namespace v1;
    int cell_x = 2 * (blockIdx.x * blockDim.x + threadIdx.x) + offset[0];
    int cell_y = 2 * (blockIdx.y * blockDim.y + threadIdx.y) + offset[1];
    int cell_z = 2 * (blockIdx.z * blockDim.z + threadIdx.z) + offset[2];
    int cell_index = get_cell_index(cell_x, cell_y, cell_z);
    __shared__ float D_sh[nmax * 3 * v1::CPS3 / 8];
    int index = threadIdx.x + threadIdx.y * blockDim.x + threadIdx.z * blockDim.x * blockDim.x;
    int atom_counts = n[cell_index];
    curandState_t localRandomState;
    curand_init(seed, (unsigned long long)index + 8ull * launch, 0, &localRandomState);
    cpy_to_Dsh(D_sh, disk, cell_index, atom_counts, index);     // holds a __syncthreads: every thread calls it
    if (atom_counts == 0) return;
    for (int a = atom_counts - 1; a >= 1; a--) {                // random_shuffle subsweep.h:50-58 as intended
        const int b = (int)(curand(&localRandomState) % (unsigned)(a + 1));
        for (int dim = 0; dim < 3; dim++) v1::swap(D_sh + 3 * nmax * index + dim * nmax + a, D_sh + 3 * nmax * index + dim * nmax + b);
    }
    int i = 0;
    float proposed_move[3];
    unsigned long long acc = 0;
    for (int s = 0; s < kTrials; s++) {
        v1::make_move(proposed_move, D_sh, i, &localRandomState, index);
        if (accept_move(proposed_move, cell_x, cell_y, cell_z, disk, D_sh, i, &localRandomState, atom_counts, n, index)) {
            cpy_proposed_to_D_sh(D_sh, proposed_move, i, index);
            acc++;
        }
        i += 1;
        if (i >= atom_counts) i = 0;
    }
    cpy_D_sh_to_Disk(D_sh, disk, cell_index, atom_counts, index);
    atomicAdd(&g_trials, (unsigned long long)kTrials);
    atomicAdd(&g_accepts, acc);
}

// calc_energy kernel.cu:452-470 on the cell arrays
static double host_energy(const std::vector<float> &disk, const std::vector<short> &n, long long *count)
{
    std::vector<float> x;
    for (int c = 0; c < CPS3; c++)
        for (int s = 0; s < n[c]; s++)
            for (int dim = 0; dim < 3; dim++) x.push_back(disk[c * 3 * nmax + dim * nmax + s]);
    const long long N = (long long)x.size() / 3;
    *count = N;
    double e = 0.0;
    for (long long i = 0; i < N; i++)
        for (long long j = i + 1; j < N; j++) {
            float dist = 0.0f;
            for (int dim = 0; dim < 3; dim++) {
                float del = fabsf(x[i * 3 + dim] - x[j * 3 + dim]);
                if (del > L / 2) del -= L;
                dist += del * del;
            }
            dist = sqrtf(dist);
            if (dist <= rc) { float p6 = powf(dist, -6.0f); e += 4.0f * (p6 * p6 - p6); }
        }
    return e;
}

static unsigned long long rng_state;
static unsigned rng() { rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(rng_state >> 33); }

int main(int argc, char **argv)
{
    const int n_seeds = argc > 1 ? atoi(argv[1]) : 8, burn = argc > 2 ? atoi(argv[2]) : 300, sample = argc > 3 ? atoi(argv[3]) : 200;
    const size_t rsize = sizeof(float) * 3 * N_ATOMS, dsize = sizeof(float) * 3 * nmax * CPS3, nsize = sizeof(short) * CPS3;
    std::vector<float> parked(3 * N_ATOMS, 1000.0f);           // outside every cell: dropped by the membership test kernel.cu:134
    float *d_r, *d_disk;
    short *d_n;
    int *d_off;
    cudaMalloc(&d_r, rsize); cudaMalloc(&d_disk, dsize); cudaMalloc(&d_n, nsize); cudaMalloc(&d_off, 3 * sizeof(int));
    std::vector<float> disk(3 * nmax * CPS3);
    std::vector<short> n(CPS3);
    printf("{\"source\": \"reference device functions (subsweep.h V1) + assign + V2 shiftCells in a bug-fixed loop, run on a B200\",\n");
    printf(" \"params\": {\"N_ATOMS\": %d, \"L\": %g, \"beta\": %g, \"cellsPerSide\": %d, \"w\": %g, \"nmax\": %d, \"n_M\": %d, \"sigma\": %g, "
           "\"burn_sweeps\": %d, \"sample_sweeps\": %d},\n \"runs\": [\n", kReal, (double)L, (double)beta, cellsPerSide, (double)w, nmax, kTrials, (double)sigma, burn, sample);
    for (int sd = 0; sd < n_seeds; sd++) {
        rng_state = 88172645463325252ull + 1000003ull * (unsigned long long)sd;
        const int N_cube = 8;
        cudaMemset(d_disk, 0, dsize);
        cudaMemcpy(d_r, parked.data(), rsize, cudaMemcpyHostToDevice);
        v2::init_r<<<1, dim3(N_cube, N_cube, N_cube)>>>(d_r, N_cube);                         // kernel.cu:629 (first 512 atoms)
        v2::assign<<<int(ceil(float(CPS3) / BLOCK_SIZE)), BLOCK_SIZE>>>(d_r, d_disk, d_n);     // kernel.cu:645
        unsigned long long launch = 0, zero = 0, tr = 0, ac = 0;
        double esum = 0.0;
        int esamples = 0, max_n = 0;
        long long count = 0;
        for (int step = 0; step < burn + sample; step++) {
            if (step == burn) { cudaMemcpyToSymbol(g_trials, &zero, 8); cudaMemcpyToSymbol(g_accepts, &zero, 8); }
            int order[8] = { 0, 1, 2, 3, 4, 5, 6, 7 };
            for (int i = 7; i >= 1; i--) { int j = (int)(((unsigned long long)rng() * (unsigned)(i + 1)) >> 31); if (j > i) j = i; int t = order[i]; order[i] = order[j]; order[j] = t; }
            for (int k = 0; k < 8; k++) {
                int off[3] = { (order[k] / 4) % 2, (order[k] / 2) % 2, order[k] % 2 };        // itoa start.cu:153-157
                cudaMemcpy(d_off, off, sizeof(off), cudaMemcpyHostToDevice);               // start.cu:242
                subsweep_fixed<<<1, dim3(cellsPerSide / 2, cellsPerSide / 2, cellsPerSide / 2)>>>(d_disk, d_n, d_off, 1234ull + (unsigned long long)sd, launch++);
            }
            const int f = (int)(rng() % 3u);                                                // kernel.cu:683
            const float d = (float)rng() / 2147483648.0f * w - w / 2.0f;                    // kernel.cu:684
            v2::shiftCells<<<1, dim3(cellsPerSide, cellsPerSide, cellsPerSide)>>>(d_disk, d_n, f, d);   // kernel.cu:687
            if (step >= burn && (step - burn) % 10 == 9) {
                cudaMemcpy(disk.data(), d_disk, dsize, cudaMemcpyDeviceToHost);
                cudaMemcpy(n.data(), d_n, nsize, cudaMemcpyDeviceToHost);
                esum += host_energy(disk, n, &count) / kReal;
                esamples++;
                for (int c = 0; c < CPS3; c++) if (n[c] > max_n) max_n = n[c];
            }
        }
        cudaError_t st = cudaDeviceSynchronize();
        if (st != cudaSuccess) { fprintf(stderr, "run failed: %s\n", cudaGetErrorString(st)); return 2; }
        cudaMemcpyFromSymbol(&tr, g_trials, 8); cudaMemcpyFromSymbol(&ac, g_accepts, 8);
        printf("  {\"seed\": %d, \"acceptance\": %.9g, \"energy_per_particle\": %.9g, \"particles\": %lld, \"max_cell_count\": %d}%s\n",
               sd, (double)ac / (double)tr, esum / esamples, count, max_n, sd + 1 < n_seeds ? "," : "");
    }
    printf(" ]}\n");
    return 0;
}
