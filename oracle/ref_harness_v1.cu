// ref_harness_v1.cu -- drives the REFERENCE's own sub-sweep device functions (V1, subsweep.h)
//   out_of_bound :73-88, get_neighbors :119-137, apply_PBC :139-151, calculate_pair_energy :90-103,
//   calculate_energy_in_cell :105-117, calculate_energy_in_neighbors :153-172,
//   cpy_to_Dsh :18-27
// on fixed (state, proposal) probes and prints what they return, so that the CPU oracle's per-trial
// decision (oracle_trial) - and through it the CUDA kernels - is pinned to what the reference itself
// computes.  TEST INFRASTRUCTURE ONLY.
//
// subsweep.h is compiled from where it lies under /root/reference (REF_SUBSWEEP_H, given by
// oracle/Makefile); nothing is copied into this repository.  The #define block below is the one of
// start.cu:14-27 (only N_ATOMS differs; it sizes nothing in subsweep.h).  None of the cuRAND-driven
// functions (random_shuffle, make_move, accept_move's Metropolis draw) is called: the probes carry the
// proposal.
//
// How a hard-disk verdict is read off the reference's Lennard-Jones energy: a probe is evaluated once
// per OTHER particle of its 3 x 3 x 3 neighbourhood with every remaining particle parked far outside
// the cut-off (calculate_pair_energy returns exactly 0 beyond r = w, subsweep.h:98-100), slots and
// counts unchanged.  The energy of that evaluation is the single pair energy 4 (r^-12 - r^-6), which is
// > 0 exactly when r < 1 = sigma_d.  hit = any pair energy > 0.  (The all-particles energy is printed too.)
// Particles live in the bottom z layer (cz = 0, z = -3.75) of the reference's 4 x 4 x 4 box.
//
// input (stdin, text):  n_states, then per state: n_cells_used, then per cell: cell_index count x.. y..
//                       n_probes, then per probe: state cx cy slot px py  cnt own_x.. own_y..
//   (global coordinates; the own cell given with the probe replaces the state's copy: the trial sees the
//   cell as the earlier trials of the same sub-sweep left it, subsweep.h:279-297)
// output (stdout): JSON, one record per probe.   GPU required.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "cuda_runtime.h"
#include "math.h"
#include <device_launch_parameters.h>
#include <curand.h>
#include <curand_kernel.h>

#define N_ATOMS 64
#define L 10.0f
#define beta 0.3
#define cellsPerSide 4
#define w 2.5f
#define nmax 10
#define BLOCK_SIZE 1024
#define n_M 10
#define sigma 0.5f
#define dimCB 8
#define MCpasses 1000

const int CPS2 = cellsPerSide * cellsPerSide;
const int CPS3 = CPS2 * cellsPerSide;

#include REF_SUBSWEEP_H

#undef L
#undef w
#undef beta
#undef sigma

static const float kZ = -3.75f;         // centre of layer cz = 0: (-5, -2.5]
static const float kFar = 1000.0f;      // parked particles: beyond the cut-off of every probe

struct Probe {
    int state, cx, cy, slot, cnt;
    float px, py;
    float own_x[8], own_y[8];
};
struct ProbeOut {
    int oob;                // out_of_bound(proposal, cx, cy, 0)
    int hit;                // some single-pair energy > 0
    int n_pairs;            // pairs evaluated
    float e_max;            // largest single-pair energy
    int arg_cell, arg_slot; // where
    float e_full_cell, e_full_nbrs;     // all particles present (the reference's new_energy terms)
    int neighbors[26];
};

// one thread block of ONE thread per probe (cpy_to_Dsh holds a __syncthreads)
__global__ void probe_kernel(const float *states_disk, const short *states_n, const Probe *probes,
                             ProbeOut *out, float *scratch_disk)
{
    const int p = blockIdx.x;
    const Probe pr = probes[p];
    const int DS = CPS3 * 3 * nmax;
    const float *sdisk = states_disk + (size_t)pr.state * DS;
    const short *sn = states_n + (size_t)pr.state * CPS3;
    float *disk = scratch_disk + (size_t)p * DS;       // private working copy of the state
    short n[CPS3];
    for (int c = 0; c < CPS3; c++) n[c] = sn[c];
    for (int i = 0; i < DS; i++) disk[i] = sdisk[i];
    const int cell = get_cell_index(pr.cx, pr.cy, 0);
    // the own cell as this trial sees it
    n[cell] = (short)pr.cnt;
    for (int s = 0; s < pr.cnt; s++) {
        disk[cell * 3 * nmax + s] = pr.own_x[s];
        disk[cell * 3 * nmax + nmax + s] = pr.own_y[s];
        disk[cell * 3 * nmax + 2 * nmax + s] = kZ;
    }
    float proposed[3] = { pr.px, pr.py, kZ };
    ProbeOut o;
    o.oob = out_of_bound(proposed, pr.cx, pr.cy, 0) ? 1 : 0;
    get_neighbors(o.neighbors, pr.cx, pr.cy, 0);
    float D_sh[3 * nmax];
    // --- all particles present: the two terms of calculate_new_energy (subsweep.h:186-191)
    cpy_to_Dsh(D_sh, disk, cell, pr.cnt, 0);
    o.e_full_cell = calculate_energy_in_cell(D_sh, proposed, pr.slot, pr.cnt, 0);
    o.e_full_nbrs = calculate_energy_in_neighbors(disk, proposed, n, pr.cx, pr.cy, 0);
    // --- one other particle at a time, the rest parked (slots and counts unchanged)
    o.hit = 0; o.n_pairs = 0; o.e_max = -1.0e30f; o.arg_cell = -1; o.arg_slot = -1;
    for (int c = 0; c < CPS2; c++) {                    // layer cz = 0 holds every particle
        for (int s = 0; s < n[c]; s++) {
            if (c == cell && s == pr.slot) continue;   // the moving particle itself stays where it is
            // park everything except (c, s) and the moving particle's own slot
            for (int c2 = 0; c2 < CPS2; c2++)
                for (int s2 = 0; s2 < n[c2]; s2++) {
                    const bool keep = (c2 == c && s2 == s) || (c2 == cell && s2 == pr.slot);
                    const float *src = (c2 == cell) ? nullptr : sdisk;
                    float x, y;
                    if (c2 == cell) { x = pr.own_x[s2]; y = pr.own_y[s2]; }
                    else { x = src[c2 * 3 * nmax + s2]; y = src[c2 * 3 * nmax + nmax + s2]; }
                    disk[c2 * 3 * nmax + s2] = keep ? x : kFar;
                    disk[c2 * 3 * nmax + nmax + s2] = keep ? y : kFar;
                    disk[c2 * 3 * nmax + 2 * nmax + s2] = keep ? kZ : kFar;
                }
            cpy_to_Dsh(D_sh, disk, cell, pr.cnt, 0);
            const float e = calculate_energy_in_cell(D_sh, proposed, pr.slot, pr.cnt, 0) +
                            calculate_energy_in_neighbors(disk, proposed, n, pr.cx, pr.cy, 0);
            o.n_pairs++;
            if (e > 0.0f) o.hit = 1;
            if (e > o.e_max) { o.e_max = e; o.arg_cell = c; o.arg_slot = s; }
        }
    }
    out[p] = o;
}

int main()
{
    int n_states = 0;
    if (scanf("%d", &n_states) != 1 || n_states < 1) { fprintf(stderr, "bad input\n"); return 1; }
    const int DS = CPS3 * 3 * nmax;
    std::vector<float> sdisk((size_t)n_states * DS, 0.0f);
    std::vector<short> sn((size_t)n_states * CPS3, 0);
    for (int st = 0; st < n_states; st++) {
        int used = 0;
        if (scanf("%d", &used) != 1) return 1;
        for (int u = 0; u < used; u++) {
            int c = 0, cnt = 0;
            if (scanf("%d %d", &c, &cnt) != 2 || c < 0 || c >= CPS2 || cnt < 0 || cnt > nmax) return 1;
            sn[(size_t)st * CPS3 + c] = (short)cnt;
            float *cellp = sdisk.data() + (size_t)st * DS + (size_t)c * 3 * nmax;
            for (int s = 0; s < cnt; s++) if (scanf("%f", cellp + s) != 1) return 1;
            for (int s = 0; s < cnt; s++) if (scanf("%f", cellp + nmax + s) != 1) return 1;
            for (int s = 0; s < cnt; s++) cellp[2 * nmax + s] = kZ;
        }
    }
    int n_probes = 0;
    if (scanf("%d", &n_probes) != 1 || n_probes < 1) return 1;
    std::vector<Probe> probes(n_probes);
    for (int p = 0; p < n_probes; p++) {
        Probe &q = probes[p];
        if (scanf("%d %d %d %d %f %f %d", &q.state, &q.cx, &q.cy, &q.slot, &q.px, &q.py, &q.cnt) != 7) return 1;
        if (q.state < 0 || q.state >= n_states || q.cnt < 1 || q.cnt > 8 || q.slot < 0 || q.slot >= q.cnt) return 1;
        for (int s = 0; s < 8; s++) { q.own_x[s] = 0.f; q.own_y[s] = 0.f; }
        for (int s = 0; s < q.cnt; s++) if (scanf("%f", &q.own_x[s]) != 1) return 1;
        for (int s = 0; s < q.cnt; s++) if (scanf("%f", &q.own_y[s]) != 1) return 1;
    }
    float *d_sdisk, *d_scratch;
    short *d_sn;
    Probe *d_probes;
    ProbeOut *d_out;
    cudaMalloc(&d_sdisk, sdisk.size() * sizeof(float));
    cudaMalloc(&d_sn, sn.size() * sizeof(short));
    cudaMalloc(&d_probes, probes.size() * sizeof(Probe));
    cudaMalloc(&d_out, probes.size() * sizeof(ProbeOut));
    cudaMalloc(&d_scratch, (size_t)n_probes * DS * sizeof(float));
    cudaMemcpy(d_sdisk, sdisk.data(), sdisk.size() * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(d_sn, sn.data(), sn.size() * sizeof(short), cudaMemcpyHostToDevice);
    cudaMemcpy(d_probes, probes.data(), probes.size() * sizeof(Probe), cudaMemcpyHostToDevice);
    probe_kernel<<<n_probes, 1>>>(d_sdisk, d_sn, d_probes, d_out, d_scratch);
    cudaError_t st = cudaDeviceSynchronize();
    if (st != cudaSuccess) { fprintf(stderr, "probe kernel failed: %s\n", cudaGetErrorString(st)); return 2; }
    std::vector<ProbeOut> out(n_probes);
    cudaMemcpy(out.data(), d_out, out.size() * sizeof(ProbeOut), cudaMemcpyDeviceToHost);
    printf("{\"source\": \"reference device functions out_of_bound, get_neighbors, apply_PBC, calculate_pair_energy, "
           "calculate_energy_in_cell, calculate_energy_in_neighbors (subsweep.h:73-172), compiled unmodified and run on a B200\",\n");
    printf(" \"params\": {\"L\": 10, \"cellsPerSide\": 4, \"w\": 2.5, \"nmax\": 10, \"layer_z\": -3.75},\n \"probes\": [\n");
    for (int p = 0; p < n_probes; p++) {
        const ProbeOut &o = out[p];
        printf("  {\"oob\": %d, \"hit\": %d, \"n_pairs\": %d, \"e_max\": %.9g, \"arg_cell\": %d, \"arg_slot\": %d, "
               "\"e_full_cell\": %.9g, \"e_full_nbrs\": %.9g, \"neighbors\": [",
               o.oob, o.hit, o.n_pairs, o.e_max, o.arg_cell, o.arg_slot, o.e_full_cell, o.e_full_nbrs);
        for (int k = 0; k < 26; k++) printf("%d%s", o.neighbors[k], k < 25 ? ", " : "");
        printf("]}%s\n", p + 1 < n_probes ? "," : "");
    }
    printf(" ]}\n");
    return 0;
}
