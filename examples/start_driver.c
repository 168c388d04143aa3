/* start_driver.c -- the reference's main() (start.cu:169-272) on top of the C-ABI.
 * Build:  gcc -Iinclude -I/usr/local/cuda/include examples/start_driver.c \
 *             -Lparallel-monte-carlo_b200 -lpmc_b200 -L/usr/local/cuda/lib64 -lcudart -o start_driver
 * usage: start_driver [N [MCpasses [print]]]
 * Prints the acceptance ratio and the invariants; with a third argument `print` also every position in
 * the format of the reference's host_print_disk (start.cu:159-166; global coordinates, 2-D).
 * Exit code 0 iff the per-call protocol and the fused pmc_sweep agree bit for bit. */
#include "pmc.h"
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv)
{
    pmc_params p;
    memset(&p, 0, sizeof(p));
    p.n_particles = argc > 1 ? atoll(argv[1]) : 16384;   /* N_ATOMS            start.cu:14 */
    p.phi = 0.70f;                                       /* -> L               start.cu:15 */
    p.sigma_d = 1.0f;
    p.cell_w = 2.0f;                                     /* w                  start.cu:18 */
    p.nmax = 8;                                          /* nmax               start.cu:19 */
    p.n_M = 4;                                           /* n_M                start.cu:21 */
    p.move_delta = 0.1f;                                 /* sigma              start.cu:22 */
    p.seed = 1234;                                       /* subsweep.h:259 */
    p.device = -1;
    p.n_ranks = 1;
    const int MCpasses = argc > 2 ? atoi(argv[2]) : 10;  /* MCpasses           start.cu:24 */

    pmc_handle *h;
    int rc = pmc_create(&p, &h);
    if (rc) { printf("pmc_create: %s\n", pmc_error_string(rc)); return 1; }

    float *d_r, *d_disk, *d_disk2;
    int16_t *d_n, *d_n2;
    cudaMalloc((void **)&d_r, pmc_r_bytes(h));           /* start.cu:202-205 */
    cudaMalloc((void **)&d_disk, pmc_disk_bytes(h));
    cudaMalloc((void **)&d_n, pmc_n_bytes(h));
    cudaMalloc((void **)&d_disk2, pmc_disk_bytes(h));
    cudaMalloc((void **)&d_n2, pmc_n_bytes(h));

    rc = pmc_init_r(h, d_r);                             /* init_r<<<>>>       start.cu:212 */
    if (rc) printf("init_r: %s\n", pmc_error_string(rc));
    rc = pmc_assign(h, d_r, d_disk, d_n);                /* assign<<<>>>       start.cu:227 */
    if (rc) printf("assign: %s\n", pmc_error_string(rc));
    pmc_assign(h, d_r, d_disk2, d_n2);

    for (int MC_step = 0; MC_step < MCpasses; MC_step++) {      /* start.cu:237 */
        int order[4], f, off[2];
        float d;
        pmc_schedule(h, (uint64_t)MC_step, order, &f, &d);      /* FY_Shuffle :238, (f, d) :251-252 */
        for (int i = 0; i < 4; i++) {                           /* :239 */
            pmc_colour_to_off(order[i], off);                   /* itoa :241 */
            rc = pmc_subsweep(h, d_disk, d_n, off, (uint64_t)MC_step);   /* :242-245 */
            if (rc) printf("subsweep: %s\n", pmc_error_string(rc));
        }
        rc = pmc_shift_cells(h, d_disk, d_n, f, d);             /* shiftCells<<<>>> :255 */
        if (rc) printf("shiftCells: %s\n", pmc_error_string(rc));
    }
    uint64_t trials, accepted, lost;
    uint32_t status;
    pmc_get_counters(h, &trials, &accepted, &lost, &status);

    /* the same loop as one call: one fused kernel per sweep */
    rc = pmc_sweep(h, d_disk2, d_n2, 0, MCpasses);
    if (rc) printf("sweep: %s\n", pmc_error_string(rc));

    size_t db = pmc_disk_bytes(h), nb = pmc_n_bytes(h);
    float *a = (float *)malloc(db), *b = (float *)malloc(db);
    int16_t *na = (int16_t *)malloc(nb), *nb2 = (int16_t *)malloc(nb);
    cudaMemcpy(a, d_disk, db, cudaMemcpyDeviceToHost);          /* start.cu:261-262 */
    cudaMemcpy(b, d_disk2, db, cudaMemcpyDeviceToHost);
    cudaMemcpy(na, d_n, nb, cudaMemcpyDeviceToHost);
    cudaMemcpy(nb2, d_n2, nb, cudaMemcpyDeviceToHost);
    int same = memcmp(a, b, db) == 0 && memcmp(na, nb2, nb) == 0;

    int64_t inv[4];
    float min_d2;
    pmc_check(h, d_disk, d_n, inv, &min_d2);
    printf("N=%lld sweeps=%d trials=%llu accepted=%llu acceptance=%.4f lost=%llu status=%u\n",
           (long long)p.n_particles, MCpasses, (unsigned long long)trials, (unsigned long long)accepted,
           trials ? (double)accepted / (double)trials : 0.0, (unsigned long long)lost, status);
    printf("particles=%lld out_of_cell=%lld min_d2=%.7f fused_equals_per_call=%d\n",
           (long long)inv[0], (long long)inv[1], min_d2, same);

    if (argc > 3 && strcmp(argv[3], "print") == 0) {            /* host_print_disk start.cu:159-166, :263 */
        pmc_geometry g;
        pmc_get_geometry(h, &g);
        for (long long c = 0; c < g.n_cells; c++) {
            const float x0 = (float)(c % g.cps) * g.w - 0.5f * g.L, y0 = (float)(c / g.cps) * g.w - 0.5f * g.L;
            for (int j = 0; j < na[c]; j++)                     /* disk[cell][dim][slot], cell-local -> global */
                printf("Position of atom %i in cell %lld: %f\t%f\n", j, c, x0 + a[16 * c + j], y0 + a[16 * c + 8 + j]);
        }
    }

    cudaFree(d_r); cudaFree(d_disk); cudaFree(d_n); cudaFree(d_disk2); cudaFree(d_n2);   /* :266-269 */
    pmc_destroy(h);
    free(a); free(b); free(na); free(nb2);
    return same && inv[0] == p.n_particles && inv[1] == 0 && status == 0 ? 0 : 2;
}
