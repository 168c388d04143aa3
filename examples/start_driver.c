/* start_driver.c -- the reference's main() (start.cu:169-272) on top of the C-ABI, with a command line in
 * place of "edit the #defines and recompile" (start.cu:14-24; V2 kernel.cu:17-34).
 * Build:  gcc -Iinclude -I/usr/local/cuda/include examples/start_driver.c \
 *             -Lparallel-monte-carlo_b200 -lpmc_b200 -L/usr/local/cuda/lib64 -lcudart -o start_driver
 *
 * usage: start_driver [N [MCpasses [print]]] [flags]          (positional form kept from round 1)
 *   --N <n>            N_ATOMS       start.cu:14     particles (a perfect square unless --rsa)
 *   --phi <f>          -> L          start.cu:15     packing fraction; L = sqrt(N pi sigma_d^2 / 4 phi)
 *   --sigma-d <f>                                     disk diameter
 *   --w <f>            w             start.cu:18     target cell width (cellsPerSide = L / w, start.cu:17)
 *   --nmax <n>         nmax          start.cu:19     slots per cell (this build: 8)
 *   --n-M <n>          n_M           start.cu:21     trials per active cell per sub-sweep
 *   --delta <f>        sigma         start.cu:22     proposal half-width
 *   --passes <n>       MCpasses      start.cu:24     sweeps
 *   --seed <n>         1234          subsweep.h:259
 *   --proposal <uniform|gaussian>    subsweep.h:64   trial displacement: uniform square (bit-exact, fast kernel) or the
 *                                                     reference's curand_normal * sigma (generic kernel)
 *   --rsa                                             random-sequential-addition start instead of init_r's lattice
 *   --fused                                           one pmc_sweep call per trace interval instead of the per-call protocol
 *   --trace [k]                                       every k sweeps (default 1): "%i: %f\n" sweep and acceptance ratio since
 *                                                     the last line, in the shape of the V2 energy trace (kernel.cu:695)
 *   --dump <file> [--dump-every k]                    trajectory in the reference's format (create_dump kernel.cu:510-536)
 *   --checkpoint <file>                               write (params, sweep, counters, disk, n) at the end
 *   --resume <file>                                   continue the chain of a checkpoint (same parameters required)
 *   --print                                           every position like host_print_disk (start.cu:159-166)
 *   --verify                                          also run the other protocol from the same start and compare bit for bit
 * Exit code 0 iff the invariants hold (and, with --verify, both protocols agree bit for bit). */
#include "pmc.h"
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int run_per_call(pmc_handle *h, float *d_disk, int16_t *d_n, uint64_t sweep0, int n)
{
    int worst = 0;
    for (int k = 0; k < n; k++) {                                   /* start.cu:237 */
        const uint64_t MC_step = sweep0 + (uint64_t)k;
        int order[4], f, off[2], rc;
        float d;
        pmc_schedule(h, MC_step, order, &f, &d);                    /* FY_Shuffle :238, (f, d) :251-252 */
        for (int i = 0; i < 4; i++) {                               /* :239 */
            pmc_colour_to_off(order[i], off);                       /* itoa :241 */
            rc = pmc_subsweep(h, d_disk, d_n, off, MC_step);        /* :242-245 */
            if (rc) { printf("subsweep: %s\n", pmc_error_string(rc)); worst = rc; }     /* printed, the run continues (:246-249) */
        }
        rc = pmc_shift_cells(h, d_disk, d_n, f, d);                 /* shiftCells<<<>>> :255 */
        if (rc) { printf("shiftCells: %s\n", pmc_error_string(rc)); worst = rc; }
    }
    return worst;
}

int main(int argc, char **argv)
{
    pmc_params p;
    memset(&p, 0, sizeof(p));
    p.n_particles = 16384;                               /* N_ATOMS            start.cu:14 */
    p.phi = 0.70f;                                       /* -> L               start.cu:15 */
    p.sigma_d = 1.0f;
    p.cell_w = 2.0f;                                     /* w                  start.cu:18 */
    p.nmax = 8;                                          /* nmax               start.cu:19 */
    p.n_M = 4;                                           /* n_M                start.cu:21 */
    p.move_delta = 0.1f;                                 /* sigma              start.cu:22 */
    p.seed = 1234;                                       /* subsweep.h:259 */
    p.device = -1;
    p.n_ranks = 1;
    int MCpasses = 10;                                   /* MCpasses           start.cu:24 */
    int print = 0, rsa = 0, fused = 0, trace = 0, verify = 0, dump_every = 1, npos = 0;
    const char *dump = NULL, *ckpt = NULL, *resume = NULL;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        const char *v = i + 1 < argc ? argv[i + 1] : NULL;
        if (a[0] != '-') {
            if (npos == 0) p.n_particles = atoll(a);
            else if (npos == 1) MCpasses = atoi(a);
            else if (npos == 2 && !strcmp(a, "print")) print = 1;
            else { printf("unexpected argument %s\n", a); return 1; }
            npos++;
            if (npos <= 2) verify = 1;                   /* the round-1 form: run both protocols and compare */
            continue;
        }
#define NEED if (!v) { printf("%s needs a value\n", a); return 1; } i++
        if (!strcmp(a, "--N")) { NEED; p.n_particles = atoll(v); }
        else if (!strcmp(a, "--phi")) { NEED; p.phi = (float)atof(v); }
        else if (!strcmp(a, "--sigma-d")) { NEED; p.sigma_d = (float)atof(v); }
        else if (!strcmp(a, "--w")) { NEED; p.cell_w = (float)atof(v); }
        else if (!strcmp(a, "--nmax")) { NEED; p.nmax = atoi(v); }
        else if (!strcmp(a, "--n-M")) { NEED; p.n_M = atoi(v); }
        else if (!strcmp(a, "--delta")) { NEED; p.move_delta = (float)atof(v); }
        else if (!strcmp(a, "--passes")) { NEED; MCpasses = atoi(v); }
        else if (!strcmp(a, "--seed")) { NEED; p.seed = strtoull(v, NULL, 10); }
        else if (!strcmp(a, "--proposal")) { NEED; p.proposal = !strcmp(v, "gaussian") ? PMC_PROPOSAL_GAUSSIAN : PMC_PROPOSAL_UNIFORM; }
        else if (!strcmp(a, "--dump")) { NEED; dump = v; }
        else if (!strcmp(a, "--dump-every")) { NEED; dump_every = atoi(v) > 0 ? atoi(v) : 1; }
        else if (!strcmp(a, "--checkpoint")) { NEED; ckpt = v; }
        else if (!strcmp(a, "--resume")) { NEED; resume = v; }
        else if (!strcmp(a, "--trace")) { trace = 1; if (v && v[0] != '-') { trace = atoi(v) > 0 ? atoi(v) : 1; i++; } }
        else if (!strcmp(a, "--rsa")) rsa = 1;
        else if (!strcmp(a, "--fused")) fused = 1;
        else if (!strcmp(a, "--print")) print = 1;
        else if (!strcmp(a, "--verify")) verify = 1;
        else { printf("unknown flag %s\n", a); return 1; }
    }

    pmc_handle *h;
    int rc = pmc_create(&p, &h);
    if (rc) { printf("pmc_create: %s\n", pmc_error_string(rc)); return 1; }
    pmc_geometry g;
    pmc_get_geometry(h, &g);
    printf("N_ATOMS=%lld L=%.6f cellsPerSide=%d w=%.7f nmax=%d n_M=%d sigma(move)=%.7f seed=%llu MCpasses=%d phi=%.4f\n",
           (long long)g.n_particles, g.L, g.cps, g.w, g.nmax, g.n_M, g.move_delta, (unsigned long long)p.seed, MCpasses, p.phi);

    float *d_r, *d_disk, *d_disk2 = NULL;
    int16_t *d_n, *d_n2 = NULL;
    cudaMalloc((void **)&d_r, pmc_r_bytes(h));           /* start.cu:202-205 */
    cudaMalloc((void **)&d_disk, pmc_disk_bytes(h));
    cudaMalloc((void **)&d_n, pmc_n_bytes(h));

    uint64_t sweep0 = 0;
    if (resume) {
        rc = pmc_load_checkpoint(h, resume, d_disk, d_n, &sweep0);
        if (rc) { printf("resume %s: %s\n", resume, pmc_error_string(rc)); return 1; }
        printf("resumed at sweep %llu\n", (unsigned long long)sweep0);
    } else {
        if (rsa) {
            float *r_host = (float *)malloc(pmc_r_bytes(h));
            int64_t attempts = 0;
            rc = pmc_rsa_host(&p, p.seed, r_host, &attempts);
            if (rc) { printf("rsa: %s\n", pmc_error_string(rc)); return 1; }
            cudaMemcpy(d_r, r_host, pmc_r_bytes(h), cudaMemcpyHostToDevice);
            free(r_host);
        } else {
            rc = pmc_init_r(h, d_r);                     /* init_r<<<>>>       start.cu:212 */
            if (rc) { printf("init_r: %s\n", pmc_error_string(rc)); return 1; }
        }
        rc = pmc_assign(h, d_r, d_disk, d_n);            /* assign<<<>>>       start.cu:227 */
        if (rc) printf("assign: %s\n", pmc_error_string(rc));
    }
    if (verify) {
        cudaMalloc((void **)&d_disk2, pmc_disk_bytes(h));
        cudaMalloc((void **)&d_n2, pmc_n_bytes(h));
        cudaMemcpy(d_disk2, d_disk, pmc_disk_bytes(h), cudaMemcpyDeviceToDevice);
        cudaMemcpy(d_n2, d_n, pmc_n_bytes(h), cudaMemcpyDeviceToDevice);
    }
    if (dump) pmc_write_dump(h, d_disk, d_n, dump, (int)sweep0, 0);

    /* the loop of start.cu:237-260, cut at every trace / dump point */
    uint64_t t_prev = 0, a_prev = 0, trials = 0, accepted = 0, lost = 0;
    uint32_t status = 0;
    for (int done = 0; done < MCpasses; ) {
        int n = MCpasses - done;
        if (trace && trace - done % trace < n) n = trace - done % trace;
        if (dump && dump_every - done % dump_every < n) n = dump_every - done % dump_every;
        rc = fused ? pmc_sweep(h, d_disk, d_n, sweep0 + (uint64_t)done, n)
                   : run_per_call(h, d_disk, d_n, sweep0 + (uint64_t)done, n);
        if (rc && fused) printf("sweep: %s\n", pmc_error_string(rc));
        done += n;
        pmc_get_counters(h, &trials, &accepted, &lost, &status);
        if (trace && (done % trace == 0 || done == MCpasses)) {
            printf("%i: %f\n", (int)(sweep0 + (uint64_t)done), trials > t_prev ? (double)(accepted - a_prev) / (double)(trials - t_prev) : 0.0);
            t_prev = trials; a_prev = accepted;
        }
        if (dump && (done % dump_every == 0 || done == MCpasses))
            pmc_write_dump(h, d_disk, d_n, dump, (int)(sweep0 + (uint64_t)done), 1);
    }
    pmc_get_counters(h, &trials, &accepted, &lost, &status);

    int same = 1;
    size_t db = pmc_disk_bytes(h), nb = pmc_n_bytes(h);
    float *a = (float *)malloc(db);
    int16_t *na = (int16_t *)malloc(nb);
    cudaMemcpy(a, d_disk, db, cudaMemcpyDeviceToHost);             /* start.cu:261-262 */
    cudaMemcpy(na, d_n, nb, cudaMemcpyDeviceToHost);
    if (verify) {
        /* the other protocol from the same start: per-call <-> one fused kernel per sweep */
        rc = fused ? run_per_call(h, d_disk2, d_n2, sweep0, MCpasses) : pmc_sweep(h, d_disk2, d_n2, sweep0, MCpasses);
        if (rc) printf("verify run: %s\n", pmc_error_string(rc));
        float *b = (float *)malloc(db);
        int16_t *nb2 = (int16_t *)malloc(nb);
        cudaMemcpy(b, d_disk2, db, cudaMemcpyDeviceToHost);
        cudaMemcpy(nb2, d_n2, nb, cudaMemcpyDeviceToHost);
        same = memcmp(a, b, db) == 0 && memcmp(na, nb2, nb) == 0;
        free(b); free(nb2);
    }

    int64_t inv[4];
    float min_d2;
    pmc_check(h, d_disk, d_n, inv, &min_d2);
    printf("N=%lld sweeps=%d trials=%llu accepted=%llu acceptance=%.4f lost=%llu status=%u\n",
           (long long)p.n_particles, MCpasses, (unsigned long long)trials, (unsigned long long)accepted,
           trials ? (double)accepted / (double)trials : 0.0, (unsigned long long)lost, status);
    printf("particles=%lld out_of_cell=%lld overlaps=%lld min_d2=%.7f fused_equals_per_call=%d\n",
           (long long)inv[0], (long long)inv[1], (long long)inv[2], min_d2, same);
    if (ckpt) {
        rc = pmc_save_checkpoint(h, d_disk, d_n, sweep0 + (uint64_t)MCpasses, ckpt);
        printf("checkpoint %s at sweep %llu: %s\n", ckpt, (unsigned long long)(sweep0 + (uint64_t)MCpasses), pmc_error_string(rc));
    }

    if (print) {                                                    /* host_print_disk start.cu:159-166, :263 */
        for (long long c = 0; c < g.n_cells; c++) {
            const float x0 = (float)(c % g.cps) * g.w - 0.5f * g.L, y0 = (float)(c / g.cps) * g.w - 0.5f * g.L;
            for (int j = 0; j < na[c]; j++)                         /* disk[cell][dim][slot], cell-local -> global */
                printf("Position of atom %i in cell %lld: %f\t%f\n", j, c, x0 + a[16 * c + j], y0 + a[16 * c + 8 + j]);
        }
    }

    cudaFree(d_r); cudaFree(d_disk); cudaFree(d_n); cudaFree(d_disk2); cudaFree(d_n2);   /* :266-269 */
    pmc_destroy(h);
    free(a); free(na);
    return same && inv[0] == p.n_particles && inv[1] == 0 && inv[2] == 0 && status == 0 ? 0 : 2;
}
