/*
 * main.c -- the xsbench host driver.
 *
 * Same sequence as the reference driver (cuda/Main.cu:3-109): parse CLI, print inputs,
 * generate (or read) the problem, move it to the device, run the selected simulation,
 * apply the final "% 999983", print the result block, exit status = is_invalid_result.
 * The device work goes through the C ABI in include/xs_gpu.h only.
 *
 * Differences that are deliberate:
 *   - the FOM is computed from device (cudaEvent) time; host wall time is printed too
 *     (the reference times with a host clock: cuda/Main.cu:59-97, cuda/XSutils.cu:77-79);
 *   - history mode runs on the GPU (the reference's CUDA build refuses it: cuda/Main.cu:84-88);
 *   - --gpus/--reps/--json/--dump-xs (see cli.c).
 */
#include "xs_host.h"

#include <stdio.h>
#include <stdlib.h>

static int cmp_double(const void *a, const void *b)
{
    const double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

int main(int argc, char *argv[])
{
    const int mype = 0;
    xs_driver_opts opt;
    if (xs_strip_driver_opts(&argc, argv, &opt) != 0)
        print_CLI_error();
    Inputs in = read_CLI(argc, argv);

    print_inputs(in, opt.gpus, XS_VERSION);

    SimulationData SD;
    if (in.binary_mode == XS_BINARY_READ) SD = binary_read(in);
    else if (opt.device_init) {
        printf("Building nuclide grids and acceleration structure on the GPU...\n");
        SD = xs_materials_only(in);
    } else SD = grid_init_do_not_profile(in, mype);
    if (in.binary_mode == XS_BINARY_WRITE && !opt.device_init)
        binary_write(in, SD);

    printf("Allocating and moving simulation data to GPU memory space...\n");
    xs_gpu_ctx *ctx = NULL;
    if (xs_gpu_init(&in, &SD, opt.gpus, &ctx) != XS_OK) {
        fprintf(stderr, "xs_gpu_init failed: %s\n", xs_gpu_last_error());
        return 2;
    }
    xs_gpu_info info;
    xs_gpu_get_info(ctx, &info);
    printf("GPU Intialization complete. Allocated %.0lf MB of data on each GPU.\n",
           info.resident_bytes / 1024.0 / 1024.0);
    if (in.binary_mode == XS_BINARY_WRITE && opt.device_init) {
        /* the problem exists on the device only: fetch the three big arrays, then write the file */
        const long n_points = in.n_isotopes * in.n_gridpoints;
        SD.length_nuclide_grid = (int)n_points;
        SD.nuclide_grid = malloc((size_t)n_points * sizeof(NuclideGridPoint));
        SD.length_unionized_energy_array = in.grid_type == XS_UNIONIZED ? (int)n_points : 0;
        SD.length_index_grid = in.grid_type == XS_UNIONIZED ? n_points * in.n_isotopes
                             : in.grid_type == XS_HASH ? (long)in.hash_bins * in.n_isotopes : 0;
        if (SD.length_unionized_energy_array)
            SD.unionized_energy_array = malloc((size_t)n_points * sizeof(double));
        if (SD.length_index_grid)
            SD.index_grid = malloc((size_t)SD.length_index_grid * sizeof(int));
        int rc = SD.nuclide_grid ? xs_gpu_read_array(ctx, XS_ARRAY_NUCLIDE_GRID, 0, n_points * 48, SD.nuclide_grid) : XS_ERR_ARG;
        if (rc == XS_OK && SD.length_unionized_energy_array)
            rc = SD.unionized_energy_array ? xs_gpu_read_array(ctx, XS_ARRAY_UNIONIZED_ENERGY, 0, n_points * 8, SD.unionized_energy_array) : XS_ERR_ARG;
        if (rc == XS_OK && SD.length_index_grid)
            rc = SD.index_grid ? xs_gpu_read_array(ctx, XS_ARRAY_INDEX_GRID, 0, SD.length_index_grid * 4, SD.index_grid) : XS_ERR_ARG;
        if (rc != XS_OK) {
            fprintf(stderr, "-b write with --device-init: cannot fetch the generated problem: %s\n", xs_gpu_last_error());
            return 2;
        }
        binary_write(in, SD);
    }
    xs_free_simulation_data(&SD);     /* the device copy is self-contained */

    printf("\n");
    border_print();
    center_print("SIMULATION", 79);
    border_print();

    if (opt.dump_xs > 0) {
        long n = opt.dump_xs;
        double *e = malloc(n * sizeof *e), *xs = malloc(5 * n * sizeof *xs);
        int *mat = malloc(n * sizeof *mat);
        if (xs_gpu_dump(ctx, 0, n, e, mat, xs, NULL) != XS_OK) {
            fprintf(stderr, "xs_gpu_dump failed: %s\n", xs_gpu_last_error());
            return 2;
        }
        for (long i = 0; i < n; i++)
            printf("lookup %ld: E = %.17g mat = %d macro_xs = [%.17g, %.17g, %.17g, %.17g, %.17g]\n",
                   i, e[i], mat[i], xs[5*i], xs[5*i+1], xs[5*i+2], xs[5*i+3], xs[5*i+4]);
        free(e); free(xs); free(mat);
    }

    xs_gpu_result res;
    int total_runs = opt.reps > 1 ? opt.reps + 1 : 1;      /* one warm-up when repeating */
    double *times = malloc((size_t)total_runs * sizeof *times);
    double host_start = get_time();
    for (int r = 0; r < total_runs; r++) {
        int rc = xs_gpu_run(ctx, &in, &res);
        if (rc != XS_OK) {
            if (rc == XS_ERR_ARG) printf("Error: No kernel ID %d found!\n", in.kernel_id);
            fprintf(stderr, "xs_gpu_run failed: %s\n", xs_gpu_last_error());
            return 1;
        }
        times[r] = res.device_seconds;
        if (r == 0) host_start = get_time() - res.host_seconds;
    }
    double host_seconds = res.host_seconds;
    (void)host_start;

    double best = times[total_runs - 1], median = best;
    if (total_runs > 1) {
        qsort(times + 1, (size_t)opt.reps, sizeof *times, cmp_double);
        best = times[1];
        median = times[1 + opt.reps / 2];
    }
    printf("\nSimulation complete.\n");
    printf("Device time: %.6f s (best of %d), median %.6f s; host wall time of last run %.6f s\n",
           best, opt.reps, median, host_seconds);
    printf("Phases (last run): sample %.6f  sort %.6f  lookup %.6f  reduce %.6f s\n",
           res.phase_seconds[XS_PHASE_SAMPLE], res.phase_seconds[XS_PHASE_SORT],
           res.phase_seconds[XS_PHASE_LOOKUP], res.phase_seconds[XS_PHASE_REDUCE]);

    xs_gpu_finalize(ctx);

    unsigned long long verification = res.verification % XS_HASH_MODULUS;
    int is_invalid_result = print_results(in, mype, best, opt.gpus, verification);

    if (opt.json)
        printf("{\"lookups\": %llu, \"device_seconds\": %.9f, \"median_device_seconds\": %.9f, "
               "\"host_seconds\": %.9f, \"lookups_per_sec\": %.1f, \"checksum\": %llu, "
               "\"valid\": %s, \"n_gpus\": %d, \"kernel_id\": %d, \"launches\": %d}\n",
               res.n_lookups, best, median, host_seconds, (double)res.n_lookups / best,
               verification, is_invalid_result ? "false" : "true", opt.gpus, in.kernel_id,
               res.gpu_launches);
    free(times);
    return is_invalid_result;
}
