/*
 * xs_host.h -- host-side (C) half of the B200-native XSBench: CLI, data model, synthetic
 * data generator, result report.  It mirrors the reference's host interface for this path
 * (same function names, argument meaning and error behaviour; citations relative to
 * ANL-CESAR/XSBench v20) so a reference user finds the same entry points:
 *
 *   read_CLI                    cuda/io.cu:226-441      (+ "-t" from openmp-threading/io.c:66-73)
 *   print_inputs/print_results  cuda/io.cu:116-185, 35-114
 *   grid_init_do_not_profile    cuda/GridInit.cu:90-262
 *   load_num_nucs/mats/concs    cuda/Materials.cu:7-116
 *   LCG_random_double, fast_forward_LCG, pick_mat
 *                               cuda/Simulation.cu:326-362, 287-324
 *   binary_read / binary_write  cuda/io.cu:443-495
 *
 * The GPU work itself is behind include/xs_gpu.h.
 */
#ifndef XS_HOST_H
#define XS_HOST_H

#include <stddef.h>
#include <stdint.h>
#include "xs_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define XS_VERSION 20      /* reference version whose behaviour is reproduced (cuda/Main.cu:8) */

/* ---- random numbers ---------------------------------------------------------------- */
double   LCG_random_double(uint64_t *seed);
uint64_t fast_forward_LCG(uint64_t seed, uint64_t n);
int      pick_mat(uint64_t *seed);
/* The 12 cumulative thresholds pick_mat compares against, computed with the reference's
 * summation order (cuda/Simulation.cu:314-321).  Uploaded to the device by xs_gpu_init. */
void     xs_material_thresholds(double thr[XS_NUM_MATERIALS]);

/* ---- CLI --------------------------------------------------------------------------- */
/* Parses argv into *out.  Returns 0, or -1 with a message in err (no exit).  Keeps the
 * reference's argument-order quirk for "-m event" (cuda/io.cu:302-311). */
int    xs_parse_cli(int argc, char *argv[], Inputs *out, char *err, size_t errlen);
/* Reference-compatible wrapper: prints usage and exit(4) on error (cuda/io.cu:208-224). */
Inputs read_CLI(int argc, char *argv[]);
void   print_CLI_error(void);
/* Extra long options of this driver (--gpus N, --reps N, --json, --dump-xs N, --device-init), stripped
 * from argv before read_CLI sees it. */
typedef struct { int gpus; int reps; int json; long dump_xs; int device_init; } xs_driver_opts;
int    xs_strip_driver_opts(int *argc, char *argv[], xs_driver_opts *o);

/* ---- data generation --------------------------------------------------------------- */
SimulationData grid_init_do_not_profile(Inputs in, int mype);
/* Only the material tables (num_nucs, mats, concs); the three big arrays stay NULL, which asks
 * xs_gpu_init to build them on the device. */
SimulationData xs_materials_only(Inputs in);
void   xs_free_simulation_data(SimulationData *sd);
int   *load_num_nucs(long n_isotopes);
int   *load_mats(int *num_nucs, long n_isotopes, int *max_num_nucs);
double *load_concs(int *num_nucs, int max_num_nucs);
int    NGP_compare(const void *a, const void *b);
int    double_compare(const void *a, const void *b);
size_t estimate_mem_usage(Inputs in);
double get_time(void);

/* ---- report ------------------------------------------------------------------------ */
void logo(int version);
void center_print(const char *s, int width);
void border_print(void);
void fancy_int(long a);
void print_inputs(Inputs in, int nprocs, int version);
/* Returns is_invalid_result (0 = checksum matches the table), like cuda/io.cu:35-114.
 * Beyond the reference's 4-entry table it also knows the golden vectors listed in
 * xs_expected_checksum(). lookups/s is printed as a 64-bit value. */
int  print_results(Inputs in, int mype, double runtime, int nprocs, unsigned long long vhash);
/* Expected checksum for a configuration, or -1 when none is known. */
long xs_expected_checksum(const Inputs *in);

/* ---- binary file mode -------------------------------------------------------------- */
void           binary_write(Inputs in, SimulationData SD);
SimulationData binary_read(Inputs in);

#ifdef __cplusplus
}
#endif
#endif
