/*
 * cli.c -- command line of the host driver.
 *
 * Behaviour reproduced from the reference (cuda/io.cu:226-441, plus "-t" from
 * openmp-threading/io.c:66-73):
 *   defaults  -m history -s large -l 34 -p 500000 -G unionized -h 10000 -k 0
 *   "-m event" turns the default 34 x 500000 into lookups = 17,000,000, particles = 0 --
 *   but only if neither -l nor -p has been seen EARLIER on the command line
 *   (cuda/io.cu:302-311: the CLI is argument-order dependent; kept on purpose).
 *   -s is validated case-insensitively; small => 68 nuclides; XL/XXL only change
 *   n_gridpoints (238847 and (long)(238847*2.1) = 501578) and only without -g
 *   (cuda/io.cu:424-437).
 *   Any unknown flag, missing value or failed validation is a usage error: the reference
 *   prints the usage text and exit(4) (cuda/io.cu:208-224).
 */
#include "xs_host.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <unistd.h>

static int fail(char *err, size_t errlen, const char *what, const char *arg)
{
    if (err && errlen) snprintf(err, errlen, "%s%s%s", what, arg ? ": " : "", arg ? arg : "");
    return -1;
}

static void set_defaults(Inputs *in)
{
    static char default_size[] = "large";
    memset(in, 0, sizeof *in);
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    in->nthreads          = ncpu > 0 ? (int)ncpu : 1;
    in->n_isotopes        = 355;
    in->n_gridpoints      = 11303;
    in->particles         = 500000;
    in->lookups           = 34;
    in->grid_type         = XS_UNIONIZED;
    in->hash_bins         = 10000;
    in->binary_mode       = XS_BINARY_NONE;
    in->kernel_id         = 0;
    in->simulation_method = XS_HISTORY_BASED;
    in->HM                = default_size;
}

static int keyword(const char *v, const char *const names[], const int codes[], int n)
{
    for (int i = 0; i < n; i++)
        if (strcmp(v, names[i]) == 0) return codes[i];
    return -1;
}

int xs_parse_cli(int argc, char *argv[], Inputs *in, char *err, size_t errlen)
{
    static const char *const method_names[] = { "history", "event" };
    static const int         method_codes[] = { XS_HISTORY_BASED, XS_EVENT_BASED };
    static const char *const grid_names[]   = { "unionized", "nuclide", "hash" };
    static const int         grid_codes[]   = { XS_UNIONIZED, XS_NUCLIDE, XS_HASH };
    static const char *const bin_names[]    = { "read", "write" };
    static const int         bin_codes[]    = { XS_BINARY_READ, XS_BINARY_WRITE };

    set_defaults(in);
    int g_given = 0, l_given = 0, p_given = 0;

    for (int i = 1; i < argc; i++) {
        const char *flag = argv[i];
        if (flag[0] != '-' || flag[1] == '\0' || flag[2] != '\0' || !strchr("tgmlhpsGbk", flag[1]))
            return fail(err, errlen, "unknown option", flag);
        if (++i >= argc)
            return fail(err, errlen, "missing value for", flag);
        char *val = argv[i];
        int code;
        switch (flag[1]) {
        case 't': in->nthreads = atoi(val); break;
        case 'g': in->n_gridpoints = atol(val); g_given = 1; break;
        case 'l': in->lookups = atoi(val); l_given = 1; break;
        case 'p': in->particles = atoi(val); p_given = 1; break;
        case 'h': in->hash_bins = atoi(val); break;
        case 'k': in->kernel_id = atoi(val); break;
        case 's': in->HM = val; break;
        case 'm':
            if ((code = keyword(val, method_names, method_codes, 2)) < 0)
                return fail(err, errlen, "bad simulation method", val);
            in->simulation_method = code;
            if (code == XS_EVENT_BASED && !l_given && !p_given) {
                in->lookups *= in->particles;      /* 34 * 500000 */
                in->particles = 0;
            }
            break;
        case 'G':
            if ((code = keyword(val, grid_names, grid_codes, 3)) < 0)
                return fail(err, errlen, "bad grid type", val);
            in->grid_type = code;
            break;
        case 'b':
            if ((code = keyword(val, bin_names, bin_codes, 2)) < 0)
                return fail(err, errlen, "bad binary mode", val);
            in->binary_mode = code;
            break;
        }
    }

    if (in->nthreads < 1)     return fail(err, errlen, "nthreads must be >= 1", NULL);
    if (in->n_isotopes < 1)   return fail(err, errlen, "n_isotopes must be >= 1", NULL);
    if (in->n_gridpoints < 1) return fail(err, errlen, "n_gridpoints must be >= 1", NULL);
    if (in->lookups < 1)      return fail(err, errlen, "lookups must be >= 1", NULL);
    if (in->hash_bins < 1)    return fail(err, errlen, "hash_bins must be >= 1", NULL);

    if (strcasecmp(in->HM, "small") == 0)
        in->n_isotopes = 68;
    else if (strcasecmp(in->HM, "large") == 0)
        ;
    else if (strcasecmp(in->HM, "XL") == 0) {
        if (!g_given) in->n_gridpoints = 238847;
    } else if (strcasecmp(in->HM, "XXL") == 0) {
        if (!g_given) in->n_gridpoints = (long)(238847 * 2.1);
    } else
        return fail(err, errlen, "bad problem size", in->HM);
    return 0;
}

void print_CLI_error(void)
{
    fputs("Usage: ./xsbench <options>\n"
          "Options include:\n"
          "  -m <simulation method>   Simulation method (history, event)\n"
          "  -s <size>                Size of H-M Benchmark to run (small, large, XL, XXL)\n"
          "  -g <gridpoints>          Number of gridpoints per nuclide (overrides -s defaults)\n"
          "  -G <grid type>           Grid search type (unionized, nuclide, hash). Defaults to unionized.\n"
          "  -p <particles>           Number of particle histories\n"
          "  -l <lookups>             History Based: Number of Cross-section (XS) lookups per particle. "
          "Event Based: Total number of XS lookups.\n"
          "  -h <hash bins>           Number of hash bins (only relevant when used with \"-G hash\")\n"
          "  -b <binary mode>         Read or write all data structures to file. If reading, this will "
          "skip initialization phase. (read, write)\n"
          "  -k <kernel ID>           Specifies which kernel to run. 0 is baseline, 1, 2, etc are "
          "optimized variants. (0 is default.)\n"
          "  -t <threads>             Host threads used by the data generator\n"
          "  --gpus <n>               GPUs of this node to partition the lookups over (default 1)\n"
          "  --reps <n>               Timed repetitions after one warm-up (default 1, no warm-up)\n"
          "  --json                   Also print one machine-readable JSON line\n"
          "  --dump-xs <n>            Print energy, material and macro_xs of the first n lookups\n"
          "  --device-init            Build the synthetic problem on the GPU instead of the host\n"
          "Default is equivalent to: -m history -s large -l 34 -p 500000 -G unionized -k 0\n"
          "See readme for full description of default run values\n", stdout);
    exit(4);
}

Inputs read_CLI(int argc, char *argv[])
{
    Inputs in;
    char why[128];
    if (xs_parse_cli(argc, argv, &in, why, sizeof why) != 0)
        print_CLI_error();
    return in;
}

/* Removes this driver's long options from argv (the reference would reject them). */
int xs_strip_driver_opts(int *argc, char *argv[], xs_driver_opts *o)
{
    o->gpus = 1; o->reps = 1; o->json = 0; o->dump_xs = 0; o->device_init = 0;
    int w = 1;
    for (int r = 1; r < *argc; r++) {
        const char *a = argv[r];
        int has_val = r + 1 < *argc;
        if (strcmp(a, "--json") == 0)                     o->json = 1;
        else if (strcmp(a, "--device-init") == 0)         o->device_init = 1;
        else if (strcmp(a, "--gpus") == 0 && has_val)     o->gpus = atoi(argv[++r]);
        else if (strcmp(a, "--reps") == 0 && has_val)     o->reps = atoi(argv[++r]);
        else if (strcmp(a, "--dump-xs") == 0 && has_val)  o->dump_xs = atol(argv[++r]);
        else if (strncmp(a, "--", 2) == 0)                return -1;
        else argv[w++] = argv[r];
    }
    *argc = w;
    if (o->gpus < 1 || o->reps < 1 || o->dump_xs < 0) return -1;
    return 0;
}
