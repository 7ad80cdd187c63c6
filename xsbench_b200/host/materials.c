/*
 * materials.c -- the 12 Hoogenboom-Martin materials: how many nuclides each holds, which
 * ones, and their (synthetic) concentrations.
 *
 * Must stay value-identical to the reference tables (cuda/Materials.cu:7-31 counts, :34-96
 * nuclide ids, :99-116 concentrations): the verification checksum depends on them.
 * Layout is the reference's: dense 12 x max_num_nucs row-major, rows padded.
 */
#include "xs_host.h"

#include <stdlib.h>
#include <string.h>

/* Composition of each material as runs "first,count" of consecutive nuclide ids would be
 * shorter, but an explicit list is easier to audit against the reference. */
static const int FUEL_SMALL[34] = { 58, 59, 60, 61, 40, 42, 43, 44, 45, 46,  1,  2,  3,  7,
                                     8,  9, 10, 29, 57, 47, 48,  0, 62, 15, 33, 34, 52, 53,
                                    54, 55, 56, 18, 23, 41 };
static const int CLADDING[5]    = { 63, 64, 65, 66, 67 };
static const int WATER[4]       = { 24, 41,  4,  5 };                /* cold and hot borated */
static const int RPV[27]        = { 19, 20, 21, 22, 35, 36, 37, 38, 39, 25, 27, 28, 29,
                                    30, 31, 32, 26, 49, 50, 51, 11, 12, 13, 14,  6, 16, 17 };
static const int STRUCTURE[21]  = { 24, 41,  4,  5, 19, 20, 21, 22, 35, 36, 37, 38, 39, 25,
                                    49, 50, 51, 11, 12, 13, 14 };   /* reflectors, plates, nozzles */
static const int ASSEMBLY_END[9]= { 24, 41,  4,  5, 63, 64, 65, 66, 67 };

enum { N_FUEL_SMALL = 34, N_FUEL_LARGE = 321, FIRST_EXTRA_FUEL_NUCLIDE = 68 };

static const struct { const int *ids; int n; } NON_FUEL[XS_NUM_MATERIALS] = {
    { NULL, 0 },           /* 0  fuel: depends on problem size                */
    { CLADDING, 5 },       /* 1  cladding                                     */
    { WATER, 4 },          /* 2  cold borated water                           */
    { WATER, 4 },          /* 3  hot borated water                            */
    { RPV, 27 },           /* 4  reactor pressure vessel                      */
    { STRUCTURE, 21 },     /* 5  lower radial reflector                       */
    { STRUCTURE, 21 },     /* 6  upper reflector / top plate                  */
    { STRUCTURE, 21 },     /* 7  bottom plate                                 */
    { STRUCTURE, 21 },     /* 8  bottom nozzle                                */
    { STRUCTURE, 21 },     /* 9  top nozzle                                   */
    { ASSEMBLY_END, 9 },   /* 10 top of fuel assemblies                       */
    { ASSEMBLY_END, 9 },   /* 11 bottom of fuel assemblies                    */
};

int *load_num_nucs(long n_isotopes)
{
    int *num_nucs = (int *)malloc(XS_NUM_MATERIALS * sizeof(int));
    if (!num_nucs) return NULL;
    /* Only the 68-nuclide "small" problem has the short fuel list; every other size uses
     * the 321-nuclide fuel (cuda/Materials.cu:13-16). */
    num_nucs[0] = (n_isotopes == 68) ? N_FUEL_SMALL : N_FUEL_LARGE;
    for (int m = 1; m < XS_NUM_MATERIALS; m++)
        num_nucs[m] = NON_FUEL[m].n;
    return num_nucs;
}

int *load_mats(int *num_nucs, long n_isotopes, int *max_num_nucs)
{
    int widest = 0;
    for (int m = 0; m < XS_NUM_MATERIALS; m++)
        if (num_nucs[m] > widest) widest = num_nucs[m];
    *max_num_nucs = widest;

    /* Padding entries are zeroed (the reference leaves them uninitialised). */
    int *mats = (int *)calloc((size_t)XS_NUM_MATERIALS * widest, sizeof(int));
    if (!mats) return NULL;

    /* Fuel: the 34 actinides/fission products of H-M small, then -- for the large fuel --
     * every nuclide id from 68 upwards (cuda/Materials.cu:45-52). */
    int n_fuel = num_nucs[0];
    for (int j = 0; j < n_fuel && j < N_FUEL_SMALL; j++)
        mats[j] = FUEL_SMALL[j];
    for (int j = N_FUEL_SMALL; j < n_fuel; j++)
        mats[j] = FIRST_EXTRA_FUEL_NUCLIDE + (j - N_FUEL_SMALL);
    (void)n_isotopes;

    for (int m = 1; m < XS_NUM_MATERIALS; m++)
        memcpy(mats + (size_t)m * widest, NON_FUEL[m].ids, (size_t)num_nucs[m] * sizeof(int));
    return mats;
}

double *load_concs(int *num_nucs, int max_num_nucs)
{
    double *concs = (double *)calloc((size_t)XS_NUM_MATERIALS * max_num_nucs, sizeof(double));
    if (!concs) return NULL;
    /* One LCG stream seeded with 1070^2, consumed in (material, j) order. */
    uint64_t seed = (uint64_t)XS_STARTING_SEED * XS_STARTING_SEED;
    for (int m = 0; m < XS_NUM_MATERIALS; m++)
        for (int j = 0; j < num_nucs[m]; j++)
            concs[(size_t)m * max_num_nucs + j] = LCG_random_double(&seed);
    return concs;
}
