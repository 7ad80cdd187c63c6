/*
 * lcg.c -- the 63-bit linear congruential generator every XSBench stream is built from,
 * and the material sampler.  Host copies; the device versions live in csrc/xs_device.cuh.
 *
 * Reference behaviour: cuda/Simulation.cu:326-334 (LCG_random_double), :336-362
 * (fast_forward_LCG), :287-324 (pick_mat).
 */
#include "xs_host.h"

static const uint64_t LCG_MULT = 2806196910506780709ULL;
static const uint64_t LCG_MOD_MASK = (1ULL << 63) - 1ULL;   /* x mod 2^63 */

/* x <- (a*x + 1) mod 2^63, returned as x / 2^63 in [0, 1]. */
double LCG_random_double(uint64_t *seed)
{
    uint64_t x = (LCG_MULT * (*seed) + 1ULL) & LCG_MOD_MASK;
    *seed = x;
    return (double)x * 0x1p-63;   /* int->f64 rounds once; the 2^-63 scale is exact */
}

/* State after n further steps.  An n-step jump is the affine map x -> M x + C; build it
 * from the binary digits of n by doubling the one-step map (all arithmetic wraps mod 2^64,
 * the final mask reduces to mod 2^63, which divides 2^64). */
uint64_t fast_forward_LCG(uint64_t seed, uint64_t n)
{
    uint64_t step_m = LCG_MULT, step_c = 1ULL;
    uint64_t M = 1ULL, C = 0ULL;
    for (n &= LCG_MOD_MASK; n != 0; n >>= 1) {
        if (n & 1ULL) {
            M *= step_m;
            C = C * step_m + step_c;
        }
        step_c *= step_m + 1ULL;
        step_m *= step_m;
    }
    return (M * seed + C) & LCG_MOD_MASK;
}

/* Volume fractions of the 12 Hoogenboom-Martin materials (fuel first). */
static const double MATERIAL_FRACTION[XS_NUM_MATERIALS] = {
    0.140, 0.052, 0.275, 0.134, 0.154, 0.064, 0.066, 0.055, 0.008, 0.015, 0.025, 0.013
};

/* thr[i] = frac[i] + frac[i-1] + ... + frac[1], added in exactly that order (the reference
 * re-accumulates from i downwards for every candidate, so the rounding of each threshold
 * depends on this order); thr[0] = 0, i.e. fuel is only reachable as the fall-through. */
void xs_material_thresholds(double thr[XS_NUM_MATERIALS])
{
    for (int i = 0; i < XS_NUM_MATERIALS; i++) {
        double acc = 0.0;
        for (int j = i; j >= 1; j--)
            acc += MATERIAL_FRACTION[j];
        thr[i] = acc;
    }
}

int pick_mat(uint64_t *seed)
{
    double thr[XS_NUM_MATERIALS];
    xs_material_thresholds(thr);

    const double roll = LCG_random_double(seed);
    for (int i = 0; i < XS_NUM_MATERIALS; i++)
        if (roll < thr[i])
            return i;
    return 0;
}
