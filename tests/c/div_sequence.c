/* CPU check of the division sequence adc_div4 (adcensus_b200/csrc/k_aggregate.cu) runs on the device:
 *     r1 = fma(r0, fma(-n, r0, 1), r0);  q0 = r1 * x;  e = fma(-n, q0, x);  q = fma(r1, e, q0)
 * with r0 = the hardware's approximate reciprocal of n.  MUFU.RCP is not available here, so the check is made for every
 * r0 within `spread` ulps of the correctly rounded 1/n: for all divisors the kernel can meet (support counts 1..65535)
 * and a dense sample of numerators, the sequence must give exactly the IEEE quotient x / n.  fmaf() is exact on the CPU as
 * FFMA.RN is on the GPU.  Prints the number of mismatches per perturbation. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static float nudge(float v, int ulps) { uint32_t u; memcpy(&u, &v, 4); u += ulps; memcpy(&v, &u, 4); return v; }

int main(int argc, char** argv) {
    const int per_n = argc > 1 ? atoi(argv[1]) : 400, spread = argc > 2 ? atoi(argv[2]) : 2;
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    long bad[9] = {0};
    for (int n = 1; n <= 65535; n++) {
        const float fn = (float)n;
        const float exact_r = 1.0f / fn;
        for (int k = 0; k < per_n; k++) {
            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
            uint32_t bits = (uint32_t)(rng >> 32);
            float x;
            if (k < 8) { static const float fixed[8] = {0.0f, 5.9604645e-8f, 1.0f, 1.8734f, 1e-6f, 3.0e-27f, 9000.0f, 0.33333334f}; x = fixed[k]; }
            else { x = ldexpf((float)(bits & 0xffffff) / 16777216.0f + 1.0f, (int)((bits >> 24) % 36) - 22); }   /* 2^-22 .. 2^14 */
            const float want = x / fn;
            for (int d = -spread; d <= spread; d++) {
                const float r0 = nudge(exact_r, d);
                const float r1 = fmaf(r0, fmaf(-fn, r0, 1.0f), r0);
                const float q0 = fmaf(r1, x, 0.0f);
                const float e = fmaf(-fn, q0, x);
                const float q = fmaf(r1, e, q0);
                if (memcmp(&q, &want, 4) != 0) bad[d + spread]++;
            }
        }
    }
    long total = 0;
    for (int d = -spread; d <= spread; d++) { printf("r0 = RN(1/n) %+d ulp: %ld mismatches\n", d, bad[d + spread]); total += bad[d + spread]; }
    printf("total %ld\n", total);
    return total != 0;
}
