// k_wta.cu -- stage 4: winner-takes-all with parabola refinement for the left view and, from the
// same volume, for the right view (reference: ADCensusStereo.cpp:188-243 and :245-310).
#include "adc_common.cuh"

// Parabola through (best-1, best, best+1), ADCensusStereo.cpp:234-240.  Explicit _rn intrinsics keep
// nvcc from contracting c1 + c2 - 2*min into an FMA.
__device__ __forceinline__ float adc_subpixel(float c1, float c2, float cmin, int best) {
    const float denom = __fsub_rn(__fadd_rn(c1, c2), __fmul_rn(2.0f, cmin));
    if (denom != 0.0f) return __fadd_rn((float)best, __fdiv_rn(__fsub_rn(c1, c2), __fmul_rn(denom, 2.0f)));
    return (float)best;
}

// Left view: one thread per pixel scans its D costs (strict '>' so the first minimum wins).
// Best at either end of the range -> Invalid_Float.
__global__ void __launch_bounds__(128)
k_wta_left(AdcDims dm, const float* __restrict__ vol, float* __restrict__ disp_l) {
    const int pair = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dm.N) return;
    const float* v = vol + (size_t)pair * dm.vol_stride + (size_t)i * dm.Dp;
    float best_cost = ADC_LARGE_F;
    int best = 0;
    const int Q = dm.Dp >> 2;
    for (int q = 0; q < Q; q++) {
        const float4 c = __ldg(reinterpret_cast<const float4*>(v) + q);
        const float cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int di = 4 * q + j;
            if (di < dm.D && best_cost > cc[j]) { best_cost = cc[j]; best = dm.dmin + di; }
        }
    }
    float out;
    if (best == dm.dmin || best == dm.dmax - 1) out = ADC_INVALID_F;
    else out = adc_subpixel(__ldg(v + best - 1 - dm.dmin), __ldg(v + best + 1 - dm.dmin), best_cost, best);
    disp_l[(size_t)pair * dm.N + i] = out;
}

// Right view: cost_R(x,d) = cost_L(x+d,d); columns outside the image are skipped for the minimum
// but count as Large_Float for the parabola (:277-286); a best at either end of the range gives
// the integer disparity, not Invalid (:290-293).  `best` starts at 0, not dmin, as in the reference.
__global__ void __launch_bounds__(128)
k_wta_right(AdcDims dm, const float* __restrict__ vol, float* __restrict__ disp_r) {
    const int pair = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dm.N) return;
    const int y = i / dm.W, x = i - y * dm.W;
    const float* row = vol + (size_t)pair * dm.vol_stride + (size_t)y * dm.W * dm.Dp;
    float best_cost = ADC_LARGE_F;
    int best = 0;
    for (int di = 0; di < dm.D; di++) {
        const int xl = x + dm.dmin + di;
        if (xl >= 0 && xl < dm.W) {
            const float c = __ldg(row + (size_t)xl * dm.Dp + di);
            if (best_cost > c) { best_cost = c; best = dm.dmin + di; }
        }
    }
    float out = (float)best;
    const int i1 = best - 1 - dm.dmin, i2 = best + 1 - dm.dmin;
    if (best != dm.dmin && best != dm.dmax - 1 && i1 >= 0 && i2 < dm.D) {
        const int x1 = x + best - 1, x2 = x + best + 1;
        const float c1 = (x1 >= 0 && x1 < dm.W) ? __ldg(row + (size_t)x1 * dm.Dp + i1) : ADC_LARGE_F;
        const float c2 = (x2 >= 0 && x2 < dm.W) ? __ldg(row + (size_t)x2 * dm.Dp + i2) : ADC_LARGE_F;
        out = adc_subpixel(c1, c2, best_cost, best);
    }
    disp_r[(size_t)pair * dm.N + i] = out;
}

void adc_launch_wta(const AdcParams& P, const AdcWave& w, const float* vol, cudaStream_t st, unsigned long long* launches) {
    dim3 grid((P.dm.N + 127) / 128, w.S);
    k_wta_left<<<grid, 128, 0, st>>>(P.dm, vol, w.disp_l);
    k_wta_right<<<grid, 128, 0, st>>>(P.dm, vol, w.disp_r);
    *launches += 2;
}
