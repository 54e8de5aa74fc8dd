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

// ---------------------------------------------------------------------------------------------
// One kernel, both views, any disparity range.  A CTA owns WT_PX neighbouring pixels of one image row -- as pixels of
// the left view AND as pixels of the right view -- and sweeps the disparity range in chunks of WT_DC.  Per chunk two
// small tiles are staged in shared memory (coalesced 128-bit loads, 128 contiguous bytes per column):
//   left tile   L[t][k] = cost(x0 + t, d0 + k)                          the chunk of the CTA's own columns
//   right tile  R[k][r] = cost(x0 + r + dmin + d0 + k, d0 + k)          cost_R(xr, d) = cost_L(xr + d, d) (:262-287),
//                                                                       stored skewed: the diagonal a right pixel walks is a
//                                                                       column of R, Large_Float where the column is outside
// and every thread scans its two chunk vectors sequentially with the reference's strict '>' (first minimum wins),
// carrying (minimum, argmin, the two parabola neighbours, the previous cost) in registers from chunk to chunk.
// A column outside the image is Large_Float: never the minimum (the running minimum starts there), and exactly what the
// reference's cost_local holds for the parabola (:277-286).  Shared memory per CTA is independent of the disparity
// range (33 KB), so the kernel keeps its six CTAs per SM at D = 64 as at D = 256 -- the first version staged
// (WT_PX + D - 1) whole columns, 172 KB for a 64-thread CTA at D = 192.
// The right tile's columns are the left tile of the neighbouring CTAs: they come out of L2.
// ---------------------------------------------------------------------------------------------
#define WT_PX 128
#define WT_DC 32
#define WT_LS (WT_DC + 1)      // row stride of the left tile (odd: thread t walks bank t + k)

struct WtaRun {                // running state of one view's scan
    float best, c1, c2, prev;
    int bd;                    // argmin as index d - dmin; -2 = none yet (so that "bd + 1" never matches)
};

__device__ __forceinline__ void wta_step(WtaRun& s, float c, int di) {
    if (di == s.bd + 1) s.c2 = c;                       // the cost right after the current minimum
    if (s.best > c) { s.best = c; s.bd = di; s.c1 = s.prev; }
    s.prev = c;
}

__global__ void __launch_bounds__(WT_PX)
k_wta(AdcDims dm, const float* __restrict__ vol, float* __restrict__ disp_l, float* __restrict__ disp_r) {
    __shared__ float tl[WT_PX * WT_LS];
    __shared__ float tr[WT_DC * WT_PX];
    const int pair = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * WT_PX;
    const int W = dm.W, D = dm.D, Dp = dm.Dp;
    const float* rowv = vol + (size_t)pair * dm.vol_stride + (size_t)y * W * Dp;
    const int t = threadIdx.x;
    const int kq = t & 7, cj = t >> 3;                  // staging role: float4 kq of the chunk, 16 columns per trip
    WtaRun sl, sr;
    sl.best = sr.best = ADC_LARGE_F;                    // min_cost starts at Large_Float (:209, :266)
    sl.c1 = sl.c2 = sr.c1 = sr.c2 = ADC_LARGE_F;
    sl.prev = sr.prev = ADC_LARGE_F;
    sl.bd = sr.bd = -2;
    for (int d0 = 0; d0 < D; d0 += WT_DC) {
        const int dn = min(WT_DC, D - d0);
        const bool qin = d0 + 4 * kq < Dp;              // this float4 exists (Dp is a multiple of 4)
        // ---- left tile: columns x0 .. x0 + WT_PX - 1
#pragma unroll 4
        for (int c = cj; c < WT_PX; c += WT_PX / 8) {
            const int x = x0 + c;
            float4 v = make_float4(ADC_LARGE_F, ADC_LARGE_F, ADC_LARGE_F, ADC_LARGE_F);
            if (qin && x < W) v = __ldg(reinterpret_cast<const float4*>(rowv + (size_t)x * Dp + d0) + kq);
            float* o = tl + c * WT_LS + 4 * kq;
            o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        }
        // ---- right tile: column j of the tile is image column cb + j; its element k belongs to right pixel r = j - k
        const int cb = x0 + dm.dmin + d0;
#pragma unroll 4
        for (int j = cj; j < WT_PX + WT_DC - 1; j += WT_PX / 8) {
            const int x = cb + j;
            float4 v = make_float4(ADC_LARGE_F, ADC_LARGE_F, ADC_LARGE_F, ADC_LARGE_F);
            if (qin && x >= 0 && x < W) v = __ldg(reinterpret_cast<const float4*>(rowv + (size_t)x * Dp + d0) + kq);
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int k = 4 * kq + i, r = j - k;
                if (r >= 0 && r < WT_PX) tr[k * WT_PX + r] = e[i];
            }
        }
        __syncthreads();
        // ---- scans: thread t = left pixel x0 + t and right pixel x0 + t
        const float* pl = tl + t * WT_LS;
        const float* pr = tr + t;
        if (dn == WT_DC) {
#pragma unroll 8
            for (int k = 0; k < WT_DC; k++) {
                wta_step(sl, pl[k], d0 + k);
                wta_step(sr, pr[k * WT_PX], d0 + k);
            }
        } else {
            for (int k = 0; k < dn; k++) {
                wta_step(sl, pl[k], d0 + k);
                wta_step(sr, pr[k * WT_PX], d0 + k);
            }
        }
        __syncthreads();
    }
    const int x = x0 + t;
    if (x >= W) return;
    const size_t o = (size_t)pair * dm.N + (size_t)y * W + x;
    {   // left view: a minimum at either end of the range (or none) is Invalid (ADCensusStereo.cpp:224-227)
        float out = ADC_INVALID_F;
        if (sl.bd > 0 && sl.bd < D - 1) out = adc_subpixel(sl.c1, sl.c2, sl.best, dm.dmin + sl.bd);
        disp_l[o] = out;
    }
    {   // right view: a minimum at either end gives the integer disparity, not Invalid (:290-293); `best` starts at 0
        // (not dmin) when no column was valid, as in the reference (:271)
        float out = 0.0f;
        if (sr.bd >= 0) {
            const int best = dm.dmin + sr.bd;
            out = (sr.bd > 0 && sr.bd < D - 1) ? adc_subpixel(sr.c1, sr.c2, sr.best, best) : (float)best;
        }
        disp_r[o] = out;
    }
}

int adc_launch_wta(const AdcParams& P, const AdcWave& w, const float* vol, cudaStream_t st, unsigned long long* launches) {
    dim3 grid((P.dm.W + WT_PX - 1) / WT_PX, P.dm.H, w.S);
    k_wta<<<grid, WT_PX, 0, st>>>(P.dm, vol, w.disp_l, w.disp_r);
    ++*launches;
    return 0;
}
