// k_wta.cu -- stage 4: winner-takes-all with parabola refinement for the left view and, from the
// same volume, for the right view (reference: ADCensusStereo.cpp:188-243 and :245-310).
#include "adc_common.cuh"
#include <stdlib.h>

// Parabola through (best-1, best, best+1), ADCensusStereo.cpp:234-240.  Explicit _rn intrinsics keep
// nvcc from contracting c1 + c2 - 2*min into an FMA.
__device__ __forceinline__ float adc_subpixel(float c1, float c2, float cmin, int best) {
    const float denom = __fsub_rn(__fadd_rn(c1, c2), __fmul_rn(2.0f, cmin));
    if (denom != 0.0f) return __fadd_rn((float)best, __fdiv_rn(__fsub_rn(c1, c2), __fmul_rn(denom, 2.0f)));
    return (float)best;
}

// One pass over the volume produces both views.  A CTA takes WT_PX pixels of one row; every thread
// holds four consecutive costs of one pixel (one 128-bit load, the warp reads 512 contiguous bytes).
//   left view : the pixel's minimum is a shared-memory atomicMin over 64-bit keys (ordered cost bits
//               << 32 | disparity index), which is exactly "strict >, first minimum wins";
//   right view: cost_R(xr, d) = cost_L(xr + d, d), so the same cost is also a candidate for right
//               pixel x - d: a second atomicMin into a CTA-local array over the right pixels the CTA
//               can touch, flushed with one global atomicMin per touched right pixel.
// A small second kernel turns the right view's keys into disparities (two gathers for the parabola).
#define WT_PX 64

__device__ __forceinline__ unsigned long long wta_key(float c, int di) {
    return ((unsigned long long)adc_f2key(c) << 32) | (unsigned)di;
}

__global__ void __launch_bounds__(1024)
k_wta_scan(AdcDims dm, const float* __restrict__ vol, float* __restrict__ disp_l, unsigned long long* __restrict__ rkey) {
    extern __shared__ unsigned long long wt_smem[];
    const int pair = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * WT_PX;
    const int Q = dm.Dp >> 2;
    const int ppi = blockDim.x / Q;                  // pixels per inner iteration
    const int span = WT_PX + dm.D - 1;
    const int xr_base = x0 - (dm.dmax - 1);
    unsigned long long* s_left = wt_smem;            // [ppi]
    unsigned long long* s_right = wt_smem + ppi;     // [span]
    const unsigned long long NONE = ~0ull;
    for (int i = threadIdx.x; i < span; i += blockDim.x) s_right[i] = NONE;
    const int p = threadIdx.x / Q, q = threadIdx.x - p * Q;
    const float* rowv = vol + (size_t)pair * dm.vol_stride + (size_t)y * dm.W * dm.Dp;
    for (int it = 0; it < WT_PX; it += ppi) {
        if (threadIdx.x < ppi) s_left[threadIdx.x] = NONE;
        __syncthreads();
        const int x = x0 + it + p;
        if (p < ppi && x < dm.W) {
            const float4 c4 = __ldg(reinterpret_cast<const float4*>(rowv + (size_t)x * dm.Dp) + q);
            const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
            unsigned long long best = NONE;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int di = 4 * q + j;
                if (di < dm.D && cc[j] < ADC_LARGE_F) {          // min_cost starts at Large_Float, strict '>'
                    const unsigned long long k = wta_key(cc[j], di);
                    best = min(best, k);
                    const int xr = x - (dm.dmin + di);
                    if (xr >= 0 && xr < dm.W) atomicMin(&s_right[xr - xr_base], k);
                }
            }
            if (best != NONE) atomicMin(&s_left[p], best);
        }
        __syncthreads();
        if (threadIdx.x < ppi) {
            const int xx = x0 + it + threadIdx.x;
            if (xx < dm.W) {
                const unsigned long long k = s_left[threadIdx.x];
                float out = ADC_INVALID_F;
                if (k != NONE) {
                    const int di = (int)(unsigned)k, best = dm.dmin + di;
                    if (best != dm.dmin && best != dm.dmax - 1) {   // ends of the range -> Invalid (ADCensusStereo.cpp:224-227)
                        const float* v = rowv + (size_t)xx * dm.Dp;
                        out = adc_subpixel(__ldg(v + di - 1), __ldg(v + di + 1), adc_key2f((unsigned)(k >> 32)), best);
                    }
                }
                disp_l[(size_t)pair * dm.N + y * dm.W + xx] = out;
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < span; i += blockDim.x) {
        const unsigned long long k = s_right[i];
        if (k != NONE) atomicMin(rkey + (size_t)pair * dm.N + y * dm.W + xr_base + i, k);
    }
}

// Right view finish: columns outside the image count as Large_Float for the parabola (:277-286); a
// best at either end of the range gives the integer disparity, not Invalid (:290-293); `best` starts
// at 0 (not dmin) when no column was valid, as in the reference.
__global__ void __launch_bounds__(256)
k_wta_right_finish(AdcDims dm, const float* __restrict__ vol, const unsigned long long* __restrict__ rkey,
                   float* __restrict__ disp_r) {
    const int pair = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dm.N) return;
    const int y = i / dm.W, x = i - y * dm.W;
    const float* row = vol + (size_t)pair * dm.vol_stride + (size_t)y * dm.W * dm.Dp;
    const unsigned long long k = rkey[(size_t)pair * dm.N + i];
    int best = 0;
    float best_cost = ADC_LARGE_F;
    if (k != ~0ull) { best = dm.dmin + (int)(unsigned)k; best_cost = adc_key2f((unsigned)(k >> 32)); }
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

// ---------------------------------------------------------------------------------------------
// Row-walking kernel (experimental, ADC_WTA_MODE=1; 490 us per wave of 16 Cone pairs against 312 us of the tile kernel,
// both issue-bound).  k_wta_tile below turned out to be bound by instruction issue
// (one thread per pixel scanning its D costs one by one out of shared memory, ~1400 thread instructions per
// pixel for the two views).  Here a warp walks along an image row, lane = disparity:
//   left view : the pixel's D costs sit in the lanes (d = lane + 32 j); non-negative floats order like their
//               bit patterns, so the minimum is one REDUX.MIN and "first minimum wins" is the lowest set bit of a
//               ballot (ADCensusStereo.cpp:211-222 -- a cost that is not below Large_Float is never taken);
//   right view: cost_R(xr, d) = cost_L(xr + d + dmin, d) (:262-287).  Lane d holds the running (min, argmin) of the
//               right pixel xr = x - dmin - d it currently serves; when the walk advances to x + 1 that right pixel
//               is served by lane d + 1, so the running pairs move up one lane per step (a systolic chain through the
//               warp, 32 disparities per register; the last lane of chain j feeds lane 0 of chain j + 1).  A right
//               pixel leaves the chain complete at d = D - 1.  Columns outside the image simply contribute
//               nothing, as in the reference, and the walk runs D - 1 steps past the row end to drain the chain.
// Every cost is read exactly once, coalesced (the two parabola neighbours of each view are re-read from L1/L2).
// ---------------------------------------------------------------------------------------------
template <int NCH>   // NCH = ceil(D / 32) chains
__global__ void __launch_bounds__(128)
k_wta_walk(AdcDims dm, int n_pairs, int seg_len, int n_seg, const float* __restrict__ vol, float* __restrict__ disp_l,
           float* __restrict__ disp_r) {
    const int lane = threadIdx.x & 31;
    long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int y = (int)(gw % dm.H); gw /= dm.H;
    const int seg = (int)(gw % n_seg);
    const int pair = (int)(gw / n_seg);
    if (pair >= n_pairs) return;
    const int W = dm.W, D = dm.D, Dp = dm.Dp, dmin = dm.dmin;
    const int a = seg * seg_len, b = min(W, a + seg_len);            // this warp's output columns (both views)
    const float* rowv = vol + (size_t)pair * dm.vol_stride + (size_t)y * W * Dp;
    float* out_l = disp_l + (size_t)pair * dm.N + (size_t)y * W;
    float* out_r = disp_r + (size_t)pair * dm.N + (size_t)y * W;
    const unsigned LARGE_BITS = __float_as_uint(ADC_LARGE_F);
    // the right pixels a..b-1 collect their candidates at x = xr + dmin + d, d = 0..D-1; the left outputs need x = a..b-1
    const int x_begin = min(a, a + dmin), x_end = max(b - 1, b - 1 + dmin + D - 1);
    float rc[NCH];      // running minimum of the right pixel served by (chain j, this lane)
    int rb[NCH];        // its disparity index
#pragma unroll
    for (int j = 0; j < NCH; j++) { rc[j] = ADC_LARGE_F; rb[j] = -1; }
    const int last_j = (D - 1) >> 5, last_lane = (D - 1) & 31;
    float pl_cost = ADC_LARGE_F, pr_cost = ADC_LARGE_F;   // parked results of this lane's column (left / right view)
    int pl_best = 0, pr_best = -1;
    // four columns per trip: their loads go out together (nothing of a later column's load depends on the chains),
    // and the columns of the trip after that are pulled into L2 meanwhile
    constexpr int UX = 4;
    for (int xb = x_begin; xb <= x_end; xb += UX) {
      float cc[UX][NCH];
#pragma unroll
      for (int u = 0; u < UX; u++) {
          const int xx = xb + u;
          const bool in2 = xx >= 0 && xx < W && xx <= x_end;
#pragma unroll
          for (int j = 0; j < NCH; j++) {
              const int d = lane + 32 * j;
              cc[u][j] = (in2 && d < D) ? __ldg(rowv + (size_t)xx * Dp + d) : ADC_INVALID_F;
          }
      }
      {
          const int xp = xb + 4 * UX + (lane >> 3);                 // 4 columns further on, 8 lanes (x 32 floats) per column
          if (xp >= 0 && xp < W && (lane & 7) * 32 < Dp) asm volatile("prefetch.global.L2 [%0];" ::"l"(rowv + (size_t)xp * Dp + (lane & 7) * 32));
      }
#pragma unroll
      for (int u = 0; u < UX; u++) {
        const int x = xb + u;
        if (x > x_end) break;
        const bool inside = x >= 0 && x < W;
        float c[NCH];
#pragma unroll
        for (int j = 0; j < NCH; j++) c[j] = cc[u][j];
        // ---- right view: shift the chains one lane up, then merge this column's candidates
        float carry_c = ADC_LARGE_F;   // enters lane 0 of chain 0: a fresh right pixel
        int carry_b = -1;
#pragma unroll
        for (int j = 0; j < NCH; j++) {
            const float tc = __shfl_sync(0xffffffffu, rc[j], (lane + 31) & 31);
            const int tb = __shfl_sync(0xffffffffu, rb[j], (lane + 31) & 31);
            rc[j] = lane == 0 ? carry_c : tc;
            rb[j] = lane == 0 ? carry_b : tb;
            carry_c = tc; carry_b = tb;              // (only lane 0's copy is used: it received lane 31's pair)
            if (inside && lane + 32 * j < D && rc[j] > c[j]) { rc[j] = c[j]; rb[j] = lane + 32 * j; }   // strict '>'
        }
        // ---- results are parked in the lanes (lane = output column mod 32) and finished 32 at a time by the whole warp:
        //      the parabola, its two neighbour loads and the store would otherwise run on a single lane every step
        {   // the right pixel that just took its last candidate (d = D - 1)
            const int xr = x - dmin - (D - 1);
            if (xr >= a && xr < b) {
                float rcl = rc[0];
                int rbl = rb[0];
#pragma unroll
                for (int j = 1; j < NCH; j++) if (j == last_j) { rcl = rc[j]; rbl = rb[j]; }
                rcl = __shfl_sync(0xffffffffu, rcl, last_lane);
                rbl = __shfl_sync(0xffffffffu, rbl, last_lane);
                const int k = (xr - a) & 31;
                if (lane == k) { pr_cost = rcl; pr_best = rbl; }
                if (k == 31 || xr == b - 1) {
                    const int xo = xr - k + lane;                     // this lane's right pixel
                    if (lane <= k) {
                        // best starts at 0 (not dmin) when no column was valid, as in the reference (:271)
                        const int best = pr_best >= 0 ? dmin + pr_best : 0;
                        const float best_cost = pr_best >= 0 ? pr_cost : ADC_LARGE_F;
                        float o = (float)best;
                        const int i1 = best - 1 - dmin, i2 = best + 1 - dmin;
                        if (best != dmin && best != dm.dmax - 1 && i1 >= 0 && i2 < D) {
                            const int x1 = xo + best - 1, x2 = xo + best + 1;
                            const float c1 = (x1 >= 0 && x1 < W) ? __ldg(rowv + (size_t)x1 * Dp + i1) : ADC_LARGE_F;   // (:277-286)
                            const float c2 = (x2 >= 0 && x2 < W) ? __ldg(rowv + (size_t)x2 * Dp + i2) : ADC_LARGE_F;
                            o = adc_subpixel(c1, c2, best_cost, best);
                        }
                        out_r[xo] = o;
                    }
                }
            }
        }
        // ---- left view of column x
        if (x >= a && x < b) {
            unsigned um = 0xffffffffu;
#pragma unroll
            for (int j = 0; j < NCH; j++) um = min(um, (lane + 32 * j < D) ? __float_as_uint(c[j]) : 0xffffffffu);
            const unsigned m = __reduce_min_sync(0xffffffffu, um);
            int best = 0;                                 // stays 0 when no cost is below Large_Float (:209, :218)
            float best_cost = ADC_LARGE_F;
            if (m < LARGE_BITS) {
                int di = -1;
#pragma unroll
                for (int j = 0; j < NCH; j++) {
                    const unsigned bal = __ballot_sync(0xffffffffu, lane + 32 * j < D && __float_as_uint(c[j]) == m);
                    if (di < 0 && bal) di = 32 * j + __ffs(bal) - 1;
                }
                best = dmin + di;
                best_cost = __uint_as_float(m);
            }
            const int k = (x - a) & 31;
            if (lane == k) { pl_cost = best_cost; pl_best = best; }
            if (k == 31 || x == b - 1) {
                const int xo = x - k + lane;
                if (lane <= k) {
                    float o = ADC_INVALID_F;
                    const int i1 = pl_best - 1 - dmin, i2 = pl_best + 1 - dmin;
                    if (pl_best != dmin && pl_best != dm.dmax - 1 && i1 >= 0 && i2 < D) {
                        const float* v = rowv + (size_t)xo * Dp;
                        o = adc_subpixel(__ldg(v + i1), __ldg(v + i2), pl_cost, pl_best);
                    }
                    out_l[xo] = o;
                }
            }
        }
      }
    }
}

template <int NCH>
static void launch_wta_walk(const AdcParams& P, const AdcWave& w, const float* vol, cudaStream_t st) {
    const long long rows = (long long)w.S * P.dm.H;
    int n_seg = (int)((148 * 40 + rows / 2) / rows);          // ~40 warps per SM over the whole launch
    if (n_seg < 1) n_seg = 1;
    if (n_seg > (P.dm.W + 63) / 64) n_seg = (P.dm.W + 63) / 64;
    const int seg_len = (P.dm.W + n_seg - 1) / n_seg;
    n_seg = (P.dm.W + seg_len - 1) / seg_len;
    const long long warps = rows * n_seg;
    k_wta_walk<NCH><<<(unsigned)((warps + 3) / 4), 128, 0, st>>>(P.dm, w.S, seg_len, n_seg, vol, w.disp_l, w.disp_r);
}

// ---------------------------------------------------------------------------------------------
// Fast path (Dp <= 192): no atomics.  A CTA stages the costs of WT_PX + D - 1 neighbouring columns of
// one row in shared memory (coalesced 128-bit loads, row stride Dp+1 words so that both scans below
// are bank-conflict free), then one thread per pixel scans d = 0..D-1 sequentially -- for the left
// view down its own column vector, for the right view along the diagonal cost_L(xr + d, d) -- with
// the reference's strict '>' comparison.  The halo columns are read twice (second time from L2).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_wta_tile(AdcDims dm, int wpx, int3 pf, const float* __restrict__ vol, float* __restrict__ disp_l, float* __restrict__ disp_r) {
    extern __shared__ float wt_tile[];
    const int pair = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * wpx;
    const int Q = dm.Dp >> 2, DS = dm.Dp + 1;
    const int col_lo = x0 + min(0, dm.dmin);
    const int col_hi = x0 + wpx - 1 + max(0, dm.dmax - 1);
    const int ncols = col_hi - col_lo + 1;
    const float* rowv = vol + (size_t)pair * dm.vol_stride + (size_t)y * dm.W * dm.Dp;
    {   // warm L2 with the core columns of the CTA that runs ~one wave of CTAs later in launch order (pf = that
        // displacement decomposed into block coordinates by the host: three carries instead of 64-bit divisions)
        int bx2 = blockIdx.x + pf.x, by2 = blockIdx.y + pf.y, bz2 = blockIdx.z + pf.z;
        if (bx2 >= (int)gridDim.x) { bx2 -= gridDim.x; by2++; }
        if (by2 >= (int)gridDim.y) { by2 -= gridDim.y; bz2++; }
        if (pf.x >= 0 && bz2 < (int)gridDim.z) {
            const float* r2 = vol + (size_t)bz2 * dm.vol_stride + ((size_t)by2 * dm.W + (size_t)bx2 * wpx) * dm.Dp;
            const int lines = min(wpx, dm.W - bx2 * wpx) * dm.Dp / 32;      // 128-byte lines of the core tile
            for (int i = threadIdx.x; i < lines; i += blockDim.x) asm volatile("prefetch.global.L2 [%0];" ::"l"(r2 + (size_t)i * 32));
        }
    }
    {   // one division per thread instead of one per element: a thread keeps its quad and strides over the columns
        const bool even = (int)blockDim.x % Q == 0;
        const int cstep = even ? (int)blockDim.x / Q : 1;
        for (int i = threadIdx.x; i < ncols * Q; i += even ? cstep * Q : (int)blockDim.x) {
            const int c = i / Q, q = i - c * Q;
            for (int cc = c; cc < (even ? ncols : c + 1); cc += cstep) {
                const int x = col_lo + cc;
                if (x >= 0 && x < dm.W) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(rowv + (size_t)x * dm.Dp) + q);
                    float* t = wt_tile + cc * DS + 4 * q;
                    t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
                }
            }
            if (even) break;
        }
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t < wpx) {                                   // ---- left view (ADCensusStereo.cpp:188-243)
        const int x = x0 + t;
        if (x < dm.W) {
            const float* v = wt_tile + (x - col_lo) * DS;
            float best_cost = ADC_LARGE_F;
            int best = 0;
#pragma unroll 8
            for (int di = 0; di < dm.D; di++) {
                const float c = v[di];
                if (best_cost > c) { best_cost = c; best = dm.dmin + di; }
            }
            float out = ADC_INVALID_F;
            const int i1 = best - 1 - dm.dmin, i2 = best + 1 - dm.dmin;
            if (best != dm.dmin && best != dm.dmax - 1 && i1 >= 0 && i2 < dm.D)
                out = adc_subpixel(v[i1], v[i2], best_cost, best);
            disp_l[(size_t)pair * dm.N + y * dm.W + x] = out;
        }
    } else {                                           // ---- right view (ADCensusStereo.cpp:245-310)
        const int x = x0 + t - wpx;
        if (x < dm.W) {
            float best_cost = ADC_LARGE_F;
            int best = 0;
#pragma unroll 8
            for (int di = 0; di < dm.D; di++) {
                const int xl = x + dm.dmin + di;
                if (xl >= 0 && xl < dm.W) {
                    const float c = wt_tile[(xl - col_lo) * DS + di];
                    if (best_cost > c) { best_cost = c; best = dm.dmin + di; }
                }
            }
            float out = (float)best;
            const int i1 = best - 1 - dm.dmin, i2 = best + 1 - dm.dmin;
            if (best != dm.dmin && best != dm.dmax - 1 && i1 >= 0 && i2 < dm.D) {
                const int x1 = x + best - 1, x2 = x + best + 1;
                const float c1 = (x1 >= 0 && x1 < dm.W) ? wt_tile[(x1 - col_lo) * DS + i1] : ADC_LARGE_F;
                const float c2 = (x2 >= 0 && x2 < dm.W) ? wt_tile[(x2 - col_lo) * DS + i2] : ADC_LARGE_F;
                out = adc_subpixel(c1, c2, best_cost, best);
            }
            disp_r[(size_t)pair * dm.N + y * dm.W + x] = out;
        }
    }
}

int adc_launch_wta(const AdcParams& P, const AdcWave& w, const float* vol, cudaStream_t st, unsigned long long* launches) {
    static int mode = -1;   // development switch ADC_WTA_MODE: 0 = tile / atomic kernels (default), 1 = row-walking kernel (measured slower so far)
    if (mode < 0) { const char* m = getenv("ADC_WTA_MODE"); mode = m ? atoi(m) : 0; }
    if (mode == 1 && P.dm.D <= 256) {
        switch ((P.dm.D + 31) / 32) {
            case 1: launch_wta_walk<1>(P, w, vol, st); break;
            case 2: launch_wta_walk<2>(P, w, vol, st); break;
            case 3: launch_wta_walk<3>(P, w, vol, st); break;
            case 4: launch_wta_walk<4>(P, w, vol, st); break;
            case 5: launch_wta_walk<5>(P, w, vol, st); break;
            case 6: launch_wta_walk<6>(P, w, vol, st); break;
            case 7: launch_wta_walk<7>(P, w, vol, st); break;
            default: launch_wta_walk<8>(P, w, vol, st); break;
        }
        ++*launches;
        return 0;
    }
    const int extra = (P.dm.dmax - 1 > 0 ? P.dm.dmax - 1 : 0) - (P.dm.dmin < 0 ? P.dm.dmin : 0);
    int wpx = 128;                                   // output pixels per CTA: as many as keep the tile small
    while (wpx > 32 && (size_t)(wpx + extra) * (P.dm.Dp + 1) * sizeof(float) > 64 * 1024) wpx >>= 1;
    const size_t tile_bytes = (size_t)(wpx + extra) * (P.dm.Dp + 1) * sizeof(float);
    if (tile_bytes <= 200 * 1024) {
        static bool attr_done[64] = {};
        if (adc_first_time_on_device(attr_done)) {
            cudaFuncSetAttribute(k_wta_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        }
        dim3 grid((P.dm.W + wpx - 1) / wpx, P.dm.H, w.S);
        const long long pfd = 148 * 4;
        const int3 pf = make_int3((int)(pfd % grid.x), (int)((pfd / grid.x) % grid.y), (int)(pfd / grid.x / grid.y));
        k_wta_tile<<<grid, 2 * wpx, tile_bytes, st>>>(P.dm, wpx, pf, vol, w.disp_l, w.disp_r);
        ++*launches;
        return 0;
    }
    // wide disparity ranges: atomic-key version
    const int Q = P.dm.Dp / 4;
    int ppi = 1024 / Q;
    if (ppi > WT_PX) ppi = WT_PX;
    if (ppi < 1) return 1;
    while (WT_PX % ppi) ppi--;                       // WT_PX is a power of two; keeps the iteration count whole
    const int threads = ppi * Q;
    const size_t smem = (size_t)(ppi + WT_PX + P.dm.D - 1) * sizeof(unsigned long long);
    if (cudaMemsetAsync(w.wta_key, 0xff, (size_t)w.S * P.dm.N * sizeof(unsigned long long), st) != cudaSuccess) return 1;
    dim3 grid((P.dm.W + WT_PX - 1) / WT_PX, P.dm.H, w.S);
    k_wta_scan<<<grid, threads, smem, st>>>(P.dm, vol, w.disp_l, w.wta_key);
    dim3 grid2((P.dm.N + 255) / 256, w.S);
    k_wta_right_finish<<<grid2, 256, 0, st>>>(P.dm, vol, w.wta_key, w.disp_r);
    *launches += 2;
    return 0;
}
