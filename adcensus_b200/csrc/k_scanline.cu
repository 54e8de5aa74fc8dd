// k_scanline.cu -- stage 3: the four chained scanline-optimisation passes
// (reference: scanline_optimizer.cpp:40-61 sequencing, :63-171 horizontal, :173-279 vertical).
//
// One warp owns one scanline (a row for the +-x passes, a column for the +-y passes) and walks it
// serially; the D disparities of a pixel are spread over the 32 lanes, K consecutive ones per
// lane, and live in registers from one step to the next.  Per step the recurrence is
//     L(d) = ( C(d) + min( Lp(d), Lp(d-1)+P1, Lp(d+1)+P1, minLp+P2 ) ) / 2
// (no "- minLp" term, and a "/2": scanline_optimizer.cpp:144-151), with Lp padded by Large_Float
// on both ends and the running minimum taken over the pads too (:96,:107-110).  Neighbour values
// cross lanes through two shuffles; the minimum over d is one REDUX on order-preserving integer
// keys.  Only add/min/exact scalings occur, so the result is bit-identical to the CPU path.
//
// P1/P2 depend on d1 = Dc(left[p], left[p_prev]) and d2 = Dc(right[xr], right[xr_prev]) with
// xr = x - d - dmin.  The reference declares d2 once per pixel (initialised to d1) and only
// overwrites it while 0 < xr < W-1, so for disparities past the valid interval it keeps the value of
// the last valid one ("sticky d2", :116-121).  Closed form used here: valid d form the interval
// [lo,hi] = [max(0,x-dmin-(W-2)), min(D-1,x-dmin-1)]; d<lo -> d1, d in [lo,hi] -> map(x-d-dmin),
// d>hi -> map(x-hi-dmin), empty interval -> d1.
#include "adc_common.cuh"

template <int K>
struct Piece { static constexpr int G = (K % 4 == 0) ? 4 : ((K % 2 == 0) ? 2 : 1); static constexpr int NP = K / G; };

template <int K>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, int lane, int Dp, float (&v)[K]) {
    constexpr int G = Piece<K>::G, NP = Piece<K>::NP;
#pragma unroll
    for (int j = 0; j < NP; j++) {
        const int d = lane * K + j * G;
        if (d < Dp) {
            if (G == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(p + d)); v[j*G] = t.x; v[j*G+1] = t.y; v[j*G+2] = t.z; v[j*G+3] = t.w; }
            else if (G == 2) { const float2 t = __ldg(reinterpret_cast<const float2*>(p + d)); v[j*G] = t.x; v[j*G+1] = t.y; }
            else v[j] = __ldg(p + d);
        } else {
#pragma unroll
            for (int g = 0; g < G; g++) v[j * G + g] = 0.f;
        }
    }
}

template <int K>
__device__ __forceinline__ void store_vec(float* __restrict__ p, int lane, int Dp, const float (&v)[K]) {
    constexpr int G = Piece<K>::G, NP = Piece<K>::NP;
#pragma unroll
    for (int j = 0; j < NP; j++) {
        const int d = lane * K + j * G;
        if (d < Dp) {
            if (G == 4) *reinterpret_cast<float4*>(p + d) = make_float4(v[j*G], v[j*G+1], v[j*G+2], v[j*G+3]);
            else if (G == 2) *reinterpret_cast<float2*>(p + d) = make_float2(v[j*G], v[j*G+1]);
            else p[d] = v[j];
        }
    }
}

#define SO_WARPS 4
#define SO_PF 4   // steps of cost loads kept in flight per warp

template <int K>
__global__ void __launch_bounds__(SO_WARPS * 32)
k_scanline(AdcParams P, const float* __restrict__ src, float* __restrict__ dst,
           const uint8_t* __restrict__ dmap, int sx, int sy) {
    const AdcDims& dm = P.dm;
    const int lane = threadIdx.x & 31;
    const int line = blockIdx.x * SO_WARPS + (threadIdx.x >> 5);
    const int pair = blockIdx.y;
    const int n_lines = sx ? dm.H : dm.W, n_steps = sx ? dm.W : dm.H;
    if (line >= n_lines) return;
    const int W = dm.W, D = dm.D, Dp = dm.Dp, dmin = dm.dmin;
    const bool fwd = (sx + sy) > 0;
    const int pstep = sx + sy * W;  // signed pixel stride along the path
    const uint8_t* ml = dmap + ((size_t)pair * 4 + (sx ? 0 : 1)) * dm.N;
    const uint8_t* mr = dmap + ((size_t)pair * 4 + (sx ? 2 : 3)) * dm.N;
    const float* S = src + (size_t)pair * dm.vol_stride;
    float* O = dst + (size_t)pair * dm.vol_stride;

    int x = sx ? (sx > 0 ? 0 : W - 1) : line;
    int y = sy ? (sy > 0 ? 0 : dm.H - 1) : line;
    int pi = y * W + x;

    bool valid[K];
#pragma unroll
    for (int k = 0; k < K; k++) valid[k] = (lane * K + k) < D;

    // path head: L = C  (scanline_optimizer.cpp:99-100)
    float L[K];
    load_vec<K>(S + (size_t)pi * Dp, lane, Dp, L);
    store_vec<K>(O + (size_t)pi * Dp, lane, Dp, L);
    unsigned key = adc_f2key(ADC_LARGE_F);
#pragma unroll
    for (int k = 0; k < K; k++) {
        if (!valid[k]) L[k] = ADC_LARGE_F;
        key = min(key, adc_f2key(L[k]));
    }
    float minL = adc_key2f(__reduce_min_sync(0xffffffffu, key));

    // software pipeline of the cost loads
    float buf[SO_PF][K];
#pragma unroll
    for (int j = 0; j < SO_PF; j++)
        if (1 + j < n_steps) load_vec<K>(S + (size_t)(pi + (long long)(1 + j) * pstep) * Dp, lane, Dp, buf[j]);

    for (int base = 1; base < n_steps; base += SO_PF) {
#pragma unroll
        for (int j = 0; j < SO_PF; j++) {
            const int step = base + j;
            if (step >= n_steps) break;
            float C[K];
#pragma unroll
            for (int k = 0; k < K; k++) C[k] = buf[j][k];
            if (step + SO_PF < n_steps)
                load_vec<K>(S + (size_t)(pi + (long long)(SO_PF + 1) * pstep) * Dp, lane, Dp, buf[j]);

            const int pi_prev = pi;
            x += sx; y += sy; pi += pstep;
            const int d1 = __ldg(ml + (fwd ? pi : pi_prev));
            const int lo = max(0, x - dmin - (W - 2));
            const int hi = min(D - 1, x - dmin - 1);
            const bool a1 = d1 < P.tso;

            const float up = __shfl_up_sync(0xffffffffu, L[K - 1], 1);
            const float down = __shfl_down_sync(0xffffffffu, L[0], 1);
            const float left_in = lane == 0 ? ADC_LARGE_F : up;
            const float right_in = lane == 31 ? ADC_LARGE_F : down;

            float Ln[K];
            unsigned kmin = adc_f2key(ADC_LARGE_F);
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int d = lane * K + k;
                int d2 = d1;
                if (lo <= hi && d >= lo) {
                    const int xr = x - min(d, hi) - dmin;
                    const int ri = y * W + xr;
                    d2 = __ldg(mr + (fwd ? ri : ri - pstep));
                }
                const bool a2 = d2 < P.tso;
                const float P1 = (a1 && a2) ? P.p1 : ((a1 || a2) ? P.p1_4 : P.p1_10);
                const float P2 = (a1 && a2) ? P.p2 : ((a1 || a2) ? P.p2_4 : P.p2_10);
                const float l1 = L[k];
                const float l2 = __fadd_rn(k > 0 ? L[k - 1] : left_in, P1);
                const float l3 = __fadd_rn(k < K - 1 ? L[k + 1] : right_in, P1);
                const float l4 = __fadd_rn(minL, P2);
                float v = __fadd_rn(C[k], fminf(fminf(l1, l2), fminf(l3, l4)));
                v = __fmul_rn(v, 0.5f);
                Ln[k] = v;
                if (valid[k]) kmin = min(kmin, adc_f2key(v));
            }
            store_vec<K>(O + (size_t)pi * Dp, lane, Dp, Ln);
#pragma unroll
            for (int k = 0; k < K; k++) L[k] = valid[k] ? Ln[k] : ADC_LARGE_F;
            minL = adc_key2f(__reduce_min_sync(0xffffffffu, kmin));
        }
    }
}

int adc_launch_scanline(const AdcParams& P, const AdcWave& w, const float* src, float* dst, int sx, int sy,
                        cudaStream_t st, unsigned long long* launches) {
    const int n_lines = sx ? P.dm.H : P.dm.W;
    dim3 grid((n_lines + SO_WARPS - 1) / SO_WARPS, w.S);
    const int K = (P.dm.Dp + 31) / 32;
    switch (K) {
#define SO_CASE(KK) case KK: k_scanline<KK><<<grid, SO_WARPS * 32, 0, st>>>(P, src, dst, w.dmap, sx, sy); break;
        SO_CASE(1) SO_CASE(2) SO_CASE(3) SO_CASE(4) SO_CASE(5) SO_CASE(6) SO_CASE(7) SO_CASE(8)
#undef SO_CASE
        default: return 1;  // D > 256 not supported by the warp-per-line kernel
    }
    ++*launches;
    return 0;
}
