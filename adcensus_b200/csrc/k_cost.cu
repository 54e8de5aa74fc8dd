// k_cost.cu -- stage 1 of the pipeline: gray conversion, 9x7 census transform and the AD-census
// cost volume (reference: cost_computor.cpp:58-121, adcensus_util.cpp:10-53), plus the colour
// difference maps the scanline optimiser consumes (scanline_optimizer.cpp:113-121, 224-235).
#include "adc_common.cuh"

// ---------------------------------------------------------------------------------------------
// gray + census.  One CTA = 32x8 output pixels of one image; the (8+8)x(32+6) gray tile lives in
// shared memory so each gray value is converted once and compared 63 times from on-chip memory.
// Gray is double precision without contraction (r*0.299 + g*0.587 + b*0.114, truncated), which is
// what cost_computor.cpp:69 evaluates; FMA contraction would flip 2933 of the 2^24 inputs.
// ---------------------------------------------------------------------------------------------
#define CT_W 32
#define CT_H 8
#define CT_HX 3
#define CT_HY 4

// (A table of the 3 x 256 possible products, which leaves two double adds per pixel, was measured: the per-CTA copy of the
//  table into shared memory costs more than the three multiplies it saves -- 220 us against 201 us per wave of 32 pairs.)
__device__ __forceinline__ uint8_t gray_of(const uint8_t* __restrict__ px) {
    const double b = (double)__ldg(px), g = (double)__ldg(px + 1), r = (double)__ldg(px + 2);
    const double v = __dadd_rn(__dadd_rn(__dmul_rn(r, 0.299), __dmul_rn(g, 0.587)), __dmul_rn(b, 0.114));
    return (uint8_t)__double2int_rz(v);
}

__global__ void __launch_bounds__(CT_W* CT_H)
k_gray_census(AdcDims dm, const uint8_t* __restrict__ bgr, uint8_t* __restrict__ gray,
              unsigned long long* __restrict__ census, unsigned* __restrict__ bgrx) {
    __shared__ uint8_t tile[CT_H + 2 * CT_HY][CT_W + 2 * CT_HX + 2];
    const int img = blockIdx.z;  // pair*2 + view
    const uint8_t* src = bgr + (size_t)img * dm.N * 3;
    uint8_t* g_out = gray + (size_t)img * dm.N;
    unsigned long long* c_out = census + (size_t)img * dm.N;
    const int x0 = blockIdx.x * CT_W, y0 = blockIdx.y * CT_H;
    const int tid = threadIdx.y * CT_W + threadIdx.x;
    constexpr int TW = CT_W + 2 * CT_HX, TH = CT_H + 2 * CT_HY;
    for (int i = tid; i < TW * TH; i += CT_W * CT_H) {
        const int ty = i / TW, tx = i - ty * TW;
        const int gx = x0 + tx - CT_HX, gy = y0 + ty - CT_HY;
        uint8_t v = 0;
        if (gx >= 0 && gx < dm.W && gy >= 0 && gy < dm.H) v = gray_of(src + ((size_t)gy * dm.W + gx) * 3);
        tile[ty][tx] = v;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= dm.W || y >= dm.H) return;
    const int tx = threadIdx.x + CT_HX, ty = threadIdx.y + CT_HY;
    const uint8_t centre = tile[ty][tx];
    g_out[(size_t)y * dm.W + x] = centre;
    {   // packed copy of the pixel (B | G<<8 | R<<16) for the kernels that compare colours
        const uint8_t* px = src + ((size_t)y * dm.W + x) * 3;
        bgrx[(size_t)img * dm.N + (size_t)y * dm.W + x] = (unsigned)__ldg(px) | ((unsigned)__ldg(px + 1) << 8) | ((unsigned)__ldg(px + 2) << 16);
    }
    unsigned long long bits = 0ull;
    // border pixels keep 0 and tiny images are skipped entirely (adcensus_util.cpp:12,17-18)
    if (dm.W > 9 && dm.H > 7 && y >= 4 && y < dm.H - 4 && x >= 3 && x < dm.W - 3) {
#pragma unroll
        for (int dy = -CT_HY; dy <= CT_HY; dy++)
#pragma unroll
            for (int dx = -CT_HX; dx <= CT_HX; dx++)
                bits = (bits << 1) | (unsigned long long)(tile[ty + dy][tx + dx] < centre);
    }
    c_out[(size_t)y * dm.W + x] = bits;
}

void adc_launch_gray_census(const AdcParams& P, const AdcWave& w, cudaStream_t st, unsigned long long* launches) {
    dim3 grid((P.dm.W + CT_W - 1) / CT_W, (P.dm.H + CT_H - 1) / CT_H, w.S * 2), block(CT_W, CT_H);
    k_gray_census<<<grid, block, 0, st>>>(P.dm, w.bgr, w.gray, w.census, w.bgrx);
    ++*launches;
}

// ---------------------------------------------------------------------------------------------
// AD-census cost volume.  One thread = one pixel x four consecutive disparities -> one 128-bit
// store; consecutive threads cover consecutive disparity quads of the same pixel, then the next
// pixel, so a warp writes 512 contiguous bytes.  The two exp() factors have tiny integer domains
// (sum of abs differences 0..765, Hamming 0..63): they come from tables built on the host with the
// host's libm expf, evaluated in the reference's order  ((1 - e_ad) + 1) - e_cen
// (cost_computor.cpp:110-117), so the volume is bit-identical to the CPU path by construction.
//
// One CTA walks a whole image row in chunks of `ppc` pixels.  Per chunk the right-image span the
// chunk can match is staged in shared memory as packed BGR words + census words, split into four
// arrays by (index mod 4) so that the stride-4 walk of a thread quad is conflict-free.  The tables
// are staged once per CTA, replicated (x32 for the 64-entry census table, x8 for the 766-entry AD
// table) so that the data-dependent lookups of a warp spread over the banks: the first two versions
// of this kernel were bound by L1 / shared-memory bank conflicts on exactly these gathers.
// (Measured and rejected in round 2: lanes = 32 consecutive disparities of one pixel -- half the index arithmetic, but
//  32-bit stores and data-dependent table addresses that differ in every lane: 631 us against 486 us per wave of 32 Cone pairs.)
// ---------------------------------------------------------------------------------------------
#define CV_AD_REP 8

__global__ void __launch_bounds__(1024)
k_cost_volume(AdcDims dm, int ppc, const unsigned* __restrict__ bgrx,
              const unsigned long long* __restrict__ census, float* __restrict__ vol,
              const float* __restrict__ lut_ad, const float* __restrict__ lut_cen) {
    extern __shared__ __align__(16) unsigned char cv_smem[];
    const int pair = blockIdx.y, y = blockIdx.x;
    const int Q = dm.Dp >> 2;                       // threads per pixel
    const int span = dm.W + dm.D - 1;               // right-image columns -(dmax-1)-dmin .. : every xr any pixel of the row can ask for
    const int sq = (span + 3) / 4 + 1;              // entries per residue array (padded)
    const int xr_base = -(dm.D - 1) - dm.dmin;      // image column of staged entry 0 (xr = x - dmin - di, x = 0, di = D-1)
    float* s_ce = reinterpret_cast<float*>(cv_smem);                                      // [64][32]
    float* s_ad = s_ce + 64 * 32;                                                         // [766][CV_AD_REP]
    unsigned long long* s_cen = reinterpret_cast<unsigned long long*>(s_ad + 766 * CV_AD_REP);  // [4][sq]
    unsigned* s_bgr = reinterpret_cast<unsigned*>(s_cen + 4 * sq);                        // [4][sq]
    const unsigned* left = bgrx + (size_t)pair * 2 * dm.N;
    const unsigned* right = left + (size_t)dm.N;
    const unsigned long long* cen_l = census + (size_t)pair * 2 * dm.N;
    const unsigned long long* cen_r = cen_l + dm.N;
    const int row = y * dm.W;
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) s_ce[i] = __ldg(lut_cen + (i >> 5));
    for (int i = threadIdx.x; i < 766 * CV_AD_REP; i += blockDim.x) s_ad[i] = __ldg(lut_ad + i / CV_AD_REP);
    // the whole right-image row, once: packed BGR + census, split by (index mod 4)
    for (int i = threadIdx.x; i < span; i += blockDim.x) {
        const int xr = xr_base + i;
        unsigned long long c = 0ull;
        unsigned pix = 0xffffffffu;                 // marker: outside the image
        if (xr >= 0 && xr < dm.W) {
            c = __ldg(cen_r + row + xr);
            pix = __ldg(right + row + xr);
        }
        s_cen[(i & 3) * sq + (i >> 2)] = c;
        s_bgr[(i & 3) * sq + (i >> 2)] = pix;
    }
    __syncthreads();
    const int p = threadIdx.x / Q, q = threadIdx.x - p * Q;
    if (p >= ppc) return;
    for (int x = p; x < dm.W; x += ppc) {
        const unsigned cl = __ldg(left + row + x);
        const unsigned long long bl = __ldg(cen_l + row + x);
        float out[4];
        const int i0 = (x - dm.dmin - 4 * q) - xr_base;   // staged index of xr = x - dmin - di for di = 4q; in [0, W+D-2] for real disparities
#pragma unroll
        for (int j = 0; j < 4; j++) {
            // branch-free: padding lanes (di >= D) and out-of-image matches compute on clamped operands and are
            // overwritten by selects -- the per-disparity branches used to cost more than the arithmetic
            const int di = 4 * q + j;
            const int i = max(i0 - j, 0);
            const int si = (i & 3) * sq + (i >> 2);
            const unsigned pix = s_bgr[si];
            const int sad = min((int)__vsadu4(cl, pix), 765);           // |dB| + |dG| + |dR| (4th byte is 0 in both)
            const int ham = __popcll(bl ^ s_cen[si]) & 63;
            float c = __fsub_rn(s_ad[sad * CV_AD_REP + (lane & (CV_AD_REP - 1))], s_ce[ham * 32 + lane]);
            c = pix == 0xffffffffu ? 1.0f : c;                          // out-of-image match: cost_computor.cpp:101-104
            out[j] = di < dm.D ? c : 0.0f;                              // padding lane, never read as a cost
        }
        float4* dst = reinterpret_cast<float4*>(vol + (size_t)pair * dm.vol_stride + ((size_t)row + x) * dm.Dp) + q;
        *dst = make_float4(out[0], out[1], out[2], out[3]);
    }
}

void adc_launch_cost(const AdcParams& P, const AdcWave& w, float* vol, cudaStream_t st, unsigned long long* launches) {
    const int Q = P.dm.Dp / 4;
    int ppc = 512 / Q;                              // pixels in flight per CTA
    if (ppc > 32) ppc = 32;
    if (ppc < 1) ppc = 1;
    const int threads = ppc * Q;
    const int span = P.dm.W + P.dm.D - 1, sq = (span + 3) / 4 + 1;
    const size_t smem = (size_t)(64 * 32 + 766 * CV_AD_REP) * 4 + (size_t)4 * sq * 12;
    static AdcOnce attr_once;
    if (adc_once_needed(attr_once)) {
        cudaFuncSetAttribute(k_cost_volume, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        adc_once_done(attr_once);
    }
    dim3 grid(P.dm.H, w.S);
    k_cost_volume<<<grid, threads, smem, st>>>(P.dm, ppc, w.bgrx, w.census, vol, w.lut_ad, w.lut_cen);
    ++*launches;
}

// ---------------------------------------------------------------------------------------------
// Colour-difference maps for the scanline optimiser: max-channel distance between a pixel and
// its predecessor along x (h) or y (v), for the left and the right image.  A forward pass reads
// map[cur]; a backward pass reads map[pixel it came from] (same two pixels, see k_scanline.cu).
// ---------------------------------------------------------------------------------------------
__global__ void k_diffmaps(AdcDims dm, const uint8_t* __restrict__ bgr, uint8_t* __restrict__ dmap) {
    const int pair = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dm.N) return;
    const int y = i / dm.W, x = i - y * dm.W;
    uint8_t* out = dmap + (size_t)pair * 4 * dm.N;
#pragma unroll
    for (int v = 0; v < 2; v++) {
        const uint8_t* img = bgr + ((size_t)pair * 2 + v) * dm.N * 3;
        const uchar3 c = adc_load_bgr(img, i);
        out[(size_t)(2 * v) * dm.N + i] = x > 0 ? (uint8_t)adc_colour_dist(c, adc_load_bgr(img, i - 1)) : 0;
        out[(size_t)(2 * v + 1) * dm.N + i] = y > 0 ? (uint8_t)adc_colour_dist(c, adc_load_bgr(img, i - dm.W)) : 0;
    }
}

void adc_launch_diffmaps(const AdcParams& P, const AdcWave& w, cudaStream_t st, unsigned long long* launches) {
    dim3 grid((P.dm.N + 255) / 256, w.S);
    k_diffmaps<<<grid, 256, 0, st>>>(P.dm, w.bgr, w.dmap);
    ++*launches;
}
