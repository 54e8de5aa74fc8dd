// A caller written against the REFERENCE's public interface (ADCensusStereo.h:14-95, usage as in main.cpp:80-118),
// built against this repo's include/ and lib by the test-suite.
//   dropin_main                                   truth table only (no image files): used by the CPU-side test
//   dropin_main left.bgr right.bgr W H dmin dmax out.f32
//                                                 packed BGR u8 in (main.cpp:61-76), float32 map out: the GPU test feeds the
//                                                 Cone pair and hashes the map against the reference's
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ADCensusStereo.h"

static bool read_file(const char* path, std::vector<uint8>& buf) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    const size_t n = std::fread(buf.data(), 1, buf.size(), f);
    std::fclose(f);
    return n == buf.size();
}

int main(int argc, char** argv) {
    sint32 width = 64, height = 48;
    ADCensusOption ad_option;                 // defaults must be the reference's
    if (ad_option.max_disparity != 64 || ad_option.cross_L1 != 34 || ad_option.so_p2 != 3.0f || !ad_option.do_filling) return 10;
    ad_option.min_disparity = 0;
    ad_option.max_disparity = 16;
    const bool files = argc >= 8;
    if (files) {
        width = std::atoi(argv[3]); height = std::atoi(argv[4]);
        ad_option.min_disparity = std::atoi(argv[5]); ad_option.max_disparity = std::atoi(argv[6]);
    }
    std::vector<uint8> left((size_t)width * height * 3, 90), right((size_t)width * height * 3, 90);
    if (files && (!read_file(argv[1], left) || !read_file(argv[2], right))) return 20;
    std::vector<float32> disparity((size_t)width * height, -1.0f);
    ADCensusStereo ad_census;
    if (ad_census.Match(left.data(), right.data(), disparity.data())) return 11;      // before Initialize -> false
    ADCensusOption bad = ad_option; bad.max_disparity = bad.min_disparity;
    if (ad_census.Initialize(width, height, bad)) return 12;                          // empty range -> false
    if (ad_census.Initialize(0, height, ad_option)) return 13;                        // bad size -> false
    const bool have_gpu = ad_census.Initialize(width, height, ad_option);
    if (!have_gpu) { std::printf("DROPIN_NO_GPU\n"); return 0; }                      // no device: Initialize is false, no fallback
    if (ad_census.Match(nullptr, right.data(), disparity.data())) return 14;          // null pointer -> false
    if (!ad_census.Match(left.data(), right.data(), disparity.data())) return 15;
    if (files) {
        FILE* f = std::fopen(argv[7], "wb");
        if (!f || std::fwrite(disparity.data(), sizeof(float32), disparity.size(), f) != disparity.size()) return 21;
        std::fclose(f);
    }
    if (!ad_census.Reset(width, height, ad_option)) return 16;
    std::printf("DROPIN_OK %f\n", disparity[(size_t)width * height / 2]);
    return 0;
}
