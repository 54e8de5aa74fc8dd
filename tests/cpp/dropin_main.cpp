// Compiled by tests/test_host_logic.py: a caller written against the REFERENCE's public interface
// (ADCensusStereo.h:14-95, usage as in main.cpp:80-118), built against this repo's include/ and lib.
#include <cstdio>
#include <cstring>
#include <vector>

#include "ADCensusStereo.h"

int main(int argc, char** argv) {
    const sint32 width = 64, height = 48;
    ADCensusOption ad_option;                 // defaults must be the reference's
    if (ad_option.max_disparity != 64 || ad_option.cross_L1 != 34 || ad_option.so_p2 != 3.0f || !ad_option.do_filling) return 10;
    ad_option.min_disparity = 0;
    ad_option.max_disparity = 16;
    std::vector<uint8> left(width * height * 3, 90), right(width * height * 3, 90);
    std::vector<float32> disparity(width * height, -1.0f);
    ADCensusStereo ad_census;
    if (ad_census.Match(left.data(), right.data(), disparity.data())) return 11;      // before Initialize -> false
    ADCensusOption bad = ad_option; bad.max_disparity = bad.min_disparity;
    if (ad_census.Initialize(width, height, bad)) return 12;                          // empty range -> false
    if (ad_census.Initialize(0, height, ad_option)) return 13;                        // bad size -> false
    const bool have_gpu = ad_census.Initialize(width, height, ad_option);
    if (!have_gpu) { std::printf("DROPIN_NO_GPU\n"); return 0; }                      // no device: Initialize is false, no fallback
    if (ad_census.Match(nullptr, right.data(), disparity.data())) return 14;          // null pointer -> false
    if (!ad_census.Match(left.data(), right.data(), disparity.data())) return 15;
    if (!ad_census.Reset(width, height, ad_option)) return 16;
    std::printf("DROPIN_OK %f\n", disparity[width * height / 2]);
    return 0;
}
