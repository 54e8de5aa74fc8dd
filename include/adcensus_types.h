// include/adcensus_types.h -- type vocabulary of the drop-in shim.
//
// Source-compatible with the reference's adcensus_types.h (typedef names, ADCensusOption field
// order/types/defaults, Invalid_Float, ADColor) so that code written against
// ethan-li-coding/AD-Census recompiles unchanged against this repo's include/ directory.
// The option block is layout-checked against the C ABI's adc_option (60 bytes).
#ifndef ADCENSUS_B200_TYPES_H_
#define ADCENSUS_B200_TYPES_H_

#include <cstddef>
#include <cstdint>
#include <limits>
#include <utility>
#include <vector>

using std::pair;
using std::vector;

using sint8 = int8_t;    using uint8 = uint8_t;
using sint16 = int16_t;  using uint16 = uint16_t;
using sint32 = int32_t;  using uint32 = uint32_t;
using sint64 = int64_t;  using uint64 = uint64_t;
using float32 = float;   using float64 = double;

// markers used in disparity maps and cost buffers (reference adcensus_types.h:33-36)
constexpr float32 Invalid_Float = std::numeric_limits<float32>::infinity();
constexpr float32 Large_Float = 99999.0f;
constexpr float32 Small_Float = -99999.0f;

enum CensusSize { Census5x5 = 0, Census9x7 };

// Algorithm parameters; field order and types are fixed by the reference (adcensus_types.h:45-75)
// because callers fill the struct member by member and the C ABI reads it as 60 raw bytes.
struct ADCensusOption {
    sint32 min_disparity = 0;     // inclusive
    sint32 max_disparity = 64;    // exclusive
    sint32 lambda_ad = 10;        // AD cost scale
    sint32 lambda_census = 30;    // census cost scale
    sint32 cross_L1 = 34;         // max arm length
    sint32 cross_L2 = 17;         // arm length beyond which the tighter colour threshold applies
    sint32 cross_t1 = 20;         // colour threshold
    sint32 cross_t2 = 6;          // tighter colour threshold
    float32 so_p1 = 1.0f;         // scanline penalties
    float32 so_p2 = 3.0f;
    sint32 so_tso = 15;           // scanline colour threshold
    sint32 irv_ts = 20;           // region voting: minimum support
    float32 irv_th = 0.4f;        // region voting: minimum peak ratio
    float32 lrcheck_thres = 1.0f; // left/right consistency threshold (pixels)
    bool do_lr_check = true;
    bool do_filling = true;
    bool do_discontinuity_adjustment = false;
};
static_assert(sizeof(ADCensusOption) == 60 && alignof(ADCensusOption) == 4, "ADCensusOption ABI");
static_assert(offsetof(ADCensusOption, so_p1) == 32 && offsetof(ADCensusOption, do_lr_check) == 56, "ADCensusOption ABI");

// BGR-constructed colour triple (reference adcensus_types.h:80-86)
struct ADColor {
    uint8 r = 0, g = 0, b = 0;
    ADColor() = default;
    ADColor(uint8 blue, uint8 green, uint8 red) : r(red), g(green), b(blue) {}
};

#endif
