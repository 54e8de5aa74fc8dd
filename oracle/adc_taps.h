/* oracle/adc_taps.h -- TEST INFRASTRUCTURE ONLY.
 * Stage and tap numbering shared by the two CPU checkers (ref_harness.cpp around the real
 * reference, adc_oracle.c our restatement).  The numbers are kept equal to the debug enums of
 * the product header include/adcensus_b200.h so that a parity test can use one id for all three.
 */
#ifndef ADC_TAPS_H_
#define ADC_TAPS_H_

/* pipeline stages in execution order (ADCensusStereo.cpp:69-132) */
enum {
    ADC_STAGE_COST = 0,   /* gray + census + AD-census volume        cost_computor.cpp:123-137 */
    ADC_STAGE_ARMS = 1,   /* cross arms + support counts, aggr<-init cross_aggregator.cpp:98-108 */
    ADC_STAGE_AGG1 = 2,   /* aggregation iteration 1 (H then V)      cross_aggregator.cpp:111-117 */
    ADC_STAGE_AGG2 = 3,   /* iteration 2 (V then H) */
    ADC_STAGE_AGG3 = 4,
    ADC_STAGE_AGG4 = 5,
    ADC_STAGE_SO1 = 6,    /* left->right   aggr -> init              scanline_optimizer.cpp:54 */
    ADC_STAGE_SO2 = 7,    /* right->left   init -> aggr              :56 */
    ADC_STAGE_SO3 = 8,    /* top->bottom   aggr -> init              :58 */
    ADC_STAGE_SO4 = 9,    /* bottom->top   init -> aggr              :60 */
    ADC_STAGE_WTA = 10,   /* left + right disparity                  ADCensusStereo.cpp:108-109 */
    ADC_STAGE_OUTLIER = 11,
    ADC_STAGE_VOTE = 12,
    ADC_STAGE_INTERP = 13,
    ADC_STAGE_DISC = 14,
    ADC_STAGE_MEDIAN = 15,
    ADC_STAGE_COUNT = 16
};

/* live buffers */
enum {
    ADC_TAP_GRAY_L = 0,     /* u8  [H][W] */
    ADC_TAP_GRAY_R = 1,
    ADC_TAP_CENSUS_L = 2,   /* u64 [H][W] */
    ADC_TAP_CENSUS_R = 3,
    ADC_TAP_VOL_INIT = 4,   /* f32 [H][W][D]  the reference's cost_init_ (also SO scratch) */
    ADC_TAP_VOL_AGGR = 5,   /* f32 [H][W][D]  the reference's cost_aggr_ */
    ADC_TAP_ARMS = 6,       /* u8  [H][W][4]  left,right,top,bottom */
    ADC_TAP_SUPCNT_H = 7,   /* u16 [H][W]     horizontal-first support size */
    ADC_TAP_SUPCNT_V = 8,   /* u16 [H][W]     vertical-first support size */
    ADC_TAP_DISP_L = 9,     /* f32 [H][W] */
    ADC_TAP_DISP_R = 10,    /* f32 [H][W] */
    ADC_TAP_MISMATCHES = 11,/* i32 [n][2] (x,y) in list order */
    ADC_TAP_OCCLUSIONS = 12
};

#endif
