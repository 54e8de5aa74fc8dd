/* oracle/adc_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * C-ABI of the CPU restatement of the reference's AD-Census hot path (oracle/adc_oracle.c).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product library (adcensus_b200/csrc) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py compares every stage of this restatement
 * bit-for-bit with the real reference (oracle/_ref, built from /root/reference by oracle/Makefile)
 * on Cone and on synthetic pairs, and tests/test_oracle_golden.py compares it with the committed
 * golden vectors in tests/golden/ that were produced by the real reference (tools/make_golden.py).
 */
#ifndef ADC_ORACLE_H_
#define ADC_ORACLE_H_

#include <stddef.h>
#include <stdint.h>
#include "adc_taps.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Same 60-byte layout as the reference's ADCensusOption (adcensus_types.h:45-75). */
typedef struct orc_option {
    int32_t min_disparity, max_disparity;
    int32_t lambda_ad, lambda_census;
    int32_t cross_L1, cross_L2, cross_t1, cross_t2;
    float so_p1, so_p2;
    int32_t so_tso;
    int32_t irv_ts;
    float irv_th;
    float lrcheck_thres;
    uint8_t do_lr_check, do_filling, do_discontinuity_adjustment, pad_;
} orc_option;

typedef struct orc_ctx orc_ctx;

void orc_default_option(orc_option* o);
orc_ctx* orc_create(int width, int height, const orc_option* opt);
void orc_destroy(orc_ctx* c);
/* whole pipeline, like ADCensusStereo::Match; returns 1 on success */
int orc_match(orc_ctx* c, const uint8_t* left, const uint8_t* right, float* disp_left);
/* staged runner, mirrors oracle/ref_harness.cpp */
int orc_begin(orc_ctx* c, const uint8_t* left, const uint8_t* right);
int orc_step(orc_ctx* c);
size_t orc_tap(orc_ctx* c, int tap, void* dst, size_t cap);
double orc_time_match(orc_ctx* c, const uint8_t* left, const uint8_t* right, float* disp, int iters);

/* leaf functions exposed for exhaustive micro-tests */
uint8_t orc_gray(uint8_t b, uint8_t g, uint8_t r);
int orc_hamming64(uint64_t a, uint64_t b);
float orc_cost_value(int sum_abs_diff, int hamming, int lambda_ad, int lambda_census);

#ifdef __cplusplus
}
#endif
#endif
