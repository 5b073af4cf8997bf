/* facerec_b200 — debug-only additions to the C ABI (include/facerec_b200.h).
 *
 * Timing and tracing experiments used while tuning conv_tma_kernel (profiles/r01_*).  They are NOT part of the
 * drop-in boundary: the product build of libfacerec_b200.so (python -m facerec_b200.build) refuses these options with
 * FRB_ERR_STATE and does not contain the stall-accounting kernel variants; build with
 *     FRB_DEBUG_BUILD=1 python -m facerec_b200.build --force
 * to get a library that honours them.  Some of them produce WRONG RESULTS by design.
 */
#ifndef FACEREC_B200_DEBUG_H_
#define FACEREC_B200_DEBUG_H_

#include "facerec_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

enum frb_debug_option {
  FRB_OPT_FORCE_GATHER = 2, /* 1: never feed the conv A operand by TMA (A/B testing of the two conv kernels)            */
  FRB_OPT_TRACE = 4,        /* non-zero: every CTA of a conv_tma launch writes 16 stall-cycle counters of its three roles
                               at h_out[1 + 16 * cta] (read with frb_debug_trace)                                      */
  FRB_OPT_MAX_STAGES = 7,   /* cap on the conv_tma stage ring depth (>= 2; latency-sensitivity experiments)             */
  FRB_OPT_SKIP_LOADS = 8    /* 1: TIMING EXPERIMENT ONLY - the A/B tiles of the per-tap path are never loaded (wrong results) */
};

/* FRB_OPT_TRACE: copies the stall counters of the LAST conv_tma launch (1 + 3 * max_records words, 16 per CTA from
 * h_out[1]).  Product build: returns FRB_ERR_STATE. */
int frb_debug_trace(frb_ctx* ctx, long long* h_out, int max_records);

#ifdef __cplusplus
}
#endif
#endif /* FACEREC_B200_DEBUG_H_ */
