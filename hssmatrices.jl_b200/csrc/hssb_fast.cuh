// hssb_fast.cuh — fixed-shape FP64 tensor-core (DMMA) kernels for uniform trees.
#pragma once

#include "hssb_internal.h"

namespace hssb {

// Tags the phases of a uniform tree that a fixed-shape kernel can run.
static void plan_fast_phases(hssb_matrix* H) { (void)H; }
static bool fast_phase_supported(const hssb_matrix* H, const Phase& ph, const CallParams& cp) {
  (void)H; (void)ph; (void)cp;
  return false;
}
static int launch_fast(hssb_matrix* H, const Phase& ph, const CallParams& cp, cudaStream_t st) {
  (void)H; (void)ph; (void)cp; (void)st;
  return HSSB_ERR_STATE;
}
static void free_fast(hssb_matrix* H) { (void)H; }

}  // namespace hssb
