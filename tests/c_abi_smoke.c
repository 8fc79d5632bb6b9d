/* Plain C99 client of include/hssb200.h (no CUDA, no C++): proves that the drop-in boundary is a C ABI.
 * Builds a plan-only synthetic matrix (host only), queries it and checks that the compute entry points
 * refuse to run without a device.  Compiled and run by tests/test_abi.py. */
#include <stdio.h>
#include <string.h>

#include "hssb200.h"

int main(void) {
  hssb_matrix* h = NULL;
  hssb_info_t info;
  hssb_ulv_info_t ui;
  int64_t nt = 0, np = 0, pl = 0;
  double x[4] = {0, 0, 0, 0}, y[4];
  if (hssb_version() != HSSB_VERSION) return 1;
  if (hssb_plan_only_synthetic(2048, 128, 16, 7u, 0, 1, &h) != HSSB_OK) { fprintf(stderr, "%s\n", hssb_last_error()); return 2; }
  if (hssb_info(h, &info) != HSSB_OK || info.m != 2048 || info.n != 2048 || info.n_leaves != 16 || info.uniform != 1) return 3;
  if (hssb_ulv_info(h, &ui) != HSSB_OK || ui.supported != 1 || ui.factored != 0) return 4;
  if (hssb_debug_counts(h, &nt, &np, &pl) != HSSB_OK || nt <= 0 || np <= 0 || pl <= 0) return 5;
  /* no CPU fallback: a plan-only handle cannot multiply or solve */
  if (hssb_matmul(h, 2048, 2048, 1, x, 2048, y, 2048, 1.0, 0.0) != HSSB_ERR_CUDA) return 7;
  if (hssb_solve(h, 2048, 1, x, 2048, y, 2048) != HSSB_ERR_CUDA) return 8;
  if (hssb_matmul(h, 2047, 2048, 1, x, 2048, y, 2048, 1.0, 0.0) != HSSB_ERR_DIM) return 9;      /* DimensionMismatch */
  if (strstr(hssb_last_error(), "DimensionMismatch") == NULL) return 10;
  if (hssb_destroy(h) != HSSB_OK) return 11;
  printf("C_ABI_OK\n");
  return 0;
}
