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
  /* the packer front end from C: a 2-leaf tree (3 x 3 and 2 x 2 leaves, rank 1), post-order registration */
  {
    hssb_builder* b = NULL;
    const double D1[9] = {1, 2, 3, 4, 5, 6, 7, 8, 10}, U1[3] = {1, 0, -1}, V1[3] = {2, 1, 0};
    const double D2[4] = {3, 1, 1, 2}, U2[2] = {1, 1}, V2[2] = {0.5, -0.5};
    const double B12[1] = {0.25}, B21[1] = {-0.75};
    double back[9];
    hssb_node_t nd;
    int64_t l, r, root;
    if (hssb_builder_create(&b) != HSSB_OK) return 12;
    l = hssb_builder_add_leaf(b, 3, 3, 1, 1, D1, 3, U1, 3, V1, 3);
    r = hssb_builder_add_leaf(b, 2, 2, 1, 1, D2, 2, U2, 2, V2, 2);
    root = hssb_builder_add_branch(b, l, r, 0, 0, B12, 1, B21, 1, NULL, 1, NULL, 1, NULL, 1, NULL, 1);
    if (l < 0 || r < 0 || root < 0) { fprintf(stderr, "%s\n", hssb_last_error()); return 13; }
    if (hssb_builder_add_leaf(b, 3, 3, 1, 1, D1, 2, U1, 3, V1, 3) != HSSB_ERR_DIM) return 14;   /* ld < rows */
    if (hssb_plan_only(b, root, 0, 1, &h) != HSSB_OK) { fprintf(stderr, "%s\n", hssb_last_error()); return 15; }
    hssb_builder_destroy(b);
    if (hssb_info(h, &info) != HSSB_OK || info.m != 5 || info.n != 5 || info.n_leaves != 2 || info.n_nodes != 3) return 16;
    if (hssb_node_info(h, 0, &nd) != HSSB_OK || nd.is_leaf || nd.left != 1 || nd.right != 2) return 17;
    if (hssb_get_block(h, 1, 0, back, 9) != HSSB_OK || memcmp(back, D1, sizeof(D1)) != 0) return 18;   /* D of the left leaf */
    if (hssb_get_block(h, 0, 3, back, 9) != HSSB_OK || back[0] != 0.25) return 19;                       /* B12 of the root */
    if (hssb_ulv_info(h, &ui) != HSSB_OK || ui.supported != 1) return 20;
    if (hssb_destroy(h) != HSSB_OK) return 21;
  }
  printf("C_ABI_OK\n");
  return 0;
}
