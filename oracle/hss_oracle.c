/* hss_oracle.c — plain-C twin of the CPU oracle (no BLAS, naive loops).
 *
 * TEST INFRASTRUCTURE ONLY: loaded by tests/ (and __graft_entry__.smoke()) to
 * cross-check oracle/hss_oracle.py with an independent arithmetic path.  The
 * shipped library never links or loads it.
 *
 * PARITY UNPINNED (see oracle/hss_oracle.py): the reference has no golden
 * vectors for this path and Julia is not available in the image.
 *
 * Restates, with the reference's recursion order:
 *   mul!         src/matmul.jl:18-28
 *   _matmatup    src/matmul.jl:32-42   (post-order; Z kept per node)
 *   _matmatdown! src/matmul.jl:44-62   (pre-order; F passed down)
 * All matrices column-major doubles with explicit leading dimensions.
 */
#include <stdlib.h>
#include <string.h>

typedef struct hsso_node {
  int leaf;         /* isleaf, src/hssmatrix.jl:88 */
  int left, right;  /* A11, A22 (indices into the node array) */
  int m, n;         /* size(), src/hssmatrix.jl:94 */
  int kr, kw;       /* gensize(), src/hssmatrix.jl:254-262 (leaf: cols of U, V) */
  const double *D, *U, *V;         /* leaf: m x n, m x kr, n x kw */
  const double *B12, *B21;         /* branch: kr(l) x kw(r), kr(r) x kw(l) */
  const double *R1, *W1, *R2, *W2; /* branch: kr(l) x kr, kw(l) x kw, kr(r) x kr, kw(r) x kw */
} hsso_node;

/* C[m x n] = alpha * op(A) * B + beta * C; beta == 0 never reads C (BLAS dgemm). */
static void gemm(int transa, int m, int n, int k, double alpha, const double* A, long lda, const double* B, long ldb,
                 double beta, double* C, long ldc) {
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) {
      double s = 0.0;
      for (int p = 0; p < k; ++p) s += (transa ? A[(long)i * lda + p] : A[(long)p * lda + i]) * B[(long)j * ldb + p];
      double* c = &C[(long)j * ldc + i];
      *c = (beta == 0.0) ? alpha * s : alpha * s + beta * (*c);
    }
}

/* _matmatup: Z[node] (kw x nrhs, ld = max(kw,1)) for every node below `t`. */
static void matmatup(const hsso_node* nd, int t, int isroot, const double* B, long ldb, int nrhs, double** Z) {
  const hsso_node* h = &nd[t];
  const int kw = isroot ? 0 : h->kw;
  Z[t] = (double*)calloc((size_t)(kw > 0 ? kw : 1) * (size_t)nrhs, sizeof(double));
  if (h->leaf) { /* matmul.jl:34: V' * B */
    if (kw > 0) gemm(1, kw, nrhs, h->n, 1.0, h->V, h->n > 0 ? h->n : 1, B, ldb, 0.0, Z[t], kw);
    return;
  }
  const hsso_node* l = &nd[h->left];
  matmatup(nd, h->left, 0, B, ldb, nrhs, Z);          /* matmul.jl:37 */
  matmatup(nd, h->right, 0, B + l->n, ldb, nrhs, Z);  /* matmul.jl:38 */
  if (kw > 0) {                                       /* matmul.jl:39: W1'*Z1 .+ W2'*Z2 */
    const hsso_node* r = &nd[h->right];
    gemm(1, kw, nrhs, l->kw, 1.0, h->W1, l->kw > 0 ? l->kw : 1, Z[h->left], l->kw > 0 ? l->kw : 1, 0.0, Z[t], kw);
    gemm(1, kw, nrhs, r->kw, 1.0, h->W2, r->kw > 0 ? r->kw : 1, Z[h->right], r->kw > 0 ? r->kw : 1, 1.0, Z[t], kw);
  }
}

/* _matmatdown!: F is kr x nrhs (NULL at the root, matmul.jl:26). */
static void matmatdown(const hsso_node* nd, int t, double* C, long ldc, const double* B, long ldb, int nrhs, double** Z,
                       const double* F, int krF, double alpha, double beta) {
  const hsso_node* h = &nd[t];
  if (h->leaf) {
    gemm(0, h->m, nrhs, h->n, alpha, h->D, h->m > 0 ? h->m : 1, B, ldb, beta, C, ldc); /* matmul.jl:46 */
    if (F && krF > 0) gemm(0, h->m, nrhs, krF, alpha, h->U, h->m > 0 ? h->m : 1, F, krF, 1.0, C, ldc); /* :47 */
    return;
  }
  const hsso_node* l = &nd[h->left];
  const hsso_node* r = &nd[h->right];
  const int ldl = l->kr > 0 ? l->kr : 1, ldr = r->kr > 0 ? r->kr : 1;
  double* F1 = (double*)calloc((size_t)ldl * (size_t)nrhs, sizeof(double));
  double* F2 = (double*)calloc((size_t)ldr * (size_t)nrhs, sizeof(double));
  /* matmul.jl:52-56 */
  gemm(0, l->kr, nrhs, r->kw, 1.0, h->B12, ldl, Z[h->right], r->kw > 0 ? r->kw : 1, 0.0, F1, ldl);
  gemm(0, r->kr, nrhs, l->kw, 1.0, h->B21, ldr, Z[h->left], l->kw > 0 ? l->kw : 1, 0.0, F2, ldr);
  if (F && krF > 0) {
    gemm(0, l->kr, nrhs, krF, 1.0, h->R1, ldl, F, krF, 1.0, F1, ldl);
    gemm(0, r->kr, nrhs, krF, 1.0, h->R2, ldr, F, krF, 1.0, F2, ldr);
  }
  matmatdown(nd, h->left, C, ldc, B, ldb, nrhs, Z, F1, l->kr, alpha, beta);                 /* :58 */
  matmatdown(nd, h->right, C + l->m, ldc, B + l->n, ldb, nrhs, Z, F2, r->kr, alpha, beta);  /* :59 */
  free(F1);
  free(F2);
}

/* mul!(C, hssA, B, alpha, beta), src/matmul.jl:18-28.  Returns 0, or -2 on a
 * DimensionMismatch (matmul.jl:19-20). */
int hsso_mul(const hsso_node* nd, int n_nodes, int root, int rows_c, int rows_b, int nrhs, const double* B, long ldb,
             double* C, long ldc, double alpha, double beta) {
  const hsso_node* h = &nd[root];
  if (h->n != rows_b || h->m != rows_c) return -2;
  if (h->leaf) { /* matmul.jl:21-22 */
    gemm(0, h->m, nrhs, h->n, alpha, h->D, h->m > 0 ? h->m : 1, B, ldb, beta, C, ldc);
    return 0;
  }
  double** Z = (double**)calloc((size_t)n_nodes, sizeof(double*));
  matmatup(nd, root, 1, B, ldb, nrhs, Z);                                 /* rooted(): matmul.jl:24-25 */
  matmatdown(nd, root, C, ldc, B, ldb, nrhs, Z, NULL, 0, alpha, beta);    /* matmul.jl:26 */
  for (int i = 0; i < n_nodes; ++i) free(Z[i]);
  free(Z);
  return 0;
}

/* Host twin of the synthetic generator (oracle/hss_oracle.py synth_values). */
static unsigned long long splitmix64(unsigned long long x) {
  unsigned long long z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

void hsso_synth_values(unsigned long long seed, unsigned long long heap_id, int kind, unsigned long long start,
                       long count, double c, double* out) {
  const unsigned long long key = splitmix64(seed ^ splitmix64(heap_id * 8ull + (unsigned long long)kind));
  for (long i = 0; i < count; ++i) {
    const unsigned long long h = splitmix64(key + start + (unsigned long long)i);
    const long long s = (long long)((h & 0xFFFFull) + ((h >> 16) & 0xFFFFull) + ((h >> 32) & 0xFFFFull) + (h >> 48)) - 131070;
    out[i] = (double)s * c;
  }
}
