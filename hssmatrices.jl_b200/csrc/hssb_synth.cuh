// hssb_synth.cuh — on-device counter-based generator for the synthetic HSS
// matrices of BASELINE.json configs 3-5.  Bit-identical twin of
// oracle/hss_oracle.py (synth_key / synth_values): splitmix64 hash, sum of four
// 16-bit fields (Irwin-Hall 4), one int->double conversion and one multiply.
#pragma once

#include "hssb_internal.h"

namespace hssb {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__host__ __device__ __forceinline__ uint64_t synth_key(uint64_t seed, uint64_t heap_id, int kind) {
  return splitmix64(seed ^ splitmix64(heap_id * 8ull + (uint64_t)kind));
}

__host__ __device__ __forceinline__ double synth_value(uint64_t key, uint64_t idx, double c) {
  const uint64_t h = splitmix64(key + idx);
  const int64_t s =
      (int64_t)((h & 0xFFFFull) + ((h >> 16) & 0xFFFFull) + ((h >> 32) & 0xFFFFull) + (h >> 48)) - 131070;
  return (double)s * c;
}

// 1/sqrt(Var) of the sum of four uniform integers on [0, 65535]; the literal is
// repr(float(np.sqrt(3.0 / (65536.0**2 - 1.0)))) so that host and device agree.
constexpr double IH4_SCALE = 2.6428997921303018e-05;
constexpr int KIND_X = 7;

struct SynthBlock {
  int64_t off;   // pool offset (doubles)
  uint64_t key;
  int32_t rows, cols, ld;  // as stored
  int32_t transposed;      // 1: stored(i, j) = logical(j, i), the logical block being cols x rows
  double c;                // IH4_SCALE * scale
};

// One CTA per block descriptor (grid-stride), threads sweep the elements.
__global__ void synth_fill_kernel(const SynthBlock* __restrict__ blocks, int64_t nblocks, double* __restrict__ pool) {
  for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
    const SynthBlock sb = blocks[b];
    const int64_t total = (int64_t)sb.rows * sb.cols;
    for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
      const int64_t j = e / sb.rows, i = e - j * sb.rows;
      const int64_t idx = sb.transposed ? i * sb.cols + j : e;
      pool[sb.off + j * sb.ld + i] = synth_value(sb.key, (uint64_t)idx, sb.c);
    }
  }
}

// Rows [row0, row0+rows) of the n x nrhs right-hand side.
__global__ void synth_rhs_kernel(uint64_t key, int64_t n, int64_t nrhs, int64_t row0, int64_t rows, double* X,
                                 int64_t ldx) {
  const int64_t total = rows * nrhs;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = e / rows, i = e - j * rows;
    X[j * ldx + i] = synth_value(key, (uint64_t)(j * n + row0 + i), IH4_SCALE);
  }
}

}  // namespace hssb
