// hssb_diag.cu — measurement helpers behind hssb_measure_peak (bench.py's
// roofline denominators).  MEASURED_PEAKS.json carries HBM and bf16 numbers but
// no FP64 figure, and this path is FP64: the DFMA / DMMA peaks are measured live
// with register-resident loops, the copy bandwidth with a plain 16-byte copy.
#include <cuda_runtime.h>

#include "hssb_internal.h"
#include "hssb_mma.cuh"

namespace hssb {

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed) {
  double a[8], x = seed + threadIdx.x * 1e-9, y = 1.0 - 1e-9;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed * (i + 1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], y, x);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 123.456) out[0] = s;  // never true; keeps the loop alive
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double seed) {
  double c[8][2];
  const double a = seed + threadIdx.x * 1e-9, b = 1e-3 * seed;
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) mma_m8n8k4(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

// DMMA throughput with ONE CTA per SM of blockDim.x threads and 16 independent accumulator tiles
// per warp (the leaf-down inner loop shape): how many warps per scheduler saturate the FP64 pipe?
__global__ void __launch_bounds__(128) dmma_occupancy_kernel(double* out, int iters, double seed) {
  double c[16][2];
  const double a = seed + threadIdx.x * 1e-9, b = 1e-3 * seed;
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) mma_m8n8k4(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

// The leaf-down inner loop in isolation: 32x32 warp tile (4x4 DMMA tiles), fragments from shared
// memory with the kernel's conflict-free leading dimensions, no global traffic, no barriers.
// variant 0: fragments reloaded every k-step (as the kernel does); 1: same with 2 k-steps of
// fragments loaded as one 16-byte LDS per operand.
template <int VARIANT>
__global__ void __launch_bounds__(256, 1) dmma_loop_kernel(double* out, int iters) {
  extern __shared__ __align__(16) double sm[];
  constexpr int LDA = 132, LDX = 132, KC = 64;
  double* As = sm;              // [KC][LDA]
  double* Xs = sm + KC * LDA;   // [64][LDX]
  // iters < 0: random operands (does the FP64 tensor rate depend on the data / power?)
  for (int i = threadIdx.x; i < KC * LDA + 64 * LDX; i += blockDim.x) {
    unsigned long long h = (i + 1) * 0x9E3779B97F4A7C15ull + blockIdx.x * 0xBF58476D1CE4E5B9ull;
    h ^= h >> 31; h *= 0x94D049BB133111EBull; h ^= h >> 29;
    sm[i] = iters < 0 ? ((double)(long long)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5) * 1e-2 : 1e-3 * (i % 7);
  }
  if (iters < 0) iters = -iters;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int wm = warp % 4, wn = warp / 4;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
    if (VARIANT == 2) {
      // chunked like the kernel: 4 k-steps (64 DMMAs) per chunk, the chunk base changes at run time
#pragma unroll 1
      for (int c = 0; c < KC / 16; ++c) {
        const double* A = As + (c * 16) * LDA + wm * 32 + g + t * LDA;
        const double* B = Xs + (wn * 32 + g) * LDX + t + c * 16;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          double a[4], b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = A[kk * 4 * LDA + i * 8];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = B[j * 8 * LDX + kk * 4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        asm volatile("" ::: "memory");
      }
    } else if (VARIANT == 0) {
      const double* A = As + wm * 32 + g + t * LDA;
      const double* B = Xs + (wn * 32 + g) * LDX + t;
#pragma unroll
      for (int kk = 0; kk < KC / 4; ++kk) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = A[kk * 4 * LDA + i * 8];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = B[j * 8 * LDX + kk * 4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) mma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    } else {
      // k permuted so that one lane needs k = 2t, 2t+1 of each group of 8: B via one LDS.128
      const double* A = As + wm * 32 + g + 2 * t * LDA;
      const double* B = Xs + (wn * 32 + g) * LDX + 2 * t;
#pragma unroll
      for (int kk = 0; kk < KC / 8; ++kk) {
        double a0[4], a1[4];
        double2 b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a0[i] = A[kk * 8 * LDA + i * 8]; a1[i] = A[(kk * 8 + 1) * LDA + i * 8]; }
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const double2*>(B + j * 8 * LDX + kk * 8);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) mma_m8n8k4(acc[i][j][0], acc[i][j][1], a0[i], b[j].x);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) mma_m8n8k4(acc[i][j][0], acc[i][j][1], a1[i], b[j].y);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
  if (s == 123.456) out[0] = s;
}

// Do DMMA (tensor path) and DFMA (FP64 pipe) share execution resources?  Even warps run the DMMA
// loop, odd warps the DFMA loop; out[1..2] report nothing, the host derives the combined rate.
__global__ void __launch_bounds__(256) mixed_fp64_kernel(double* out, int iters, double seed, int mode) {
  const int warp = threadIdx.x >> 5;
  const bool do_mma = mode == 0 ? true : mode == 1 ? false : (warp & 1) == 0;
  double s = 0;
  if (do_mma) {
    double c[8][2];
    const double a = seed + threadIdx.x * 1e-9, b = 1e-3 * seed;
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) mma_m8n8k4(c[i][0], c[i][1], a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  } else {
    double a[8], x = seed + threadIdx.x * 1e-9, y = 1.0 - 1e-9;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (i + 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], y, x);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
  }
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

}  // namespace hssb

using namespace hssb;

extern "C" int hssb_measure_peak(int device, int kind, int64_t arg, double* out) {
  if (!out) HSSB_FAIL(HSSB_ERR_ARG, "hssb_measure_peak: out is NULL");
  *out = 0.0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    HSSB_FAIL(HSSB_ERR_CUDA, "hssb_measure_peak: no usable CUDA device");
  }
  int prev = 0;
  cudaGetDevice(&prev);
  HSSB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  HSSB_CUDA(cudaGetDeviceProperties(&prop, device));
  cudaEvent_t e0, e1;
  HSSB_CUDA(cudaEventCreate(&e0));
  HSSB_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  int rc = HSSB_OK;
  if (kind == 0 || kind == 1) {
    const int iters = arg > 0 ? (int)arg : 20000;
    const int grid = prop.multiProcessorCount * 4;
    double* d = nullptr;
    HSSB_CUDA(cudaMalloc(&d, 64));
    for (int rep = 0; rep < 4; ++rep) {
      HSSB_CUDA(cudaEventRecord(e0));
      if (kind == 0) dfma_peak_kernel<<<grid, 256>>>(d, iters, 0.5);
      else dmma_peak_kernel<<<grid, 256>>>(d, iters, 0.5);
      HSSB_CUDA(cudaEventRecord(e1));
      HSSB_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      HSSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      // DFMA: 8 FMAs/thread/iter; DMMA: 8 m8n8k4 (= 512 flops) per warp per iter
      const double flops = kind == 0 ? 2.0 * 8 * iters * 256.0 * grid : 8.0 * 512.0 * iters * 8.0 * grid;
      if (rep > 0) best = std::max(best, flops / (ms * 1e-3) * 1e-12);
    }
    cudaFree(d);
  } else if (kind == 7) {
    // arg: 0 = all warps DMMA, 1 = all warps DFMA, 2 = half and half; returns combined TFLOP/s
    const int mode = (int)arg, iters = 20000, grid = prop.multiProcessorCount * 4;
    double* d = nullptr;
    HSSB_CUDA(cudaMalloc(&d, 64));
    for (int rep = 0; rep < 4; ++rep) {
      HSSB_CUDA(cudaEventRecord(e0));
      mixed_fp64_kernel<<<grid, 256>>>(d, iters, 0.5, mode);
      HSSB_CUDA(cudaEventRecord(e1));
      HSSB_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      HSSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      const double mma_warps = mode == 0 ? 8 : mode == 1 ? 0 : 4, fma_warps = 8 - mma_warps;
      // per iteration: a DMMA warp does 8 x 512 flops, a DFMA warp 8 x 64 flops
      const double flops = (mma_warps * 8 * 512.0 + fma_warps * 8 * 64.0) * iters * (double)grid;
      if (rep > 0) best = std::max(best, flops / (ms * 1e-3) * 1e-12);
    }
    cudaFree(d);
  } else if (kind == 3) {
    // arg = warps per SM sub-partition: CTAs of 128 threads (one warp per scheduler), arg CTAs per SM
    const int threads = 128;
    const int iters = 4000;
    const int grid = prop.multiProcessorCount * (int)((arg >= 1 && arg <= 16) ? arg : 1);
    double* d = nullptr;
    HSSB_CUDA(cudaMalloc(&d, 64));
    for (int rep = 0; rep < 4; ++rep) {
      HSSB_CUDA(cudaEventRecord(e0));
      dmma_occupancy_kernel<<<grid, threads>>>(d, iters, 0.5);
      HSSB_CUDA(cudaEventRecord(e1));
      HSSB_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      HSSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      const double flops = 16.0 * 512.0 * iters * (threads / 32.0) * grid;
      if (rep > 0) best = std::max(best, flops / (ms * 1e-3) * 1e-12);
    }
    cudaFree(d);
  } else if (kind == 4 || kind == 5 || kind == 6) {
    // arg = CTAs per SM (1 or 2)
    const int per_sm = 1;
    const int iters = arg < 0 ? -2000 : 2000, grid = prop.multiProcessorCount * per_sm;  // arg < 0: random operands
    const size_t smem = sizeof(double) * (64 * 132 + 64 * 132);
    double* d = nullptr;
    HSSB_CUDA(cudaMalloc(&d, 64));
    HSSB_CUDA(cudaFuncSetAttribute(dmma_loop_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HSSB_CUDA(cudaFuncSetAttribute(dmma_loop_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HSSB_CUDA(cudaFuncSetAttribute(dmma_loop_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int rep = 0; rep < 4; ++rep) {
      HSSB_CUDA(cudaEventRecord(e0));
      if (kind == 4) dmma_loop_kernel<0><<<grid, 256, smem>>>(d, iters);
      else if (kind == 5) dmma_loop_kernel<1><<<grid, 256, smem>>>(d, iters);
      else dmma_loop_kernel<2><<<grid, 256, smem>>>(d, iters);
      HSSB_CUDA(cudaEventRecord(e1));
      HSSB_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      HSSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      const double flops = 2.0 * 128 * 64 * 64 * (double)(iters < 0 ? -iters : iters) * grid;  // 128x64 tile, K = 64 per iteration
      if (rep > 0) best = std::max(best, flops / (ms * 1e-3) * 1e-12);
    }
    cudaFree(d);
  } else if (kind == 2) {
    const int64_t bytes = arg > 0 ? arg : ((int64_t)1 << 30);
    const int64_t n2 = bytes / 16;
    double2 *a = nullptr, *b = nullptr;
    if (cudaMalloc(&a, (size_t)n2 * 16) != cudaSuccess || cudaMalloc(&b, (size_t)n2 * 16) != cudaSuccess) {
      cudaGetLastError();
      cudaFree(a);
      set_error("hssb_measure_peak: allocation failed");
      rc = HSSB_ERR_ALLOC;
    } else {
      cudaMemset(a, 0, (size_t)n2 * 16);
      for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        copy_kernel<<<prop.multiProcessorCount * 16, 256>>>(a, b, n2);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0) best = std::max(best, 2.0 * n2 * 16.0 / (ms * 1e-3) * 1e-9);
      }
      cudaFree(a);
      cudaFree(b);
    }
  } else {
    set_error("hssb_measure_peak: unknown kind %d", kind);
    rc = HSSB_ERR_ARG;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaSetDevice(prev);
  if (cudaGetLastError() != cudaSuccess && rc == HSSB_OK) {
    set_error("hssb_measure_peak: CUDA failure");
    rc = HSSB_ERR_CUDA;
  }
  *out = best;
  return rc;
}
