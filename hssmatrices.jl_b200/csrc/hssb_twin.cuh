// hssb_twin.cuh — adjoint twin pool: the generators of A' in the layout of the primary pool, built on the device,
// so that A' X runs the forward plan and the fixed-shape kernels.  Included by hssb_api.cu.
#pragma once

namespace hssb {

// --------------------------------------------------------- adjoint twin pool ---
// A' is the HSS matrix with generators D', U <-> V, B12 <-> B21', R <-> W (hssmatrix.jl:165-180).
// On a uniform tree (square leaves, one rank) those blocks have the stored shapes of the blocks they
// replace (V and W are stored transposed), so the twin pool keeps the layout of the primary pool
// and every twin block is the transpose of one stored primary block.
static inline int twin_partner(int kind) {
  switch (kind) {
    case BK_U: return BK_V;
    case BK_V: return BK_U;
    case BK_B12: return BK_B21;
    case BK_B21: return BK_B12;
    case BK_R: return BK_W;
    case BK_W: return BK_R;
    default: return BK_D;
  }
}

static bool twin_blocks(const hssb_matrix* H, std::vector<TwinBlock>& out) {
  out.clear();
  if (!H->padded) return false;
  for (auto& t : H->nodes)
    for (int k = 0; k < BK_COUNT; ++k) {
      const int s = twin_partner(k);
      if ((t.off[k] < 0) != (t.off[s] < 0)) return false;
      if (t.off[k] < 0) continue;
      if (t.rows[k] != t.cols[s] || t.cols[k] != t.rows[s]) return false;
      TwinBlock b;
      b.src = t.off[s]; b.dst = t.off[k];
      b.rows = (int32_t)t.rows[s]; b.cols = (int32_t)t.cols[s];
      b.ld_src = t.ld[s]; b.ld_dst = t.ld[k];
      out.push_back(b);
    }
  return true;
}

// One CTA per block (grid-stride), 32x32 tiles through shared memory: both sides coalesced.
__global__ void __launch_bounds__(256)
twin_transpose_kernel(const TwinBlock* __restrict__ blocks, int64_t nblocks, const double* __restrict__ pool,
                      double* __restrict__ twin) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
    const TwinBlock tb = blocks[b];
    const int tr = (tb.rows + 31) / 32, tc = (tb.cols + 31) / 32;
    for (int t = 0; t < tr * tc; ++t) {
      const int r0 = (t % tr) * 32, c0 = (t / tr) * 32;
#pragma unroll
      for (int j = ty; j < 32; j += 8)
        if (r0 + tx < tb.rows && c0 + j < tb.cols) tile[j][tx] = pool[tb.src + (int64_t)(c0 + j) * tb.ld_src + r0 + tx];
      __syncthreads();
#pragma unroll
      for (int j = ty; j < 32; j += 8)  // dst(c, r) = src(r, c): dst column r0 + j, dst row c0 + tx
        if (c0 + tx < tb.cols && r0 + j < tb.rows) twin[tb.dst + (int64_t)(r0 + j) * tb.ld_dst + c0 + tx] = tile[tx][j];
      __syncthreads();
    }
  }
}

// 0: the twin is ready, 1: not available (caller falls back to the any-shape transposed plan), < 0: error
static int ensure_twin(hssb_matrix* H) {
  if (H->pool_t_dev) return 0;
  if (!H->adjoint_twin || H->twin_unavailable) return 1;
  std::vector<TwinBlock> tb;
  if (!twin_blocks(H, tb) || tb.empty()) { H->twin_unavailable = true; return 1; }
  size_t free_b = 0, total_b = 0;
  const size_t need = (size_t)H->pool_len * sizeof(double);
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || free_b < need + ((size_t)1 << 30)) {  // keep 1 GiB for workspaces / staging
    cudaGetLastError();
    H->twin_unavailable = true;
    return 1;
  }
  if (cudaMalloc(&H->pool_t_dev, need) != cudaSuccess) {
    cudaGetLastError();
    H->pool_t_dev = nullptr;
    H->twin_unavailable = true;
    return 1;
  }
  TwinBlock* dtb = nullptr;
  auto fail = [&]() { cudaFree(dtb); cudaFree(H->pool_t_dev); H->pool_t_dev = nullptr; };
  cudaError_t e = cudaMalloc(&dtb, tb.size() * sizeof(TwinBlock));
  if (e == cudaSuccess) e = cudaMemsetAsync(H->pool_t_dev, 0, need, H->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dtb, tb.data(), tb.size() * sizeof(TwinBlock), cudaMemcpyHostToDevice, H->stream);
  if (e == cudaSuccess) {
    const int grid = (int)std::min<size_t>(tb.size(), 148 * 8);
    twin_transpose_kernel<<<grid, 256, 0, H->stream>>>(dtb, (int64_t)tb.size(), H->pool_dev, H->pool_t_dev);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(H->stream);
  if (e != cudaSuccess) {
    fail();
    HSSB_FAIL(HSSB_ERR_CUDA, "building the adjoint twin pool failed: %s", cudaGetErrorString(e));
  }
  cudaFree(dtb);
  return 0;
}

static void drop_twin(hssb_matrix* H) {
  if (H->pool_t_dev) cudaFree(H->pool_t_dev);
  H->pool_t_dev = nullptr;
  H->twin_unavailable = false;
}

}  // namespace hssb
