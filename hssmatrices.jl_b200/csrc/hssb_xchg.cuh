// hssb_xchg.cuh — the one exchange step of the sharded product (subtree-root Z blocks): NVLink peer stores into
// the peers' workspaces (CUDA IPC across processes, direct pointers inside one process) and the NCCL all-gather
// fallback (NCCL is dlopen'ed).  Included by hssb_api.cu.
#pragma once

namespace hssb {

// ---------------------------------------------------------------- NCCL glue ---
// Loaded lazily with dlopen so that single-GPU use never needs NCCL.
struct Id128 { char b[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /* ncclUniqueId by value: 128 bytes */ Id128, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.lib) return HSSB_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* nm : names) {
    lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) HSSB_FAIL(HSSB_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(lib, "ncclCommInitRank");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy)
    HSSB_FAIL(HSSB_ERR_COMM, "libnccl is missing required symbols");
  g_nccl.lib = lib;
  return HSSB_OK;
}
#define HSSB_NCCL(expr)                                                                                   \
  do {                                                                                                    \
    int _r = (expr);                                                                                      \
    if (_r != 0)                                                                                          \
      HSSB_FAIL(HSSB_ERR_COMM, "NCCL error %d (%s) at %s:%d", _r,                                         \
                g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?", __FILE__, __LINE__);             \
  } while (0)

// --------------------------------------------------- peer-memory exchange ---
// One-shot all-gather of the subtree-root Z blocks over NVLink peer stores, replacing the NCCL
// call on the critical path: every rank writes its slot straight into the Z workspace of every
// peer, then raises a flag there; the consumer spins on its own flag block.  An acknowledgement
// flag (written after the last top-tree phase) keeps a fast rank from overwriting a slot a slow
// peer is still reading.  Epochs live in device memory so that the kernels replay inside a graph.
struct XchgParams {
  double* z[hssb_matrix::MAX_PEERS];
  unsigned long long* flags[hssb_matrix::MAX_PEERS];
  int rank, nranks;
  long long slot_off;    // element offset of slot 0 in every Z workspace
  long long slot_elems;  // elements per slot
};

__device__ __forceinline__ unsigned long long ld_volatile_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void spin_until_ge(const unsigned long long* p, unsigned long long want, unsigned long long what) {
  const long long t0 = clock64();
  while (ld_volatile_sys(p) < want) {
    if (clock64() - t0 > 8000000000ll) trap_report(what, want, ld_volatile_sys(p));  // a lost peer must surface as an error, not a hang
  }
}

__global__ void __launch_bounds__(256) xchg_push_kernel(XchgParams q) {
  const int P = q.nranks, me = q.rank;
  unsigned long long* mine = q.flags[me];
  __shared__ unsigned long long s_epoch;
  if (threadIdx.x == 0) s_epoch = ld_volatile_sys(mine + 2 * P) + 1;
  __syncthreads();
  const unsigned long long e = s_epoch;
  // peers must have consumed the previous epoch before their copy of my slot is overwritten
  if ((int)threadIdx.x < P && (int)threadIdx.x != me) spin_until_ge(mine + P + threadIdx.x, e - 1, TRAP_PEER_ACK);
  __syncthreads();
  const double2* src = reinterpret_cast<const double2*>(q.z[me] + q.slot_off + (long long)me * q.slot_elems);
  const long long n2 = q.slot_elems / 2;
  for (int r = 0; r < P; ++r) {
    if (r == me) continue;
    double2* dst = reinterpret_cast<double2*>(q.z[r] + q.slot_off + (long long)me * q.slot_elems);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool s_last;
  if (threadIdx.x == 0) s_last = atomicAdd(mine + 2 * P + 1, 1ull) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {  // every CTA of this rank has pushed: publish, then wait for everybody else's slot
    if (threadIdx.x == 0) mine[2 * P + 1] = 0;
    if ((int)threadIdx.x < P && (int)threadIdx.x != me) {
      __threadfence_system();
      asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(q.flags[threadIdx.x] + me), "l"(e) : "memory");
      spin_until_ge(mine + threadIdx.x, e, TRAP_PEER_DATA);
    }
    __syncthreads();
    __threadfence_system();
    if (threadIdx.x == 0) mine[2 * P] = e;
  }
}

__global__ void xchg_ack_kernel(XchgParams q) {
  const int P = q.nranks, me = q.rank;
  const unsigned long long e = ld_volatile_sys(q.flags[me] + 2 * P);
  if ((int)threadIdx.x < P && (int)threadIdx.x != me)
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(q.flags[threadIdx.x] + P + me), "l"(e) : "memory");
}

static XchgParams xchg_params(const hssb_matrix* H, const CallParams& cp) {
  XchgParams q;
  memset(&q, 0, sizeof(q));
  for (int r = 0; r < H->n_shards; ++r) { q.z[r] = H->peer_z[r]; q.flags[r] = H->peer_flags[r]; }
  q.rank = H->shard_rank; q.nranks = H->n_shards;
  q.slot_off = H->xchg_zoff * (long long)cp.nrhs;
  q.slot_elems = H->xchg_slot_rows * (long long)cp.nrhs;
  return q;
}

}  // namespace hssb
