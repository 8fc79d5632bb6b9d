// hssb_hostpipe.h — host side of hssb_matmul for PAGEABLE caller memory.
//
// The Julia drop-in hands the library ordinary Matrix{Float64} storage (matmul.jl:13 allocates C with
// `similar`): pageable memory.  cudaMemcpy*Async from pageable memory is staged by the driver through
// one internal pinned buffer, synchronously and on one thread (~8 GB/s measured against ~55 GB/s of
// PCIe Gen5 from pinned memory), which made the product 6x slower end to end than from pinned buffers.
//
// Here the staging is done by the library: two rings of pinned 2 MiB slots (host -> device, device ->
// host) and two small pools of worker threads.  An IN piece is copied by a worker into a free slot and
// sent with cudaMemcpyAsync; an OUT piece is DMAed into a slot and copied by a worker into the caller's
// matrix.  Many pieces are in flight, so the memcpy of one piece overlaps the DMA of others and the
// device work of earlier column blocks; the PCIe link, not one core's memcpy rate, becomes the limit.
//
// Host-only code (no device functions); included by hssb_api.cu.
#pragma once

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif
#if defined(__linux__)
#include <sched.h>
#endif

#include "hssb_internal.h"

namespace hssb {

// Minimal FIFO thread pool.  The two process-wide pools are created on first use and never destroyed
// (worker threads must not outlive static destruction order games with the CUDA runtime at exit).
class HostPool {
 public:
  explicit HostPool(int n) {
    for (int i = 0; i < n; ++i) std::thread([this] { run(); }).detach();
    nthreads_ = n;
  }
  void submit(std::function<void()> f) {
    {
      std::lock_guard<std::mutex> lk(m_);
      q_.push_back(std::move(f));
    }
    cv_.notify_one();
  }
  int size() const { return nthreads_; }

 private:
  void run() {
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [this] { return !q_.empty(); });
        f = std::move(q_.front());
        q_.pop_front();
      }
      f();
    }
  }
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<std::function<void()>> q_;
  int nthreads_ = 0;
};

static int host_threads_per_pool() {
  if (const char* e = getenv("HSSB_HOST_THREADS")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 64) return v;
  }
  // Two pools (in / out) per process, and on a multi-GPU box one process per GPU is the usual arrangement:
  // share the cores the process may run on between the GPUs of the box (8 GPUs on 32 cores: 2 threads per
  // direction and process; one GPU on 16 cores: 8).
  int cores = (int)std::thread::hardware_concurrency();
#if defined(__linux__)
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) cores = CPU_COUNT(&set);
#endif
  int gpus = 1;
  if (cudaGetDeviceCount(&gpus) != cudaSuccess || gpus < 1) { cudaGetLastError(); gpus = 1; }
  const int n = cores > 0 ? cores / (2 * gpus) : 4;
  return n < 2 ? 2 : (n > 8 ? 8 : n);
}
static HostPool& host_pool(int dir) {
  static HostPool* pools[2] = {new HostPool(host_threads_per_pool()), new HostPool(host_threads_per_pool())};
  return *pools[dir];
}

// Streaming copy: the destination is written with non-temporal stores (no read-for-ownership of the
// destination lines, no cache pollution): a staging copy touches every byte exactly once, and the DMA
// engine or the caller reads it from memory anyway.  ~1.5x the rate of memcpy() per core at 2 MiB.
static inline void stream_copy(void* dst, const void* src, size_t bytes) {
#if defined(__x86_64__)
  char* d = (char*)dst;
  const char* s = (const char*)src;
  const size_t head = ((uintptr_t)d & 15) ? 16 - ((uintptr_t)d & 15) : 0;
  if (bytes < 256 + head) { memcpy(d, s, bytes); return; }
  if (head) { memcpy(d, s, head); d += head; s += head; bytes -= head; }
  size_t n64 = bytes / 64;
  for (; n64; --n64, d += 64, s += 64) {
    const __m128i a = _mm_loadu_si128((const __m128i*)s), b = _mm_loadu_si128((const __m128i*)(s + 16));
    const __m128i c = _mm_loadu_si128((const __m128i*)(s + 32)), e = _mm_loadu_si128((const __m128i*)(s + 48));
    _mm_stream_si128((__m128i*)d, a);
    _mm_stream_si128((__m128i*)(d + 16), b);
    _mm_stream_si128((__m128i*)(d + 32), c);
    _mm_stream_si128((__m128i*)(d + 48), e);
  }
  _mm_sfence();
  if (bytes & 63) memcpy(d, s, bytes & 63);
#else
  memcpy(dst, src, bytes);
#endif
}
static bool env_flag(const char* name, bool dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) != 0 : dflt;
}

// Counts outstanding pieces; remembers the first error.
struct Latch {
  std::mutex m;
  std::condition_variable cv;
  int64_t pending = 0;
  cudaError_t err = cudaSuccess;
  void add(int64_t n) {
    std::lock_guard<std::mutex> lk(m);
    pending += n;
  }
  void done(cudaError_t e) {
    std::lock_guard<std::mutex> lk(m);
    if (e != cudaSuccess && err == cudaSuccess) err = e;
    if (--pending == 0) cv.notify_all();
  }
  cudaError_t wait() {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [this] { return pending == 0; });
    return err;
  }
};

// Pinned slot rings of one handle.  Pieces are numbered per direction for the lifetime of the handle;
// piece p uses slot p % NSLOT once the slot's previous use (number p / NSLOT - 1) is complete.
struct Bounce {
  static constexpr int NSLOT = 32;
  static constexpr size_t SLOT = (size_t)2 << 20;
  char* pin[2] = {nullptr, nullptr};  // [0] host -> device, [1] device -> host
  cudaEvent_t ev[2][NSLOT] = {};
  std::atomic<int64_t> uses[2][NSLOT];  // completed uses of the slot
  int64_t next_piece[2] = {0, 0};
  int device = 0;
  bool nt = true;

  int init(int dev) {
    device = dev;
    nt = env_flag("HSSB_BOUNCE_NT", true);
    const bool spin = env_flag("HSSB_BOUNCE_SPIN", true);  // measured: blocking waits cost ~10 % end to end
    for (int d = 0; d < 2; ++d) {
      HSSB_CUDA(cudaHostAlloc((void**)&pin[d], NSLOT * SLOT, cudaHostAllocDefault));
      for (int s = 0; s < NSLOT; ++s) {
        // device -> host events are waited on by a worker (HSSB_BOUNCE_SPIN=0: block instead of spinning on a core)
        HSSB_CUDA(cudaEventCreateWithFlags(&ev[d][s], cudaEventDisableTiming | (d && !spin ? cudaEventBlockingSync : 0)));
        uses[d][s].store(0);
      }
    }
    return HSSB_OK;
  }
  ~Bounce() {
    for (int d = 0; d < 2; ++d) {
      if (pin[d]) cudaFreeHost(pin[d]);
      for (int s = 0; s < NSLOT; ++s)
        if (ev[d][s]) cudaEventDestroy(ev[d][s]);
    }
  }
  void wait_slot(int d, int slot, int64_t use) {
    while (uses[d][slot].load(std::memory_order_acquire) != use) std::this_thread::yield();
  }

  // host (pageable) -> device through the ring; `latch` is released when the DMA has been queued
  void submit_in(const char* src, char* dst_dev, size_t bytes, cudaStream_t st, Latch* latch) {
    const int64_t p = next_piece[0]++;
    latch->add(1);
    host_pool(0).submit([=] {
      const int slot = (int)(p % NSLOT);
      const int64_t use = p / NSLOT;
      cudaError_t e = cudaSetDevice(device);
      wait_slot(0, slot, use);
      if (e == cudaSuccess && use > 0) e = cudaEventSynchronize(ev[0][slot]);  // the slot's previous DMA has read it
      char* s = pin[0] + (size_t)slot * SLOT;
      if (nt) stream_copy(s, src, bytes); else memcpy(s, src, bytes);
      if (e == cudaSuccess) e = cudaMemcpyAsync(dst_dev, s, bytes, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) e = cudaEventRecord(ev[0][slot], st);
      uses[0][slot].store(use + 1, std::memory_order_release);
      latch->done(e);
    });
  }
  // device -> host (pageable) through the ring; `latch` is released when the bytes are in `dst`
  void submit_out(const char* src_dev, char* dst, size_t bytes, cudaStream_t st, Latch* latch) {
    const int64_t p = next_piece[1]++;
    latch->add(1);
    host_pool(1).submit([=] {
      const int slot = (int)(p % NSLOT);
      const int64_t use = p / NSLOT;
      cudaError_t e = cudaSetDevice(device);
      wait_slot(1, slot, use);  // the previous piece has been copied out of the slot
      char* s = pin[1] + (size_t)slot * SLOT;
      if (e == cudaSuccess) e = cudaMemcpyAsync(s, src_dev, bytes, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaEventRecord(ev[1][slot], st);
      if (e == cudaSuccess) e = cudaEventSynchronize(ev[1][slot]);
      if (e == cudaSuccess) { if (nt) stream_copy(dst, s, bytes); else memcpy(dst, s, bytes); }
      uses[1][slot].store(use + 1, std::memory_order_release);
      latch->done(e);
    });
  }

  // A rows x ncols column-major panel: contiguous runs of at most SLOT bytes (whole panel if it is dense
  // on both sides, else column by column).
  template <class F>
  static void for_pieces(int64_t rows, int64_t ncols, int64_t ld_host, int64_t ld_dev, F&& f) {
    if (rows <= 0 || ncols <= 0) return;
    if (ld_host == rows && ld_dev == rows) {
      const size_t total = (size_t)rows * (size_t)ncols * 8;
      for (size_t o = 0; o < total; o += SLOT) f(o, o, std::min(SLOT, total - o));
    } else {
      const size_t col = (size_t)rows * 8;
      for (int64_t c = 0; c < ncols; ++c)
        for (size_t o = 0; o < col; o += SLOT) f((size_t)c * ld_host * 8 + o, (size_t)c * ld_dev * 8 + o, std::min(SLOT, col - o));
    }
  }
};

// Pageable (unregistered) host memory?  Pinned / registered / managed memory goes the direct way.
static bool host_ptr_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

}  // namespace hssb
