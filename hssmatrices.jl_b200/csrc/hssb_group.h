// hssb_group.h — ONE process, ONE call, P devices ("a single ccall from one Julia thread
// drives all GPUs").
//
// A group is P sharded handles of the same matrix (shard g = the g-th subtree at depth log2 P plus the
// replicated top tree), one per device, wired to each other IN PROCESS: every device enables
// peer access to the others and the exchange kernels store straight into the peers' Z workspaces and
// flag blocks (no CUDA IPC, no NCCL, no second process).  The group entry points take the WHOLE X and Y:
//   hssb_group_matmul      host pointers: shard g runs the pipelined host entry on its row block, all shards
//                          concurrently (one short-lived host thread per shard, because every shard's call
//                          blocks until its own copies are done while its kernels wait for the peers)
//   hssb_group_matmul_dev  per-device pointers to the local row blocks, asynchronous: queued device by
//                          device from the calling thread
// Shards may share a device (devices = {0, 0}): the exchange then runs between two streams of one GPU,
// which is how the sharded path is exercised on a single-GPU box.
//
// Host-only code; included at the end of hssb_api.cu.
#pragma once

struct hssb_group {
  std::vector<hssb_matrix*> shard;
  std::vector<double*> wired_z;  // Z workspaces the peer tables currently point at
};

namespace hssb {

static int group_enable_peers(hssb_group* g) {
  const int P = (int)g->shard.size();
  for (int a = 0; a < P; ++a)
    for (int b = 0; b < P; ++b) {
      const int da = g->shard[(size_t)a]->device, db = g->shard[(size_t)b]->device;
      if (da == db) continue;
      int can = 0;
      HSSB_CUDA(cudaDeviceCanAccessPeer(&can, da, db));
      if (!can) HSSB_FAIL(HSSB_ERR_COMM, "device %d cannot access device %d: no peer path for the in-process exchange", da, db);
      DeviceGuard dg(da);
      const cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else HSSB_CUDA(e);
    }
  return HSSB_OK;
}

// Workspaces for nrhs columns on every shard, then (re)build the peer tables if a workspace moved.
static int group_prepare(hssb_group* g, int64_t nrhs) {
  const int P = (int)g->shard.size();
  if (P == 1) return HSSB_OK;
  bool moved = g->wired_z.size() != (size_t)P;
  for (int r = 0; r < P; ++r) {
    hssb_matrix* h = g->shard[(size_t)r];
    DeviceGuard dg(h->device);
    if (int rc = ensure_workspace(h, nrhs)) return rc;
    if (!h->my_flags) {
      HSSB_CUDA(cudaMalloc(&h->my_flags, XCHG_FLAG_WORDS * sizeof(unsigned long long)));
      HSSB_CUDA(cudaMemset(h->my_flags, 0, XCHG_FLAG_WORDS * sizeof(unsigned long long)));
      HSSB_CUDA(cudaDeviceSynchronize());
      moved = true;
    }
    if (!moved && g->wired_z[(size_t)r] != h->z_dev) moved = true;
  }
  if (!moved) return HSSB_OK;
  // nobody may still be pushing into a workspace that has just been replaced
  for (hssb_matrix* h : g->shard) {
    DeviceGuard dg(h->device);
    HSSB_CUDA(cudaDeviceSynchronize());
  }
  g->wired_z.assign((size_t)P, nullptr);
  for (int r = 0; r < P; ++r) g->wired_z[(size_t)r] = g->shard[(size_t)r]->z_dev;
  for (hssb_matrix* h : g->shard) {
    for (int r = 0; r < P; ++r) {
      h->peer_z[r] = g->shard[(size_t)r]->z_dev;
      h->peer_flags[r] = g->shard[(size_t)r]->my_flags;
    }
    h->peer_xchg = true;
    h->peer_inprocess = true;
    DeviceGuard dg(h->device);
    invalidate_graphs(h);
  }
  return HSSB_OK;
}

static int group_check(const hssb_group* g, const char* who) {
  if (!g || g->shard.empty()) HSSB_FAIL(HSSB_ERR_ARG, "%s: NULL or empty group", who);
  return HSSB_OK;
}

static int group_wire(std::unique_ptr<hssb_group>& g, hssb_group** out) {
  if (g->shard.size() > 1) {
    if (int rc = group_enable_peers(g.get())) return rc;
  }
  *out = g.release();
  return HSSB_OK;
}

static int group_devices_ok(const int* devices, int n_devices, const char* who) {
  if (!devices || n_devices < 1 || n_devices > hssb_matrix::MAX_PEERS || !is_pow2(n_devices))
    HSSB_FAIL(HSSB_ERR_ARG, "%s: the number of devices must be a power of two in 1..%d", who, hssb_matrix::MAX_PEERS);
  // Shards that share a GPU wait for each other INSIDE kernels (the exchange spins on a flag), so nothing one of
  // them needs may ever be queued behind the other's kernels.  A device has 8 hardware work queues by default and
  // the runtime maps streams onto them: with three streams per shard (plus the application's own) two shards can
  // land on one queue and deadlock.  Sharing a device is therefore a TEST facility (it runs the sharded path on a
  // single-GPU box) and is only accepted with CUDA_DEVICE_MAX_CONNECTIONS >= 16 and CUDA_MODULE_LOADING=EAGER in the
  // environment (both must be set before the first CUDA call of the process; with lazy loading the first launch of
  // a kernel loads its code, which can synchronise the device while the other shard's kernel is already waiting)
  // and at most two shards per device.
  for (int a = 0; a < n_devices; ++a) {
    int same = 0;
    for (int b = 0; b < n_devices; ++b) same += devices[b] == devices[a];
    if (same > 2) HSSB_FAIL(HSSB_ERR_ARG, "%s: at most two shards may share a device (device %d is listed %d times)", who, devices[a], same);
    if (same > 1) {
      const char* e = getenv("CUDA_DEVICE_MAX_CONNECTIONS");
      const char* ml = getenv("CUDA_MODULE_LOADING");
      if (!e || atoi(e) < 16 || !ml || strcmp(ml, "EAGER") != 0)
        HSSB_FAIL(HSSB_ERR_ARG, "%s: device %d is listed twice; shards sharing a device need CUDA_DEVICE_MAX_CONNECTIONS >= 16 and "
                                "CUDA_MODULE_LOADING=EAGER in the environment before the first CUDA call (their kernels wait for "
                                "each other)", who, devices[a]);
    }
  }
  return HSSB_OK;
}

static int group_host(hssb_group* g, int trans, int64_t rows_y, int64_t rows_x, int64_t nrhs, const double* X, int64_t ldx, double* Y,
                      int64_t ldy, double alpha, double beta) {
  if (int rc = group_check(g, "hssb_group_matmul")) return rc;
  const hssb_matrix* h0 = g->shard[0];
  const int P = (int)g->shard.size();
  if (P == 1) return matmul_host_impl(g->shard[0], trans, rows_y, rows_x, nrhs, X, ldx, Y, ldy, alpha, beta);
  // DimensionMismatch checks of matmul.jl:19-20 against the WHOLE matrix
  const int64_t need_x = trans ? h0->m : h0->n, need_y = trans ? h0->n : h0->m;
  if (rows_x != need_x)
    HSSB_FAIL(HSSB_ERR_DIM, "DimensionMismatch: first dimension of B (%lld) does not match second dimension of A (%lld)",
              (long long)rows_x, (long long)need_x);
  if (rows_y != need_y)
    HSSB_FAIL(HSSB_ERR_DIM, "DimensionMismatch: dimensions of C (%lld rows) don't match up with A (%lld rows)", (long long)rows_y,
              (long long)need_y);
  if (nrhs < 0 || nrhs > INT32_MAX) HSSB_FAIL(HSSB_ERR_ARG, "hssb_group_matmul: bad nrhs");
  if (nrhs == 0 || rows_y == 0) return HSSB_OK;
  if (ldx < std::max<int64_t>(rows_x, 1) || ldy < rows_y) HSSB_FAIL(HSSB_ERR_DIM, "hssb_group_matmul: leading dimension too small");
  if ((rows_x > 0 && !X) || !Y) HSSB_FAIL(HSSB_ERR_ARG, "hssb_group_matmul: NULL matrix pointer");
  // every shard cuts the call into the same column blocks; size the workspaces for the widest block up front
  // so that no shard reallocates (and invalidates the peer tables) while another one is already running
  if (int rc = group_prepare(g, nrhs)) return rc;
  // Everything that allocates, frees (cudaFree synchronises the device) or can fail is done HERE, shard by shard,
  // before any kernel can start waiting for a peer: once the shards run concurrently, a synchronising call of one
  // shard would wait for a kernel of another shard of the same device that in turn waits for this shard's push.
  for (hssb_matrix* h : g->shard) {
    DeviceGuard dg(h->device);
    const int64_t x0 = trans ? h->local_row0 : h->local_col0, xr = trans ? h->local_m : h->local_n;
    const int64_t y0 = trans ? h->local_col0 : h->local_row0, yr = trans ? h->local_n : h->local_m;
    if (int rc = matmul_host_impl(h, trans, yr, xr, nrhs, X + x0, ldx, Y + y0, ldy, alpha, beta, /*prepare=*/true)) return rc;
  }
  std::vector<int> rcs((size_t)P, HSSB_OK);
  std::vector<std::string> errs((size_t)P);
  std::vector<std::thread> th;
  th.reserve((size_t)P);
  for (int r = 0; r < P; ++r)
    th.emplace_back([&, r] {
      hssb_matrix* h = g->shard[(size_t)r];
      const int64_t x0 = trans ? h->local_row0 : h->local_col0, xr = trans ? h->local_m : h->local_n;
      const int64_t y0 = trans ? h->local_col0 : h->local_row0, yr = trans ? h->local_n : h->local_m;
      rcs[(size_t)r] = matmul_host_impl(h, trans, yr, xr, nrhs, X + x0, ldx, Y + y0, ldy, alpha, beta);
      if (rcs[(size_t)r]) errs[(size_t)r] = g_err;  // the message lives in this thread's buffer
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < P; ++r)
    if (rcs[(size_t)r]) HSSB_FAIL(rcs[(size_t)r], "shard %d: %s", r, errs[(size_t)r].c_str());
  return HSSB_OK;
}

static int group_dev(hssb_group* g, int trans, int64_t nrhs, const double* const* dX, int64_t ldx, double* const* dY, int64_t ldy,
                     double alpha, double beta, void* const* streams) {
  if (int rc = group_check(g, "hssb_group_matmul_dev")) return rc;
  if (!dX || !dY) HSSB_FAIL(HSSB_ERR_ARG, "hssb_group_matmul_dev: NULL pointer table");
  if (int rc = group_prepare(g, nrhs)) return rc;
  const int P = (int)g->shard.size();
  // pass 0 sets every shard up (workspaces, twin pool, graph capture + upload) without launching, pass 1 queues the
  // products: once a shard's kernels are waiting for their peers, nothing that could synchronise a device follows
  for (int pass = 0; pass < 2; ++pass)
    for (int r = 0; r < P; ++r) {
      hssb_matrix* h = g->shard[(size_t)r];
      const int64_t xr = trans ? h->local_m : h->local_n, yr = trans ? h->local_n : h->local_m;
      h->prepare_only = pass == 0;
      const int rc = matmul_dev_impl(h, trans, yr, xr, nrhs, dX[r], ldx, dY[r], ldy, alpha, beta, streams ? streams[r] : (void*)h->stream);
      h->prepare_only = false;
      if (rc) HSSB_FAIL(rc, "shard %d: %s", r, std::string(g_err).c_str());
    }
  return HSSB_OK;
}

}  // namespace hssb

extern "C" {

int hssb_group_create_synthetic(int64_t n, int64_t leafsize, int64_t rank, uint64_t seed, const int* devices, int n_devices,
                                hssb_group** out) {
  return guarded<int>([&]() -> int {
  if (!out) HSSB_FAIL(HSSB_ERR_ARG, "hssb_group_create_synthetic: out is NULL");
  *out = nullptr;
  if (int rc = group_devices_ok(devices, n_devices, "hssb_group_create_synthetic")) return rc;
  std::unique_ptr<hssb_group> g(new (std::nothrow) hssb_group());
  if (!g) HSSB_FAIL(HSSB_ERR_ALLOC, "hssb_group_create_synthetic: out of memory");
  for (int r = 0; r < n_devices; ++r) {
    hssb_matrix* h = nullptr;
    const int rc = hssb_create_synthetic(n, leafsize, rank, seed, devices[r], r, n_devices, &h);
    if (rc) { hssb_group_destroy(g.release()); return rc; }
    g->shard.push_back(h);
  }
  const int rc = group_wire(g, out);
  if (rc) hssb_group_destroy(g.release());
  return rc;
  });
}

int hssb_group_finalize(hssb_builder* b, int64_t root, const int* devices, int n_devices, hssb_group** out) {
  return guarded<int>([&]() -> int {
  if (!out) HSSB_FAIL(HSSB_ERR_ARG, "hssb_group_finalize: out is NULL");
  *out = nullptr;
  if (int rc = group_devices_ok(devices, n_devices, "hssb_group_finalize")) return rc;
  std::unique_ptr<hssb_group> g(new (std::nothrow) hssb_group());
  if (!g) HSSB_FAIL(HSSB_ERR_ALLOC, "hssb_group_finalize: out of memory");
  for (int r = 0; r < n_devices; ++r) {
    hssb_matrix* h = nullptr;
    const int rc = hssb_builder_finalize(b, root, devices[r], r, n_devices, &h);  // prunes the other shards' subtrees
    if (rc) { hssb_group_destroy(g.release()); return rc; }
    g->shard.push_back(h);
  }
  const int rc = group_wire(g, out);
  if (rc) hssb_group_destroy(g.release());
  return rc;
  });
}

int hssb_group_destroy(hssb_group* g) {
  return guarded<int>([&]() -> int {
  if (!g) return HSSB_OK;
  // quiesce every device before any workspace a peer may still be writing to goes away
  for (hssb_matrix* h : g->shard)
    if (h && h->device >= 0) {
      DeviceGuard dg(h->device);
      cudaDeviceSynchronize();
    }
  for (hssb_matrix* h : g->shard) hssb_destroy(h);
  delete g;
  return HSSB_OK;
  });
}

int hssb_group_size(const hssb_group* g) { return g ? (int)g->shard.size() : 0; }

hssb_matrix* hssb_group_shard(hssb_group* g, int i) {
  if (!g || i < 0 || i >= (int)g->shard.size()) return nullptr;
  return g->shard[(size_t)i];
}

int hssb_group_reserve(hssb_group* g, int64_t max_nrhs) {
  return guarded<int>([&]() -> int {
  if (int rc = group_check(g, "hssb_group_reserve")) return rc;
  if (max_nrhs < 0) HSSB_FAIL(HSSB_ERR_ARG, "hssb_group_reserve: bad argument");
  if (g->shard.size() == 1) return hssb_reserve(g->shard[0], max_nrhs);
  return group_prepare(g, max_nrhs);
  });
}

int hssb_group_matmul(hssb_group* g, int64_t rows_y, int64_t rows_x, int64_t nrhs, const double* X, int64_t ldx, double* Y, int64_t ldy,
                      double alpha, double beta) {
  return guarded<int>([&]() -> int {
    return group_host(g, 0, rows_y, rows_x, nrhs, X, ldx, Y, ldy, alpha, beta);
  });
}

int hssb_group_matmul_t(hssb_group* g, int64_t rows_y, int64_t rows_x, int64_t nrhs, const double* X, int64_t ldx, double* Y,
                        int64_t ldy, double alpha, double beta) {
  return guarded<int>([&]() -> int {
    return group_host(g, 1, rows_y, rows_x, nrhs, X, ldx, Y, ldy, alpha, beta);
  });
}

int hssb_group_matmul_dev(hssb_group* g, int64_t nrhs, const double* const* dX, int64_t ldx, double* const* dY, int64_t ldy, double alpha,
                          double beta, void* const* streams) {
  return guarded<int>([&]() -> int {
    return group_dev(g, 0, nrhs, dX, ldx, dY, ldy, alpha, beta, streams);
  });
}

int hssb_group_sync(hssb_group* g) {
  return guarded<int>([&]() -> int {
  if (int rc = group_check(g, "hssb_group_sync")) return rc;
  for (hssb_matrix* h : g->shard)
    if (int rc = hssb_sync(h)) return rc;
  return HSSB_OK;
  });
}

}  // extern "C"
