// hssb_file.h — the packed format as a versioned file (hssb_save / hssb_load).  The file is
// untrusted input: sizes are checked against the file length and the node table must be a tree in BFS order.
// Included by hssb_api.cu inside its extern "C" block.
#pragma once

// ---- packed format on disk -------------------------------
// The reference has no serialisation; the packed format (tree shape + level-ordered pool) is the
// natural file format: fixtures become reproducible without Julia, and a packed matrix can be
// checkpointed / reloaded without re-walking the pointer tree.
struct FileHeader {
  char magic[8];
  uint32_t version, node_words;
  int64_t n_nodes, pool_len;
  int32_t shard_rank, n_shards, synthetic, padded;
  uint64_t seed;
  int64_t synth_rank;
};
static const char kMagic[8] = {'H', 'S', 'S', 'B', '2', '0', '0', 0};
static const uint32_t kFileVersion = 2;  // bump whenever layout_pool / stored_transposed change
static const uint32_t kNodeWords = 9;

int hssb_save(const hssb_matrix* h, const char* path) {
  return guarded<int>([&]() -> int {
  if (!h || !path) HSSB_FAIL(HSSB_ERR_ARG, "hssb_save: NULL argument");
  FILE* fp = fopen(path, "wb");
  if (!fp) HSSB_FAIL(HSSB_ERR_ARG, "hssb_save: cannot open %s for writing", path);
  FileHeader hd;
  memset(&hd, 0, sizeof(hd));
  memcpy(hd.magic, kMagic, 8);
  hd.version = kFileVersion; hd.node_words = kNodeWords;
  hd.n_nodes = (int64_t)h->nodes.size(); hd.pool_len = h->pool_len;
  hd.shard_rank = h->shard_rank; hd.n_shards = h->n_shards; hd.synthetic = h->synthetic; hd.padded = h->padded;
  hd.seed = h->seed; hd.synth_rank = h->synth_rank;
  bool ok = fwrite(&hd, sizeof(hd), 1, fp) == 1;
  for (const Node& t : h->nodes) {
    const int64_t w[kNodeWords] = {t.left, t.right, t.leaf, t.remote, t.m, t.n, t.kr, t.kw, (int64_t)t.heap_id};
    ok = ok && fwrite(w, sizeof(int64_t), kNodeWords, fp) == kNodeWords;
  }
  if (ok && h->device < 0) {
    ok = fwrite(h->pool_host.data(), sizeof(double), (size_t)h->pool_len, fp) == (size_t)h->pool_len;
  } else if (ok) {
    DeviceGuard dg(h->device);
    const size_t CH = (size_t)1 << 22;
    std::vector<double> buf(CH);
    for (int64_t off = 0; ok && off < h->pool_len; off += (int64_t)CH) {
      const size_t cnt = (size_t)std::min<int64_t>((int64_t)CH, h->pool_len - off);
      if (cudaMemcpy(buf.data(), h->pool_dev + off, cnt * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
      ok = fwrite(buf.data(), sizeof(double), cnt, fp) == cnt;
    }
  }
  ok = (fclose(fp) == 0) && ok;
  if (!ok) HSSB_FAIL(HSSB_ERR_ARG, "hssb_save: write to %s failed", path);
  return HSSB_OK;
  });
}

// device >= 0: load onto that GPU; device < 0: host-only (plan-only) handle for CPU-side inspection.
int hssb_load(const char* path, int device, hssb_matrix** out) {
  return guarded<int>([&]() -> int {
  if (!path || !out) HSSB_FAIL(HSSB_ERR_ARG, "hssb_load: NULL argument");
  *out = nullptr;
  if (device >= 0) { int rc = check_device(device); if (rc) return rc; }
  FILE* fp = fopen(path, "rb");
  if (!fp) HSSB_FAIL(HSSB_ERR_ARG, "hssb_load: cannot open %s", path);
  struct Closer { FILE* f; ~Closer() { if (f) fclose(f); } } closer{fp};
  FileHeader hd;
  if (fread(&hd, sizeof(hd), 1, fp) != 1 || memcmp(hd.magic, kMagic, 8) != 0) HSSB_FAIL(HSSB_ERR_ARG, "hssb_load: %s is not an hssb200 file", path);
  if (hd.version != kFileVersion || hd.node_words != kNodeWords) HSSB_FAIL(HSSB_ERR_ARG, "hssb_load: file version %u, library reads %u", hd.version, kFileVersion);
  // The file is untrusted input: every size is checked against the file length before anything is allocated, and
  // the node table must be a proper tree in BFS order (children of the k-th branch are nodes 2k+1, 2k+2 of the
  // branch sequence: consecutive, after their parent, each referenced exactly once) -- a cycle would send the
  // planner's tree walks into an endless loop.
  if (hd.n_nodes <= 0 || hd.pool_len <= 0 || hd.n_nodes > (int64_t)1 << 40 || hd.pool_len > (int64_t)1 << 48)
    HSSB_FAIL(HSSB_ERR_ARG, "hssb_load: corrupt header");
  if (hd.n_shards < 1 || hd.n_shards > hssb_matrix::MAX_PEERS || !is_pow2(hd.n_shards) || hd.shard_rank < 0 || hd.shard_rank >= hd.n_shards)
    HSSB_FAIL(HSSB_ERR_ARG, "hssb_load: corrupt header (shard %d of %d)", hd.shard_rank, hd.n_shards);
  {
    const long pos = ftell(fp);
    if (pos < 0 || fseek(fp, 0, SEEK_END) != 0) HSSB_FAIL(HSSB_ERR_ARG, "hssb_load: cannot seek in %s", path);
    const long fsize = ftell(fp);
    if (fseek(fp, pos, SEEK_SET) != 0) HSSB_FAIL(HSSB_ERR_ARG, "hssb_load: cannot seek in %s", path);
    const double need = (double)pos + (double)hd.n_nodes * kNodeWords * 8.0 + (double)hd.pool_len * 8.0;
    if ((double)fsize < need) HSSB_FAIL(HSSB_ERR_ARG, "hssb_load: file is truncated (%ld bytes, header asks for %.0f)", fsize, need);
  }
  std::unique_ptr<hssb_matrix> H(new (std::nothrow) hssb_matrix());
  if (!H) HSSB_FAIL(HSSB_ERR_ALLOC, "hssb_load: out of memory");
  H->device = device < 0 ? -1 : device;
  H->shard_rank = hd.shard_rank; H->n_shards = hd.n_shards;
  H->synthetic = hd.synthetic != 0; H->seed = hd.seed; H->synth_rank = hd.synth_rank;
  H->nodes.resize((size_t)hd.n_nodes);
  int64_t next_child = 1;
  for (Node& t : H->nodes) {
    int64_t w[kNodeWords];
    if (fread(w, sizeof(int64_t), kNodeWords, fp) != kNodeWords) HSSB_FAIL(HSSB_ERR_ARG, "hssb_load: file is truncated");
    t.left = w[0]; t.right = w[1]; t.leaf = w[2] != 0; t.remote = w[3] != 0;
    t.m = w[4]; t.n = w[5]; t.kr = w[6]; t.kw = w[7]; t.heap_id = (uint64_t)w[8];
    if (t.m < 0 || t.n < 0 || t.kr < 0 || t.kw < 0 || t.m > (int64_t)1 << 40 || t.n > (int64_t)1 << 40 || t.kr > INT32_MAX || t.kw > INT32_MAX)
      HSSB_FAIL(HSSB_ERR_ARG, "hssb_load: corrupt node table (sizes)");
    if (!t.leaf && !t.remote) {
      if (t.left != next_child || t.right != next_child + 1 || t.right >= hd.n_nodes)
        HSSB_FAIL(HSSB_ERR_ARG, "hssb_load: corrupt node table (not a tree in breadth-first order)");
      next_child += 2;
    } else {
      t.left = t.right = -1;
    }
  }
  if (next_child != hd.n_nodes) HSSB_FAIL(HSSB_ERR_ARG, "hssb_load: corrupt node table (%lld nodes, %lld reachable)", (long long)hd.n_nodes, (long long)next_child);
  int rc;
  if (device < 0) {
    rc = plan_matrix(H.get());
    if (!rc && H->pool_len != hd.pool_len) { set_error("hssb_load: pool layout mismatch"); rc = HSSB_ERR_ARG; }
    if (!rc) {
      H->pool_host.resize((size_t)H->pool_len);
      if (fread(H->pool_host.data(), sizeof(double), (size_t)H->pool_len, fp) != (size_t)H->pool_len) { set_error("hssb_load: file is truncated"); rc = HSSB_ERR_ARG; }
    }
    if (rc) return rc;
  } else {
    // plan first (host), so that a layout mismatch is caught before anything is uploaded
    {
      hssb_matrix probe;
      probe.device = -1; probe.shard_rank = H->shard_rank; probe.n_shards = H->n_shards; probe.nodes = H->nodes;
      rc = plan_matrix(&probe);
      if (rc) return rc;
      if (probe.pool_len != hd.pool_len) HSSB_FAIL(HSSB_ERR_ARG, "hssb_load: pool layout mismatch (file %lld, library %lld doubles)", (long long)hd.pool_len, (long long)probe.pool_len);
    }
    rc = finish_matrix(H.get(), nullptr, fp);
    if (rc) { hssb_destroy(H.release()); return rc; }
  }
  *out = H.release();
  return HSSB_OK;
  });
}
