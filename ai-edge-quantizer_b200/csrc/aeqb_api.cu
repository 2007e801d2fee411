// extern "C" surface of libaeqb200.so: argument validation, error strings,
// device-attribute caching.  All arithmetic lives in the kernel translation units.
#include <stdarg.h>
#include <stdio.h>

#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/aeqb200.h"
#include "aeqb_kernels.h"

namespace {

thread_local char g_err[512] = "";

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

int check(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  return fail("%s: %s", what, cudaGetErrorString(e));
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool bits_ok(int bits) { return bits == 2 || bits == 4 || bits == 8; }

}  // namespace

namespace aeqb {
std::atomic<long long> g_launches{0};
int host_fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
int host_check(cudaError_t e, const char* what) { return check(e, what); }
int sm_count_cached() { return sm_count(); }
}  // namespace aeqb

extern "C" {

int aeqb_version(void) { return AEQB_VERSION; }
const char* aeqb_last_error(void) { return g_err; }
int64_t aeqb_launch_count(void) { return aeqb::g_launches.load(); }

// ---- shared batching logic --------------------------------------------------
namespace {

// ---- device job tables -----------------------------------------------------------------------
// More than kMaxInlineJobs tensors of one kernel class go out as ONE persistent launch whose job
// table lives in device memory.  The table is staged in a small ring of pinned host buffers (a
// slot is reused only after the copy that read it has completed) and copied on the caller's
// stream into a stream-ordered allocation that is freed, again in stream order, behind the
// kernel — so the call stays asynchronous and nothing outlives it.
struct TableRing {
  static constexpr int kSlots = 8;
  void* host[kSlots] = {};
  size_t cap[kSlots] = {};
  cudaEvent_t ev[kSlots] = {};
  bool used[kSlots] = {};
  int next = 0;
};
std::mutex g_table_mu;
TableRing g_table_ring[64];

bool use_job_tables() {
  static const bool on = !(getenv("AEQB_NO_JOB_TABLE") && atoi(getenv("AEQB_NO_JOB_TABLE")));
  return on;
}

int upload_table(const void* src, size_t bytes, cudaStream_t st, void** d_table) {
  int dev = 0;
  if (int rc = check(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (dev < 0 || dev >= 64) return fail("device index %d out of range", dev);
  std::lock_guard<std::mutex> lock(g_table_mu);
  TableRing& r = g_table_ring[dev];
  const int i = r.next;
  r.next = (r.next + 1) % TableRing::kSlots;
  if (!r.ev[i]) { if (int rc = check(cudaEventCreateWithFlags(&r.ev[i], cudaEventDisableTiming), "cudaEventCreate")) return rc; }
  if (r.used[i]) { if (int rc = check(cudaEventSynchronize(r.ev[i]), "cudaEventSynchronize")) return rc; }
  if (r.cap[i] < bytes) {
    if (r.host[i]) cudaFreeHost(r.host[i]);
    r.host[i] = nullptr; r.cap[i] = 0;
    const size_t want = (bytes + 65535) & ~size_t(65535);
    if (int rc = check(cudaHostAlloc(&r.host[i], want, cudaHostAllocDefault), "cudaHostAlloc")) return rc;
    r.cap[i] = want;
  }
  std::memcpy(r.host[i], src, bytes);
  if (int rc = check(cudaMallocAsync(d_table, bytes, st), "cudaMallocAsync")) return rc;
  if (int rc = check(cudaMemcpyAsync(*d_table, r.host[i], bytes, cudaMemcpyHostToDevice, st), "job table H2D")) {
    cudaFreeAsync(*d_table, st);
    return rc;
  }
  r.used[i] = true;
  return check(cudaEventRecord(r.ev[i], st), "cudaEventRecord");
}

struct RowsOpts { int bits, symmetric; };

// Runs every job: stream-class jobs are grouped into <= kMaxInlineJobs batches per
// class (one persistent launch each), the rest go through the generic kernel.
// Per-channel scales are 4 bytes per row: stored into the peers from the requantisation kernel one
// row at a time they are one NVLink packet each (13.7 M packets per rank and step for the 8 B-parameter
// set at N = 8: 8.53 ms per step against 6.66 without the exchange).  So the kernel writes its own
// copy only and a second, tiny launch on the same stream pushes every job's scale vector to the
// peers with 16-byte stores.  AEQB_MIRROR_IN_KERNEL=1 restores the in-kernel stores (A/B).
bool mirror_in_kernel() {
  static const bool v = getenv("AEQB_MIRROR_IN_KERNEL") && atoi(getenv("AEQB_MIRROR_IN_KERNEL"));
  return v;
}

int run_rows(const aeqb::RowsJob* jobs, int64_t n, RowsOpts o, cudaStream_t st, const char* who,
             const aeqb::PeerMirror* peers_in = nullptr) {
  const int sms = sm_count();
  const aeqb::PeerMirror* peers = (peers_in && !mirror_in_kernel()) ? nullptr : peers_in;
  if (peers_in && peers_in->n > 0 && peers == nullptr) {  // mirror afterwards, whatever kernel a tensor takes
    if (int rc = run_rows(jobs, n, o, st, who, nullptr)) return rc;
    std::vector<aeqb::MirrorSpan> spans;
    long long max_n = 0;
    for (int64_t i = 0; i < n; ++i) {
      if (jobs[i].rows <= 0 || jobs[i].cols <= 0 || !jobs[i].scale) continue;
      const long long cnt = jobs[i].out_stride ? jobs[i].rows : 1;
      spans.push_back(aeqb::MirrorSpan{jobs[i].scale, cnt});
      if (cnt > max_n) max_n = cnt;
    }
    if (spans.empty()) return 0;
    void* d_spans = nullptr;
    if (int rc = upload_table(spans.data(), spans.size() * sizeof(aeqb::MirrorSpan), st, &d_spans)) return rc;
    const int rc = check(aeqb::launch_mirror_f32(static_cast<const aeqb::MirrorSpan*>(d_spans),
                                                 static_cast<int>(spans.size()), max_n, *peers_in, st), who);
    cudaFreeAsync(d_spans, st);
    return rc;
  }
  long long batch_bytes = 0;  // the class of a job depends on the size of the batch it travels in
  for (int64_t i = 0; i < n; ++i)
    if (jobs[i].rows > 0 && jobs[i].cols > 0) batch_bytes += jobs[i].rows * static_cast<long long>(jobs[i].cols) * 4;
  bool done_class[5] = {false, false, false, false, false};
  if (peers && peers->n > 0) {  // only the tile-stream kernels mirror their scales
    for (int64_t i = 0; i < n; ++i)
      if (jobs[i].rows > 0 && jobs[i].cols > 0 && aeqb::rows_job_class(jobs[i], o.bits, batch_bytes) == 0)
        return fail("%s: tensor %lld ([%lld, %d]) does not take the tile-stream kernel; gather its "
                    "scales with a collective", who, (long long)i, (long long)jobs[i].rows, jobs[i].cols);
  }
  for (int klass = 1; klass <= 4 && use_job_tables(); ++klass) {  // one launch per class when it is big
    std::vector<aeqb::RowsJob> tab;
    long long n_tiles = 0;
    bool rich = false;
    for (int64_t i = 0; i < n; ++i) {
      aeqb::RowsJob j = jobs[i];
      if (j.rows <= 0 || j.cols <= 0 || aeqb::rows_job_class(j, o.bits, batch_bytes) != klass) continue;
      j.rows_per_tile = aeqb::rows_job_rows_per_tile(j, klass);
      j.cpr_magic = static_cast<unsigned>(((1u << 20) + (j.cols / 128) - 1) / (j.cols / 128));
      j.tile0 = n_tiles;
      j.tile_end = j.tile0 + (j.rows + j.rows_per_tile - 1) / j.rows_per_tile;
      n_tiles = j.tile_end;
      rich |= j.mse_k != 0.0f || j.clip != nullptr || j.given_scale != nullptr;
      tab.push_back(j);
    }
    if (tab.size() <= static_cast<size_t>(aeqb::kMaxInlineJobs)) continue;  // the inline path below takes it
    void* d_table = nullptr;
    const size_t ends_bytes = (tab.size() * sizeof(long long) + 127) & ~size_t(127);
    std::vector<unsigned char> blob(ends_bytes + tab.size() * sizeof(aeqb::RowsJob));
    for (size_t t = 0; t < tab.size(); ++t) reinterpret_cast<long long*>(blob.data())[t] = tab[t].tile_end;
    std::memcpy(blob.data() + ends_bytes, tab.data(), tab.size() * sizeof(aeqb::RowsJob));
    if (int rc = upload_table(blob.data(), blob.size(), st, &d_table)) return rc;
    aeqb::RowsBatch b{};
    b.bits = o.bits; b.symmetric = o.symmetric;
    if (peers) b.peers = *peers;
    b.table_ends = static_cast<const long long*>(d_table);
    b.table = reinterpret_cast<const aeqb::RowsJob*>(static_cast<unsigned char*>(d_table) + ends_bytes);
    b.rich = rich ? 1 : 0;
    b.n_jobs = static_cast<int>(tab.size());
    b.n_tiles = n_tiles;
    const int rc = check(aeqb::launch_requant_rows_stream(b, klass, sms, st), who);
    cudaFreeAsync(d_table, st);
    if (rc) return rc;
    done_class[klass] = true;
  }
  for (int klass = 1; klass <= 4; ++klass) {
    if (done_class[klass]) continue;
    aeqb::RowsBatch b{};
    b.bits = o.bits; b.symmetric = o.symmetric;
    if (peers) b.peers = *peers;
    auto flush = [&]() -> int {
      if (b.n_jobs == 0) return 0;
      int rc = check(aeqb::launch_requant_rows_stream(b, klass, sms, st), who);
      b.n_jobs = 0; b.n_tiles = 0;
      return rc;
    };
    for (int64_t i = 0; i < n; ++i) {
      aeqb::RowsJob j = jobs[i];
      if (j.rows <= 0 || j.cols <= 0 || aeqb::rows_job_class(j, o.bits, batch_bytes) != klass) continue;
      j.rows_per_tile = aeqb::rows_job_rows_per_tile(j, klass);
      j.cpr_magic = static_cast<unsigned>(((1u << 20) + (j.cols / 128) - 1) / (j.cols / 128));
      j.tile0 = b.n_tiles;
      j.tile_end = j.tile0 + (j.rows + j.rows_per_tile - 1) / j.rows_per_tile;
      b.n_tiles = j.tile_end;
      b.jobs[b.n_jobs++] = j;
      if (b.n_jobs == aeqb::kMaxInlineJobs) { if (int rc = flush()) return rc; }
    }
    if (int rc = flush()) return rc;
  }
  for (int64_t i = 0; i < n; ++i) {
    const aeqb::RowsJob& j = jobs[i];
    if (j.rows <= 0 || j.cols <= 0 || aeqb::rows_job_class(j, o.bits, batch_bytes) != 0) continue;
    if (j.packed && (j.cols % (8 / o.bits) != 0))
      return fail("%s: packed output of a [%lld, %d] tensor would straddle rows; pack separately",
                  who, (long long)j.rows, j.cols);
    if (int rc = check(aeqb::launch_requant_rows_generic(j, o.bits, o.symmetric, st), who)) return rc;
  }
  return 0;
}

int run_blocks(const aeqb::BlocksJob* jobs, int64_t n, int block, int bits, cudaStream_t st,
               const char* who, const aeqb::PeerMirror* peers = nullptr) {
  const int sms = sm_count();
  aeqb::BlocksBatch b{};
  b.block = block; b.bits = bits;
  if (peers && peers->n > 0) {  // only the tile-stream kernel mirrors, and only packed + fp16 scale jobs
    b.peers = *peers;
    for (int64_t i = 0; i < n; ++i) {
      const aeqb::BlocksJob& j = jobs[i];
      if (j.n <= 0) continue;
      if (!aeqb::blocks_job_streamable(j) || j.q || !j.packed || !j.scale_f16 ||
          (reinterpret_cast<uintptr_t>(j.scale_f16) & 3))
        return fail("%s: tensor %lld cannot mirror its scales (needs 16-byte aligned x, packed + fp16 "
                    "scale outputs only, 4-byte aligned scale_f16)", who, (long long)i);
    }
  }
  bool bq = false, bp = false;
  // One launch for a model of more than kMaxInlineJobs streamable tensors with one output set.
  if (use_job_tables()) {
    std::vector<aeqb::BlocksJob> tab;
    long long n_tiles = 0;
    bool uniform = true, tq = false, tp = false;
    for (int64_t i = 0; i < n && uniform; ++i) {
      aeqb::BlocksJob j = jobs[i];
      if (j.n <= 0) continue;
      const bool jq = j.q != nullptr, jp = j.packed != nullptr;
      if (!aeqb::blocks_job_streamable(j) || (!tab.empty() && (jq != tq || jp != tp))) { uniform = false; break; }
      tq = jq; tp = jp;
      j.tile0 = n_tiles;
      j.tile_end = j.tile0 + aeqb::blocks_job_tiles(j.n);
      n_tiles = j.tile_end;
      tab.push_back(j);
    }
    if (uniform && tab.size() > static_cast<size_t>(aeqb::kMaxInlineJobs)) {
      void* d_table = nullptr;
      const size_t ends_bytes = (tab.size() * sizeof(long long) + 127) & ~size_t(127);
      std::vector<unsigned char> blob(ends_bytes + tab.size() * sizeof(aeqb::BlocksJob));
      for (size_t t = 0; t < tab.size(); ++t) reinterpret_cast<long long*>(blob.data())[t] = tab[t].tile_end;
      std::memcpy(blob.data() + ends_bytes, tab.data(), tab.size() * sizeof(aeqb::BlocksJob));
      if (int rc = upload_table(blob.data(), blob.size(), st, &d_table)) return rc;
      b.table_ends = static_cast<const long long*>(d_table);
      b.table = reinterpret_cast<const aeqb::BlocksJob*>(static_cast<unsigned char*>(d_table) + ends_bytes);
      b.n_jobs = static_cast<int>(tab.size());
      b.n_tiles = n_tiles;
      const int rc = check(aeqb::launch_requant_blocks_stream(b, tq, tp, sms, st), who);
      cudaFreeAsync(d_table, st);
      return rc;
    }
  }
  auto flush = [&]() -> int {
    if (b.n_jobs == 0) return 0;
    int rc = check(aeqb::launch_requant_blocks_stream(b, bq, bp, sms, st), who);
    b.n_jobs = 0; b.n_tiles = 0;
    return rc;
  };
  for (int64_t i = 0; i < n; ++i) {
    aeqb::BlocksJob j = jobs[i];
    if (j.n <= 0) continue;
    if (!aeqb::blocks_job_streamable(j)) {
      if (int rc = check(aeqb::launch_requant_blocks_generic(j, block, bits, st), who)) return rc;
      continue;
    }
    const bool jq = j.q != nullptr, jp = j.packed != nullptr;
    if (b.n_jobs && (jq != bq || jp != bp)) { if (int rc = flush()) return rc; }
    bq = jq; bp = jp;
    j.tile0 = b.n_tiles;
    j.tile_end = j.tile0 + aeqb::blocks_job_tiles(j.n);
    b.n_tiles = j.tile_end;
    b.jobs[b.n_jobs++] = j;
    if (b.n_jobs == aeqb::kMaxInlineJobs) { if (int rc = flush()) return rc; }
  }
  return flush();
}

int rows_args_ok(int64_t rows, int64_t cols, int bits, const void* packed, const void* x) {
  if (rows < 0 || cols < 0 || cols > 0x7fffffff) return fail("bad shape [%lld, %lld]", (long long)rows, (long long)cols);
  if (!bits_ok(bits)) return fail("unsupported num_bits %d (2, 4 or 8)", bits);
  if (packed && bits == 8) return fail("packed output needs num_bits 2 or 4");
  if (rows * cols > 0 && !x) return fail("x is NULL");
  return 0;
}

int blocks_args_ok(int64_t rows, int64_t cols, int block, int bits, const void* packed, const void* x) {
  if (rows < 0 || cols < 0) return fail("bad shape [%lld, %lld]", (long long)rows, (long long)cols);
  if (block != 32 && block != 64 && block != 128 && block != 256) return fail("unsupported block size %d", block);
  if (cols % block)
    return fail("Quantized dimension %lld is not divisible by block size %d.", (long long)cols, block);
  if (!bits_ok(bits)) return fail("unsupported num_bits %d (2, 4 or 8)", bits);
  if (packed && bits != 4) return fail("fused packed output needs num_bits 4");
  if (rows * cols > 0 && !x) return fail("x is NULL");
  return 0;
}

}  // namespace

int aeqb_requant_rows_f32(const float* x, int64_t rows, int64_t cols, int bits, int symmetric,
                          const float* clip, int8_t* q, uint8_t* packed, float* scale,
                          int32_t* zp, void* stream) {
  if (int rc = rows_args_ok(rows, cols, bits, packed, x)) return rc;
  aeqb::RowsJob j{};
  j.x = x; j.q = q; j.packed = packed; j.scale = scale; j.zp = zp; j.clip = clip;
  j.rows = rows; j.cols = static_cast<int>(cols);
  j.mm_stride = 1; j.clip_stride = 1; j.out_stride = 1;
  return run_rows(&j, 1, {bits, symmetric ? 1 : 0}, static_cast<cudaStream_t>(stream),
                  "aeqb_requant_rows_f32");
}

int aeqb_requant_rows_batch_f32(const aeqb_rows_job* jobs, int64_t n_jobs, int bits, int symmetric,
                                void* stream) {
  if (n_jobs < 0 || (n_jobs > 0 && !jobs)) return fail("bad job list");
  std::vector<aeqb::RowsJob> v(static_cast<size_t>(n_jobs));
  for (int64_t i = 0; i < n_jobs; ++i) {
    const aeqb_rows_job& a = jobs[i];
    if (int rc = rows_args_ok(a.rows, a.cols, bits, a.packed, a.x)) return rc;
    aeqb::RowsJob j{};
    j.x = a.x; j.q = a.q; j.packed = a.packed; j.scale = a.scale; j.zp = a.zp; j.clip = a.clip;
    j.rows = a.rows; j.cols = static_cast<int>(a.cols);
    j.mm_stride = 1; j.clip_stride = 1; j.out_stride = 1;
    v[static_cast<size_t>(i)] = j;
  }
  return run_rows(v.data(), n_jobs, {bits, symmetric ? 1 : 0}, static_cast<cudaStream_t>(stream),
                  "aeqb_requant_rows_batch_f32");
}

int aeqb_requant_rows_batch_mirror_f32(const aeqb_rows_job* jobs, int64_t n_jobs, int bits,
                                       int symmetric, const int64_t* peer_delta_bytes, int n_peers,
                                       void* stream) {
  if (n_jobs < 0 || (n_jobs > 0 && !jobs)) return fail("bad job list");
  if (n_peers < 0 || n_peers > aeqb::kMaxPeers || (n_peers > 0 && !peer_delta_bytes))
    return fail("bad peer list (0..%d peers)", aeqb::kMaxPeers);
  aeqb::PeerMirror pm{};
  pm.n = n_peers;
  for (int i = 0; i < n_peers; ++i) {
    if (peer_delta_bytes[i] % 4) return fail("peer offset %d is not a multiple of 4 bytes", i);
    pm.delta[i] = peer_delta_bytes[i];
  }
  std::vector<aeqb::RowsJob> v(static_cast<size_t>(n_jobs));
  for (int64_t i = 0; i < n_jobs; ++i) {
    const aeqb_rows_job& a = jobs[i];
    if (int rc = rows_args_ok(a.rows, a.cols, bits, a.packed, a.x)) return rc;
    if (n_peers > 0 && !a.scale) return fail("job %lld has no scale output to mirror", (long long)i);
    aeqb::RowsJob j{};
    j.x = a.x; j.q = a.q; j.packed = a.packed; j.scale = a.scale; j.zp = a.zp; j.clip = a.clip;
    j.rows = a.rows; j.cols = static_cast<int>(a.cols);
    j.mm_stride = 1; j.clip_stride = 1; j.out_stride = 1;
    v[static_cast<size_t>(i)] = j;
  }
  return run_rows(v.data(), n_jobs, {bits, symmetric ? 1 : 0}, static_cast<cudaStream_t>(stream),
                  "aeqb_requant_rows_batch_mirror_f32", &pm);
}

// ---- peer-visible device buffers (CUDA IPC): the gathered scale buffer every rank maps
int aeqb_peer_alloc(size_t bytes, void** ptr, void* handle64) {
  if (!ptr || !handle64 || bytes == 0) return fail("aeqb_peer_alloc: bad arguments");
  void* p = nullptr;
  if (int rc = check(cudaMalloc(&p, bytes), "aeqb_peer_alloc")) return rc;
  if (int rc = check(cudaMemset(p, 0, bytes), "aeqb_peer_alloc")) { cudaFree(p); return rc; }
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
  if (int rc = check(cudaIpcGetMemHandle(&h, p), "aeqb_peer_alloc")) { cudaFree(p); return rc; }
  std::memcpy(handle64, &h, sizeof(h));
  *ptr = p;
  return 0;
}

int aeqb_peer_open(const void* handle64, void** ptr) {
  if (!ptr || !handle64) return fail("aeqb_peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, sizeof(h));
  return check(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess), "aeqb_peer_open");
}

int aeqb_peer_close(void* ptr) { return check(cudaIpcCloseMemHandle(ptr), "aeqb_peer_close"); }
int aeqb_peer_free(void* ptr) { return check(cudaFree(ptr), "aeqb_peer_free"); }

int aeqb_requant_given_minmax_f32(const float* x, int64_t rows, int64_t cols, int bits,
                                  int symmetric, const float* mn, const float* mx,
                                  const float* clip, int per_row, int8_t* q, uint8_t* packed,
                                  float* scale, int32_t* zp, void* stream) {
  if (int rc = rows_args_ok(rows, cols, bits, packed, x)) return rc;
  if (!mn || !mx) return fail("min / max are NULL");
  aeqb::RowsJob j{};
  j.x = x; j.q = q; j.packed = packed; j.scale = scale; j.zp = zp; j.clip = clip;
  j.given_min = mn; j.given_max = mx;
  j.rows = rows; j.cols = static_cast<int>(cols);
  j.mm_stride = j.clip_stride = j.out_stride = per_row ? 1 : 0;
  return run_rows(&j, 1, {bits, symmetric ? 1 : 0}, static_cast<cudaStream_t>(stream),
                  "aeqb_requant_given_minmax_f32");
}

int aeqb_requant_blocks_f32(const float* x, int64_t rows, int64_t cols, int block, int bits,
                            const float* clip, int8_t* q, uint8_t* packed, float* scale,
                            uint16_t* scale_f16, void* stream) {
  if (int rc = blocks_args_ok(rows, cols, block, bits, packed, x)) return rc;
  aeqb::BlocksJob j{};
  j.x = x; j.q = q; j.packed = packed; j.scale = scale; j.scale_f16 = scale_f16; j.clip = clip;
  j.n = rows * cols;
  return run_blocks(&j, 1, block, bits, static_cast<cudaStream_t>(stream), "aeqb_requant_blocks_f32");
}

int aeqb_requant_blocks_batch_f32(const aeqb_blocks_job* jobs, int64_t n_jobs, int block, int bits,
                                  void* stream) {
  if (n_jobs < 0 || (n_jobs > 0 && !jobs)) return fail("bad job list");
  std::vector<aeqb::BlocksJob> v(static_cast<size_t>(n_jobs));
  for (int64_t i = 0; i < n_jobs; ++i) {
    const aeqb_blocks_job& a = jobs[i];
    if (int rc = blocks_args_ok(a.rows, a.cols, block, bits, a.packed, a.x)) return rc;
    aeqb::BlocksJob j{};
    j.x = a.x; j.q = a.q; j.packed = a.packed; j.scale = a.scale; j.scale_f16 = a.scale_f16;
    j.clip = a.clip; j.n = a.rows * a.cols;
    v[static_cast<size_t>(i)] = j;
  }
  return run_blocks(v.data(), n_jobs, block, bits, static_cast<cudaStream_t>(stream),
                    "aeqb_requant_blocks_batch_f32");
}

int aeqb_requant_blocks_batch_mirror_f32(const aeqb_blocks_job* jobs, int64_t n_jobs, int block,
                                         int bits, const int64_t* peer_delta_bytes, int n_peers,
                                         void* stream) {
  if (n_jobs < 0 || (n_jobs > 0 && !jobs)) return fail("bad job list");
  if (n_peers < 0 || n_peers > aeqb::kMaxPeers || (n_peers > 0 && !peer_delta_bytes))
    return fail("bad peer list (0..%d peers)", aeqb::kMaxPeers);
  aeqb::PeerMirror pm{};
  pm.n = n_peers;
  for (int i = 0; i < n_peers; ++i) {
    if (peer_delta_bytes[i] % 4) return fail("peer offset %d is not a multiple of 4 bytes", i);
    pm.delta[i] = peer_delta_bytes[i];
  }
  std::vector<aeqb::BlocksJob> v(static_cast<size_t>(n_jobs));
  for (int64_t i = 0; i < n_jobs; ++i) {
    const aeqb_blocks_job& a = jobs[i];
    if (int rc = blocks_args_ok(a.rows, a.cols, block, bits, a.packed, a.x)) return rc;
    aeqb::BlocksJob j{};
    j.x = a.x; j.q = a.q; j.packed = a.packed; j.scale = a.scale; j.scale_f16 = a.scale_f16;
    j.clip = a.clip; j.n = a.rows * a.cols;
    v[static_cast<size_t>(i)] = j;
  }
  return run_blocks(v.data(), n_jobs, block, bits, static_cast<cudaStream_t>(stream),
                    "aeqb_requant_blocks_batch_mirror_f32", &pm);
}

int aeqb_ema_sequence_f32(const float* pairs, int64_t n, double smoothing, float* out2, void* stream) {
  if (n < 0) return fail("negative batch count");
  if (!out2 || (n > 0 && !pairs)) return fail("pairs / out2 are NULL");
  return check(aeqb::launch_ema_sequence(pairs, n, smoothing, out2, static_cast<cudaStream_t>(stream)),
               "aeqb_ema_sequence_f32");
}

size_t aeqb_minmax_workspace_bytes(void) { return aeqb::minmax_workspace_bytes(); }

int aeqb_minmax_tensors_f32(const aeqb_minmax_job* jobs, int64_t n_jobs, float lo, float hi,
                            int use_lo, int use_hi, void* ws, void* stream) {
  if (n_jobs < 0 || (n_jobs > 0 && !jobs)) return fail("bad job list");
  if (n_jobs > 0 && !ws) return fail("ws is NULL");
  for (int64_t i0 = 0; i0 < n_jobs; i0 += aeqb::kMaxInlineJobs) {
    aeqb::MinmaxBatch b{};
    for (int64_t i = i0; i < n_jobs && i < i0 + aeqb::kMaxInlineJobs; ++i) {
      if (jobs[i].n < 0) return fail("negative element count");
      if (!jobs[i].out2) return fail("out2 is NULL");
      if (jobs[i].n > 0 && !jobs[i].x) return fail("x is NULL");
      aeqb::MinmaxJob& m = b.jobs[b.n_jobs++];
      m.x = jobs[i].x; m.n = jobs[i].n; m.out2 = jobs[i].out2;
    }
    if (int rc = check(aeqb::launch_minmax_tensors(b, lo, hi, use_lo, use_hi, ws, sm_count(),
                                                   static_cast<cudaStream_t>(stream)),
                       "aeqb_minmax_tensors_f32"))
      return rc;
  }
  return 0;
}

int aeqb_minmax_tensor_f32(const float* x, int64_t n, float lo, float hi, int use_lo, int use_hi,
                           float* out2, void* ws, void* stream) {
  aeqb_minmax_job j;
  j.x = x; j.n = n; j.out2 = out2;
  return aeqb_minmax_tensors_f32(&j, 1, lo, hi, use_lo, use_hi, ws, stream);
}

int aeqb_hist_accumulate_f32(const float* x, int64_t n, float lower_bound, float bin_width, int nbins,
                             int finite_only, int64_t* counts, void* stream) {
  if (n < 0) return fail("negative element count");
  if (nbins < 1 || nbins > 12288) return fail("nbins must be in [1, 12288], got %d", nbins);
  if (n > 0 && (!x || !counts)) return fail("x / counts are NULL");
  return check(aeqb::launch_hist(x, n, lower_bound, bin_width, nbins, finite_only,
                                 reinterpret_cast<long long*>(counts), sm_count(),
                                 static_cast<cudaStream_t>(stream)),
               "aeqb_hist_accumulate_f32");
}

int aeqb_row_stats_f32(const float* x, int64_t rows, int64_t cols, float* mn, float* mx,
                       float* sumsq, void* stream) {
  if (rows < 0 || cols < 0 || cols > 0x7fffffff) return fail("bad shape [%lld, %lld]", (long long)rows, (long long)cols);
  return check(aeqb::launch_row_stats(x, rows, static_cast<int>(cols), mn, mx, sumsq,
                                      static_cast<cudaStream_t>(stream)),
               "aeqb_row_stats_f32");
}

int aeqb_minmax_blocks_f32(const float* x, int64_t rows, int64_t cols, int block, float* mn,
                           float* mx, void* stream) {
  if (block != 32 && block != 64 && block != 128 && block != 256) return fail("unsupported block size %d", block);
  if (cols % block)
    return fail("Quantized dimension %lld is not divisible by block size %d.", (long long)cols, block);
  if (!mn || !mx) return fail("min / max are NULL");
  return check(aeqb::launch_block_minmax(x, rows * cols, block, mn, mx,
                                         static_cast<cudaStream_t>(stream)),
               "aeqb_minmax_blocks_f32");
}

size_t aeqb_octav_workspace_bytes(int64_t groups, int max_iterations) {
  if (groups < 0 || max_iterations < 1) return 0;
  return aeqb::octav_workspace_bytes(groups, max_iterations);
}

int aeqb_octav_clip_rows_f32(const float* x, int64_t rows, int64_t cols, int bits,
                             int max_iterations, float exponent_divisor, int early_stop,
                             float* clip, void* ws, void* stream) {
  if (rows < 0 || cols < 0) return fail("bad shape [%lld, %lld]", (long long)rows, (long long)cols);
  if (bits < 1 || bits > 16) return fail("unsupported num_bits %d", bits);
  if (max_iterations < 1 || max_iterations > 32) return fail("max_iterations must be in [1, 32]");
  if (rows * cols > 0 && (!x || !clip || !ws)) return fail("x / clip / ws are NULL");
  return check(aeqb::launch_octav_rows(x, rows, cols, bits, max_iterations, exponent_divisor,
                                       early_stop, clip, ws, sm_count(),
                                       static_cast<cudaStream_t>(stream)),
               "aeqb_octav_clip_rows_f32");
}

int aeqb_octav_clip_blocks_f32(const float* x, int64_t rows, int64_t cols, int block, int bits,
                               int max_iterations, float exponent_divisor, int early_stop,
                               float* clip, void* ws, void* stream) {
  if (rows < 0 || cols < 0) return fail("bad shape [%lld, %lld]", (long long)rows, (long long)cols);
  if (block != 32 && block != 64 && block != 128 && block != 256) return fail("unsupported block size %d", block);
  if (cols % block)
    return fail("Quantized dimension %lld is not divisible by block size %d.", (long long)cols, block);
  if (bits < 1 || bits > 16) return fail("unsupported num_bits %d", bits);
  if (max_iterations < 1 || max_iterations > 32) return fail("max_iterations must be in [1, 32]");
  if (rows * cols > 0 && (!x || !clip || !ws)) return fail("x / clip / ws are NULL");
  return check(aeqb::launch_octav_blocks(x, rows * cols, block, bits, max_iterations,
                                         exponent_divisor, early_stop, clip, ws, sm_count(),
                                         static_cast<cudaStream_t>(stream)),
               "aeqb_octav_clip_blocks_f32");
}

size_t aeqb_mse_workspace_bytes(void) { return aeqb::mse_workspace_bytes(); }

int aeqb_mse_scale_rows_f32(const float* x, int64_t rows, int64_t cols, float k, float* scale,
                            void* ws, void* stream) {
  if (rows < 0 || cols <= 0) return fail("bad shape [%lld, %lld]", (long long)rows, (long long)cols);
  if (rows > 0 && (!x || !scale)) return fail("x / scale are NULL");
  return check(aeqb::launch_mse_scale_rows(x, rows, cols, k, scale, ws, sm_count(),
                                           static_cast<cudaStream_t>(stream)),
               "aeqb_mse_scale_rows_f32");
}

int aeqb_hadamard_rows_f32(const float* x, int64_t rows, int64_t cols, int64_t n, float* out,
                           void* stream) {
  if (rows < 0 || cols < 0) return fail("bad shape [%lld, %lld]", (long long)rows, (long long)cols);
  if (n < 2 || (n & (n - 1))) return fail("Hadamard matrix size must be a power of 2. ");
  if (cols % n) return fail("hadamard size %lld does not divide the last dimension %lld", (long long)n, (long long)cols);
  if (rows * cols > 0 && (!x || !out)) return fail("x / out are NULL");
  return check(aeqb::launch_hadamard_rows(x, rows, cols, n, out, sm_count(),
                                          static_cast<cudaStream_t>(stream)),
               "aeqb_hadamard_rows_f32");
}

size_t aeqb_xtx_workspace_bytes(int64_t tokens, int64_t k) {
  return aeqb::xtx_workspace_bytes(tokens, k, sm_count());
}

int aeqb_xtx_f32(const float* x, int64_t tokens, int64_t k, double alpha, double* hessian, void* ws,
                 void* stream) {
  if (tokens < 0 || k < 0 || k > 0x7fffffff) return fail("bad shape [%lld, %lld]", (long long)tokens, (long long)k);
  if (k > 0 && !hessian) return fail("hessian is NULL");
  if (tokens * k > 0 && !x) return fail("x is NULL");
  return check(aeqb::launch_xtx_f64(x, tokens, k, alpha, hessian, ws, sm_count(),
                                    static_cast<cudaStream_t>(stream)),
               "aeqb_xtx_f32");
}

size_t aeqb_hessian_inverse_workspace_bytes(int64_t k) {
  return k > 0 ? aeqb::hessian_inverse_workspace_bytes(k) : 0;
}

int aeqb_hessian_inverse_f64(double* hessian, int64_t k, double damp, int keep_damped_diagonal,
                             float* hinv, void* ws, int* info, void* stream) {
  if (k < 0 || k > 0x7fffffff) return fail("bad Hessian order %lld", (long long)k);
  if (k > 0 && (!hessian || !hinv || !ws)) return fail("hessian / hinv / ws are NULL");
  return check(aeqb::launch_hessian_inverse(hessian, k, damp, keep_damped_diagonal, hinv, ws, info,
                                            sm_count(), static_cast<cudaStream_t>(stream)),
               "aeqb_hessian_inverse_f64");
}

size_t aeqb_gptq_workspace_bytes(int64_t rows, int64_t k) {
  return (rows > 0 && k > 0) ? aeqb::gptq_workspace_bytes(rows, k) : 0;
}

int aeqb_gptq_quantize_f32(float* w_work, int64_t rows, int64_t k, const float* hinv,
                           const float* scale, const int32_t* zp, int64_t scale_cols, int block,
                           int bits, int symmetric, int blocksize, int8_t* q, void* ws, void* stream) {
  if (rows < 0 || k < 0 || rows > 0x7fffffff || k > 0x7fffffff)
    return fail("bad shape [%lld, %lld]", (long long)rows, (long long)k);
  if (blocksize != 64) return fail("GPTQ blocksize must be 64 (gptq.py:136), got %d", blocksize);
  if (bits < 2 || bits > 8) return fail("unsupported num_bits %d (2..8)", bits);
  if (block != 0 && block != 32 && block != 64 && block != 128 && block != 256)
    return fail("unsupported block size %d", block);
  if (block && (k % block || scale_cols != k / block))
    return fail("blockwise scales must be [rows, k / block]");
  if (!block && scale_cols != 0 && scale_cols != 1) return fail("scale_cols must be 0 or 1 without blocks");
  if (rows * k > 0 && (!w_work || !hinv || !scale || !q || !ws)) return fail("w_work / hinv / scale / q / ws are NULL");
  return check(aeqb::launch_gptq_quantize(w_work, rows, k, hinv, scale, zp,
                                          static_cast<int>(scale_cols), block, bits,
                                          symmetric ? 1 : 0, q, ws, sm_count(),
                                          static_cast<cudaStream_t>(stream)),
               "aeqb_gptq_quantize_f32");
}

int aeqb_hessian_merge_f64(const double* a, double wa, const double* b, double wb, double* out,
                           int64_t n, void* stream) {
  if (n < 0) return fail("negative element count");
  if (n > 0 && (!a || !b || !out)) return fail("a / b / out are NULL");
  return check(aeqb::launch_weighted_mean_f64(a, wa, b, wb, out, n, sm_count(),
                                              static_cast<cudaStream_t>(stream)),
               "aeqb_hessian_merge_f64");
}

int aeqb_scale_zp_from_minmax(const float* mn, const float* mx, const float* clip, int64_t n,
                              int bits, int symmetric, int blockwise, float* scale, int32_t* zp,
                              uint16_t* scale_f16, void* stream) {
  if (bits < 2 || bits > 16) return fail("unsupported num_bits %d", bits);
  if (n > 0 && (!mn || !mx || !scale)) return fail("min / max / scale are NULL");
  return check(aeqb::launch_scale_zp(mn, mx, clip, n, bits, symmetric, blockwise, scale, zp,
                                     scale_f16, static_cast<cudaStream_t>(stream)),
               "aeqb_scale_zp_from_minmax");
}

int aeqb_quantize_f32(const float* x, int64_t n, int64_t channels, int64_t inner,
                      const float* scale, const int32_t* zp, int param_stride, int bits,
                      int symmetric, void* q, void* stream) {
  if (bits < 2 || bits > 16) return fail("unsupported num_bits %d", bits);
  if (channels <= 0 || inner <= 0) return fail("channels / inner must be positive");
  if (n > 0 && (!x || !scale || !q)) return fail("x / scale / q are NULL");
  // [channels, inner] matrices with one parameter per row take the tile-stream kernel (one pass,
  // 128-bit loads, packed stores); every other layout takes the generic element-wise kernel.
  if ((bits == 2 || bits == 4 || bits == 8) && n == channels * inner && inner <= 0x7fffffff &&
      (param_stride == 1 || param_stride == 0)) {
    aeqb::RowsJob j{};
    j.x = x; j.q = static_cast<int8_t*>(q);
    j.given_scale = scale; j.given_zp = zp;
    j.rows = channels; j.cols = static_cast<int>(inner);
    j.mm_stride = param_stride; j.clip_stride = 0; j.out_stride = 0;
    if (aeqb::rows_job_class(j, bits) != 0)
      return run_rows(&j, 1, {bits, symmetric ? 1 : 0}, static_cast<cudaStream_t>(stream),
                      "aeqb_quantize_f32");
  }
  return check(aeqb::launch_quantize(x, n, channels, inner, scale, zp, param_stride, bits,
                                     symmetric, q, sm_count(), static_cast<cudaStream_t>(stream)),
               "aeqb_quantize_f32");
}

int aeqb_dequantize_f32(const void* q, int q_bytes, int64_t n, int64_t channels, int64_t inner,
                        const float* scale, const int32_t* zp, int param_stride, int wrap8,
                        float* out, void* stream) {
  if (q_bytes != 1 && q_bytes != 2 && q_bytes != 4) return fail("q_bytes must be 1, 2 or 4");
  if (channels <= 0 || inner <= 0) return fail("channels / inner must be positive");
  if (n > 0 && (!q || !scale || !out)) return fail("q / scale / out are NULL");
  return check(aeqb::launch_dequantize(q, q_bytes, n, channels, inner, scale, zp, param_stride,
                                       wrap8, out, sm_count(), static_cast<cudaStream_t>(stream)),
               "aeqb_dequantize_f32");
}

int aeqb_pack_bits(const int8_t* q, int64_t n, int bits, uint8_t* out, void* stream) {
  if (bits != 2 && bits != 4) return fail("pack_bits needs num_bits 2 or 4");
  if (n > 0 && (!q || !out)) return fail("q / out are NULL");
  return check(aeqb::launch_pack(q, n, bits, out, sm_count(), static_cast<cudaStream_t>(stream)),
               "aeqb_pack_bits");
}

int aeqb_swap_axes(const void* x, int64_t a, int64_t b, int64_t inner, int elem_bytes, void* out,
                   void* stream) {
  if (elem_bytes != 1 && elem_bytes != 4) return fail("elem_bytes must be 1 or 4");
  if (a < 0 || b < 0 || inner < 0) return fail("bad shape [%lld, %lld, %lld]", (long long)a, (long long)b, (long long)inner);
  if (a * b * inner > 0 && (!x || !out)) return fail("x / out are NULL");
  if (x == out && a * b * inner > 0) return fail("swap_axes cannot run in place");
  return check(aeqb::launch_swap_axes(x, a, b, inner, elem_bytes, out, sm_count(),
                                      static_cast<cudaStream_t>(stream)), "aeqb_swap_axes");
}

size_t aeqb_dwr_workspace_bytes(int64_t n_groups, int64_t group_len) {
  return (n_groups > 0 && group_len > 0) ? aeqb::dwr_workspace_bytes(n_groups, group_len) : 0;
}

int aeqb_dwr_scales_f32(const float* x, int64_t n_groups, int64_t group_len, float* scale, void* ws,
                        void* stream) {
  if (n_groups < 0 || group_len <= 0) return fail("bad group shape [%lld x %lld]", (long long)n_groups, (long long)group_len);
  if (n_groups == 0) return 0;
  if (!x || !scale) return fail("x / scale are NULL");
  if (group_len > 16384 && n_groups != 1)
    return fail("groups longer than 16384 values are supported for a single group only, got %lld groups of %lld",
                (long long)n_groups, (long long)group_len);
  if (group_len > 16384 && !ws) return fail("ws is NULL (aeqb_dwr_workspace_bytes)");
  return check(aeqb::launch_dwr_scales(x, n_groups, group_len, scale, ws, sm_count(),
                                       static_cast<cudaStream_t>(stream)),
               "aeqb_dwr_scales_f32");
}

int aeqb_max_abs_diff_f32(const float* a, const float* b, int64_t n, float* out, void* ws, void* stream) {
  if (n < 0) return fail("bad length %lld", (long long)n);
  if (!out || !ws) return fail("out / ws are NULL");
  if (n > 0 && (!a || !b)) return fail("a / b are NULL");
  return check(aeqb::launch_max_abs_diff(a, b, n, out, ws, sm_count(), static_cast<cudaStream_t>(stream)),
               "aeqb_max_abs_diff_f32");
}

int aeqb_cast_f32_f16(const float* x, int64_t n, uint16_t* out, void* stream) {
  if (n < 0) return fail("bad length %lld", (long long)n);
  if (n > 0 && (!x || !out)) return fail("x / out are NULL");
  return check(aeqb::launch_cast_f16(x, n, out, sm_count(), static_cast<cudaStream_t>(stream)),
               "aeqb_cast_f32_f16");
}

size_t aeqb_colsq_workspace_bytes(int64_t n, int64_t d) {
  return d > 0 ? aeqb::colsq_workspace_bytes(n, d, sm_count()) : 0;
}

int aeqb_colsq_f64(const float* x, int64_t n, int64_t d, double alpha, double* out, void* ws,
                   void* stream) {
  if (n < 0 || d < 0 || d > 0x7fffffff) return fail("bad shape [%lld, %lld]", (long long)n, (long long)d);
  if (d == 0) return 0;
  if (!out || !ws || (n > 0 && !x)) return fail("x / out / ws are NULL");
  return check(aeqb::launch_colsq(x, n, d, alpha, out, ws, sm_count(), static_cast<cudaStream_t>(stream)),
               "aeqb_colsq_f64");
}

static int oscar_group_ok(int64_t n, int64_t d, int64_t g, bool allow_tensor) {
  if (n <= 0 || d <= 0 || d > 0x7fffffff) return fail("bad shape [%lld, %lld]", (long long)n, (long long)d);
  if (g == d || (allow_tensor && g == n * d)) return 0;
  if ((g == 32 || g == 64 || g == 128 || g == 256) && d % g == 0) return 0;
  return fail("group length %lld is not the row, a block size dividing %lld, or the tensor",
              (long long)g, (long long)d);
}

size_t aeqb_oscar_pass_workspace_bytes(int64_t n, int64_t d, int64_t g) {
  return (n > 0 && d > 0 && g > 0 && d % g == 0) ? aeqb::oscar_pass_workspace_bytes(n, d, g, sm_count()) : 0;
}

int aeqb_oscar_pass_f32(const float* w, int64_t n, int64_t d, int64_t g, const double* s,
                        double* group_sq, double* a_eff, void* ws, void* stream) {
  if (int rc = oscar_group_ok(n, d, g, false)) return rc;
  if (!w || !s || !group_sq || !ws) return fail("w / s / group_sq / ws are NULL");
  return check(aeqb::launch_oscar_pass(w, n, d, g, s, group_sq, a_eff, ws, sm_count(),
                                       static_cast<cudaStream_t>(stream)),
               "aeqb_oscar_pass_f32");
}

size_t aeqb_oscar_clip_workspace_bytes(int64_t n, int64_t d, int64_t g) {
  return (n > 0 && d > 0) ? aeqb::oscar_clip_workspace_bytes(n, d, g) : 0;
}

int aeqb_oscar_clip_f32(const float* w, int64_t n, int64_t d, int64_t g, const double* s,
                        const double* m, const double* mass_dev, double mass0, int qmax,
                        double* bound, void* ws, void* stream) {
  if (int rc = oscar_group_ok(n, d, g, true)) return rc;
  if (!w || !s || !m || !bound) return fail("w / s / m / bound are NULL");
  if (qmax <= 0) return fail("bad qmax %d", qmax);
  const bool tensor_form = g == n * d && (n > 1 || d > 16384);
  if (tensor_form && !ws) return fail("ws is NULL (aeqb_oscar_clip_workspace_bytes)");
  if (tensor_form && n * d > 0xffffffffLL) return fail("tensor too large for 32-bit sort indices");
  if (!tensor_form && g == d && d > 16384)
    return fail("rows longer than 16384 columns are not supported (got %lld)", (long long)d);
  if (!tensor_form && g != d && g != n * d && !mass_dev) return fail("mass_dev is NULL");
  return check(aeqb::launch_oscar_clip(w, n, d, g, s, m, mass_dev, mass0, qmax, bound, ws, sm_count(),
                                       static_cast<cudaStream_t>(stream)),
               "aeqb_oscar_clip_f32");
}

int aeqb_oscar_scale_f64(const double* bound, int64_t n, int qmax, int blockwise, double* scale,
                         void* stream) {
  if (n < 0 || qmax <= 0) return fail("bad arguments");
  if (n > 0 && (!bound || !scale)) return fail("bound / scale are NULL");
  return check(aeqb::launch_oscar_scale(bound, n, qmax, blockwise, scale, static_cast<cudaStream_t>(stream)),
               "aeqb_oscar_scale_f64");
}

int aeqb_oscar_quantize_f32(const float* w, int64_t n, int64_t d, int64_t group_len, const double* s,
                            const double* scale, int bits, int8_t* q, void* stream) {
  if (n < 0 || d < 0 || d > 0x7fffffff || group_len <= 0) return fail("bad shape");
  if (!bits_ok(bits)) return fail("num_bits must be 2, 4 or 8, got %d", bits);
  if (n * d == 0) return 0;
  if (!w || !s || !scale || !q) return fail("w / s / scale / q are NULL");
  return check(aeqb::launch_oscar_quantize(w, n, d, group_len, s, scale, bits, q, sm_count(),
                                           static_cast<cudaStream_t>(stream)),
               "aeqb_oscar_quantize_f32");
}

int aeqb_requant_mse_rows_f32(const float* x, int64_t rows, int64_t cols, int bits, float multiplier,
                              int8_t* q, uint8_t* packed, float* scale, int32_t* zp, void* stream) {
  if (!bits_ok(bits)) return fail("num_bits must be 2, 4 or 8, got %d", bits);
  if (rows < 0 || cols < 0 || cols > 0x7fffffff) return fail("bad shape [%lld, %lld]", (long long)rows, (long long)cols);
  if (rows == 0 || cols == 0) return 0;
  if (!x || !scale) return fail("x / scale are NULL");
  if (!(multiplier > 0.0f)) return fail("the MSE multiplier must be positive");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  aeqb::RowsJob j{};
  j.x = x; j.q = q; j.packed = packed; j.scale = scale; j.zp = zp;
  j.rows = rows; j.cols = static_cast<int>(cols);
  j.mm_stride = j.clip_stride = j.out_stride = 1;
  j.mse_k = multiplier;
  if (aeqb::rows_job_class(j, bits) != 0)
    return run_rows(&j, 1, {bits, 1}, st, "aeqb_requant_mse_rows_f32");
  // unaligned / odd shapes: the two unfused kernels
  if (packed) return fail("packed output needs 16-byte aligned rows of a multiple of 128 values");
  if (int rc = check(aeqb::launch_mse_scale_rows(x, rows, cols, multiplier, scale, nullptr, sm_count(), st),
                     "aeqb_requant_mse_rows_f32"))
    return rc;
  if (zp) {
    if (int rc = check(cudaMemsetAsync(zp, 0, static_cast<size_t>(rows) * sizeof(int32_t), st),
                       "aeqb_requant_mse_rows_f32"))
      return rc;
  }
  if (!q) return 0;
  return check(aeqb::launch_quantize(x, rows * cols, rows, cols, scale, nullptr, 1, bits, 1, q, sm_count(), st),
               "aeqb_requant_mse_rows_f32");
}

}  // extern "C"
