// Host-buffer entry points: the drop-in call a NumPy caller makes.
//
// Inputs are host arrays (the reference hands out read-only NumPy views onto the
// mmap'd flatbuffer, utils/tfl_flatbuffer_utils.py:254-263; alignment is not
// guaranteed), outputs are host arrays owned by the caller.  A batch of tensors
// is cut into row chunks of ~32 MiB of fp32 and pushed through a ring of slots PER
// DEVICE:
//     [worker threads: pageable -> pinned]  ->  H2D  ->  fused kernel  ->  D2H  ->
//     [worker threads: pinned -> user memory]
// on one stream per slot, so the staging of chunk i+2, the upload of chunk i+1, the
// kernel of chunk i and the download / copy-out of older chunks all overlap and
// both PCIe directions stay busy.  Host ranges that are already page-locked
// (cudaHostAlloc / cudaHostRegister, e.g. aeqb_host_alloc) are DMA'd in place.
// Row chunks are independent for per-channel and blockwise granularity, so
// chunking does not change any result.
//
// Staging is the part a single thread cannot do at link speed (one core copies
// ~10 GB/s, the link moves ~55): every device owns a small pool of worker threads,
// bound to the CPUs of the GPU's own NUMA node (sysfs local_cpulist) together with
// the pinned buffers they first touch, and a chunk is staged as several parallel
// memcpy pieces.  All CUDA calls stay on the calling thread.
//
// One call can fan out over several GPUs of the node (aeqb_host_set_devices):
// chunks are dealt round-robin to the devices, each with its own ring, streams and
// workers.  The default is the current device only — under one-process-per-GPU
// launchers every rank sees every GPU and must not spill onto its neighbours'.
#include <ctype.h>
#include <immintrin.h>
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/aeqb200.h"
#include "aeqb_kernels.h"

namespace aeqb {
int host_fail(const char* fmt, ...);  // aeqb_api.cu
int host_check(cudaError_t e, const char* what);
int sm_count_cached();
}  // namespace aeqb

namespace {

constexpr int kSlots = 6;                   // chunks in flight per device
constexpr int kLookahead = 2;               // chunks whose staging runs ahead of their upload
constexpr size_t kChunkBytes = 32u << 20;   // fp32 input bytes per slot
constexpr size_t kPieceBytes = 2u << 20;    // one worker task copies this much
constexpr size_t kMaxParams = kChunkBytes / 4 / 32;  // blockwise-32 scales per chunk (largest case)
constexpr size_t kOutStage = kChunkBytes / 4 + kChunkBytes / 8 + kMaxParams * 10 + 4096;

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// ---------------------------------------------------------------- the copy itself
// Both staging directions write memory the CPU will not read again (pinned staging is read by
// the DMA engine, results by the caller much later), so the stores bypass the cache: a cached
// store first READS the destination line (read-for-ownership), i.e. 3 bytes of DRAM traffic per
// byte copied instead of 2 — and DRAM traffic is what bounds sixteen workers.  glibc's memcpy
// only switches to streaming stores far above the 2 MiB pieces used here.
__attribute__((target("avx2"))) void copy_stream_avx2(char* dst, const char* src, size_t n) {
  const size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
  if (head >= n) { memcpy(dst, src, n); return; }
  memcpy(dst, src, head);
  dst += head; src += head; n -= head;
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
  }
  _mm_sfence();
  if (i < n) memcpy(dst + i, src + i, n - i);
}

bool use_stream_copy() {
  static const bool on = env_int("AEQB_HOST_NT", 1) != 0 && __builtin_cpu_supports("avx2");
  return on;
}

inline void copy_bytes(void* dst, const void* src, size_t n) {
  if (n >= 4096 && use_stream_copy())
    copy_stream_avx2(static_cast<char*>(dst), static_cast<const char*>(src), n);
  else
    memcpy(dst, src, n);
}

// Fresh NumPy outputs are untouched anonymous memory: with 4 KiB pages the copy-out spends more
// time in page faults than in copying.  Where transparent huge pages are available on request
// (THP "madvise" or "always") the 2 MiB-aligned interior of a large output is faulted 2 MiB at
// a time instead.  Advisory: failures are ignored.
void advise_huge(void* p, size_t bytes) {
  static const bool on = env_int("AEQB_HOST_THP", 1) != 0;
  if (!on || !p || bytes < (8u << 20)) return;
  const uintptr_t a = (reinterpret_cast<uintptr_t>(p) + (2u << 20) - 1) & ~uintptr_t((2u << 20) - 1);
  const uintptr_t e = (reinterpret_cast<uintptr_t>(p) + bytes) & ~uintptr_t((2u << 20) - 1);
  if (e > a) madvise(reinterpret_cast<void*>(a), e - a, MADV_HUGEPAGE);
}

// ---------------------------------------------------------------- worker pool
// A counter the submitter can wait on: "all copy tasks of this chunk are done".
struct Latch {
  std::mutex mu;
  std::condition_variable cv;
  int pending = 0;
  void add(int n) { std::lock_guard<std::mutex> l(mu); pending += n; }
  void done() {
    std::lock_guard<std::mutex> l(mu);
    if (--pending == 0) cv.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> l(mu);
    cv.wait(l, [&] { return pending == 0; });
  }
};

struct CopyTask { void* dst; const void* src; size_t bytes; Latch* latch; bool populate; };

// Sources are often views of an mmap'd model file: mapping its page-cache pages one fault at a
// time costs more than copying them.  MADV_POPULATE_READ (Linux 5.14+) maps a whole piece in one
// call; on older kernels or odd mappings it fails and the copy faults the pages as before.
#ifndef MADV_POPULATE_READ
#define MADV_POPULATE_READ 22
#endif
inline void populate_source(const void* src, size_t bytes) {
  static const bool on = env_int("AEQB_HOST_POPULATE", 1) != 0;
  if (!on) return;
  const uintptr_t a = reinterpret_cast<uintptr_t>(src) & ~uintptr_t(4095);
  const uintptr_t e = (reinterpret_cast<uintptr_t>(src) + bytes + 4095) & ~uintptr_t(4095);
  madvise(reinterpret_cast<void*>(a), e - a, MADV_POPULATE_READ);
}

class Pool {
 public:
  Pool(int n_threads, const std::vector<int>& cpus) {
    for (int i = 0; i < n_threads; ++i) threads_.emplace_back([this, cpus] { run(cpus); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> l(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  // Splits one copy into kPieceBytes pieces (64-byte aligned cuts) and queues them.
  void copy(void* dst, const void* src, size_t bytes, Latch* latch, bool populate = false) {
    if (bytes == 0) return;
    const size_t pieces = (bytes + kPieceBytes - 1) / kPieceBytes;
    latch->add(static_cast<int>(pieces));
    {
      std::lock_guard<std::mutex> l(mu_);
      for (size_t p = 0; p < pieces; ++p) {
        const size_t off = p * kPieceBytes;
        q_.push_back({static_cast<char*>(dst) + off, static_cast<const char*>(src) + off,
                      std::min(kPieceBytes, bytes - off), latch, populate});
      }
    }
    cv_.notify_all();
  }
  int size() const { return static_cast<int>(threads_.size()); }

 private:
  void run(std::vector<int> cpus) {
    if (!cpus.empty()) {
      cpu_set_t set;
      CPU_ZERO(&set);
      for (int c : cpus) if (c >= 0 && c < CPU_SETSIZE) CPU_SET(c, &set);
      pthread_setaffinity_np(pthread_self(), sizeof(set), &set);  // best effort
    }
    for (;;) {
      CopyTask t;
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return stop_ || !q_.empty(); });
        if (q_.empty()) return;
        t = q_.front();
        q_.pop_front();
      }
      if (t.populate) populate_source(t.src, t.bytes);
      copy_bytes(t.dst, t.src, t.bytes);
      t.latch->done();
    }
  }
  std::vector<std::thread> threads_;
  std::deque<CopyTask> q_;
  std::mutex mu_;
  std::condition_variable cv_;
  bool stop_ = false;
};

// "0-15,32-47" -> cpu ids; empty when the file is missing (no binding then).
std::vector<int> local_cpus(int device) {
  std::vector<int> cpus;
  char bus[32] = "";
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
    cudaGetLastError();
    return cpus;
  }
  for (char* p = bus; *p; ++p) *p = static_cast<char>(tolower(*p));
  char path[128];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
  FILE* f = fopen(path, "r");
  if (!f) return cpus;
  char line[1024] = "";
  if (fgets(line, sizeof(line), f)) {
    for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
      int a = 0, b = 0;
      const int n = sscanf(tok, "%d-%d", &a, &b);
      if (n == 1) b = a;
      if (n >= 1) for (int c = a; c <= b && c < a + 4096; ++c) cpus.push_back(c);
    }
  }
  fclose(f);
  // keep only CPUs this process may run on (cgroup / taskset); none left -> no binding
  cpu_set_t allowed;
  CPU_ZERO(&allowed);
  if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
    std::vector<int> ok;
    for (int c : cpus) if (c < CPU_SETSIZE && CPU_ISSET(c, &allowed)) ok.push_back(c);
    cpus.swap(ok);
  }
  return cpus;
}

// ---------------------------------------------------------------- per-device ring
struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  float* d_in = nullptr;
  int8_t* d_q = nullptr;        // kChunkBytes / 4
  uint8_t* d_packed = nullptr;  // kChunkBytes / 8
  float* d_scale = nullptr;     // per row or per block (<= kMaxParams floats)
  int32_t* d_zp = nullptr;
  uint16_t* d_f16 = nullptr;
  float* h_in = nullptr;           // pinned staging, allocated on first pageable use
  unsigned char* h_out = nullptr;  // q | packed | scale | zp | f16
  struct Pending { void* dst; const void* src; size_t bytes; };
  std::vector<Pending> pending;    // staged outputs of the chunk in flight
  Latch staged_in, copied_out;     // worker tasks of this slot
  bool in_flight = false;          // kernel + copies enqueued, event recorded
};

struct Ring {
  int device = -1;
  Slot slots[kSlots];
  std::unique_ptr<Pool> pool;
  int next = 0;
  bool ready = false;
};

std::mutex g_mu;
std::vector<std::unique_ptr<Ring>> g_rings;  // one per device ever used
std::vector<int> g_devices;                  // fan-out list; empty = current device

int ring_init(Ring& r, int dev) {
  if (int rc = aeqb::host_check(cudaSetDevice(dev), "cudaSetDevice")) return rc;
  const std::vector<int> cpus = local_cpus(dev);
  int hw = static_cast<int>(cpus.empty() ? std::thread::hardware_concurrency() : cpus.size());
  if (hw <= 0) hw = 4;
  const int n_threads = std::max(1, std::min(env_int("AEQB_HOST_THREADS", 16), hw));
  r.pool.reset(new Pool(n_threads, env_int("AEQB_HOST_NO_BIND", 0) ? std::vector<int>() : cpus));
  for (Slot& s : r.slots) {
    if (int rc = aeqb::host_check(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate")) return rc;
    if (int rc = aeqb::host_check(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming), "cudaEventCreate")) return rc;
    if (int rc = aeqb::host_check(cudaMalloc(&s.d_in, kChunkBytes), "cudaMalloc")) return rc;
    if (int rc = aeqb::host_check(cudaMalloc(&s.d_q, kChunkBytes / 4), "cudaMalloc")) return rc;
    if (int rc = aeqb::host_check(cudaMalloc(&s.d_packed, kChunkBytes / 8), "cudaMalloc")) return rc;
    if (int rc = aeqb::host_check(cudaMalloc(&s.d_scale, kMaxParams * 4), "cudaMalloc")) return rc;
    if (int rc = aeqb::host_check(cudaMalloc(&s.d_zp, kMaxParams * 4), "cudaMalloc")) return rc;
    if (int rc = aeqb::host_check(cudaMalloc(&s.d_f16, kMaxParams * 2), "cudaMalloc")) return rc;
  }
  r.device = dev;
  r.ready = true;
  return 0;
}

int ring_for(int dev, Ring** out) {
  for (auto& r : g_rings)
    if (r->device == dev) { *out = r.get(); return 0; }
  std::unique_ptr<Ring> r(new Ring());
  if (int rc = ring_init(*r, dev)) return rc;  // a half-built ring is dropped with its pool
  g_rings.push_back(std::move(r));
  *out = g_rings.back().get();
  return 0;
}

// Pinned staging is allocated while the calling thread sits on the device's local CPUs, so the
// pages land on the GPU's own NUMA node (first touch at pin time).
int ensure_staging(Ring& r, Slot& s) {
  if (s.h_in) return 0;
  cpu_set_t old;
  CPU_ZERO(&old);
  bool moved = false;
  const std::vector<int> cpus = env_int("AEQB_HOST_NO_BIND", 0) ? std::vector<int>() : local_cpus(r.device);
  if (!cpus.empty() && pthread_getaffinity_np(pthread_self(), sizeof(old), &old) == 0) {
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int c : cpus) if (c < CPU_SETSIZE) CPU_SET(c, &set);
    moved = pthread_setaffinity_np(pthread_self(), sizeof(set), &set) == 0;
  }
  int rc = aeqb::host_check(cudaHostAlloc(&s.h_in, kChunkBytes, cudaHostAllocDefault), "cudaHostAlloc");
  if (!rc) rc = aeqb::host_check(cudaHostAlloc(&s.h_out, kOutStage, cudaHostAllocDefault), "cudaHostAlloc");
  if (moved) pthread_setaffinity_np(pthread_self(), sizeof(old), &old);
  return rc;
}

bool is_pinned(const void* p) {
  if (!p) return true;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// The chunk's results are on the host (pinned staging or the user's pinned arrays): hand the
// staged ones to the workers.  Blocks until the GPU side of the chunk is done.
int slot_retire(Ring& r, Slot& s) {
  if (!s.in_flight) return 0;
  s.in_flight = false;
  if (int rc = aeqb::host_check(cudaEventSynchronize(s.done), "cudaEventSynchronize")) return rc;
  for (const Slot::Pending& p : s.pending) r.pool->copy(p.dst, p.src, p.bytes, &s.copied_out);
  s.pending.clear();
  return 0;
}

// D2H of one output: straight into pinned user memory, or via the slot's staging area.
int download(Slot& s, void* user, const void* dev, size_t bytes, bool user_pinned, size_t* stage_off) {
  if (!user || bytes == 0) return 0;
  void* dst = user;
  if (!user_pinned) {
    dst = s.h_out + *stage_off;
    s.pending.push_back({user, dst, bytes});
    *stage_off += (bytes + 255) & ~size_t(255);
  }
  return aeqb::host_check(cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, s.stream), "D2H");
}

struct HostJob {  // one tensor, host pointers
  const float* x;
  long long rows, cols;
  int8_t* q;
  uint8_t* packed;
  float* scale;
  int32_t* zp;
  uint16_t* f16;
};

struct Chunk {
  const HostJob* job;
  long long r0, nr;
  bool in_pinned, q_pinned, p_pinned, s_pinned, z_pinned, f_pinned;
  Ring* ring;
  Slot* slot;
  bool staged;  // stage-in submitted (or not needed)
};

// After a failure nothing may keep pointers into the caller's arrays: wait for the GPU and the
// workers, drop the staged outputs without copying them.
void drain_all() {
  for (auto& r : g_rings) {
    if (!r->ready) continue;
    cudaSetDevice(r->device);
    for (Slot& s : r->slots) {
      s.staged_in.wait();
      s.copied_out.wait();
      if (s.stream) cudaStreamSynchronize(s.stream);
      s.pending.clear();
      s.in_flight = false;
    }
  }
  cudaGetLastError();
}

int enqueue_chunk(Chunk& c, int mode, int bits, int symmetric, int block, float mse_k) {
  const HostJob& hj = *c.job;
  Slot& s = *c.slot;
  const long long ne = c.nr * hj.cols;
  const int sms = aeqb::sm_count_cached();
  const float* src = hj.x + c.r0 * hj.cols;
  if (!c.in_pinned) {
    s.staged_in.wait();
    src = s.h_in;
  }
  s.copied_out.wait();  // the previous chunk's staged outputs have left h_out
  if (int rc = aeqb::host_check(cudaMemcpyAsync(s.d_in, src, static_cast<size_t>(ne) * 4, cudaMemcpyHostToDevice, s.stream), "H2D")) return rc;
  long long n_params;
  if (mode == 0) {
    aeqb::RowsBatch b{};
    b.bits = bits; b.symmetric = symmetric;
    aeqb::RowsJob j{};
    j.x = s.d_in; j.q = hj.q ? s.d_q : nullptr; j.packed = hj.packed ? s.d_packed : nullptr;
    j.scale = s.d_scale; j.zp = s.d_zp;
    j.rows = c.nr; j.cols = static_cast<int>(hj.cols);
    j.mm_stride = j.clip_stride = j.out_stride = 1;
    j.mse_k = mse_k;  // != 0: scale = k * RMS(row) (mse.py:100-108); rows stay independent, so chunking is exact
    const int klass = aeqb::rows_job_class(j, bits);
    cudaError_t e;
    if (klass == 0 && mse_k != 0.0f) {
      return aeqb::host_fail("MSE rows need a multiple of 128 columns and at most 96 KiB per row");
    } else if (klass == 0) {
      e = aeqb::launch_requant_rows_generic(j, bits, symmetric, s.stream);
    } else {
      j.rows_per_tile = aeqb::rows_job_rows_per_tile(j, klass);
      j.cpr_magic = static_cast<unsigned>(((1u << 20) + (j.cols / 128) - 1) / (j.cols / 128));
      j.tile0 = 0;
      j.tile_end = (j.rows + j.rows_per_tile - 1) / j.rows_per_tile;
      b.jobs[0] = j; b.n_jobs = 1; b.n_tiles = j.tile_end;
      e = aeqb::launch_requant_rows_stream(b, klass, sms, s.stream);
    }
    if (int rc = aeqb::host_check(e, "requant_rows")) return rc;
    n_params = c.nr;
  } else {
    aeqb::BlocksJob j{};
    j.x = s.d_in; j.q = hj.q ? s.d_q : nullptr; j.packed = hj.packed ? s.d_packed : nullptr;
    j.scale = hj.scale ? s.d_scale : nullptr; j.scale_f16 = hj.f16 ? s.d_f16 : nullptr;
    j.n = ne;
    cudaError_t e;
    if (aeqb::blocks_job_streamable(j)) {
      aeqb::BlocksBatch b{};
      b.block = block; b.bits = bits;
      j.tile0 = 0; j.tile_end = aeqb::blocks_job_tiles(j.n);
      b.jobs[0] = j; b.n_jobs = 1; b.n_tiles = j.tile_end;
      e = aeqb::launch_requant_blocks_stream(b, j.q != nullptr, j.packed != nullptr, sms, s.stream);
    } else {
      e = aeqb::launch_requant_blocks_generic(j, block, bits, s.stream);
    }
    if (int rc = aeqb::host_check(e, "requant_blocks")) return rc;
    n_params = ne / block;
  }
  size_t off = 0;
  const long long p0 = mode == 0 ? c.r0 : c.r0 * hj.cols / block;
  if (int rc = download(s, hj.q ? hj.q + c.r0 * hj.cols : nullptr, s.d_q, static_cast<size_t>(ne), c.q_pinned, &off)) return rc;
  if (int rc = download(s, hj.packed ? hj.packed + (c.r0 * hj.cols * bits) / 8 : nullptr, s.d_packed, static_cast<size_t>(ne) * bits / 8, c.p_pinned, &off)) return rc;
  if (int rc = download(s, hj.scale ? hj.scale + p0 : nullptr, s.d_scale, static_cast<size_t>(n_params) * 4, c.s_pinned, &off)) return rc;
  if (int rc = download(s, hj.zp ? hj.zp + p0 : nullptr, s.d_zp, static_cast<size_t>(n_params) * 4, c.z_pinned, &off)) return rc;
  if (int rc = download(s, hj.f16 ? hj.f16 + p0 : nullptr, s.d_f16, static_cast<size_t>(n_params) * 2, c.f_pinned, &off)) return rc;
  if (int rc = aeqb::host_check(cudaEventRecord(s.done, s.stream), "cudaEventRecord")) return rc;
  s.in_flight = true;
  return 0;
}

// mode 0: per-channel rows kernel; mode 1: blockwise kernel.
int run_host_locked(const HostJob* jobs, long long n_jobs, int mode, int bits, int symmetric, int block, float mse_k) {
  int home = 0;
  if (int rc = aeqb::host_check(cudaGetDevice(&home), "cudaGetDevice")) return rc;
  // ---- validate everything before anything is enqueued
  for (long long ji = 0; ji < n_jobs; ++ji) {
    const HostJob& hj = jobs[ji];
    if (hj.rows <= 0 || hj.cols <= 0) continue;
    if (static_cast<size_t>(hj.cols) * 4 > kChunkBytes)
      return aeqb::host_fail("rows longer than %zu bytes are not supported by the host pipeline", kChunkBytes);
    if (mode == 0 && hj.packed && hj.cols % (8 / bits) != 0)
      return aeqb::host_fail("packed output of a [%lld, %lld] tensor would straddle rows; pack separately",
                             hj.rows, hj.cols);
    if (mode == 0 && mse_k != 0.0f && (hj.cols % 128 != 0 || hj.cols * 4 > 98304))
      return aeqb::host_fail("MSE rows need a multiple of 128 columns and at most 96 KiB per row, got %lld columns",
                             hj.cols);
  }
  std::vector<int> devs = g_devices;
  if (devs.empty()) devs.push_back(home);
  std::vector<Ring*> rings;
  for (int d : devs) {
    Ring* r = nullptr;
    if (int rc = ring_for(d, &r)) { cudaSetDevice(home); return rc; }
    rings.push_back(r);
  }
  // ---- chunk list, dealt round-robin to the devices
  std::vector<Chunk> chunks;
  for (long long ji = 0; ji < n_jobs; ++ji) {
    const HostJob& hj = jobs[ji];
    if (hj.rows <= 0 || hj.cols <= 0) continue;
    const size_t row_bytes = static_cast<size_t>(hj.cols) * 4;
    const long long chunk_rows = std::min<long long>(
        static_cast<long long>(kMaxParams), std::max<long long>(1, static_cast<long long>(kChunkBytes / row_bytes)));
    Chunk c{};
    c.job = &hj;
    c.in_pinned = is_pinned(hj.x);
    c.q_pinned = is_pinned(hj.q); c.p_pinned = is_pinned(hj.packed);
    c.s_pinned = is_pinned(hj.scale); c.z_pinned = is_pinned(hj.zp); c.f_pinned = is_pinned(hj.f16);
    if (!c.q_pinned) advise_huge(hj.q, static_cast<size_t>(hj.rows * hj.cols));
    if (!c.p_pinned) advise_huge(hj.packed, static_cast<size_t>(hj.rows * hj.cols) * bits / 8);
    for (long long r0 = 0; r0 < hj.rows; r0 += chunk_rows) {
      c.r0 = r0;
      c.nr = std::min(chunk_rows, hj.rows - r0);
      chunks.push_back(c);
    }
  }
  const size_t n = chunks.size();
  for (size_t k = 0; k < n; ++k) chunks[k].ring = rings[k % rings.size()];
  const size_t ahead = static_cast<size_t>(kLookahead) * rings.size();

  // Takes the ring's next slot for chunk k: retires what the slot held, starts the staging.
  auto prepare = [&](size_t k) -> int {
    Chunk& c = chunks[k];
    Ring& r = *c.ring;
    if (int rc = aeqb::host_check(cudaSetDevice(r.device), "cudaSetDevice")) return rc;
    Slot& s = r.slots[r.next];
    r.next = (r.next + 1) % kSlots;
    c.slot = &s;
    if (int rc = slot_retire(r, s)) return rc;
    const bool need_stage = !(c.in_pinned && c.q_pinned && c.p_pinned && c.s_pinned && c.z_pinned && c.f_pinned);
    if (need_stage) { if (int rc = ensure_staging(r, s)) return rc; }
    if (!c.in_pinned)
      r.pool->copy(s.h_in, c.job->x + c.r0 * c.job->cols, static_cast<size_t>(c.nr * c.job->cols) * 4, &s.staged_in, true);
    c.staged = true;
    return 0;
  };

  int rc = 0;
  for (size_t k = 0; k < n && !rc; ++k) {
    for (size_t a = k; a < std::min(n, k + ahead + 1) && !rc; ++a)
      if (!chunks[a].staged) rc = prepare(a);
    if (rc) break;
    rc = aeqb::host_check(cudaSetDevice(chunks[k].ring->device), "cudaSetDevice");
    if (!rc) rc = enqueue_chunk(chunks[k], mode, bits, symmetric, block, mse_k);
  }
  if (!rc) {
    for (Ring* r : rings) {
      if ((rc = aeqb::host_check(cudaSetDevice(r->device), "cudaSetDevice"))) break;
      for (Slot& s : r->slots)
        if ((rc = slot_retire(*r, s))) break;
      if (rc) break;
    }
  }
  if (!rc)
    for (Ring* r : rings)
      for (Slot& s : r->slots) s.copied_out.wait();
  if (rc) drain_all();
  cudaSetDevice(home);
  return rc;
}

int run_host(const HostJob* jobs, long long n_jobs, int mode, int bits, int symmetric, int block, float mse_k = 0.0f) {
  std::lock_guard<std::mutex> lock(g_mu);
  return run_host_locked(jobs, n_jobs, mode, bits, symmetric, block, mse_k);
}

// ---------------------------------------------------------------- plain staged copies
// Pageable host memory <-> device at link speed for the per-tensor algorithms (OCTAV, Hadamard,
// GPTQ, calibration batches): the same ring and workers, 8 MiB pieces.
constexpr size_t kCopyPiece = 8u << 20;

int copy_in_locked(void* dst_dev, const void* src, size_t bytes, cudaStream_t user) {
  int dev = 0;
  if (int rc = aeqb::host_check(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (bytes == 0) return 0;
  if (is_pinned(src))
    return aeqb::host_check(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, user), "H2D");
  Ring* r = nullptr;
  if (int rc = ring_for(dev, &r)) return rc;
  if (int rc = aeqb::host_check(cudaSetDevice(dev), "cudaSetDevice")) return rc;
  const size_t n = (bytes + kCopyPiece - 1) / kCopyPiece;
  std::vector<Slot*> slot_of(n, nullptr);
  int rc = 0;
  auto prepare = [&](size_t k) -> int {
    Slot& s = r->slots[r->next];
    r->next = (r->next + 1) % kSlots;
    slot_of[k] = &s;
    if (int e = slot_retire(*r, s)) return e;
    if (int e = ensure_staging(*r, s)) return e;
    s.copied_out.wait();
    if (s.stream && cudaStreamSynchronize(s.stream) != cudaSuccess) return aeqb::host_fail("stream synchronise failed");
    const size_t off = k * kCopyPiece;
    r->pool->copy(s.h_in, static_cast<const char*>(src) + off, std::min(kCopyPiece, bytes - off), &s.staged_in, true);
    return 0;
  };
  for (size_t k = 0; k < n && !rc; ++k) {
    for (size_t a = k; a < std::min(n, k + kLookahead + 1) && !rc; ++a)
      if (!slot_of[a]) rc = prepare(a);
    if (rc) break;
    Slot& s = *slot_of[k];
    s.staged_in.wait();
    const size_t off = k * kCopyPiece;
    rc = aeqb::host_check(cudaMemcpyAsync(static_cast<char*>(dst_dev) + off, s.h_in, std::min(kCopyPiece, bytes - off),
                                          cudaMemcpyHostToDevice, s.stream), "H2D");
  }
  // the data must be visible to work the caller enqueues on `user` after this returns
  for (Slot& s : r->slots)
    if (s.stream && cudaStreamSynchronize(s.stream) != cudaSuccess && !rc) rc = aeqb::host_fail("stream synchronise failed");
  if (rc) drain_all();
  return rc;
}

int copy_out_locked(void* dst, const void* src_dev, size_t bytes, cudaStream_t user) {
  int dev = 0;
  if (int rc = aeqb::host_check(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (bytes == 0) return 0;
  if (int rc = aeqb::host_check(cudaStreamSynchronize(user), "cudaStreamSynchronize")) return rc;  // producers done
  if (is_pinned(dst)) {
    if (int rc = aeqb::host_check(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, user), "D2H")) return rc;
    return aeqb::host_check(cudaStreamSynchronize(user), "cudaStreamSynchronize");
  }
  Ring* r = nullptr;
  if (int rc = ring_for(dev, &r)) return rc;
  if (int rc = aeqb::host_check(cudaSetDevice(dev), "cudaSetDevice")) return rc;
  advise_huge(dst, bytes);
  const size_t piece = std::min(kCopyPiece, kOutStage & ~size_t(255));
  const size_t n = (bytes + piece - 1) / piece;
  std::vector<Slot*> slot_of(n, nullptr);
  int rc = 0;
  for (size_t k = 0; k < n && !rc; ++k) {
    Slot& s = r->slots[r->next];
    r->next = (r->next + 1) % kSlots;
    slot_of[k] = &s;
    if ((rc = slot_retire(*r, s))) break;
    if ((rc = ensure_staging(*r, s))) break;
    s.copied_out.wait();
    const size_t off = k * piece, nb = std::min(piece, bytes - off);
    s.pending.push_back({static_cast<char*>(dst) + off, s.h_out, nb});
    rc = aeqb::host_check(cudaMemcpyAsync(s.h_out, static_cast<const char*>(src_dev) + off, nb,
                                          cudaMemcpyDeviceToHost, s.stream), "D2H");
    if (!rc) rc = aeqb::host_check(cudaEventRecord(s.done, s.stream), "cudaEventRecord");
    if (!rc) s.in_flight = true;
    // hand the piece issued two steps ago to the workers now, so that its copy runs under the
    // downloads that follow instead of when its slot comes round again
    if (!rc && k >= 2) rc = slot_retire(*r, *slot_of[k - 2]);
  }
  if (!rc)
    for (Slot& s : r->slots)
      if ((rc = slot_retire(*r, s))) break;
  if (!rc)
    for (Slot& s : r->slots) s.copied_out.wait();
  if (rc) drain_all();
  return rc;
}

}  // namespace

extern "C" {

void* aeqb_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void aeqb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int aeqb_host_set_devices(const int* devices, int n) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (n < 0 || (n > 0 && !devices)) return aeqb::host_fail("bad device list");
  int count = 0;
  if (int rc = aeqb::host_check(cudaGetDeviceCount(&count), "cudaGetDeviceCount")) return rc;
  std::vector<int> v;
  for (int i = 0; i < n; ++i) {
    if (devices[i] < 0 || devices[i] >= count)
      return aeqb::host_fail("device %d is not visible (%d devices)", devices[i], count);
    if (std::find(v.begin(), v.end(), devices[i]) == v.end()) v.push_back(devices[i]);
  }
  g_devices.swap(v);
  return 0;
}

int aeqb_host_worker_threads(void) {
  std::lock_guard<std::mutex> lock(g_mu);
  int n = 0;
  for (auto& r : g_rings) n += r->pool ? r->pool->size() : 0;
  return n;
}

int aeqb_host_copy_in(void* dst_device, const void* src_host, size_t bytes, void* stream) {
  if (bytes > 0 && (!dst_device || !src_host)) return aeqb::host_fail("dst / src are NULL");
  std::lock_guard<std::mutex> lock(g_mu);
  return copy_in_locked(dst_device, src_host, bytes, static_cast<cudaStream_t>(stream));
}

int aeqb_host_copy_out(void* dst_host, const void* src_device, size_t bytes, void* stream) {
  if (bytes > 0 && (!dst_host || !src_device)) return aeqb::host_fail("dst / src are NULL");
  std::lock_guard<std::mutex> lock(g_mu);
  return copy_out_locked(dst_host, src_device, bytes, static_cast<cudaStream_t>(stream));
}

int aeqb_host_requant_rows_batch_f32(const aeqb_rows_job* jobs, int64_t n_jobs, int bits,
                                     int symmetric) {
  if (n_jobs < 0 || (n_jobs > 0 && !jobs)) return aeqb::host_fail("bad job list");
  if (bits != 2 && bits != 4 && bits != 8) return aeqb::host_fail("unsupported num_bits %d (2, 4 or 8)", bits);
  std::vector<HostJob> v(static_cast<size_t>(n_jobs));
  for (int64_t i = 0; i < n_jobs; ++i) {
    const aeqb_rows_job& a = jobs[i];
    if (a.clip) return aeqb::host_fail("clipping constants are not supported by the host pipeline");
    if (a.packed && bits == 8) return aeqb::host_fail("packed output needs num_bits 2 or 4");
    if (a.rows * a.cols > 0 && !a.x) return aeqb::host_fail("x is NULL");
    v[static_cast<size_t>(i)] = HostJob{a.x, a.rows, a.cols, a.q, a.packed, a.scale, a.zp, nullptr};
  }
  return run_host(v.data(), n_jobs, 0, bits, symmetric ? 1 : 0, 0);
}

int aeqb_host_requant_mse_rows_batch_f32(const aeqb_rows_job* jobs, int64_t n_jobs, int bits, float k) {
  if (n_jobs < 0 || (n_jobs > 0 && !jobs)) return aeqb::host_fail("bad job list");
  if (bits != 4 && bits != 8) return aeqb::host_fail("unsupported num_bits %d (4 or 8)", bits);
  if (!(k > 0.0f)) return aeqb::host_fail("the MSE multiplier must be positive");
  std::vector<HostJob> v(static_cast<size_t>(n_jobs));
  for (int64_t i = 0; i < n_jobs; ++i) {
    const aeqb_rows_job& a = jobs[i];
    if (a.clip) return aeqb::host_fail("clipping constants are not supported by the host pipeline");
    if (a.packed && bits == 8) return aeqb::host_fail("packed output needs num_bits 4");
    if (a.rows * a.cols > 0 && !a.x) return aeqb::host_fail("x is NULL");
    v[static_cast<size_t>(i)] = HostJob{a.x, a.rows, a.cols, a.q, a.packed, a.scale, a.zp, nullptr};
  }
  return run_host(v.data(), n_jobs, 0, bits, 1, 0, k);
}

int aeqb_host_requant_blocks_batch_f32(const aeqb_blocks_job* jobs, int64_t n_jobs, int block,
                                       int bits) {
  if (n_jobs < 0 || (n_jobs > 0 && !jobs)) return aeqb::host_fail("bad job list");
  if (block != 32 && block != 64 && block != 128 && block != 256) return aeqb::host_fail("unsupported block size %d", block);
  if (bits != 2 && bits != 4 && bits != 8) return aeqb::host_fail("unsupported num_bits %d (2, 4 or 8)", bits);
  std::vector<HostJob> v(static_cast<size_t>(n_jobs));
  for (int64_t i = 0; i < n_jobs; ++i) {
    const aeqb_blocks_job& a = jobs[i];
    if (a.clip) return aeqb::host_fail("clipping constants are not supported by the host pipeline");
    if (a.cols % block)
      return aeqb::host_fail("Quantized dimension %lld is not divisible by block size %d.", (long long)a.cols, block);
    if (a.packed && bits != 4) return aeqb::host_fail("fused packed output needs num_bits 4");
    if (a.rows * a.cols > 0 && !a.x) return aeqb::host_fail("x is NULL");
    v[static_cast<size_t>(i)] = HostJob{a.x, a.rows, a.cols, a.q, a.packed, a.scale, nullptr, a.scale_f16};
  }
  return run_host(v.data(), n_jobs, 1, bits, 1, block);
}

void aeqb_host_release(void) {
  std::lock_guard<std::mutex> lock(g_mu);
  int home = 0;
  cudaGetDevice(&home);
  for (auto& rp : g_rings) {
    Ring& r = *rp;
    cudaSetDevice(r.device);
    for (Slot& s : r.slots) {
      s.staged_in.wait();
      s.copied_out.wait();
      if (s.stream) cudaStreamSynchronize(s.stream);
      cudaFree(s.d_in); cudaFree(s.d_q); cudaFree(s.d_packed); cudaFree(s.d_scale);
      cudaFree(s.d_zp); cudaFree(s.d_f16);
      if (s.h_in) cudaFreeHost(s.h_in);
      if (s.h_out) cudaFreeHost(s.h_out);
      if (s.done) cudaEventDestroy(s.done);
      if (s.stream) cudaStreamDestroy(s.stream);
    }
  }
  g_rings.clear();  // joins the workers
  cudaSetDevice(home);
  cudaGetLastError();
}

}  // extern "C"
