// Host-buffer entry points: the drop-in call a NumPy caller makes.
//
// Inputs are host arrays (the reference hands out read-only NumPy views onto the
// mmap'd flatbuffer, utils/tfl_flatbuffer_utils.py:254-263; alignment is not
// guaranteed), outputs are host arrays owned by the caller.  A batch of tensors
// is cut into row chunks of ~32 MiB of fp32 and pushed through a 4-slot ring:
//     [CPU copy -> pinned]  ->  H2D  ->  fused kernel  ->  D2H  ->  [CPU copy <- pinned]
// on one stream per slot, so the upload of chunk i+1, the kernel of chunk i and
// the download of chunk i-1 overlap and both PCIe directions stay busy.  Host
// ranges that are already page-locked (cudaHostAlloc / cudaHostRegister, e.g.
// aeqb_host_alloc) are DMA'd in place; pageable ranges are staged through pinned
// slots.  Row chunks are independent for per-channel and blockwise granularity,
// so chunking does not change any result.
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "../../include/aeqb200.h"
#include "aeqb_kernels.h"

namespace aeqb {
int host_fail(const char* fmt, ...);  // aeqb_api.cu
int host_check(cudaError_t e, const char* what);
int sm_count_cached();
}  // namespace aeqb

namespace {

constexpr int kSlots = 4;
constexpr size_t kChunkBytes = 32u << 20;  // fp32 input bytes per slot

struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  // device
  float* d_in = nullptr;
  int8_t* d_q = nullptr;       // kChunkBytes / 4
  uint8_t* d_packed = nullptr; // kChunkBytes / 8
  float* d_scale = nullptr;    // per row or per block (<= kChunkBytes / 128 floats)
  int32_t* d_zp = nullptr;
  uint16_t* d_f16 = nullptr;
  // pinned staging (allocated on first pageable use)
  float* h_in = nullptr;
  unsigned char* h_out = nullptr;  // q | packed | scale | zp | f16
  // deferred copy-out of a staged chunk
  struct Pending { void* dst; const void* src; size_t bytes; };
  std::vector<Pending> pending;
  bool busy = false;
};

struct Ring {
  int device = -1;
  Slot slots[kSlots];
  bool ready = false;
};

std::mutex g_mu;
Ring g_ring;

constexpr size_t kMaxParams = kChunkBytes / 4 / 32;  // blockwise-32 scales per chunk (largest case)

int ring_init(Ring& r) {
  int dev = 0;
  if (int rc = aeqb::host_check(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (r.ready && r.device == dev) return 0;
  if (r.ready) return aeqb::host_fail("host pipeline is bound to device %d, current device is %d", r.device, dev);
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

bool is_pinned(const void* p) {
  if (!p) return true;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

int ensure_staging(Slot& s) {
  if (!s.h_in) {
    if (int rc = aeqb::host_check(cudaHostAlloc(&s.h_in, kChunkBytes, cudaHostAllocDefault), "cudaHostAlloc")) return rc;
    const size_t out = kChunkBytes / 4 + kChunkBytes / 8 + kMaxParams * 10 + 4096;
    if (int rc = aeqb::host_check(cudaHostAlloc(&s.h_out, out, cudaHostAllocDefault), "cudaHostAlloc")) return rc;
  }
  return 0;
}

// Waits for the slot's previous chunk and lands its staged outputs in user memory.
int slot_retire(Slot& s) {
  if (!s.busy) return 0;
  if (int rc = aeqb::host_check(cudaEventSynchronize(s.done), "cudaEventSynchronize")) return rc;
  for (const Slot::Pending& p : s.pending) memcpy(p.dst, p.src, p.bytes);
  s.pending.clear();
  s.busy = false;
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

// mode 0: per-channel rows kernel; mode 1: blockwise kernel.
int run_host(const HostJob* jobs, long long n_jobs, int mode, int bits, int symmetric, int block) {
  std::lock_guard<std::mutex> lock(g_mu);
  Ring& r = g_ring;
  if (int rc = ring_init(r)) return rc;
  const int sms = aeqb::sm_count_cached();
  int slot_i = 0;
  for (long long ji = 0; ji < n_jobs; ++ji) {
    const HostJob& hj = jobs[ji];
    if (hj.rows <= 0 || hj.cols <= 0) continue;
    const size_t row_bytes = static_cast<size_t>(hj.cols) * 4;
    if (row_bytes > kChunkBytes) return aeqb::host_fail("rows longer than %zu bytes are not supported by the host pipeline", kChunkBytes);
    const long long chunk_rows = std::min<long long>(
        static_cast<long long>(kMaxParams), std::max<long long>(1, static_cast<long long>(kChunkBytes / row_bytes)));
    const bool in_pinned = is_pinned(hj.x);
    const bool q_pinned = is_pinned(hj.q), p_pinned = is_pinned(hj.packed);
    const bool s_pinned = is_pinned(hj.scale), z_pinned = is_pinned(hj.zp), f_pinned = is_pinned(hj.f16);
    const bool need_stage = !(in_pinned && q_pinned && p_pinned && s_pinned && z_pinned && f_pinned);
    for (long long r0 = 0; r0 < hj.rows; r0 += chunk_rows) {
      const long long nr = std::min(chunk_rows, hj.rows - r0);
      const long long ne = nr * hj.cols;
      Slot& s = r.slots[slot_i];
      slot_i = (slot_i + 1) % kSlots;
      if (int rc = slot_retire(s)) return rc;
      if (need_stage) { if (int rc = ensure_staging(s)) return rc; }
      // ---- upload
      const float* src = hj.x + r0 * hj.cols;
      if (!in_pinned) {
        memcpy(s.h_in, src, static_cast<size_t>(ne) * 4);
        src = s.h_in;
      }
      if (int rc = aeqb::host_check(cudaMemcpyAsync(s.d_in, src, static_cast<size_t>(ne) * 4, cudaMemcpyHostToDevice, s.stream), "H2D")) return rc;
      // ---- kernel
      long long n_params;
      if (mode == 0) {
        aeqb::RowsBatch b{};
        b.bits = bits; b.symmetric = symmetric;
        aeqb::RowsJob j{};
        j.x = s.d_in; j.q = hj.q ? s.d_q : nullptr; j.packed = hj.packed ? s.d_packed : nullptr;
        j.scale = s.d_scale; j.zp = s.d_zp;
        j.rows = nr; j.cols = static_cast<int>(hj.cols);
        j.mm_stride = j.clip_stride = j.out_stride = 1;
        const int klass = aeqb::rows_job_class(j, bits);
        cudaError_t e;
        if (klass == 0) {
          if (j.packed && (j.cols % (8 / bits) != 0)) return aeqb::host_fail("packed output would straddle rows; pack separately");
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
        n_params = nr;
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
      // ---- download
      size_t off = 0;
      const long long p0 = mode == 0 ? r0 : r0 * hj.cols / block;
      if (int rc = download(s, hj.q ? hj.q + r0 * hj.cols : nullptr, s.d_q, static_cast<size_t>(ne), q_pinned, &off)) return rc;
      if (int rc = download(s, hj.packed ? hj.packed + (r0 * hj.cols * bits) / 8 : nullptr, s.d_packed, static_cast<size_t>(ne) * bits / 8, p_pinned, &off)) return rc;
      if (int rc = download(s, hj.scale ? hj.scale + p0 : nullptr, s.d_scale, static_cast<size_t>(n_params) * 4, s_pinned, &off)) return rc;
      if (int rc = download(s, hj.zp ? hj.zp + p0 : nullptr, s.d_zp, static_cast<size_t>(n_params) * 4, z_pinned, &off)) return rc;
      if (int rc = download(s, hj.f16 ? hj.f16 + p0 : nullptr, s.d_f16, static_cast<size_t>(n_params) * 2, f_pinned, &off)) return rc;
      if (int rc = aeqb::host_check(cudaEventRecord(s.done, s.stream), "cudaEventRecord")) return rc;
      s.busy = true;
    }
  }
  for (Slot& s : r.slots)
    if (int rc = slot_retire(s)) return rc;
  return 0;
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
  Ring& r = g_ring;
  if (!r.ready) return;
  for (Slot& s : r.slots) {
    cudaStreamSynchronize(s.stream);
    cudaFree(s.d_in); cudaFree(s.d_q); cudaFree(s.d_packed); cudaFree(s.d_scale);
    cudaFree(s.d_zp); cudaFree(s.d_f16);
    if (s.h_in) cudaFreeHost(s.h_in);
    if (s.h_out) cudaFreeHost(s.h_out);
    cudaEventDestroy(s.done);
    cudaStreamDestroy(s.stream);
    s = Slot();
  }
  r.ready = false;
  r.device = -1;
}

}  // extern "C"
