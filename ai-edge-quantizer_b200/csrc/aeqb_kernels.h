// Internal launcher interface between the C-ABI (aeqb_api.cu) and the kernels.
// Not installed; the public boundary is include/aeqb200.h.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <atomic>

namespace aeqb {

// Kernel launches issued by this library since load (aeqb_launch_count()).
extern std::atomic<long long> g_launches;
inline cudaError_t count_launch(int n = 1) {
  g_launches.fetch_add(n, std::memory_order_relaxed);
  return cudaGetLastError();
}

// "Configured once" flags for cudaFuncSetAttribute, which is a PER-DEVICE setting: a process that
// drives several GPUs (aeqb_host_set_devices, or a caller switching devices) must repeat it on each.
// A process-wide flag made the first launch on a second device fail with "invalid argument".
struct PerDevice {
  bool flags[64] = {};
  static int current() {
    int d = 0;
    return (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 64) ? d : -1;
  }
  bool done() const {
    const int d = current();
    return d >= 0 && flags[d];
  }
  void set() {
    const int d = current();
    if (d >= 0) flags[d] = true;
  }
};

constexpr int kMaxInlineJobs = 64;

// Per-channel / per-tensor fused requantisation of [rows, cols] fp32 matrices.
// Up to kMaxInlineJobs tensors travel by value in the kernel parameters, so a
// batch needs no device-side table, no workspace and no extra copy.
struct RowsJob {
  const float* x;
  int8_t* q;        // [rows*cols] one value per byte, or null
  uint8_t* packed;  // [rows*cols*bits/8] INT4 / INT2 packed, or null
  float* scale;     // [rows] (out_stride 1) or [1] (out_stride 0), or null
  int32_t* zp;      // same shape as scale, or null
  const float* clip;       // optional clipping constants, index row*clip_stride
  const float* given_min;  // optional precomputed min/max (QSV / per-tensor path)
  const float* given_max;
  const float* given_scale;  // optional final scale (uniform_quantize with caller parameters);
  const int32_t* given_zp;   //   its zero point, or null for zeros.  Index row * mm_stride.
  float mse_k;               // != 0: scale = mse_k * sqrt(mean(row^2)) (mse.py:100-108), zero point 0
  long long rows;
  int cols;
  int rows_per_tile;            // launcher
  unsigned cpr_magic;           // launcher: ceil(2^20 / (cols/128)), row = (chunk*magic)>>20
  int mm_stride, clip_stride;   // 0: one value for the whole tensor, 1: per row
  int out_stride;
  long long tile0, tile_end;    // this job's tile range inside the batch (launcher)
};
// Scales (and zero points) of a sharded job set can be mirrored into the other ranks' copies of
// the gathered buffer from the kernel's own epilogue: the lane that publishes a row's scale also
// stores it at the same offset of up to kMaxPeers peer mappings (NVLink peer memory), which is the
// all-gather of per-channel scales without a collective launch.  delta[i] = peer base - local
// base in bytes; the local `scale` pointers must point into the local copy of that buffer.
constexpr int kMaxPeers = 15;
struct PeerMirror {
  int n;
  long long delta[kMaxPeers];
};
// A run of fp32 values that a rank mirrors into its peers' copies of the gathered buffer.
struct MirrorSpan {
  float* p;
  long long n;
};
// Copies every span into each peer mapping (address + pm.delta[i]) with 16-byte stores where the
// alignment allows: one NVLink packet per four scales instead of one per scale.
cudaError_t launch_mirror_f32(const MirrorSpan* d_spans, int n_spans, long long max_n, const PeerMirror& pm,
                              cudaStream_t st);

struct RowsBatch {
  RowsJob jobs[kMaxInlineJobs];
  int n_jobs;
  int bits;
  int symmetric;
  long long n_tiles;
  PeerMirror peers;
  // A model with more than kMaxInlineJobs tensors travels as ONE launch too: `table` then points
  // at n_jobs entries in device memory (uploaded by the C ABI from a pinned ring, stream-ordered)
  // and `jobs` is unused.  `rich`: what the launcher would have derived from the inline jobs.
  const RowsJob* table;
  const long long* table_ends;  // tile_end of every table entry, packed (16 per 128-byte line): the
                                // producer's walk over hundreds of small tensors reads these, not the jobs
  int rich;
  int one_poller;  // AEQB_ROWS_ONE_POLLER: only consumer warp 0 polls the full barrier, the others park on a named barrier
  int fold;        // per-row partial maxima folded in registers, no shared-memory atomics (launcher sets it)
};
// 0: generic kernel, 1 / 2 / 3: tile-stream kernel with 16 / 32 / 64 KiB stages.
// `batch_bytes`: fp32 bytes of the whole batch the job travels in (0: unknown / single tensor).
int rows_job_class(const RowsJob& j, int bits, long long batch_bytes = 0);
int rows_job_rows_per_tile(const RowsJob& j, int klass);
cudaError_t launch_requant_rows_stream(const RowsBatch& b, int klass, int sm_count, cudaStream_t st);
cudaError_t launch_requant_rows_generic(const RowsJob& j, int bits, int symmetric, cudaStream_t st);

// Blockwise symmetric requantisation: each job is a flat fp32 array cut into
// `block`-long groups (blocks never straddle rows because cols % block == 0).
// Up to kMaxInlineJobs tensors travel by value in the kernel parameters, so a
// batch needs no device-side table, no workspace and no extra copy.
struct BlocksJob {
  const float* x;
  int8_t* q;            // [n] or null
  uint8_t* packed;      // [n/2] or null (bits 4 only)
  float* scale;         // [n/block] fp32 (already bf16->fp16 rounded), or null
  uint16_t* scale_f16;  // [n/block] fp16 bit patterns, or null
  const float* clip;    // optional [n/block]
  long long n;
  long long tile0, tile_end;  // this job's tile range inside the batch (launcher)
};
struct BlocksBatch {
  BlocksJob jobs[kMaxInlineJobs];
  int n_jobs;
  int block;
  int bits;
  long long n_tiles;
  PeerMirror peers;  // n > 0: the fp16 scales are also stored into the peers' gathered buffers
  const BlocksJob* table;  // device job table for more than kMaxInlineJobs tensors (see RowsBatch)
  const long long* table_ends;
};
long long blocks_job_tiles(long long n);
bool blocks_job_streamable(const BlocksJob& j);
cudaError_t launch_requant_blocks_stream(const BlocksBatch& b, bool out_q, bool out_p,
                                         int sm_count, cudaStream_t st);
cudaError_t launch_requant_blocks_generic(const BlocksJob& j, int block, int bits,
                                          cudaStream_t st);

// Statistics (reduce.cu).
struct MinmaxJob {
  const float* x;
  float* out2;                 // [min, max]
  long long n;
  long long nvec;              // launcher: aligned float4 count
  long long tile0, tile_end;   // launcher
  int head;                    // launcher: scalars before 16-byte alignment
};
struct MinmaxBatch {
  MinmaxJob jobs[kMaxInlineJobs];
  int n_jobs;
  long long n_tiles;
};
size_t minmax_workspace_bytes();
cudaError_t launch_minmax_tensors(MinmaxBatch& b, float lo, float hi, int use_lo, int use_hi,
                                  void* ws, int sm_count, cudaStream_t st);
cudaError_t launch_row_stats(const float* x, long long rows, int cols, float* mn, float* mx,
                             float* sumsq, cudaStream_t st);
cudaError_t launch_block_minmax(const float* x, long long n, int block, float* mn, float* mx,
                                cudaStream_t st);

// qsv_utils.moving_average_update folded over n (min, max) pairs in batch order (reduce.cu).
cudaError_t launch_ema_sequence(const float* pairs, long long n, double smoothing, float* out2,
                                cudaStream_t st);
cudaError_t launch_hist(const float* x, long long n, float lb, float bw, int nbins, int finite_only,
                        long long* counts, int sm_count, cudaStream_t st);
size_t mse_workspace_bytes();
cudaError_t launch_mse_scale_rows(const float* x, long long rows, long long cols, float k,
                                  float* scale, void* ws, int sm_count, cudaStream_t st);

// OCTAV clipping search (octav.cu).  ws: octav_workspace_bytes(groups, iters) bytes.
size_t octav_workspace_bytes(long long groups, int iters);
cudaError_t launch_octav_rows(const float* x, long long rows, long long cols, int bits, int iters,
                              float divisor, int early_stop, float* clip, void* ws, int sm_count,
                              cudaStream_t st);
cudaError_t launch_octav_blocks(const float* x, long long n, int block, int bits, int iters,
                                float divisor, int early_stop, float* clip, void* ws, int sm_count,
                                cudaStream_t st);

// Block-diagonal Hadamard rotation of the last axis (hadamard.cu).
cudaError_t launch_hadamard_rows(const float* x, long long rows, long long cols, long long n,
                                 float* out, int sm_count, cudaStream_t st);

// GPTQ (gptq_hessian.cu, gptq_quantize.cu).
size_t xtx_workspace_bytes(long long T, long long K, int sm_count);
// tcgen05 3xTF32 path (xtx_tc.cu); *flag_out: device int, non-zero when the input held
// non-finite values and `out` was left untouched.
bool xtx_tc_eligible(long long T, long long K);
size_t xtx_tc_workspace_bytes(long long T, long long K);
template <typename OutT>
cudaError_t launch_xtx_tc(const float* x, long long T, long long K, double alpha, OutT* out,
                          void* ws, int sm_count, const int** flag_out, cudaStream_t st, int lower_tri = 0);
cudaError_t launch_xtx_f64(const float* x, long long T, long long K, double alpha, double* out,
                           void* ws, int sm_count, cudaStream_t st);
cudaError_t launch_xtx_f32(const float* x, long long T, long long K, double alpha, float* out,
                           void* ws, int sm_count, cudaStream_t st);
size_t hessian_inverse_workspace_bytes(long long K);
// chol_dmma.cu: blocked fp64 Cholesky with DMMA rank-128 updates and one-step lookahead
size_t cholesky_dmma_workspace_bytes();
cudaError_t launch_cholesky_dmma(double* A, int K, double* linv, int* info, cudaStream_t st, int* launches);
cudaError_t launch_hessian_inverse(double* hessian, long long K, double damp, int mutate_diagonal,
                                   float* hinv, void* ws, int* info_out, int sm_count,
                                   cudaStream_t st);
cudaError_t launch_weighted_mean_f64(const double* a, double wa, const double* b, double wb,
                                     double* out, long long n, int sm_count, cudaStream_t st);
size_t gptq_workspace_bytes(long long R, long long K);
cudaError_t launch_gptq_quantize(float* w_work, long long R, long long K, const float* hinv,
                                 const float* scale, const int32_t* zp, int row_stride, int qblock,
                                 int bits, int symmetric, int8_t* q, void* ws, int sm_count,
                                 cudaStream_t st);
// Left-looking tensor-core update (gptq_update_tc.cu).
struct alignas(64) GptqTcMapsOpaque { unsigned char bytes[512]; };  // four CUtensorMap
bool gptq_update_tc_eligible(long long R, long long K);
int gptq_update_tc_max_splits();
int gptq_update_tc_splits(long long R, int kb_total, int sm_count);
cudaError_t launch_split_planes(const float* x, long long n, float* hi, float* lo, int sm_count,
                                cudaStream_t st);
cudaError_t gptq_update_tc_prepare(GptqTcMapsOpaque* out, const float* err_hi, const float* err_lo,
                                   const float* h_hi, const float* h_lo, long long R, long long K);
cudaError_t launch_gptq_update_tc(const GptqTcMapsOpaque* maps, float* part, long long R, int c0, int L,
                                  int n_splits, cudaStream_t st, int light = 0);

// Unfused element-wise pieces (elementwise.cu).
cudaError_t launch_scale_zp(const float* mn, const float* mx, const float* clip, long long n,
                            int bits, int symmetric, int blockwise, float* scale, int32_t* zp,
                            uint16_t* scale_f16, cudaStream_t st);
cudaError_t launch_quantize(const float* x, long long n, long long channels, long long inner,
                            const float* scale, const int32_t* zp, int pstride, int bits,
                            int symmetric, void* q, int sm_count, cudaStream_t st);
cudaError_t launch_dequantize(const void* q, int q_bytes, long long n, long long channels,
                              long long inner, const float* scale, const int32_t* zp, int pstride,
                              int wrap8, float* out, int sm_count, cudaStream_t st);
// oscar.cu
size_t colsq_workspace_bytes(long long n, long long d, int sm_count);
cudaError_t launch_colsq(const float* x, long long n, long long d, double alpha, double* out,
                         void* ws, int sm_count, cudaStream_t st);
size_t oscar_pass_workspace_bytes(long long n, long long d, long long g, int sm_count);
cudaError_t launch_oscar_pass(const float* W, long long n, long long d, long long g, const double* s,
                              double* group_sq, double* a_eff, void* ws, int sm_count,
                              cudaStream_t st);
size_t oscar_clip_workspace_bytes(long long n, long long d, long long g);
cudaError_t launch_oscar_clip(const float* W, long long n, long long d, long long g,
                              const double* s, const double* m, const double* mass_dev,
                              double mass0, int qmax, double* bound, void* ws, int sm_count,
                              cudaStream_t st);
cudaError_t launch_oscar_scale(const double* bound, long long n, int qmax, int blockwise,
                               double* scale, cudaStream_t st);
cudaError_t launch_oscar_quantize(const float* W, long long n, long long d, long long glen,
                                  const double* s, const double* scale, int bits, int8_t* q,
                                  int sm_count, cudaStream_t st);
// recovery.cu
size_t dwr_workspace_bytes(long long n_groups, long long glen);
cudaError_t launch_dwr_scales(const float* x, long long n_groups, long long glen, float* scale,
                              void* ws, int sm_count, cudaStream_t st);
cudaError_t launch_max_abs_diff(const float* a, const float* b, long long n, float* out, void* ws,
                                int sm_count, cudaStream_t st);
cudaError_t launch_cast_f16(const float* x, long long n, void* out, int sm_count, cudaStream_t st);
cudaError_t launch_swap_axes(const void* in, long long a, long long b, long long inner,
                             int elem_bytes, void* out, int sm_count, cudaStream_t st);
cudaError_t launch_pack(const int8_t* q, long long n, int bits, uint8_t* out, int sm_count,
                        cudaStream_t st);

}  // namespace aeqb
