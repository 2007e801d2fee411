/* aeqb200 — C ABI of the B200-native numeric core for AI Edge Quantizer.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / NumPy
 * types.  Every entry point replaces one NumPy expression group of the
 * reference (cited per function as <file>:<lines> under
 * ai_edge_quantizer/ in google-ai-edge/ai-edge-quantizer v0.10.0; `uqt` =
 * algorithms/uniform_quantize/uniform_quantize_tensor.py).  INTEGRATION.md shows
 * the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *  - `*_f32` entry points take DEVICE pointers (caller-owned, any CUDA
 *    allocator) and a `stream` (a cudaStream_t passed as void*; NULL = legacy
 *    default stream).  They enqueue work and return without synchronising.
 *  - `aeqb_host_*` entry points take HOST pointers, stage through pinned
 *    buffers, and return after the results are in host memory.
 *  - Return value: 0 on success, non-zero on error; aeqb_last_error() gives a
 *    thread-local message.  Nothing here falls back to a CPU implementation.
 *  - Weight layout: FC / EMBEDDING weights are [rows = out_features,
 *    cols = in_features], C-contiguous fp32 (utils/tfl_flatbuffer_utils.py:95-106,
 *    :254-263).  Per-channel quantises dim 0; blockwise cuts dim 1 into blocks.
 *  - Optional outputs may be NULL.
 */
#ifndef AEQB200_H_
#define AEQB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AEQB_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define AEQB_API __attribute__((visibility("default")))
#else
#define AEQB_API
#endif

AEQB_API int aeqb_version(void);
AEQB_API const char* aeqb_last_error(void);
/* Number of CUDA kernels this library has launched since it was loaded. */
AEQB_API int64_t aeqb_launch_count(void);

/* ---------------------------------------------------------------- weights, fused
 * Per-channel min/max -> scale/zp -> quantise (-> pack) in one pass.
 * Replaces naive_min_max_quantize.get_tensor_quant_params
 * (algorithms/uniform_quantize/naive_min_max_quantize.py:34-110) =
 * common_quantize.init_tensor_min_max (common_quantize.py:1311-1359) +
 * tensor_zp_scale_from_min_max (uqt:492-586) + uniform_quantize (uqt:273-362)
 * [+ transformation_utils.pack_data, transformations/transformation_utils.py:293-353].
 *   bits: 2, 4 or 8.  symmetric: narrow range iff bits >= 8 (uqt:313-315).
 *   clip: optional [rows] clipping constants (OCTAV), NULL for plain min-max.
 *   q: [rows*cols] int8, one value per byte (UniformQuantParams.quantized_data).
 *   packed: [rows*cols*bits/8] bytes, bits 4 or 2 only.
 *   scale: [rows] fp32.  zp: [rows] int32 (zeros when symmetric). */
AEQB_API int aeqb_requant_rows_f32(const float* x, int64_t rows, int64_t cols, int bits, int symmetric,
                          const float* clip, int8_t* q, uint8_t* packed, float* scale,
                          int32_t* zp, void* stream);

/* Same arithmetic with min/max supplied by the caller instead of reduced from
 * x: the QSV path of get_tensor_quant_params (naive_min_max_quantize.py:52-75)
 * and the TENSORWISE granularity (common_quantize.py:1334-1336).
 *   per_row != 0: mn/mx/clip/scale/zp hold `rows` entries; else one entry. */
AEQB_API int aeqb_requant_given_minmax_f32(const float* x, int64_t rows, int64_t cols, int bits,
                                  int symmetric, const float* mn, const float* mx,
                                  const float* clip, int per_row, int8_t* q, uint8_t* packed,
                                  float* scale, int32_t* zp, void* stream);

/* Blockwise symmetric: |x| max per block -> scale rounded fp32->bf16->fp16
 * (uqt:577-581) -> quantise -> two nibbles per byte.  Replaces the same
 * functions with a BLOCKWISE_{32,64,128,256} granularity (common_quantize.py:1345-1358,
 * uqt:222-270) plus quantize_tensor._perform_blockwise_quantization's fp16 scale
 * tensor (transformations/quantize_tensor.py:107-147).
 *   block: 32/64/128/256, cols % block == 0.  bits: 2, 4 or 8.
 *   clip: optional [rows*cols/block].  packed: bits == 4 only.
 *   scale: [rows*cols/block] fp32 (the rounded value).  scale_f16: same, fp16 bits. */
AEQB_API int aeqb_requant_blocks_f32(const float* x, int64_t rows, int64_t cols, int block, int bits,
                            const float* clip, int8_t* q, uint8_t* packed, float* scale,
                            uint16_t* scale_f16, void* stream);

/* ---------------------------------------------------------------- weights, batched
 * A whole model's weight buffers in one call: every tensor that qualifies for
 * the tile-stream kernels is folded into persistent launches of up to 64 tensors
 * (the job table travels in the kernel parameters: no workspace, no extra copy,
 * no per-tensor launch ramp).  This is the device side of the batched driver
 * that sits behind params_generator.generate_quantization_parameters
 * (params_generator.py:69-185).  `jobs` is a HOST array; the pointers inside are
 * device pointers with the meaning of the single-tensor entry points above. */
typedef struct aeqb_rows_job {
  const float* x;
  int64_t rows, cols;
  const float* clip; /* optional [rows] */
  int8_t* q;         /* optional [rows*cols] */
  uint8_t* packed;   /* optional [rows*cols*bits/8] */
  float* scale;      /* optional [rows] */
  int32_t* zp;       /* optional [rows] */
} aeqb_rows_job;

typedef struct aeqb_blocks_job {
  const float* x;
  int64_t rows, cols;
  const float* clip;   /* optional [rows*cols/block] */
  int8_t* q;           /* optional [rows*cols] */
  uint8_t* packed;     /* optional [rows*cols/2], bits == 4 */
  float* scale;        /* optional [rows*cols/block] */
  uint16_t* scale_f16; /* optional [rows*cols/block] fp16 bits */
} aeqb_blocks_job;

AEQB_API int aeqb_requant_rows_batch_f32(const aeqb_rows_job* jobs, int64_t n_jobs, int bits,
                                         int symmetric, void* stream);
AEQB_API int aeqb_requant_blocks_batch_f32(const aeqb_blocks_job* jobs, int64_t n_jobs, int block,
                                           int bits, void* stream);

/* ---------------------------------------------------------------- weights, batched + sharded
 * The sharded form of the call above (one process per GPU, tensors partitioned across ranks): the
 * reference gathers every tensor's quantisation parameters in one process
 * (params_generator.py:110-183 fills a single dict); here each rank's kernel also stores every
 * row's scale at the same offset of `n_peers` peer mappings of the gathered scale buffer (NVLink
 * peer memory), so the all-gather of per-channel scales needs no collective launch.
 * `peer_delta_bytes[i]` = peer i's base address of that buffer minus the local base address; the
 * jobs' `scale` pointers must point into the local copy.  Remote copies are complete once the
 * writer's stream has been synchronised and the ranks have met (a barrier). */
AEQB_API int aeqb_requant_rows_batch_mirror_f32(const aeqb_rows_job* jobs, int64_t n_jobs, int bits,
                                                int symmetric, const int64_t* peer_delta_bytes,
                                                int n_peers, void* stream);
/* The blockwise form: every block's fp16 scale (2 B per `block` weights — the scale tensor
 * quantize_tensor._perform_blockwise_quantization stores, transformations/quantize_tensor.py:107-147)
 * is also stored into `n_peers` peer mappings of the gathered fp16 scale buffer, 64 B per peer and
 * 32 KiB tile as coalesced 4-byte stores.  Jobs must have packed + scale_f16 outputs only
 * (q == NULL: quantised payloads stay on the owning GPU), `scale_f16` pointing into the local copy. */
AEQB_API int aeqb_requant_blocks_batch_mirror_f32(const aeqb_blocks_job* jobs, int64_t n_jobs,
                                                  int block, int bits,
                                                  const int64_t* peer_delta_bytes, int n_peers,
                                                  void* stream);
/* A device buffer other processes on the node can map (CUDA IPC): `handle64` receives 64 opaque
 * bytes to send to the peers, which call aeqb_peer_open on them. */
AEQB_API int aeqb_peer_alloc(size_t bytes, void** ptr, void* handle64);
AEQB_API int aeqb_peer_open(const void* handle64, void** ptr);
AEQB_API int aeqb_peer_close(void* ptr);
AEQB_API int aeqb_peer_free(void* ptr);

/* ---------------------------------------------------------------- host buffers
 * The call a NumPy caller makes (naive_min_max_quantize.get_tensor_quant_params,
 * :34-110, over every weight of a model): HOST pointers in the job structs, in
 * and out.  Tensors are cut into ~32 MiB row chunks and pipelined
 * copy-in -> H2D -> fused kernel -> D2H -> copy-out over 6 slots / streams per device;
 * page-locked ranges (aeqb_host_alloc, cudaHostRegister) are DMA'd in place,
 * pageable ones (the reference's mmap views) are staged by a pool of worker threads bound to
 * the GPU's NUMA node.  Returns when every output is in host memory.  On any error the pipeline
 * is drained first: nothing keeps a pointer into the caller's arrays.
 * `clip` must be NULL here. */
AEQB_API int aeqb_host_requant_rows_batch_f32(const aeqb_rows_job* jobs, int64_t n_jobs, int bits,
                                              int symmetric);
AEQB_API int aeqb_host_requant_blocks_batch_f32(const aeqb_blocks_job* jobs, int64_t n_jobs,
                                                int block, int bits);
/* MSE per channel (mse.get_tensor_quant_params, algorithms/uniform_quantize/mse.py:36-128) through the same
 * pipeline: scale = k * sqrt(mean(row^2)), zero point 0, symmetric; rows must be a multiple of 128 columns
 * and at most 96 KiB (the fused tile-stream kernel's RMS statistic). */
AEQB_API int aeqb_host_requant_mse_rows_batch_f32(const aeqb_rows_job* jobs, int64_t n_jobs, int bits, float k);
/* Which GPUs of the node the host-buffer calls above fan out over (chunks are dealt round-robin,
 * every device has its own ring of slots, streams and staging workers).  n == 0 restores the
 * default: the calling thread's current device only — under one-process-per-GPU launchers every
 * rank sees every GPU and must stay on its own.  The reference walks one tensor at a time in one
 * process (params_generator.py:110-183); this is that process driving the whole node. */
AEQB_API int aeqb_host_set_devices(const int* devices, int n);
/* Staging worker threads currently alive (0 before the first host-buffer call). */
AEQB_API int aeqb_host_worker_threads(void);
/* Pageable host memory -> device and back through the same pinned ring and worker threads
 * (what `torch.from_numpy(x).to(device)` does with one thread): the upload of a weight for the
 * per-tensor algorithms (octav / hadamard_rotation / gptq get_tensor_quant_params take the same
 * read-only mmap views, utils/tfl_flatbuffer_utils.py:254-263) and the download of their
 * results.  copy_in returns when the bytes are in device memory; copy_out first waits for
 * `stream` (the producer of src_device) and returns when the bytes are in dst_host. */
AEQB_API int aeqb_host_copy_in(void* dst_device, const void* src_host, size_t bytes, void* stream);
AEQB_API int aeqb_host_copy_out(void* dst_host, const void* src_device, size_t bytes, void* stream);
/* Page-locked host memory for callers that want zero-copy staging. */
AEQB_API void* aeqb_host_alloc(size_t bytes);
AEQB_API void aeqb_host_free(void* p);
/* Frees the pipeline's streams, device slots and pinned staging buffers. */
AEQB_API void aeqb_host_release(void);

/* ---------------------------------------------------------------- statistics
 * Whole-tensor min/max with the open-interval validity filter and raw fallback:
 * out2[0] = min{x : x > lo} (raw NaN-propagating min when nothing passes or
 * use_lo == 0), out2[1] = max{x : x < hi} likewise.  Replaces
 * common_quantize.get_activation_min_max (common_quantize.py:1362-1413), the
 * reduction inside naive_min_max_quantize.min_max_calibrate (:181-226) and
 * gptq.calibrate's min/max (gptq.py:84-98).
 *   The batched form reduces all the activation tensors of one calibration step
 *   (calibrator.py:545-582 visits them one by one) in ONE launch per 64 tensors.
 *   ws: aeqb_minmax_workspace_bytes() bytes of device scratch, ZEROED before its
 *   first use (the kernel leaves it ready for the next call); one ws per stream. */
typedef struct aeqb_minmax_job {
  const float* x; /* device, 4-byte aligned */
  int64_t n;
  float* out2;    /* device [2] */
} aeqb_minmax_job;

AEQB_API size_t aeqb_minmax_workspace_bytes(void);
AEQB_API int aeqb_minmax_tensors_f32(const aeqb_minmax_job* jobs, int64_t n_jobs, float lo, float hi,
                                     int use_lo, int use_hi, void* ws, void* stream);
AEQB_API int aeqb_minmax_tensor_f32(const float* x, int64_t n, float lo, float hi, int use_lo,
                                    int use_hi, float* out2, void* ws, void* stream);

/* out2 = qsv_utils.moving_average_update (utils/qsv_utils.py:43-68) folded over n per-batch
 * (min, max) pairs in batch order: the first pair verbatim (calibrator.py:415-416), then
 * smoothing * old + (1 - smoothing) * new in fp32 with NumPy's weak-scalar rounding (the two
 * float64 coefficients are rounded to fp32 separately, hence `double smoothing`).  pairs:
 * DEVICE [n, 2] (e.g. the all-gathered output of aeqb_minmax_tensors_f32), out2: DEVICE [2]. */
AEQB_API int aeqb_ema_sequence_f32(const float* pairs, int64_t n, double smoothing, float* out2,
                                   void* stream);

/* counts[clip(int32(floor((x - lower_bound) / bin_width)), 0, nbins - 1)] += 1 for every element
 * (finite ones only when finite_only != 0): the bin-count step of
 * histogram_utils._DynamicHistogram1D.add (utils/histogram_utils.py:139-164) and the np.isfinite
 * filter of DynamicHistogram.add (:396-416).  counts: DEVICE int64 [nbins], accumulated into
 * (zero it for a fresh batch).  nbins <= 12288. */
AEQB_API int aeqb_hist_accumulate_f32(const float* x, int64_t n, float lower_bound, float bin_width,
                                      int nbins, int finite_only, int64_t* counts, void* stream);

/* Per-row min, max and sum of squares of a [rows, cols] matrix (any may be
 * NULL).  min/max: common_quantize.init_tensor_min_max CHANNELWISE branch
 * (common_quantize.py:1337-1344); sumsq: the reduction of mse.get_tensor_quant_params
 * (algorithms/uniform_quantize/mse.py:100-107). */
AEQB_API int aeqb_row_stats_f32(const float* x, int64_t rows, int64_t cols, float* mn, float* mx,
                                float* sumsq, void* stream);

/* Per-block min and max, shape [rows, cols/block] (common_quantize.py:1345-1358). */
AEQB_API int aeqb_minmax_blocks_f32(const float* x, int64_t rows, int64_t cols, int block,
                                    float* mn, float* mx, void* stream);

/* ---------------------------------------------------------------- OCTAV / MSE / Hadamard
 * OCTAV clipping constants, octav._guess_clipping_with_octav
 * (algorithms/uniform_quantize/octav.py:30-112): Newton iterations of eq. (6) from
 * c = 1 with the reference's GLOBAL np.allclose early stop.  One HBM pass computes
 * every group's whole trajectory; a select step picks the iteration the reference
 * stops at.  rows variant: one constant per row (rows == 1: per tensor);
 * blocks variant: one per `block`-long group, clip shaped [rows, cols/block].
 *   exponent_divisor: 3.0 (signed), max_iterations: 10 in octav.py:186-192.
 *   ws: aeqb_octav_workspace_bytes(number of constants, max_iterations) bytes. */
AEQB_API size_t aeqb_octav_workspace_bytes(int64_t groups, int max_iterations);
AEQB_API int aeqb_octav_clip_rows_f32(const float* x, int64_t rows, int64_t cols, int bits,
                                      int max_iterations, float exponent_divisor, int early_stop,
                                      float* clip, void* ws, void* stream);
AEQB_API int aeqb_octav_clip_blocks_f32(const float* x, int64_t rows, int64_t cols, int block,
                                        int bits, int max_iterations, float exponent_divisor,
                                        int early_stop, float* clip, void* ws, void* stream);

/* scale[r] = k * sqrt(mean(x[r, :]^2)), mse.get_tensor_quant_params
 * (algorithms/uniform_quantize/mse.py:100-108); rows == 1 gives the per-tensor scale.
 *   ws: aeqb_mse_workspace_bytes() bytes, or NULL (a whole-tensor reduction then runs on one CTA). */
AEQB_API size_t aeqb_mse_workspace_bytes(void);
AEQB_API int aeqb_mse_scale_rows_f32(const float* x, int64_t rows, int64_t cols, float k,
                                     float* scale, void* ws, void* stream);

/* out = x.reshape(-1, n) @ (H_n / sqrt(n)) over the last axis,
 * hadamard_rotation._rotate_with_diagonal_hadamard
 * (algorithms/uniform_quantize/hadamard_rotation.py:93-134).  n: power of two
 * dividing cols (the caller applies gcd / max_hadamard_size, :121-123).
 * out may alias x. */
AEQB_API int aeqb_hadamard_rows_f32(const float* x, int64_t rows, int64_t cols, int64_t n,
                                    float* out, void* stream);

/* ---------------------------------------------------------------- GPTQ
 * hessian[k, k] (float64) = alpha * X^T X with X = x viewed as [tokens, k], fp32
 * accumulation promoted by the float64 scalar: gptq.calibrate
 * (algorithms/uniform_quantize/gptq.py:100-106; alpha = 2 / num_samples).
 *   ws: aeqb_xtx_workspace_bytes(tokens, k) bytes (may be 0 -> NULL allowed). */
AEQB_API size_t aeqb_xtx_workspace_bytes(int64_t tokens, int64_t k);
AEQB_API int aeqb_xtx_f32(const float* x, int64_t tokens, int64_t k, double alpha, double* hessian,
                          void* ws, void* stream);

/* hinv[k, k] (float32) = inverse of the damped Hessian, gptq._prepare_hessian_inverse
 * (gptq.py:111-128): zero diagonal entries -> 1, + damp * mean(diag), float64
 * Cholesky, float32 triangular inverse, L^-T L^-1.
 *   keep_damped_diagonal != 0 reproduces the reference's side effect of leaving the
 *   damped diagonal in the caller's `hessian` (np.diag returns a view, gptq.py:114,123).
 *   info: DEVICE int, 0 on success, i > 0 if the leading minor of order i is not
 *   positive definite (np.linalg.cholesky would raise LinAlgError).
 *   ws: aeqb_hessian_inverse_workspace_bytes(k) bytes. */
AEQB_API size_t aeqb_hessian_inverse_workspace_bytes(int64_t k);
AEQB_API int aeqb_hessian_inverse_f64(double* hessian, int64_t k, double damp,
                                      int keep_damped_diagonal, float* hinv, void* ws, int* info,
                                      void* stream);

/* The 64-column lazy-block OBS loop, gptq._apply_gptq (gptq.py:131-216).
 *   w_work: [rows, k] fp32 COPY of the weight, updated in place (gptq.py:139 copies too).
 *   scale / zp: scale_cols entries per row: 1 (per channel), k / block (blockwise,
 *   gptq.py:177-189) or 0 (one entry for the whole tensor).  zp may be NULL (zeros).
 *   blocksize: 64 (the reference's only value).  q: [rows, k] int8.
 *   ws: aeqb_gptq_workspace_bytes(rows, k) bytes.  Large layers (k a multiple of 64, k >= 1024)
 *   take the LEFT-looking schedule: before block b, W[:, b] -= Err[:, :64b] @ Hinv[:64b, b] as one
 *   3xTF32 tcgen05 product per block (same numbers as gptq.py:208-214 regrouped; the error lives
 *   as two TF32 planes in ws), and w_work is then only read.  Other shapes keep the reference's
 *   right-looking order with an fp32 SIMT update of w_work. */
AEQB_API size_t aeqb_gptq_workspace_bytes(int64_t rows, int64_t k);
AEQB_API int aeqb_gptq_quantize_f32(float* w_work, int64_t rows, int64_t k, const float* hinv,
                                    const float* scale, const int32_t* zp, int64_t scale_cols,
                                    int block, int bits, int symmetric, int blocksize, int8_t* q,
                                    void* ws, void* stream);

/* out = (a * wa + b * wb) / (wa + wb), float64: the sample-weighted Hessian mean of
 * qsv_utils._gptq_merge_hessian (utils/qsv_utils.py:71-88).  out may alias a or b. */
AEQB_API int aeqb_hessian_merge_f64(const double* a, double wa, const double* b, double wb,
                                    double* out, int64_t n, void* stream);

/* ---------------------------------------------------------------- unfused pieces
 * tensor_zp_scale_from_min_max (uqt:492-586) on n (min, max[, clip]) triples.
 * blockwise != 0 applies the bf16->fp16 scale rounding (uqt:577-581) and, with
 * clip, the fp16 range limit (uqt:529-550).  zp / scale_f16 may be NULL. */
AEQB_API int aeqb_scale_zp_from_minmax(const float* mn, const float* mx, const float* clip,
                                       int64_t n, int bits, int symmetric, int blockwise,
                                       float* scale, int32_t* zp, uint16_t* scale_f16,
                                       void* stream);

/* uniform_quantize (uqt:273-362) with caller-supplied parameters.  The tensor is
 * viewed as [outer, channels, inner] (n = outer*channels*inner); element i uses
 * parameter ((i / inner) % channels) * param_stride (param_stride 0: one scale
 * for the whole tensor; blockwise: channels = n/block, inner = block).
 * bits <= 8 writes int8, bits <= 16 writes int16.  zp may be NULL (zeros). */
AEQB_API int aeqb_quantize_f32(const float* x, int64_t n, int64_t channels, int64_t inner,
                               const float* scale, const int32_t* zp, int param_stride, int bits,
                               int symmetric, void* q, void* stream);

/* uniform_dequantize (uqt:365-409): out = (q - zp) * scale in fp32.  q_bytes is
 * 1, 2 or 4 (int8/int16/int32).  wrap8 != 0 reproduces NumPy's int8 - int8
 * wrap-around of the difference. */
AEQB_API int aeqb_dequantize_f32(const void* q, int q_bytes, int64_t n, int64_t channels,
                                 int64_t inner, const float* scale, const int32_t* zp,
                                 int param_stride, int wrap8, float* out, void* stream);

/* transformation_utils.pack_data (transformations/transformation_utils.py:293-353):
 * INT4 two nibbles per byte (even index low), INT2 four crumbs per byte (index 0
 * lowest), odd tail zero padded.  out: ceil(n*bits/8) bytes.  bits 2 or 4. */
AEQB_API int aeqb_pack_bits(const int8_t* q, int64_t n, int bits, uint8_t* out, void* stream);

/* out[j, i, t] = x[i, j, t] for x viewed as [a, b, inner] (elem_bytes 4: fp32, 1: int8).  Moves
 * the quantised axis of weights whose channels are not dim 0 (DEPTHWISE_CONV_2D dim 3,
 * BATCH_MATMUL rank-1 / rank-2: utils/tfl_flatbuffer_utils.py:95-106,
 * algorithms/utils/common_utils.py:1210-1218) to the front, so that the per-channel reduction of
 * common_quantize.init_tensor_min_max (common_quantize.py:1337-1344, reduce all dims but the
 * quantised one) is the row kernels' reduction, and moves the integers back.  Not in place. */
AEQB_API int aeqb_swap_axes(const void* x, int64_t a, int64_t b, int64_t inner, int elem_bytes,
                            void* out, void* stream);

/* mse.get_tensor_quant_params fused (algorithms/uniform_quantize/mse.py:36-128, per channel):
 * scale[r] = multiplier * sqrt(mean(x[r]^2)) (fp32 squares, fp64 sum, fp32 mean / sqrt /
 * multiply), zero point 0, q = clip(rint(x / scale)) with the symmetric range, in ONE pass
 * over x (4 B read + 1 B written per weight).  q / packed / zp may be NULL. */
AEQB_API int aeqb_requant_mse_rows_f32(const float* x, int64_t rows, int64_t cols, int bits,
                                       float multiplier, int8_t* q, uint8_t* packed, float* scale,
                                       int32_t* zp, void* stream);

/* ---------------------------------------------------------------- recovery / casting
 * dequantized_weight_recovery.get_zp_scale_from_dequantized_symmetric_weights
 * (algorithms/uniform_quantize/dequantized_weight_recovery.py:132-217): for each of n_groups
 * groups of group_len consecutive floats (a row, a block, or the whole tensor as one group)
 * scale = max(smallest difference > 1e-9 between neighbours of sort(|x| U {0}), 1e-9);
 * 1e-9 when all values coincide.  group_len <= 16384, or n_groups == 1 with
 * ws = aeqb_dwr_workspace_bytes(1, group_len) bytes (radix-sort path). */
AEQB_API size_t aeqb_dwr_workspace_bytes(int64_t n_groups, int64_t group_len);
AEQB_API int aeqb_dwr_scales_f32(const float* x, int64_t n_groups, int64_t group_len, float* scale,
                                 void* ws, void* stream);

/* out[0] = max |a[i] - b[i]| (NaN propagates): the check of
 * dequantized_weight_recovery._validate_recovered_weights (:36-63).  ws: 8 bytes. */
AEQB_API int aeqb_max_abs_diff_f32(const float* a, const float* b, int64_t n, float* out, void* ws,
                                   void* stream);

/* weight.astype(np.float16), round to nearest even, overflow to inf:
 * float_casting.materialize_fc_conv (algorithms/nonlinear_quantize/float_casting.py:160-162). */
AEQB_API int aeqb_cast_f32_f16(const float* x, int64_t n, uint16_t* out, void* stream);

/* ---------------------------------------------------------------- OSCAR (float64 like the reference)
 * out[j] (float64) = alpha * sum_i x[i, j]^2 over a [n, d] fp32 matrix: oscar.calibrate's
 * mu2 = mean(x*x, axis=0) (algorithms/uniform_quantize/oscar.py:271-277, alpha = 1/n) and
 * _compute_channel_scales' a_base (:213, alpha = 1).  ws: aeqb_colsq_workspace_bytes(n, d). */
AEQB_API size_t aeqb_colsq_workspace_bytes(int64_t n, int64_t d);
AEQB_API int aeqb_colsq_f64(const float* x, int64_t n, int64_t d, double alpha, double* out,
                            void* ws, void* stream);

/* One pass of oscar._channel_scale_objective (:172-189) and of the a_eff scatter (:223-230) for
 * channel scales s[d] (float64) and column groups of g (g == d: the whole row):
 *   group_sq[d / g] = sum_i (max_{j in group} |w_ij| s_j)^2
 *   a_eff[d] += w[i, j*]^2 at the first arg-max column j* of every (row, group); the caller
 *   zeroes a_eff; NULL skips it.  ws: aeqb_oscar_pass_workspace_bytes(n, d, g). */
AEQB_API size_t aeqb_oscar_pass_workspace_bytes(int64_t n, int64_t d, int64_t g);
AEQB_API int aeqb_oscar_pass_f32(const float* w, int64_t n, int64_t d, int64_t g, const double* s,
                                 double* group_sq, double* a_eff, void* ws, void* stream);

/* oscar._optimal_group_clip (:62-104) on a = |w * s| with column masses m[d] (float64):
 * g == d: bound[n] per row; g in {32, 64, 128, 256}: bound[n * d / g] per block;
 * g == n * d: bound[1] for the whole tensor (masses tiled).  mass_dev: DEVICE array of the
 * group masses sum(m) + 1e-12 (d / g entries) for the blockwise form; mass0: the same number
 * for the row / tensor forms (host value).  d <= 16384 for the row form.
 * ws: aeqb_oscar_clip_workspace_bytes(n, d, g) (non-zero for the tensor form only). */
AEQB_API size_t aeqb_oscar_clip_workspace_bytes(int64_t n, int64_t d, int64_t g);
AEQB_API int aeqb_oscar_clip_f32(const float* w, int64_t n, int64_t d, int64_t g, const double* s,
                                 const double* m, const double* mass_dev, double mass0, int qmax,
                                 double* bound, void* ws, void* stream);

/* scale[i] (float64) = max(|bound[i]|, 1e-9) / qmax, blockwise != 0: rounded
 * float32 -> bf16 -> fp16 (tensor_zp_scale_from_min_max on float64 bounds, uqt:492-586). */
AEQB_API int aeqb_oscar_scale_f64(const double* bound, int64_t n, int qmax, int blockwise,
                                  double* scale, void* stream);

/* q[i, j] (int8) = clip(rint((w_ij * s_j) / scale[(i * d + j) / group_len])) in float64:
 * uniform_quantize of the scaled weight (oscar.py:455-457; narrow range for 8 bits). */
AEQB_API int aeqb_oscar_quantize_f32(const float* w, int64_t n, int64_t d, int64_t group_len,
                                     const double* s, const double* scale, int bits, int8_t* q,
                                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AEQB200_H_ */
