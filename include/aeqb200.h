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

#ifdef __cplusplus
}
#endif
#endif /* AEQB200_H_ */
