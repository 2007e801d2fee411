// extern "C" surface of libaeqb200.so: argument validation, error strings,
// device-attribute caching.  All arithmetic lives in the kernel translation units.
#include <stdarg.h>
#include <stdio.h>

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

extern "C" {

int aeqb_version(void) { return AEQB_VERSION; }
const char* aeqb_last_error(void) { return g_err; }

int aeqb_requant_rows_f32(const float* x, int64_t rows, int64_t cols, int bits, int symmetric,
                          const float* clip, int8_t* q, uint8_t* packed, float* scale,
                          int32_t* zp, void* stream) {
  if (rows < 0 || cols < 0 || cols > 0x7fffffff) return fail("bad shape [%lld, %lld]", (long long)rows, (long long)cols);
  if (!bits_ok(bits)) return fail("unsupported num_bits %d (2, 4 or 8)", bits);
  if (packed && bits == 8) return fail("packed output needs num_bits 2 or 4");
  if (rows * cols > 0 && !x) return fail("x is NULL");
  aeqb::RowsArgs a{};
  a.x = x; a.q = q; a.packed = packed; a.scale = scale; a.zp = zp; a.clip = clip;
  a.rows = rows; a.cols = static_cast<int>(cols); a.bits = bits; a.symmetric = symmetric ? 1 : 0;
  a.mm_stride = 1; a.clip_stride = 1; a.out_stride = 1;
  return check(aeqb::launch_requant_rows(a, sm_count(), static_cast<cudaStream_t>(stream)),
               "aeqb_requant_rows_f32");
}

int aeqb_requant_given_minmax_f32(const float* x, int64_t rows, int64_t cols, int bits,
                                  int symmetric, const float* mn, const float* mx,
                                  const float* clip, int per_row, int8_t* q, uint8_t* packed,
                                  float* scale, int32_t* zp, void* stream) {
  if (rows < 0 || cols < 0 || cols > 0x7fffffff) return fail("bad shape [%lld, %lld]", (long long)rows, (long long)cols);
  if (!bits_ok(bits)) return fail("unsupported num_bits %d (2, 4 or 8)", bits);
  if (packed && bits == 8) return fail("packed output needs num_bits 2 or 4");
  if (!mn || !mx) return fail("min / max are NULL");
  aeqb::RowsArgs a{};
  a.x = x; a.q = q; a.packed = packed; a.scale = scale; a.zp = zp; a.clip = clip;
  a.given_min = mn; a.given_max = mx;
  a.rows = rows; a.cols = static_cast<int>(cols); a.bits = bits; a.symmetric = symmetric ? 1 : 0;
  a.mm_stride = a.clip_stride = a.out_stride = per_row ? 1 : 0;
  return check(aeqb::launch_requant_rows(a, sm_count(), static_cast<cudaStream_t>(stream)),
               "aeqb_requant_given_minmax_f32");
}

int aeqb_requant_blocks_f32(const float* x, int64_t rows, int64_t cols, int block, int bits,
                            const float* clip, int8_t* q, uint8_t* packed, float* scale,
                            uint16_t* scale_f16, void* stream) {
  if (rows < 0 || cols < 0) return fail("bad shape [%lld, %lld]", (long long)rows, (long long)cols);
  if (block != 32 && block != 64 && block != 128 && block != 256) return fail("unsupported block size %d", block);
  if (cols % block)
    return fail("Quantized dimension %lld is not divisible by block size %d.", (long long)cols, block);
  if (!bits_ok(bits)) return fail("unsupported num_bits %d (2, 4 or 8)", bits);
  if (packed && bits != 4) return fail("fused packed output needs num_bits 4");
  if (rows * cols > 0 && !x) return fail("x is NULL");
  aeqb::BlocksArgs a{};
  a.x = x; a.q = q; a.packed = packed; a.scale = scale; a.scale_f16 = scale_f16; a.clip = clip;
  a.n = rows * cols; a.block = block; a.bits = bits;
  return check(aeqb::launch_requant_blocks(a, sm_count(), static_cast<cudaStream_t>(stream)),
               "aeqb_requant_blocks_f32");
}

size_t aeqb_minmax_workspace_bytes(void) { return 8 * sizeof(int); }

int aeqb_minmax_tensor_f32(const float* x, int64_t n, float lo, float hi, int use_lo, int use_hi,
                           float* out2, void* ws, void* stream) {
  if (n < 0) return fail("negative element count");
  if (!out2 || !ws) return fail("out2 / ws are NULL");
  if (n > 0 && !x) return fail("x is NULL");
  return check(aeqb::launch_minmax_tensor(x, n, lo, hi, use_lo, use_hi, out2,
                                          static_cast<int*>(ws), sm_count(),
                                          static_cast<cudaStream_t>(stream)),
               "aeqb_minmax_tensor_f32");
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

}  // extern "C"
