// Unfused building blocks with the reference's exact per-element arithmetic:
//   scale / zero-point from min/max   (uniform_quantize_tensor.tensor_zp_scale_from_min_max, uqt:492-586)
//   quantise with given scale / zp    (uniform_quantize, uqt:273-362)
//   dequantise                        (uniform_dequantize, uqt:365-409)
//   INT4 / INT2 bit packing           (transformation_utils.pack_data, transformations/transformation_utils.py:293-353)
// The fused kernels (requant_rows.cu / requant_blocks.cu) are the hot path; these
// cover arbitrary scale layouts (any quantised axis, GPTQ's per-column use,
// caller-supplied parameters) and are plain grid-stride streaming kernels.
#include "aeqb_common.cuh"
#include "aeqb_kernels.h"

namespace aeqb {

namespace {

// ---------------------------------------------------------------- scale / zp
__global__ void scale_zp_kernel(const float* __restrict__ mn, const float* __restrict__ mx,
                                const float* __restrict__ clip, long long n, int bits,
                                int symmetric, int blockwise, float* scale, int32_t* zp,
                                uint16_t* scale_f16) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const QRange qr = qrange(bits, symmetric != 0);
  float c_hi = 0.f, c_lo = 0.f;
  const bool has_clip = clip != nullptr;
  if (has_clip) {
    c_hi = clip[i];
    c_lo = -c_hi;
    if (blockwise && symmetric) {  // uqt:529-550
      c_hi = min_nan(c_hi, 65280.0f * static_cast<float>((1 << bits) - 1));
      c_lo = max_nan(c_lo, -65280.0f * static_cast<float>(1 << bits));
    }
  }
  float s, z = 0.0f;
  if (symmetric) {
    float bound = max_nan(max_nan(fabsf(mn[i]), fabsf(mx[i])), 1e-9f);
    if (has_clip) bound = min_nan(max_nan(bound, c_lo), c_hi);
    s = __fdiv_rn(bound, qr.qmax);
  } else {
    const float bmax = max_nan(mx[i], 0.0f);
    const float bmin = min_nan(mn[i], 0.0f);
    float bound = max_nan(__fsub_rn(bmax, bmin), 1e-9f);
    if (has_clip) bound = min_nan(max_nan(bound, -clip[i]), clip[i]);  // uqt:571-572
    s = __fdiv_rn(bound, __fsub_rn(qr.qmax, qr.qmin));
    z = rintf(__fsub_rn(qr.qmin, __fdiv_rn(bmin, s)));
  }
  uint16_t h = 0;
  if (blockwise) s = round_scale_bf16_f16(s, &h);
  scale[i] = s;
  if (zp) zp[i] = rni(z);
  if (scale_f16) scale_f16[i] = h;
}

// ---------------------------------------------------------------- quantise / dequantise
// Tensor viewed as [outer, channels, inner]; parameter index = channel * pstride
// (pstride 0 => one scale for everything).  Blockwise: channels = n/block,
// inner = block, outer = 1.  int64 indices throughout.
template <typename OutT>
__global__ void __launch_bounds__(256)
    quantize_kernel(const float* __restrict__ x, long long n, long long channels, long long inner,
                    const float* __restrict__ scale, const int32_t* __restrict__ zp, int pstride,
                    int lo, int hi, OutT* __restrict__ q) {
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += step) {
    const long long ch = (i / inner) % channels;
    const float s = scale[ch * pstride];
    const float z = zp ? static_cast<float>(zp[ch * pstride]) : 0.0f;
    const float t = __fadd_rn(__fdiv_rn(x[i], s), z);
    q[i] = static_cast<OutT>(clampi(rni(t), lo, hi));
  }
}

template <typename InT>
__global__ void __launch_bounds__(256)
    dequantize_kernel(const InT* __restrict__ q, long long n, long long channels, long long inner,
                      const float* __restrict__ scale, const int32_t* __restrict__ zp, int pstride,
                      int wrap8, float* __restrict__ out) {
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += step) {
    const long long ch = (i / inner) % channels;
    const int z = zp ? zp[ch * pstride] : 0;
    // NumPy evaluates `q - zp` in the operands' common integer type: int8 - int8
    // wraps modulo 256 (wrap8); wider types are exact.  One fp32 multiply follows.
    int d = static_cast<int>(q[i]) - z;
    if (wrap8) d = static_cast<int>(static_cast<int8_t>(d));
    out[i] = __fmul_rn(static_cast<float>(d), scale[ch * pstride]);
  }
}

// ---------------------------------------------------------------- bit packing
__global__ void __launch_bounds__(256)
    pack_kernel(const int8_t* __restrict__ q, long long n, int bits, uint8_t* __restrict__ out,
                long long n_out) {
  const int per = 8 / bits;
  const unsigned mask = (1u << bits) - 1u;
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long j = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; j < n_out; j += step) {
    unsigned b = 0;
    for (int k = 0; k < per; ++k) {
      const long long e = j * per + k;
      if (e < n) b |= (static_cast<unsigned>(q[e]) & mask) << (bits * k);  // zero-padded tail
    }
    out[j] = static_cast<uint8_t>(b);
  }
}

// ---------------------------------------------------------------- axis swap
// out[j, i, t] = in[i, j, t] for in viewed as [a, b, inner]: moves the quantised axis of a
// DEPTHWISE_CONV_2D ([1, H, W, C], dim 3) or BATCH_MATMUL weight to the front so the row
// kernels apply, and the integers back (tfl_flatbuffer_utils.py:95-106 lists the dims).
// inner == 1: 32 x 32 shared-memory tile transpose, both sides coalesced; inner > 1: one
// thread per output element (runs of `inner` elements stay contiguous on both sides).
template <typename T>
__global__ void __launch_bounds__(256)
    transpose_tile_kernel(const T* __restrict__ in, long long a, long long b, T* __restrict__ out) {
  __shared__ T tile[32][33];
  const long long tiles_b = (b + 31) / 32;
  const long long tb = blockIdx.x % tiles_b, ta = blockIdx.x / tiles_b;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const long long i = ta * 32 + r, j = tb * 32 + tx;
    if (i < a && j < b) tile[r][tx] = in[i * b + j];
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const long long j = tb * 32 + r, i = ta * 32 + tx;
    if (i < a && j < b) out[j * a + i] = tile[tx][r];
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    swap_axes_kernel(const T* __restrict__ in, long long a, long long b, long long inner,
                     T* __restrict__ out, long long n) {
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long o = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; o < n; o += step) {
    const long long t = o % inner, ji = o / inner;
    const long long i = ji % a, j = ji / a;
    out[o] = in[(i * b + j) * inner + t];
  }
}

unsigned grid_for(long long n, int sm_count) {
  long long g = (n + 255) / 256;
  const long long cap = static_cast<long long>(sm_count) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

}  // namespace

cudaError_t launch_scale_zp(const float* mn, const float* mx, const float* clip, long long n,
                            int bits, int symmetric, int blockwise, float* scale, int32_t* zp,
                            uint16_t* scale_f16, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  scale_zp_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
      mn, mx, clip, n, bits, symmetric, blockwise, scale, zp, scale_f16);
  return count_launch();
}

cudaError_t launch_quantize(const float* x, long long n, long long channels, long long inner,
                            const float* scale, const int32_t* zp, int pstride, int bits,
                            int symmetric, void* q, int sm_count, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const QRange qr = qrange(bits, symmetric != 0);
  const unsigned g = grid_for(n, sm_count);
  if (bits <= 8)
    quantize_kernel<int8_t><<<g, 256, 0, st>>>(x, n, channels, inner, scale, zp, pstride, qr.lo, qr.hi, static_cast<int8_t*>(q));
  else if (bits <= 16)
    quantize_kernel<int16_t><<<g, 256, 0, st>>>(x, n, channels, inner, scale, zp, pstride, qr.lo, qr.hi, static_cast<int16_t*>(q));
  else
    return cudaErrorInvalidValue;
  return count_launch();
}

cudaError_t launch_dequantize(const void* q, int q_bytes, long long n, long long channels,
                              long long inner, const float* scale, const int32_t* zp, int pstride,
                              int wrap8, float* out, int sm_count, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const unsigned g = grid_for(n, sm_count);
  if (q_bytes == 1)
    dequantize_kernel<int8_t><<<g, 256, 0, st>>>(static_cast<const int8_t*>(q), n, channels, inner, scale, zp, pstride, wrap8, out);
  else if (q_bytes == 2)
    dequantize_kernel<int16_t><<<g, 256, 0, st>>>(static_cast<const int16_t*>(q), n, channels, inner, scale, zp, pstride, wrap8, out);
  else if (q_bytes == 4)
    dequantize_kernel<int32_t><<<g, 256, 0, st>>>(static_cast<const int32_t*>(q), n, channels, inner, scale, zp, pstride, wrap8, out);
  else
    return cudaErrorInvalidValue;
  return count_launch();
}

cudaError_t launch_pack(const int8_t* q, long long n, int bits, uint8_t* out, int sm_count,
                        cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const long long n_out = (n * bits + 7) / 8;
  pack_kernel<<<grid_for(n_out, sm_count), 256, 0, st>>>(q, n, bits, out, n_out);
  return count_launch();
}

template <typename T>
static cudaError_t swap_axes_t(const T* in, long long a, long long b, long long inner, T* out,
                               int sm_count, cudaStream_t st) {
  if (inner == 1) {
    const long long tiles = ((a + 31) / 32) * ((b + 31) / 32);
    if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    transpose_tile_kernel<T><<<static_cast<unsigned>(tiles), 256, 0, st>>>(in, a, b, out);
  } else {
    const long long n = a * b * inner;
    swap_axes_kernel<T><<<grid_for(n, sm_count), 256, 0, st>>>(in, a, b, inner, out, n);
  }
  return count_launch();
}

cudaError_t launch_swap_axes(const void* in, long long a, long long b, long long inner,
                             int elem_bytes, void* out, int sm_count, cudaStream_t st) {
  if (a <= 0 || b <= 0 || inner <= 0) return cudaSuccess;
  if (elem_bytes == 4)
    return swap_axes_t(static_cast<const uint32_t*>(in), a, b, inner, static_cast<uint32_t*>(out), sm_count, st);
  if (elem_bytes == 1)
    return swap_axes_t(static_cast<const uint8_t*>(in), a, b, inner, static_cast<uint8_t*>(out), sm_count, st);
  return cudaErrorInvalidValue;
}


// ------------------------------------------------------------------ peer mirror of fp32 scale vectors
namespace {
__global__ void __launch_bounds__(256)
    mirror_f32_kernel(const MirrorSpan* __restrict__ spans, const PeerMirror pm) {
  const MirrorSpan s = spans[blockIdx.y];
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long t0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  bool vec = (reinterpret_cast<uintptr_t>(s.p) & 15) == 0;
  for (int i = 0; i < pm.n; ++i) vec = vec && (pm.delta[i] & 15) == 0;
  long long done = 0;
  if (vec) {
    const long long nv = s.n >> 2;
    const float4* p4 = reinterpret_cast<const float4*>(s.p);
    for (long long k = t0; k < nv; k += stride) {
      const float4 v = __ldcg(p4 + k);
      for (int i = 0; i < pm.n; ++i)
        *reinterpret_cast<float4*>(reinterpret_cast<char*>(const_cast<float4*>(p4 + k)) + pm.delta[i]) = v;
    }
    done = nv << 2;
  }
  for (long long k = done + t0; k < s.n; k += stride) {
    const float v = __ldcg(s.p + k);
    for (int i = 0; i < pm.n; ++i)
      *reinterpret_cast<float*>(reinterpret_cast<char*>(s.p + k) + pm.delta[i]) = v;
  }
}
}  // namespace

cudaError_t launch_mirror_f32(const MirrorSpan* d_spans, int n_spans, long long max_n, const PeerMirror& pm,
                              cudaStream_t st) {
  if (n_spans <= 0 || pm.n <= 0) return cudaSuccess;
  long long gx = (max_n / 4 + 255) / 256;
  if (gx < 1) gx = 1;
  if (gx > 16) gx = 16;
  mirror_f32_kernel<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(n_spans)), 256, 0, st>>>(d_spans, pm);
  return count_launch();
}

}  // namespace aeqb
