// Sort-based scale recovery and the fp16 cast:
//   dequantized_weight_recovery.get_zp_scale_from_dequantized_symmetric_weights
//       (algorithms/uniform_quantize/dequantized_weight_recovery.py:132-217): per group
//       (tensor / row / block) sort |x| with a 0 appended, take the smallest adjacent
//       difference > 1e-9, scale = max(that, 1e-9) (1e-9 when every value is equal)
//   dequantized_weight_recovery._validate_recovered_weights (:36-63): max |dequant - x|
//   float_casting.materialize_fc_conv (algorithms/nonlinear_quantize/float_casting.py:160-162):
//       weight.astype(np.float16)
//
// |x| >= 0, so fp32 order is the order of the raw bit patterns.  Rows and blocks are sorted
// in shared memory with a segmented bitonic network (a row of up to 16384 floats is one
// segment padded with +inf to a power of two; 32..256-wide blocks are many segments of one
// 4096-float tile).  The appended zero only contributes "smallest |x| minus 0", which is taken
// from the segment's first sorted element instead of growing the segment.  A whole tensor as
// one group goes through cub::DeviceRadixSort (library code, TENSORWISE only) and a grid-wide
// adjacent-difference minimum.
#include <cub/device/device_radix_sort.cuh>

#include "aeqb_common.cuh"
#include "aeqb_kernels.h"

namespace aeqb {

namespace {

constexpr float kDwrTol = 1e-9f;       // diffs > tolerance (:197-199)
constexpr float kDwrMinScale = 1e-9f;  // min_scale (:136)
constexpr unsigned kInfBits = 0x7f800000u;

// Ascending bitonic sort of `tile` floats in shared memory as independent segments of `seg`
// (both powers of two, seg <= tile).  All threads of the CTA participate.
__device__ __forceinline__ void bitonic_segments(float* s, int tile, int seg) {
  const int half = tile >> 1;
  for (int k = 2; k <= seg; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int p = threadIdx.x; p < half; p += blockDim.x) {
        const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
        const int q = i | j;
        const bool up = (k == seg) || ((i & k) == 0);
        const float a = s[i], b = s[q];
        if ((a > b) == up) {
          s[i] = b;
          s[q] = a;
        }
      }
      __syncthreads();
    }
  }
}

// groups of `glen` consecutive floats; seg = power of two >= glen; tile = max(seg, 4096) floats
// per CTA iteration holding tile / seg groups (glen == seg whenever tile > seg).
__global__ void __launch_bounds__(1024)
    dwr_group_scale(const float* __restrict__ x, long long n_groups, int glen, int seg, int tile,
                    float* __restrict__ scale) {
  extern __shared__ float dwr_s[];
  int* gmin = reinterpret_cast<int*>(dwr_s + tile);  // per-group min candidate (float bits)
  const int gpt = tile / seg;
  const long long n_tiles = (n_groups + gpt - 1) / gpt;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long g0 = t * gpt;
    for (int i = threadIdx.x; i < tile; i += blockDim.x) {
      const long long g = g0 + i / seg;
      const int e = i % seg;
      float v = __uint_as_float(kInfBits);
      if (g < n_groups && e < glen) v = fabsf(x[g * glen + e]);
      dwr_s[i] = v;
    }
    for (int i = threadIdx.x; i < gpt; i += blockDim.x) gmin[i] = static_cast<int>(kInfBits);
    __syncthreads();
    bitonic_segments(dwr_s, tile, seg);
    for (int i = threadIdx.x; i < tile; i += blockDim.x) {
      const int e = i % seg;
      if (e >= glen) continue;
      const float a = dwr_s[i];
      // e == 0: the difference to the appended zero; otherwise to the previous element
      const float d = (e == 0) ? a : __fsub_rn(a, dwr_s[i - 1]);
      if (d > kDwrTol && d < __uint_as_float(kInfBits)) atomicMin(&gmin[i / seg], __float_as_int(d));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < gpt; i += blockDim.x) {
      const long long g = g0 + i;
      if (g < n_groups) {
        const float m = __int_as_float(gmin[i]);
        scale[g] = (gmin[i] == static_cast<int>(kInfBits)) ? kDwrMinScale : fmaxf(m, kDwrMinScale);
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
    abs_bits_kernel(const float* __restrict__ x, long long n, unsigned* __restrict__ keys,
                    int* __restrict__ min_bits) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *min_bits = static_cast<int>(kInfBits);  // "no candidate"
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += step)
    keys[i] = __float_as_uint(x[i]) & 0x7fffffffu;
}

// Sorted bit patterns -> min adjacent difference > tol (and the first element against 0).
__global__ void __launch_bounds__(256)
    sorted_min_diff(const unsigned* __restrict__ keys, long long n, int* __restrict__ out_bits) {
  float best = __uint_as_float(kInfBits);
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += step) {
    const float a = __uint_as_float(keys[i]);
    const float d = (i == 0) ? a : __fsub_rn(a, __uint_as_float(keys[i - 1]));
    if (d > kDwrTol && d < best) best = d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0 && best < __uint_as_float(kInfBits))
    atomicMin(out_bits, __float_as_int(best));
}

__global__ void dwr_finish_tensor(const int* __restrict__ bits, float* __restrict__ scale) {
  const int b = *bits;
  scale[0] = (b == static_cast<int>(kInfBits)) ? kDwrMinScale : fmaxf(__int_as_float(b), kDwrMinScale);
}

// out[0] = max |a - b| (NaN propagates like np.max), through an ordered-int atomic.
__global__ void __launch_bounds__(256)
    max_abs_diff_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                        int* __restrict__ out_bits, int* __restrict__ nan_flag) {
  float best = 0.0f;
  bool nan = false;
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += step) {
    const float d = fabsf(__fsub_rn(a[i], b[i]));
    nan |= (d != d);
    best = fmaxf(best, d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_int(best));
  if (nan) atomicExch(nan_flag, 1);
}
__global__ void max_abs_diff_finish(const int* __restrict__ bits, const int* __restrict__ nan_flag,
                                    float* __restrict__ out) {
  out[0] = *nan_flag ? __int_as_float(0x7fc00000) : __int_as_float(*bits);
}

__global__ void __launch_bounds__(256)
    cast_f16_kernel(const float* __restrict__ x, long long n, __half* __restrict__ out) {
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long n4 = n >> 2;
  const bool vec = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 8 == 0);
  if (vec) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += step) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
      const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
      uint2 o;
      o.x = *reinterpret_cast<const unsigned*>(&lo);
      o.y = *reinterpret_cast<const unsigned*>(&hi);
      reinterpret_cast<uint2*>(out)[i] = o;
    }
  }
  const long long tail0 = vec ? n4 * 4 : 0;
  for (long long i = tail0 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += step)
    out[i] = __float2half_rn(x[i]);
}

int pow2_at_least(long long v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

struct TensorWs {
  size_t keys_in, keys_out, bits, temp, temp_bytes, total;
};
TensorWs tensor_ws(long long n) {
  TensorWs w;
  size_t temp = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, temp, static_cast<const unsigned*>(nullptr),
                                 static_cast<unsigned*>(nullptr), n, 0, 31);
  const size_t kb = (static_cast<size_t>(n) * 4 + 255) / 256 * 256;
  w.keys_in = 0;
  w.keys_out = kb;
  w.bits = 2 * kb;
  w.temp = 2 * kb + 256;
  w.temp_bytes = temp;
  w.total = w.temp + temp;
  return w;
}

}  // namespace

constexpr int kDwrMaxSeg = 16384;

size_t dwr_workspace_bytes(long long n_groups, long long glen) {
  return n_groups == 1 && glen > kDwrMaxSeg ? tensor_ws(glen).total : 0;
}

// scale[g] for n_groups groups of glen consecutive floats.  Rows longer than 16384 floats are
// only supported as a single group (TENSORWISE) through the radix-sort path (ws required).
cudaError_t launch_dwr_scales(const float* x, long long n_groups, long long glen, float* scale,
                              void* ws, int sm_count, cudaStream_t st) {
  if (n_groups <= 0) return cudaSuccess;
  if (glen <= 0) return cudaErrorInvalidValue;
  if (glen > kDwrMaxSeg) {
    if (n_groups != 1 || ws == nullptr) return cudaErrorInvalidValue;
    const TensorWs w = tensor_ws(glen);
    unsigned char* p = static_cast<unsigned char*>(ws);
    unsigned* kin = reinterpret_cast<unsigned*>(p + w.keys_in);
    unsigned* kout = reinterpret_cast<unsigned*>(p + w.keys_out);
    int* bits = reinterpret_cast<int*>(p + w.bits);
    size_t temp = w.temp_bytes;
    const unsigned grid = static_cast<unsigned>(sm_count * 8);
    abs_bits_kernel<<<grid, 256, 0, st>>>(x, glen, kin, bits);
    cudaError_t e = cub::DeviceRadixSort::SortKeys(p + w.temp, temp, kin, kout, glen, 0, 31, st);
    if (e != cudaSuccess) return e;
    sorted_min_diff<<<grid, 256, 0, st>>>(kout, glen, bits);
    dwr_finish_tensor<<<1, 1, 0, st>>>(bits, scale);
    return count_launch(4);
  }
  const int seg = pow2_at_least(glen);
  int tile = seg > 4096 ? seg : 4096;
  if (seg < tile && glen != seg) tile = seg;  // a padded short row cannot share a tile
  const int gpt = tile / seg;
  const size_t smem = static_cast<size_t>(tile) * 4 + static_cast<size_t>(gpt) * 4;
  static PerDevice attr_done;
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(dwr_group_scale, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kDwrMaxSeg * 4 + 4096);
    if (e != cudaSuccess) return e;
    attr_done.set();
  }
  const long long n_tiles = (n_groups + gpt - 1) / gpt;
  long long grid = n_tiles;
  if (grid > static_cast<long long>(sm_count) * 2) grid = static_cast<long long>(sm_count) * 2;
  const int threads = tile >= 2048 ? 1024 : (tile >= 512 ? 256 : 64);
  dwr_group_scale<<<static_cast<unsigned>(grid), threads, smem, st>>>(x, n_groups, static_cast<int>(glen),
                                                                      seg, tile, scale);
  return count_launch();
}

// out[0] = max |a - b| over n floats; ws: 8 bytes.
cudaError_t launch_max_abs_diff(const float* a, const float* b, long long n, float* out, void* ws,
                                int sm_count, cudaStream_t st) {
  int* bits = static_cast<int*>(ws);
  cudaError_t e = cudaMemsetAsync(bits, 0, 2 * sizeof(int), st);
  if (e != cudaSuccess) return e;
  if (n > 0) {
    long long g = (n + 255) / 256;
    if (g > static_cast<long long>(sm_count) * 8) g = static_cast<long long>(sm_count) * 8;
    max_abs_diff_kernel<<<static_cast<unsigned>(g), 256, 0, st>>>(a, b, n, bits, bits + 1);
  }
  max_abs_diff_finish<<<1, 1, 0, st>>>(bits, bits + 1, out);
  return count_launch(2);
}

cudaError_t launch_cast_f16(const float* x, long long n, void* out, int sm_count, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  long long g = (n / 4 + 255) / 256;
  if (g > static_cast<long long>(sm_count) * 16) g = static_cast<long long>(sm_count) * 16;
  if (g < 1) g = 1;
  cast_f16_kernel<<<static_cast<unsigned>(g), 256, 0, st>>>(x, n, static_cast<__half*>(out));
  return count_launch();
}

}  // namespace aeqb
