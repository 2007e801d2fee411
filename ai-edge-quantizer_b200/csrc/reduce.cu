// Calibration / statistics reductions (HBM-bound, read-once):
//   * whole-tensor min & max with the reference's open-interval validity filter
//     and its raw fallback  (common_quantize.get_activation_min_max,
//     algorithms/uniform_quantize/common_quantize.py:1362-1413)
//   * per-row and per-block min & max of a weight
//     (common_quantize.init_tensor_min_max, common_quantize.py:1311-1359)
//   * per-row sum of squares (mse.get_tensor_quant_params, mse.py:36-128)
// All use 128-bit ld.global.nc loads, warp-shuffle reductions with NaN
// propagation (np.min / np.max semantics) and ordered-int atomics for the
// cross-CTA merge, so results are order-free and bit-exact.
#include "aeqb_common.cuh"
#include "aeqb_kernels.h"

namespace aeqb {

namespace {

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// ---------------------------------------------------------------- whole tensor
// ws layout (ints): 0 fmin_ord, 1 fmax_ord, 2 rmin_ord, 3 rmax_ord, 4 nan flag
struct TensorAcc {
  float fmin, fmax;  // filtered (x > lo / x < hi); NaN never passes a comparison
  float rmin, rmax;  // raw, NaN tracked separately
  int nan;
};

__device__ __forceinline__ void acc1(TensorAcc& a, float v, float lo, float hi) {
  if (v > lo) a.fmin = fminf(a.fmin, v);
  if (v < hi) a.fmax = fmaxf(a.fmax, v);
  a.rmin = fminf(a.rmin, v);
  a.rmax = fmaxf(a.rmax, v);
  a.nan |= (v != v);
}

__global__ void minmax_tensor_init(int* ws) {
  ws[0] = f2ord(INFINITY);
  ws[1] = f2ord(-INFINITY);
  ws[2] = f2ord(INFINITY);
  ws[3] = f2ord(-INFINITY);
  ws[4] = 0;
}

__global__ void __launch_bounds__(256)
    minmax_tensor_kernel(const float* __restrict__ x, long long n, float lo, float hi, int* ws) {
  TensorAcc a{INFINITY, -INFINITY, INFINITY, -INFINITY, 0};
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  // scalar head until 16-byte alignment, vector body, scalar tail
  const uintptr_t addr = reinterpret_cast<uintptr_t>(x);
  long long head = ((16 - (addr & 15)) & 15) / 4;
  if (head > n) head = n;
  const long long nvec = (n - head) / 4;
  const float4* xv = reinterpret_cast<const float4*>(x + head);
  for (long long i = tid; i < head; i += nthreads) acc1(a, x[i], lo, hi);
  long long i = tid;
  for (; i + 3 * nthreads < nvec; i += 4 * nthreads) {
    const float4 v0 = ldg_stream(xv + i);
    const float4 v1 = ldg_stream(xv + i + nthreads);
    const float4 v2 = ldg_stream(xv + i + 2 * nthreads);
    const float4 v3 = ldg_stream(xv + i + 3 * nthreads);
    acc1(a, v0.x, lo, hi); acc1(a, v0.y, lo, hi); acc1(a, v0.z, lo, hi); acc1(a, v0.w, lo, hi);
    acc1(a, v1.x, lo, hi); acc1(a, v1.y, lo, hi); acc1(a, v1.z, lo, hi); acc1(a, v1.w, lo, hi);
    acc1(a, v2.x, lo, hi); acc1(a, v2.y, lo, hi); acc1(a, v2.z, lo, hi); acc1(a, v2.w, lo, hi);
    acc1(a, v3.x, lo, hi); acc1(a, v3.y, lo, hi); acc1(a, v3.z, lo, hi); acc1(a, v3.w, lo, hi);
  }
  for (; i < nvec; i += nthreads) {
    const float4 v = ldg_stream(xv + i);
    acc1(a, v.x, lo, hi); acc1(a, v.y, lo, hi); acc1(a, v.z, lo, hi); acc1(a, v.w, lo, hi);
  }
  for (long long j = head + nvec * 4 + tid; j < n; j += nthreads) acc1(a, x[j], lo, hi);

#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a.fmin = fminf(a.fmin, __shfl_xor_sync(0xffffffffu, a.fmin, o));
    a.fmax = fmaxf(a.fmax, __shfl_xor_sync(0xffffffffu, a.fmax, o));
    a.rmin = fminf(a.rmin, __shfl_xor_sync(0xffffffffu, a.rmin, o));
    a.rmax = fmaxf(a.rmax, __shfl_xor_sync(0xffffffffu, a.rmax, o));
    a.nan |= __shfl_xor_sync(0xffffffffu, a.nan, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&ws[0], f2ord(a.fmin));
    atomicMax(&ws[1], f2ord(a.fmax));
    atomicMin(&ws[2], f2ord(a.rmin));
    atomicMax(&ws[3], f2ord(a.rmax));
    if (a.nan) atomicOr(&ws[4], 1);
  }
}

// common_quantize.py:1386-1408: filtered value unless nothing passed the filter
// (then the raw, NaN-propagating np.min / np.max).
__global__ void minmax_tensor_final(const int* ws, int use_lo, int use_hi, long long n,
                                    float* out) {
  const float fmin = ord2f(ws[0]), fmax = ord2f(ws[1]);
  const float rmin = ws[4] ? NAN : ord2f(ws[2]);
  const float rmax = ws[4] ? NAN : ord2f(ws[3]);
  out[0] = (use_lo && fmin != INFINITY) ? fmin : rmin;
  out[1] = (use_hi && fmax != -INFINITY) ? fmax : rmax;
  (void)n;
}

// ---------------------------------------------------------------- per row
// One warp per row; 128-bit loads when the row start is 16-byte aligned.
__global__ void __launch_bounds__(256)
    row_stats_kernel(const float* __restrict__ x, long long rows, int cols, float* mn_out,
                     float* mx_out, float* sumsq_out) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* p = x + row * cols;
  float mn = INFINITY, mx = -INFINITY, ss = 0.0f;
  const bool vec = (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (cols % 4 == 0);
  if (vec) {
    const float4* pv = reinterpret_cast<const float4*>(p);
    for (int i = lane; i < cols / 4; i += 32) {
      const float4 v = ldg_stream(pv + i);
      mn = min_nan(min_nan(mn, v.x), min_nan(v.y, min_nan(v.z, v.w)));
      mx = max_nan(max_nan(mx, v.x), max_nan(v.y, max_nan(v.z, v.w)));
      ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
    }
  } else {
    for (int i = lane; i < cols; i += 32) {
      const float v = p[i];
      mn = min_nan(mn, v);
      mx = max_nan(mx, v);
      ss = fmaf(v, v, ss);
    }
  }
  mn = warp_min_nan(mn);
  mx = warp_max_nan(mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) {
    if (mn_out) mn_out[row] = mn;
    if (mx_out) mx_out[row] = mx;
    if (sumsq_out) sumsq_out[row] = ss;
  }
}

// ---------------------------------------------------------------- MSE scale
// mse.get_tensor_quant_params (mse.py:100-108): scale = k * sqrt(mean(x**2)) per row
// (rows == 1: per tensor).  x**2 is rounded to fp32 like the reference's temporary;
// the sum runs in fp64 and is rounded once (NumPy's pairwise fp32 sum is within a
// few ulp of that), then fp32 divide, sqrt and multiply as NumPy does.
__global__ void __launch_bounds__(1024)
    mse_scale_rows_kernel(const float* __restrict__ x, long long rows, long long cols, float k,
                          float* __restrict__ scale) {
  __shared__ double s_part[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const float* p = x + row * cols;
    double acc = 0.0;
    if ((reinterpret_cast<uintptr_t>(p) % 16 == 0) && (cols % 4 == 0)) {
      const float4* pv = reinterpret_cast<const float4*>(p);
      for (long long i = tid; i < cols / 4; i += blockDim.x) {
        const float4 v = ldg_stream(pv + i);
        acc += static_cast<double>(__fmul_rn(v.x, v.x)) + static_cast<double>(__fmul_rn(v.y, v.y));
        acc += static_cast<double>(__fmul_rn(v.z, v.z)) + static_cast<double>(__fmul_rn(v.w, v.w));
      }
    } else {
      for (long long i = tid; i < cols; i += blockDim.x) {
        const float v = p[i];
        acc += static_cast<double>(__fmul_rn(v, v));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < nw; ++w) t += s_part[w];
      const float mean = __fdiv_rn(static_cast<float>(t), static_cast<float>(cols));
      scale[row] = __fmul_rn(k, __fsqrt_rn(mean));
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- per block
// block/8 lanes share a block; each lane owns 8 consecutive floats.
template <int BLOCK>
__global__ void __launch_bounds__(256)
    block_minmax_kernel(const float* __restrict__ x, long long n, float* mn_out, float* mx_out) {
  constexpr int LPB = BLOCK / 8;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long e = t * 8;
  const bool valid = e < n;
  float mn = INFINITY, mx = -INFINITY;
  if (valid) {
    if (reinterpret_cast<uintptr_t>(x) % 16 == 0) {
      const float4 a = *reinterpret_cast<const float4*>(x + e);
      const float4 b = *reinterpret_cast<const float4*>(x + e + 4);
      mn = min_nan(min_nan(min_nan(a.x, a.y), min_nan(a.z, a.w)),
                   min_nan(min_nan(b.x, b.y), min_nan(b.z, b.w)));
      mx = max_nan(max_nan(max_nan(a.x, a.y), max_nan(a.z, a.w)),
                   max_nan(max_nan(b.x, b.y), max_nan(b.z, b.w)));
    } else {
      for (int j = 0; j < 8; ++j) {
        mn = min_nan(mn, x[e + j]);
        mx = max_nan(mx, x[e + j]);
      }
    }
  }
#pragma unroll
  for (int o = 1; o < LPB; o <<= 1) {
    mn = min_nan(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max_nan(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (valid && (threadIdx.x & (LPB - 1)) == 0) {
    mn_out[e / BLOCK] = mn;
    mx_out[e / BLOCK] = mx;
  }
}

}  // namespace

cudaError_t launch_minmax_tensor(const float* x, long long n, float lo, float hi, int use_lo,
                                 int use_hi, float* out2, int* ws, int sm_count,
                                 cudaStream_t st) {
  minmax_tensor_init<<<1, 1, 0, st>>>(ws);
  if (n > 0) {
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = static_cast<long long>(sm_count) * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    minmax_tensor_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        x, n, use_lo ? lo : -INFINITY, use_hi ? hi : INFINITY, ws);
  }
  minmax_tensor_final<<<1, 1, 0, st>>>(ws, use_lo, use_hi, n, out2);
  return count_launch(n > 0 ? 3 : 2);
}

cudaError_t launch_row_stats(const float* x, long long rows, int cols, float* mn, float* mx,
                             float* sumsq, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  const int warps = 8;
  const long long grid = (rows + warps - 1) / warps;
  row_stats_kernel<<<static_cast<unsigned>(grid), warps * 32, 0, st>>>(x, rows, cols, mn, mx, sumsq);
  return count_launch();
}

cudaError_t launch_mse_scale_rows(const float* x, long long rows, long long cols, float k,
                                  float* scale, int sm_count, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  const int threads = cols >= 16384 ? 1024 : 256;
  long long grid = static_cast<long long>(sm_count) * (threads == 1024 ? 2 : 8);
  if (grid > rows) grid = rows;
  mse_scale_rows_kernel<<<static_cast<unsigned>(grid), threads, 0, st>>>(x, rows, cols, k, scale);
  return count_launch();
}

cudaError_t launch_block_minmax(const float* x, long long n, int block, float* mn, float* mx,
                                cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const long long threads = n / 8;
  const unsigned grid = static_cast<unsigned>((threads + 255) / 256);
  switch (block) {
    case 32: block_minmax_kernel<32><<<grid, 256, 0, st>>>(x, n, mn, mx); break;
    case 64: block_minmax_kernel<64><<<grid, 256, 0, st>>>(x, n, mn, mx); break;
    case 128: block_minmax_kernel<128><<<grid, 256, 0, st>>>(x, n, mn, mx); break;
    case 256: block_minmax_kernel<256><<<grid, 256, 0, st>>>(x, n, mn, mx); break;
    default: return cudaErrorInvalidValue;
  }
  return count_launch();
}

}  // namespace aeqb
