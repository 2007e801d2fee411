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
#include <limits.h>

#include "aeqb_common.cuh"
#include "aeqb_kernels.h"

#include <cstdlib>

namespace aeqb {

namespace {

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// ---------------------------------------------------------------- whole tensors, batched
// ONE launch reduces up to kMaxInlineJobs tensors (a calibration step's activations):
// the tensors form one stream of 32 KiB tiles, a persistent CTA takes every gridDim-th
// tile with eight 128-bit loads in flight per thread (four CTAs per SM; round 2: 89.6 -> 82.3 us for eight
// 64 MiB tensors = 0.915 -> 0.996 of the measured copy peak against 16 KiB tiles, four loads and eight CTAs
// per SM, same box back to back -- the same bytes in flight per SM in larger contiguous pieces; two or
// eight CTAs per SM measure 0.985), keeps (filtered min, filtered max,
// raw min, raw max, NaN seen) in registers while it stays inside one tensor, and flushes a
// block-reduced partial into its own slot ws[cta][job] when it moves on.  No atomics on
// shared addresses (9.5 k same-address atomics cost more than the 64 MiB read); the last
// CTA to finish (one counter) folds the slots and applies the reference's raw fallback.
constexpr int kMmThreads = 256;
constexpr int kMmPer = 8;                       // 128-bit loads in flight per thread
constexpr int kMmTileVec = kMmThreads * kMmPer;  // float4 per tile
constexpr int kMmMaxGrid = 2048;

struct TensorAcc {
  float fmin, fmax;  // filtered (x > lo / x < hi); NaN never passes a comparison
  float rmin, rmax;  // raw, NaN-propagating (np.min / np.max semantics)
};

__device__ __forceinline__ void acc_reset(TensorAcc& a) {
  a.fmin = INFINITY; a.fmax = -INFINITY; a.rmin = INFINITY; a.rmax = -INFINITY;
}

__device__ __forceinline__ void acc1(TensorAcc& a, float v, float lo, float hi) {
  if (v > lo) a.fmin = fminf(a.fmin, v);
  if (v < hi) a.fmax = fmaxf(a.fmax, v);
  a.rmin = min_nan(a.rmin, v);
  a.rmax = max_nan(a.rmax, v);
}

// Sixteen elements at ~2 instructions each: NaN-propagating min / max of the group first; only
// when the group touches the filter bounds (or holds a NaN) does the per-element path run.
__device__ __forceinline__ void acc16(TensorAcc& a, const float4 (&v)[4], float lo, float hi) {
  float mn = min_nan(min_nan(v[0].x, v[0].y), min_nan(v[0].z, v[0].w));
  float mx = max_nan(max_nan(v[0].x, v[0].y), max_nan(v[0].z, v[0].w));
#pragma unroll
  for (int u = 1; u < 4; ++u) {
    mn = min_nan(mn, min_nan(min_nan(v[u].x, v[u].y), min_nan(v[u].z, v[u].w)));
    mx = max_nan(mx, max_nan(max_nan(v[u].x, v[u].y), max_nan(v[u].z, v[u].w)));
  }
  if (mn > lo && mx < hi) {  // false when mn / mx is NaN
    a.fmin = fminf(a.fmin, mn);
    a.fmax = fmaxf(a.fmax, mx);
    a.rmin = min_nan(a.rmin, mn);
    a.rmax = max_nan(a.rmax, mx);
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc1(a, v[u].x, lo, hi); acc1(a, v[u].y, lo, hi); acc1(a, v[u].z, lo, hi); acc1(a, v[u].w, lo, hi);
    }
  }
}

__device__ __forceinline__ void acc_merge(TensorAcc& a, const TensorAcc& b) {
  a.fmin = fminf(a.fmin, b.fmin); a.fmax = fmaxf(a.fmax, b.fmax);
  a.rmin = min_nan(a.rmin, b.rmin); a.rmax = max_nan(a.rmax, b.rmax);
}

__device__ __forceinline__ TensorAcc acc_warp(TensorAcc a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    TensorAcc b;
    b.fmin = __shfl_xor_sync(0xffffffffu, a.fmin, o);
    b.fmax = __shfl_xor_sync(0xffffffffu, a.fmax, o);
    b.rmin = __shfl_xor_sync(0xffffffffu, a.rmin, o);
    b.rmax = __shfl_xor_sync(0xffffffffu, a.rmax, o);
    acc_merge(a, b);
  }
  return a;
}

// ws: [0] finished-CTA counter (zero before the first launch; the kernel re-zeroes it),
//     then slots[grid][n_jobs] of 4 floats.
__global__ void __launch_bounds__(kMmThreads)
    minmax_tensors_kernel(const __grid_constant__ MinmaxBatch b, float lo, float hi, int use_lo,
                          int use_hi, float* __restrict__ ws) {
  __shared__ TensorAcc s_w[kMmThreads / 32];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  float* slots = ws + 8 + static_cast<size_t>(blockIdx.x) * b.n_jobs * 4;
  for (int j = tid; j < b.n_jobs; j += kMmThreads) {
    float* s = slots + j * 4;
    s[0] = INFINITY; s[1] = -INFINITY; s[2] = INFINITY; s[3] = -INFINITY;
  }
  __syncthreads();

  TensorAcc a;
  acc_reset(a);
  int j = 0;
  bool dirty = false;
  auto flush = [&](int job) {  // block-reduce `a` into this CTA's slot of `job`
    const TensorAcc w = acc_warp(a);
    if (lane == 0) s_w[warp] = w;
    __syncthreads();
    if (tid == 0) {
      TensorAcc t = s_w[0];
      for (int k = 1; k < kMmThreads / 32; ++k) acc_merge(t, s_w[k]);
      float* s = slots + job * 4;
      s[0] = t.fmin; s[1] = t.fmax; s[2] = t.rmin; s[3] = t.rmax;
    }
    __syncthreads();
    acc_reset(a);
  };
  for (long long tile = blockIdx.x; tile < b.n_tiles; tile += gridDim.x) {
    while (tile >= b.jobs[j].tile_end) {
      if (dirty) { flush(j); dirty = false; }
      ++j;
    }
    const MinmaxJob& job = b.jobs[j];
    const long long t = tile - job.tile0;
    const float4* xv = reinterpret_cast<const float4*>(job.x + job.head);
    const long long v0 = t * kMmTileVec;
    float4 v[kMmPer];
#pragma unroll
    for (int u = 0; u < kMmPer; ++u) {
      const long long i = v0 + u * kMmThreads + tid;
      v[u] = i < job.nvec ? ldg_stream(xv + i) : make_float4(NAN, NAN, NAN, NAN);
    }
    if (v0 + kMmTileVec <= job.nvec) {
#pragma unroll
      for (int u = 0; u < kMmPer; u += 4) {
        const float4 q[4] = {v[u], v[u + 1], v[u + 2], v[u + 3]};
        acc16(a, q, lo, hi);
      }
    } else {
#pragma unroll
      for (int u = 0; u < kMmPer; ++u) {
        const long long i = v0 + u * kMmThreads + tid;
        if (i < job.nvec) {
          acc1(a, v[u].x, lo, hi); acc1(a, v[u].y, lo, hi); acc1(a, v[u].z, lo, hi); acc1(a, v[u].w, lo, hi);
        }
      }
    }
    if (t == 0) {  // unaligned head and the < 4-element tail ride along with the first tile
      if (tid < job.head) acc1(a, job.x[tid], lo, hi);
      const long long tail0 = job.head + job.nvec * 4;
      if (tail0 + tid < job.n && tid < 4) acc1(a, job.x[tail0 + tid], lo, hi);
    }
    dirty = true;
  }
  if (dirty) flush(j);

  // ---- last CTA folds every slot
  __threadfence();
  if (tid == 0) s_last = (atomicAdd(reinterpret_cast<int*>(ws), 1) == static_cast<int>(gridDim.x) - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int job = warp; job < b.n_jobs; job += kMmThreads / 32) {
    TensorAcc t;
    acc_reset(t);
    for (int c = lane; c < static_cast<int>(gridDim.x); c += 32) {
      const float4 o4 = __ldcg(reinterpret_cast<const float4*>(
          ws + 8 + (static_cast<size_t>(c) * b.n_jobs + job) * 4));
      TensorAcc o;
      o.fmin = o4.x; o.fmax = o4.y; o.rmin = o4.z; o.rmax = o4.w;
      acc_merge(t, o);
    }
    t = acc_warp(t);
    if (lane == 0) {
      // common_quantize.py:1386-1408: the filtered value unless nothing passed the filter, then
      // the raw NaN-propagating np.min / np.max.  An empty tensor keeps +-inf.
      float* out = b.jobs[job].out2;
      out[0] = (use_lo && t.fmin != INFINITY) ? t.fmin : t.rmin;
      out[1] = (use_hi && t.fmax != -INFINITY) ? t.fmax : t.rmax;
    }
  }
  if (tid == 0) *reinterpret_cast<int*>(ws) = 0;  // ready for the next launch on this stream
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

// Whole tensor (rows == 1, TENSORWISE): per-CTA fp64 partial sums of fp32 squares, folded in fixed
// order by one thread (deterministic), then the same fp32 divide / sqrt / multiply.
constexpr int kMseMaxParts = 1024;
__global__ void __launch_bounds__(256)
    mse_tensor_partials(const float* __restrict__ x, long long n, double* __restrict__ part) {
  __shared__ double s_part[8];
  double acc = 0.0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = x[i];
    acc += static_cast<double>(__fmul_rn(v, v));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    part[blockIdx.x] = t;
  }
}

__global__ void mse_tensor_final(const double* __restrict__ part, int nparts, long long n, float k,
                                 float* __restrict__ scale) {
  double t = 0.0;
  for (int i = 0; i < nparts; ++i) t += part[i];
  const float mean = __fdiv_rn(static_cast<float>(t), static_cast<float>(n));
  scale[0] = __fmul_rn(k, __fsqrt_rn(mean));
}

// ---------------------------------------------------------------- histogram
// The bin-count step of histogram_utils._DynamicHistogram1D.add (utils/histogram_utils.py:139-164):
//   idx = clip(int32(floor((x - lower_bound) / bin_width)), 0, nbins - 1);  counts[idx] += 1
// in fp32 like NumPy does for an fp32 array with fp32 / weak scalars.  Out-of-int32-range
// quotients and NaN cast to INT_MIN on x86 (cvttss2si) and therefore clip to bin 0; mirrored.
// finite_only drops NaN / +-inf first (DynamicHistogram.add filters with np.isfinite, :401-416).
// Shared-memory privatised counters (one copy per CTA), merged with 64-bit global atomics.
__global__ void __launch_bounds__(256)
    hist_kernel(const float* __restrict__ x, long long n, float lb, float bw, int nbins,
                int finite_only, unsigned long long* __restrict__ counts) {
  extern __shared__ unsigned s_bins[];
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) s_bins[i] = 0;
  __syncthreads();
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  auto add1 = [&](float v) {
    if (finite_only && !(fabsf(v) < INFINITY)) return;
    const float t = floorf(__fdiv_rn(__fsub_rn(v, lb), bw));
    int idx = (t >= -2147483648.0f && t < 2147483648.0f) ? static_cast<int>(t) : INT_MIN;
    idx = min(max(idx, 0), nbins - 1);
    atomicAdd(&s_bins[idx], 1u);
  };
  const uintptr_t addr = reinterpret_cast<uintptr_t>(x);
  long long head = ((16 - (addr & 15)) & 15) / 4;
  if (head > n) head = n;
  const long long nvec = (n - head) / 4;
  const float4* xv = reinterpret_cast<const float4*>(x + head);
  for (long long i = tid; i < head; i += nthreads) add1(x[i]);
  for (long long i = tid; i < nvec; i += nthreads) {
    const float4 v = ldg_stream(xv + i);
    add1(v.x); add1(v.y); add1(v.z); add1(v.w);
  }
  for (long long j = head + nvec * 4 + tid; j < n; j += nthreads) add1(x[j]);
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += blockDim.x)
    if (s_bins[i]) atomicAdd(&counts[i], static_cast<unsigned long long>(s_bins[i]));
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

size_t minmax_workspace_bytes() {
  return 32 + static_cast<size_t>(kMmMaxGrid) * kMaxInlineJobs * 4 * sizeof(float);
}

// jobs: x / n / out2 filled by the caller; the rest is derived here.
cudaError_t launch_minmax_tensors(MinmaxBatch& b, float lo, float hi, int use_lo, int use_hi,
                                  void* ws, int sm_count, cudaStream_t st) {
  if (b.n_jobs <= 0) return cudaSuccess;
  long long tiles = 0;
  for (int j = 0; j < b.n_jobs; ++j) {
    MinmaxJob& m = b.jobs[j];
    const uintptr_t addr = reinterpret_cast<uintptr_t>(m.x);
    long long head = (addr % 4 == 0) ? static_cast<long long>(((16 - (addr & 15)) & 15) / 4) : m.n;
    if (head > m.n) head = m.n;
    m.head = static_cast<int>(head);
    m.nvec = (m.n - head) / 4;
    if (addr % 4 != 0) return cudaErrorMisalignedAddress;
    long long nt = (m.nvec + kMmTileVec - 1) / kMmTileVec;
    if (nt < 1) nt = 1;  // head / tail / empty tensors still get one tile (writes +-inf when empty)
    m.tile0 = tiles;
    tiles += nt;
    m.tile_end = tiles;
  }
  b.n_tiles = tiles;
  static const int per_sm = getenv("AEQB_MM_GRID") ? atoi(getenv("AEQB_MM_GRID")) : 4;
  long long grid = static_cast<long long>(sm_count) * per_sm;
  if (grid > kMmMaxGrid) grid = kMmMaxGrid;
  if (grid > tiles) grid = tiles;
  minmax_tensors_kernel<<<static_cast<unsigned>(grid), kMmThreads, 0, st>>>(
      b, use_lo ? lo : -INFINITY, use_hi ? hi : INFINITY, use_lo, use_hi, static_cast<float*>(ws));
  return count_launch();
}

cudaError_t launch_row_stats(const float* x, long long rows, int cols, float* mn, float* mx,
                             float* sumsq, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  const int warps = 8;
  const long long grid = (rows + warps - 1) / warps;
  row_stats_kernel<<<static_cast<unsigned>(grid), warps * 32, 0, st>>>(x, rows, cols, mn, mx, sumsq);
  return count_launch();
}

size_t mse_workspace_bytes() { return kMseMaxParts * sizeof(double); }

cudaError_t launch_mse_scale_rows(const float* x, long long rows, long long cols, float k,
                                  float* scale, void* ws, int sm_count, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  if (rows == 1 && cols > 65536 && ws != nullptr) {
    int grid = sm_count * 8;
    if (grid > kMseMaxParts) grid = kMseMaxParts;
    mse_tensor_partials<<<grid, 256, 0, st>>>(x, cols, static_cast<double*>(ws));
    mse_tensor_final<<<1, 1, 0, st>>>(static_cast<const double*>(ws), grid, cols, k, scale);
    return count_launch(2);
  }
  const int threads = cols >= 16384 ? 1024 : 256;
  long long grid = static_cast<long long>(sm_count) * (threads == 1024 ? 2 : 8);
  if (grid > rows) grid = rows;
  mse_scale_rows_kernel<<<static_cast<unsigned>(grid), threads, 0, st>>>(x, rows, cols, k, scale);
  return count_launch();
}

cudaError_t launch_hist(const float* x, long long n, float lb, float bw, int nbins, int finite_only,
                        long long* counts, int sm_count, cudaStream_t st) {
  if (n <= 0 || nbins <= 0) return cudaSuccess;
  if (nbins > 12288) return cudaErrorInvalidValue;  // 48 KiB of shared counters
  long long grid = (n / 4 + 255) / 256;
  if (grid > sm_count * 8LL) grid = sm_count * 8LL;
  if (grid < 1) grid = 1;
  hist_kernel<<<static_cast<unsigned>(grid), 256, nbins * sizeof(unsigned), st>>>(
      x, n, lb, bw, nbins, finite_only, reinterpret_cast<unsigned long long*>(counts));
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

// ---------------------------------------------------------------- EMA over calibration batches
// out2 = fold of qsv_utils.moving_average_update (utils/qsv_utils.py:43-68) over n per-batch
// (min, max) pairs IN BATCH ORDER: the first pair is kept verbatim (calibrator.py:415-416), then
// old = fl32(fl32(s * old) + fl32(c * new)) with s = fl32(smoothing) and c = fl32(1 - smoothing
// evaluated in float64) — NumPy 2 keeps an fp32 array fp32 against the Python-float weak
// scalars.  The recurrence is inherently serial (O(1) per batch), so one thread walks it; the
// multiply and add are kept apart (no FMA contraction) to round exactly like NumPy.
__global__ void ema_sequence_kernel(const float* __restrict__ pairs, long long n, float s, float c,
                                    float* __restrict__ out2) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  if (n <= 0) return;
  float mn = pairs[0], mx = pairs[1];
  for (long long i = 1; i < n; ++i) {
    mn = __fadd_rn(__fmul_rn(s, mn), __fmul_rn(c, pairs[2 * i]));
    mx = __fadd_rn(__fmul_rn(s, mx), __fmul_rn(c, pairs[2 * i + 1]));
  }
  out2[0] = mn;
  out2[1] = mx;
}

cudaError_t launch_ema_sequence(const float* pairs, long long n, double smoothing, float* out2,
                                cudaStream_t st) {
  // the Python floats 0.95 and (1.0 - 0.95) are rounded to fp32 separately, from float64
  const float s = static_cast<float>(smoothing);
  const float c = static_cast<float>(1.0 - smoothing);
  ema_sequence_kernel<<<1, 32, 0, st>>>(pairs, n, s, c, out2);
  return count_launch();
}

}  // namespace aeqb
