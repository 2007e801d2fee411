// OCTAV clipping-constant search (arXiv 2206.06501 eq. 6) — replaces
// octav._guess_clipping_with_octav (algorithms/uniform_quantize/octav.py:30-112).
//
// Per reduction group (a weight row, or one `block`-long group of a row) the
// reference iterates, from c = 1, at most `max_iterations` times
//     c <- ( sum{x : x >= c} - sum{x : x <= -c} )
//          / ( (1 - s) * (count{x >= c} + count{x <= -c}) + s * N ),   s = f32(4^-bits / divisor)
// and stops EARLY only when np.allclose(old, new) holds for every group of the
// tensor at once (octav.py:109).  Groups never interact except through that stop
// time, so this file computes every group's whole trajectory in ONE pass over HBM
// with the group resident in registers (4 B read per weight, no re-reads), records
// the trace [max_iterations, groups] plus a bit mask "some group was not yet
// converged after iteration i", and a tiny second kernel picks the iteration the
// reference would have stopped at.
//
// Arithmetic notes (mirroring NumPy's dtype flow, see oracle/aeq_oracle.py:octav_clip):
//  * for c > 0 the two masks are disjoint and sum_hi - sum_lo = sum{|x| : |x| >= c};
//    for c == 0 every finite element is selected and zeros are counted TWICE
//    (x >= 0 and x <= -0), which changes the denominator — reproduced here;
//  * counts are exact integers; (1 - s) * count is an fp32 product; s * N is
//    float64 (N is an np.int64 scalar) and the sum is rounded to fp32 once;
//  * the masked fp32 sums are order-dependent in NumPy; here they are an fp32
//    tree (per-thread partials -> shuffles -> fixed-order cross-warp sum), within a
//    few ulp of the exact sum.  DESIGN.md states the resulting tolerance.
#include "aeqb_common.cuh"
#include "aeqb_kernels.h"

namespace aeqb {

namespace {

constexpr int kMaxIter = 32;

struct OctavConst {
  float s;         // f32(4^-bits / divisor)
  float one_m_s;   // 1.0f - s   (fp32, octav.py:105)
  double s_n;      // double(s) * N (octav.py:106, float64 because N is np.int64)
};

__device__ __forceinline__ float octav_update(float num, int cnt, const OctavConst& k) {
  const float den0 = __fmul_rn(static_cast<float>(cnt), k.one_m_s);
  const float den = static_cast<float>(static_cast<double>(den0) + k.s_n);
  return __fdiv_rn(num, den);
}

__device__ __forceinline__ float octav_update_ll(float num, long long cnt, const OctavConst& k) {
  const float den0 = __fmul_rn(static_cast<float>(cnt), k.one_m_s);
  const float den = static_cast<float>(static_cast<double>(den0) + k.s_n);
  return __fdiv_rn(num, den);
}

// np.isclose(old, new) with rtol=1e-5, atol=1e-8 evaluated in fp32 (weak Python scalars).
__device__ __forceinline__ bool is_close(float oldv, float newv) {
  const float tol = __fadd_rn(1e-8f, __fmul_rn(1e-5f, fabsf(newv)));
  const bool fin = fabsf(newv) < INFINITY;  // isfinite
  return ((fabsf(__fsub_rn(oldv, newv)) <= tol) && fin) || (oldv == newv);
}

__device__ __forceinline__ void acc_elem(float v, float g, bool gzero, float& sum, int& cnt) {
  const float a = fabsf(v);
  if (a >= g) {  // false for NaN elements and NaN guesses
    sum += a;
    cnt += 1;
  }
  if (gzero && v == 0.0f) cnt += 1;  // x >= 0 and x <= -0 both hold for zeros
}

// ------------------------------------------------------------------ rows
// One CTA per row at a time (grid-stride over rows); thread t keeps float4
// chunks t, t + T, ... of the row in registers (NV of them), so the row is read
// from HBM exactly once and every iteration runs from registers.
template <int NV, int THREADS>
__global__ void __launch_bounds__(THREADS)
    octav_rows_trace(const float* __restrict__ x, long long rows, int cols, OctavConst k,
                     int iters, float* __restrict__ trace, unsigned* __restrict__ notclose) {
  constexpr int NW = THREADS / 32;
  __shared__ float s_sum[2][NW];
  __shared__ int s_cnt[2][NW];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nvec = cols >> 2;
  unsigned my_mask = 0;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const float4* p = reinterpret_cast<const float4*>(x + row * cols);
    float4 v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = j * THREADS + tid;
      v[j] = i < nvec ? __ldg(p + i) : make_float4(NAN, NAN, NAN, NAN);
    }
    float g = 1.0f;
    for (int it = 0; it < iters; ++it) {
      float sum = 0.0f;
      int cnt = 0;
      const bool gz = g == 0.0f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        acc_elem(v[j].x, g, gz, sum, cnt);
        acc_elem(v[j].y, g, gz, sum, cnt);
        acc_elem(v[j].z, g, gz, sum, cnt);
        acc_elem(v[j].w, g, gz, sum, cnt);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      const int b = it & 1;
      if (lane == 0) {
        s_sum[b][warp] = sum;
        s_cnt[b][warp] = cnt;
      }
      __syncthreads();  // double-buffered: one barrier per iteration
      float tsum = 0.0f;
      int tcnt = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        tsum += s_sum[b][w];
        tcnt += s_cnt[b][w];
      }
      const float ng = octav_update(tsum, tcnt, k);
      if (tid == 0) {
        trace[static_cast<long long>(it) * rows + row] = ng;
        if (!is_close(g, ng)) my_mask |= 1u << it;
      }
      g = ng;
    }
    __syncthreads();  // s_* of parity (iters-1)&1 may be rewritten by the next row's first iterations
  }
  if (tid == 0 && my_mask) atomicOr(notclose, my_mask);
}

// Any cols / alignment (and the per-tensor case: rows == 1): the row is re-read
// from global memory (L2) every iteration.
__global__ void __launch_bounds__(1024)
    octav_rows_trace_generic(const float* __restrict__ x, long long rows, long long cols,
                             OctavConst k, int iters, float* __restrict__ trace,
                             unsigned* __restrict__ notclose) {
  __shared__ double s_sum[2][32];
  __shared__ long long s_cnt[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nw = blockDim.x >> 5;
  unsigned my_mask = 0;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const float* p = x + row * cols;
    float g = 1.0f;
    for (int it = 0; it < iters; ++it) {
      double sum = 0.0;  // long per-thread runs: accumulate in fp64, round once
      int cnt = 0;
      const bool gz = g == 0.0f;
      for (long long i = tid; i < cols; i += blockDim.x) {
        const float v = p[i];
        const float a = fabsf(v);
        if (a >= g) {
          sum += static_cast<double>(a);
          cnt += 1;
        }
        if (gz && v == 0.0f) cnt += 1;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      const int b = it & 1;
      if (lane == 0) {
        s_sum[b][warp] = sum;
        s_cnt[b][warp] = cnt;
      }
      __syncthreads();
      double dsum = 0.0;
      long long tcnt = 0;
      for (int w = 0; w < nw; ++w) {
        dsum += s_sum[b][w];
        tcnt += s_cnt[b][w];
      }
      const float ng = octav_update_ll(static_cast<float>(dsum), tcnt, k);
      if (tid == 0) {
        trace[static_cast<long long>(it) * rows + row] = ng;
        if (!is_close(g, ng)) my_mask |= 1u << it;
      }
      g = ng;
    }
    __syncthreads();
  }
  if (tid == 0 && my_mask) atomicOr(notclose, my_mask);
}

// ------------------------------------------------------------------ blocks
// Each lane owns 8 consecutive floats; BLOCK/8 lanes share a group and reduce
// with xor-shuffles, so a group's whole trajectory is computed in registers.
template <int BLOCK>
__global__ void __launch_bounds__(256)
    octav_blocks_trace(const float* __restrict__ x, long long n, OctavConst k, int iters,
                       float* __restrict__ trace, unsigned* __restrict__ notclose) {
  constexpr int LPB = BLOCK / 8;
  __shared__ unsigned s_mask;
  if (threadIdx.x == 0) s_mask = 0;
  __syncthreads();
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long nblk = n / BLOCK;
  const bool aligned = reinterpret_cast<uintptr_t>(x) % 16 == 0;
  const int lane = threadIdx.x & 31;
  unsigned my_mask = 0;
  // Warp-uniform trip count (full-mask shuffles inside); lanes past the end hold NaN,
  // which no comparison selects.  n is a multiple of BLOCK, so a group is valid as a whole.
  for (long long t0 = static_cast<long long>(blockIdx.x) * blockDim.x + (threadIdx.x - lane);
       t0 * 8 < n; t0 += nthreads) {
    const long long e = (t0 + lane) * 8;
    const bool valid = e < n;
    float v[8];
    if (!valid) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = NAN;
    } else if (aligned) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(x + e));
      const float4 b = __ldg(reinterpret_cast<const float4*>(x + e + 4));
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
      v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = x[e + j];
    }
    const long long blk = e / BLOCK;
    const bool writer = valid && (lane & (LPB - 1)) == 0;
    float g = 1.0f;
    for (int it = 0; it < iters; ++it) {
      float sum = 0.0f;
      int cnt = 0;
      const bool gz = g == 0.0f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc_elem(v[j], g, gz, sum, cnt);
#pragma unroll
      for (int o = 1; o < LPB; o <<= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      }
      const float ng = octav_update(sum, cnt, k);
      if (writer) {
        trace[static_cast<long long>(it) * nblk + blk] = ng;
        if (!is_close(g, ng)) my_mask |= 1u << it;
      }
      g = ng;
    }
  }
  my_mask = __reduce_or_sync(0xffffffffu, my_mask);
  if ((threadIdx.x & 31) == 0 && my_mask) atomicOr(&s_mask, my_mask);
  __syncthreads();
  if (threadIdx.x == 0 && s_mask) atomicOr(notclose, s_mask);
}

// clip[g] = trace[stop][g], stop = first iteration after which every group was
// converged (octav.py:109), else the last iteration.
__global__ void __launch_bounds__(256)
    octav_select(const float* __restrict__ trace, const unsigned* __restrict__ notclose,
                 long long groups, int iters, int early_stop, float* __restrict__ clip) {
  int stop = iters - 1;
  if (early_stop) {
    const unsigned m = *notclose;
    for (int i = 0; i < iters; ++i)
      if (!((m >> i) & 1u)) { stop = i; break; }
  }
  const float* src = trace + static_cast<long long>(stop) * groups;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < groups;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    clip[i] = src[i];
}

OctavConst make_const(int bits, float divisor, long long n) {
  OctavConst k;
  // np.asarray(4.0 ** (-bits) / exponent_divisor, dtype=np.float32): float64 math, one rounding.
  double p = 1.0;
  for (int i = 0; i < bits; ++i) p *= 0.25;
  k.s = static_cast<float>(p / static_cast<double>(divisor));
  k.one_m_s = 1.0f - k.s;
  k.s_n = static_cast<double>(k.s) * static_cast<double>(n);
  return k;
}

template <int NV>
void launch_rows_nv(const float* x, long long rows, int cols, const OctavConst& k, int iters,
                    float* trace, unsigned* notclose, int sm_count, cudaStream_t st) {
  constexpr int T = 256;
  long long grid = static_cast<long long>(sm_count) * (NV <= 4 ? 6 : (NV <= 8 ? 4 : 2));
  if (grid > rows) grid = rows;
  octav_rows_trace<NV, T><<<static_cast<unsigned>(grid), T, 0, st>>>(x, rows, cols, k, iters, trace,
                                                                      notclose);
}

}  // namespace

size_t octav_workspace_bytes(long long groups, int iters) {
  return static_cast<size_t>(groups) * static_cast<size_t>(iters) * sizeof(float) + 256;
}

// ws layout: [0, 256) the not-close mask (4 bytes used), then the trace.
cudaError_t launch_octav_rows(const float* x, long long rows, long long cols, int bits, int iters,
                              float divisor, int early_stop, float* clip, void* ws, int sm_count,
                              cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  if (iters < 1 || iters > kMaxIter) return cudaErrorInvalidValue;
  unsigned* notclose = static_cast<unsigned*>(ws);
  float* trace = reinterpret_cast<float*>(static_cast<unsigned char*>(ws) + 256);
  cudaError_t e = cudaMemsetAsync(notclose, 0, 256, st);
  if (e != cudaSuccess) return e;
  const OctavConst k = make_const(bits, divisor, cols);
  const bool vec = cols > 0 && cols % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
                   cols <= 16384;
  if (vec) {
    const int c = static_cast<int>(cols);
    const int nv = (c / 4 + 255) / 256;  // float4 per thread
    if (nv <= 1) launch_rows_nv<1>(x, rows, c, k, iters, trace, notclose, sm_count, st);
    else if (nv <= 2) launch_rows_nv<2>(x, rows, c, k, iters, trace, notclose, sm_count, st);
    else if (nv <= 4) launch_rows_nv<4>(x, rows, c, k, iters, trace, notclose, sm_count, st);
    else if (nv <= 8) launch_rows_nv<8>(x, rows, c, k, iters, trace, notclose, sm_count, st);
    else launch_rows_nv<16>(x, rows, c, k, iters, trace, notclose, sm_count, st);
  } else {
    long long grid = static_cast<long long>(sm_count) * 2;
    if (grid > rows) grid = rows;
    const int threads = cols >= 4096 ? 1024 : 256;
    octav_rows_trace_generic<<<static_cast<unsigned>(grid), threads, 0, st>>>(
        x, rows, cols, k, iters, trace, notclose);
  }
  long long sgrid = (rows + 255) / 256;
  if (sgrid > sm_count * 4) sgrid = sm_count * 4;
  octav_select<<<static_cast<unsigned>(sgrid), 256, 0, st>>>(trace, notclose, rows, iters, early_stop,
                                                             clip);
  return count_launch(2);
}

cudaError_t launch_octav_blocks(const float* x, long long n, int block, int bits, int iters,
                                float divisor, int early_stop, float* clip, void* ws, int sm_count,
                                cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (iters < 1 || iters > kMaxIter) return cudaErrorInvalidValue;
  unsigned* notclose = static_cast<unsigned*>(ws);
  float* trace = reinterpret_cast<float*>(static_cast<unsigned char*>(ws) + 256);
  cudaError_t e = cudaMemsetAsync(notclose, 0, 256, st);
  if (e != cudaSuccess) return e;
  const OctavConst k = make_const(bits, divisor, block);
  long long grid = (n / 8 + 255) / 256;
  const long long cap = static_cast<long long>(sm_count) * 8;
  if (grid > cap) grid = cap;
  const unsigned g = static_cast<unsigned>(grid);
  switch (block) {
    case 32: octav_blocks_trace<32><<<g, 256, 0, st>>>(x, n, k, iters, trace, notclose); break;
    case 64: octav_blocks_trace<64><<<g, 256, 0, st>>>(x, n, k, iters, trace, notclose); break;
    case 128: octav_blocks_trace<128><<<g, 256, 0, st>>>(x, n, k, iters, trace, notclose); break;
    case 256: octav_blocks_trace<256><<<g, 256, 0, st>>>(x, n, k, iters, trace, notclose); break;
    default: return cudaErrorInvalidValue;
  }
  const long long groups = n / block;
  long long sgrid = (groups + 255) / 256;
  if (sgrid > sm_count * 4) sgrid = sm_count * 4;
  octav_select<<<static_cast<unsigned>(sgrid), 256, 0, st>>>(trace, notclose, groups, iters,
                                                             early_stop, clip);
  return count_launch(2);
}

}  // namespace aeqb
