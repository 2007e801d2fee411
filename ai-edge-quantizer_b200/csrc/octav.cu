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
// with the group resident in registers (4 B read per weight, no re-reads; rows of 513..4096
// floats: one warp per row fed by bulk copies, `octav_rows_warp`), records
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

#include <cstdlib>

namespace aeqb {

namespace {

constexpr int kMaxIter = 32;
constexpr int kTensorMaxParts = 1024;

struct OctavConst {
  float s;         // f32(4^-bits / divisor)
  float one_m_s;   // 1.0f - s   (fp32, octav.py:105)
  double s_n;      // double(s) * N (octav.py:106, float64 because N is np.int64)
  float s_n_f32;   // s_n when it is an fp32 value (N a power of two, ...), else NaN
};

// den = f32(f64(den0) + s_n).  When s_n is itself an fp32 value and both terms sit within 29
// binary orders of each other (make_const checks the range of N and s), the float64 sum of two
// fp32 values is exact, so rounding it to fp32 is the correctly rounded fp32 sum: one FADD
// instead of F2F.F64 + DADD + F2F.F32.  Same bits either way.
__device__ __forceinline__ float octav_update(float num, int cnt, const OctavConst& k) {
  const float den0 = __fmul_rn(static_cast<float>(cnt), k.one_m_s);
  const float den = (k.s_n_f32 == k.s_n_f32) ? __fadd_rn(den0, k.s_n_f32)
                                             : static_cast<float>(static_cast<double>(den0) + k.s_n);
  return __fdiv_rn(num, den);
}

// Same with the count already an exact fp32 integer (< 2^24): no F2I / I2F round trip.
__device__ __forceinline__ float octav_update_f(float num, float cnt, const OctavConst& k) {
  const float den0 = __fmul_rn(cnt, k.one_m_s);
  const float den = (k.s_n_f32 == k.s_n_f32) ? __fadd_rn(den0, k.s_n_f32)
                                             : static_cast<float>(static_cast<double>(den0) + k.s_n);
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

// ---- branch-free inner loop on pre-conditioned magnitudes -------------------------------
// At load time every element becomes a = |x| (NaN -> -1, which no guess >= 0 selects) and the
// zeros of the group are counted once.  An iteration is then, per element pair,
//   m = (a >= g) ? 1.0f : 0.0f   (FSET x2)    sum += a * m   (FFMA2)    cnt += m   (FADD2)
// i.e. two issue slots per element on sm_100a's packed-fp32 pipe.  a * 1.0f and + 0.0f are exact,
// so the arithmetic is the masked sum / count of the reference; counts stay exact in fp32 below
// 2^24.  A NaN guess selects nothing (handled by the caller: the loop is skipped).
__device__ __forceinline__ float mask_ge(float a, float g) {
  float m;
  asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(m) : "f"(a), "f"(g));
  return m;
}

__device__ __forceinline__ float condition(float x, int& zeros) {
  zeros += (x == 0.0f);
  return fmaxf(fabsf(x), -1.0f);  // maxNum: |x|, or -1 for a NaN (one FMNMX)
}

__device__ __forceinline__ void acc_pair(float a0, float a1, float g, float2& sum, float2& cnt) {
  const float2 m = make_float2(mask_ge(a0, g), mask_ge(a1, g));
  sum = __ffma2_rn(make_float2(a0, a1), m, sum);
  cnt = __fadd2_rn(cnt, m);
}

// ------------------------------------------------------------------ rows
// One CTA per row at a time (grid-stride over rows); thread t keeps float4 chunks t, t + T, ...
// of the row in registers (NV of them) and the NEXT row's chunks in flight, so HBM is read
// exactly once and its latency hides behind the current row's iterations.  The recurrence
// stops at its fixed point (new == old bitwise: later iterates are identical) and the rest of
// the trace is filled, which removes most of the 10 iterations on real weights.
template <int NV, int THREADS, bool PREFETCH>
__global__ void __launch_bounds__(THREADS)
    octav_rows_trace(const float* __restrict__ x, long long rows, int cols, OctavConst k,
                     int iters, float* __restrict__ trace, unsigned* __restrict__ notclose) {
  constexpr int NW = THREADS / 32;
  __shared__ float s_sum[2][NW];
  __shared__ float s_cnt[2][NW];
  __shared__ int s_zero[NW];
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int nvec = cols >> 2;
  unsigned my_mask = 0;
  float4 nxt[NV];
  auto load_row = [&](long long row) {
    const float4* p = reinterpret_cast<const float4*>(x + row * cols);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = j * THREADS + tid;
      nxt[j] = (row < rows && i < nvec) ? __ldg(p + i) : make_float4(NAN, NAN, NAN, NAN);
    }
  };
  load_row(blockIdx.x);
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    float4 a[NV];
    int zeros = 0;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      a[j].x = condition(nxt[j].x, zeros); a[j].y = condition(nxt[j].y, zeros);
      a[j].z = condition(nxt[j].z, zeros); a[j].w = condition(nxt[j].w, zeros);
    }
    if (PREFETCH) load_row(row + gridDim.x);  // in flight during the iterations below
    zeros = __reduce_add_sync(0xffffffffu, zeros);
    if (lane == 0) s_zero[warp] = zeros;
    __syncthreads();
    int row_zeros = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) row_zeros += s_zero[w];

    float g = 1.0f;
    int it = 0;
    for (; it < iters; ++it) {
      float2 sum = make_float2(0.0f, 0.0f), cnt = make_float2(0.0f, 0.0f);
      if (g == g) {  // a NaN guess selects nothing
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          acc_pair(a[j].x, a[j].y, g, sum, cnt);
          acc_pair(a[j].z, a[j].w, g, sum, cnt);
        }
      }
      float ts = sum.x + sum.y, tc = cnt.x + cnt.y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ts += __shfl_xor_sync(0xffffffffu, ts, o);
        tc += __shfl_xor_sync(0xffffffffu, tc, o);
      }
      const int b = it & 1;
      if (lane == 0) {
        s_sum[b][warp] = ts;
        s_cnt[b][warp] = tc;
      }
      __syncthreads();  // double-buffered: one barrier per iteration
      float rs = 0.0f, rc = 0.0f;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        rs += s_sum[b][w];
        rc += s_cnt[b][w];
      }
      const int cnt_i = static_cast<int>(rc) + (g == 0.0f ? row_zeros : 0);
      const float ng = octav_update(rs, cnt_i, k);
      if (tid == 0) {
        trace[static_cast<long long>(it) * rows + row] = ng;
        if (!is_close(g, ng)) my_mask |= 1u << it;
      }
      const bool fixed = (ng == g) || (ng != ng && g != g);
      g = ng;
      if (fixed) { ++it; break; }  // CTA-uniform: every thread holds the same g
    }
    if (tid == 0) {
      // Past a fixed point every iterate equals g; np.isclose(g, g) holds unless g is NaN / inf-inf.
      const bool self_close = is_close(g, g);
      for (int r = it; r < iters; ++r) {
        trace[static_cast<long long>(r) * rows + row] = g;
        if (!self_close) my_mask |= 1u << r;
      }
    }
    __syncthreads();  // s_* are rewritten by the next row
    if (!PREFETCH) load_row(row + gridDim.x);
  }
  if (tid == 0 && my_mask) atomicOr(notclose, my_mask);
}

// ------------------------------------------------------------------ rows, one warp per row
// Rows of 513..4096 floats: a warp keeps its whole row in registers (NV float4 per lane), so an
// iteration has no barrier and no shared-memory exchange, only the xor-shuffle fold, and the
// fixed per-iteration cost (fold + update, ~60 issue slots) is spread over up to 128 elements per
// lane instead of 64.  Rows arrive by 1-D bulk copies (cp.async.bulk, one mbarrier per warp): the
// warp drains its buffer into registers and lane 0 immediately queues the NEXT row into the same
// buffer, so HBM latency hides behind the ten iterations without a second register set.  Four
// independent accumulator pairs keep the packed-fp32 pipe from waiting on one FFMA2 chain.
template <int NV>
struct WarpRowsCfg {
  static constexpr int kWarps = 4;
  static constexpr int kMinBlocks = NV >= 32 ? 3 : NV >= 24 ? 4 : NV >= 16 ? 5 : 8;
};

template <int NV, bool FULL, bool SKIP>
__global__ void __launch_bounds__(WarpRowsCfg<NV>::kWarps * 32, WarpRowsCfg<NV>::kMinBlocks)
    octav_rows_warp(const float* __restrict__ x, long long rows, int cols, OctavConst k, int iters,
                    float* __restrict__ trace, unsigned* __restrict__ notclose) {
  constexpr int W = WarpRowsCfg<NV>::kWarps;
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ __align__(8) uint64_t s_bar[W];
  const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int nvec = cols >> 2;  // FULL: nvec == NV * 32, no guards
  const uint32_t row_bytes = static_cast<uint32_t>(cols) * 4u;
  float4* buf = reinterpret_cast<float4*>(s_raw + static_cast<size_t>(warp) * row_bytes);
  uint64_t* bar = &s_bar[warp];
  const long long stride = static_cast<long long>(gridDim.x) * W;
  long long row = static_cast<long long>(blockIdx.x) * W + warp;
  if (lane == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    if (row < rows) {
      mbar_arrive_expect_tx(bar, row_bytes);
      bulk_g2s(buf, x + row * cols, row_bytes, bar);
    }
  }
  __syncwarp();
  unsigned my_mask = 0, parity = 0;
  for (; row < rows; row += stride) {
    mbar_wait(bar, parity);
    parity ^= 1u;
    float4 a[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = j * 32 + lane;
      float4 v;
      if (FULL) v = buf[i];
      else v = i < nvec ? buf[i] : make_float4(NAN, NAN, NAN, NAN);
      // a = |x|, or -1 for a NaN (maxNum, one FMNMX): no guess >= 0 selects it
      a[j].x = fmaxf(fabsf(v.x), -1.0f); a[j].y = fmaxf(fabsf(v.y), -1.0f);
      a[j].z = fmaxf(fabsf(v.z), -1.0f); a[j].w = fmaxf(fabsf(v.w), -1.0f);
    }
    __syncwarp();  // every lane has read the buffer: the next row may land in it
    if (lane == 0 && row + stride < rows) {
      mbar_arrive_expect_tx(bar, row_bytes);
      bulk_g2s(buf, x + (row + stride) * cols, row_bytes, bar);
    }
    // SKIP: lane j remembers the largest magnitude of float4 slot j over the whole warp (128
    // elements; REDUX.MAX on the bit patterns of non-negative floats).  An iteration then votes
    // once which slots hold ANY element >= the guess and jumps over the others: they would add
    // exactly +0 to both accumulators, so the result is bit-identical.  The guess climbs towards
    // the tail of the row (1 -> 0 -> mean |x| -> ...): from the fourth or fifth iteration on only
    // the slots with outliers are left, and the first iteration (guess 1) is usually empty.
    float slot_max = 0.0f;
    if (SKIP) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float m = fmaxf(fmaxf(fmaxf(a[j].x, a[j].y), fmaxf(a[j].z, a[j].w)), 0.0f);
        const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
        if (lane == j) slot_max = __uint_as_float(wm);
      }
    }

    // The first two iterates are known without touching the masks on ordinary weights: the start
    // guess 1 selects nothing when the row's largest magnitude is below 1 (sum 0, count 0 ->
    // guess 0), and the guess 0 selects every element that is not a NaN, so that iteration is a
    // plain packed sum (same accumulators, same order: a * 1.0f + s == a + s, bit-identical) with
    // the count known in advance.  `row_max` / `has_nan` come from the load.
    float row_max = 0.0f;
    bool has_nan = false;
    if (SKIP) {
      row_max = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(slot_max)));
      float mn = 0.0f;
#pragma unroll
      for (int j = 0; j < NV; ++j) mn = fminf(mn, fminf(fminf(a[j].x, a[j].y), fminf(a[j].z, a[j].w)));
      has_nan = __any_sync(0xffffffffu, mn < 0.0f);  // conditioned NaNs are -1
    }

    float g = 1.0f;
    int it = 0;
    for (; it < iters; ++it) {
      float2 s0 = make_float2(0.f, 0.f), s1 = s0, c0 = s0, c1 = s0;
      const unsigned active = SKIP ? __ballot_sync(0xffffffffu, slot_max >= g) : 0xffffffffu;
      const bool nothing = SKIP && g == g && g > row_max;              // no element reaches the guess
      const bool everything = SKIP && g == 0.0f && !has_nan && FULL;   // every element is selected
      if (everything) {
#pragma unroll
        for (int j = 0; j < NV; j += 2) {
          s0 = __fadd2_rn(make_float2(a[j].x, a[j].y), s0);
          s1 = __fadd2_rn(make_float2(a[j].z, a[j].w), s1);
          s0 = __fadd2_rn(make_float2(a[j + 1].x, a[j + 1].y), s0);
          s1 = __fadd2_rn(make_float2(a[j + 1].z, a[j + 1].w), s1);
        }
        c0 = make_float2(static_cast<float>(4 * NV), 0.0f);  // this lane's 4 * NV elements, all selected
      } else if (g == g && !nothing) {  // a NaN guess selects nothing
        // Eight masks are formed before the four packed FMAs / adds that consume them: the
        // compare runs on the half-rate ALU pipe with a longer latency than the FMA pipe, and a
        // warp has at most three neighbours on its scheduler to hide that behind.
#pragma unroll
        for (int j = 0; j < NV; j += 2) {
          if (SKIP && !(active & (3u << j))) continue;  // warp-uniform
          const float2 m0 = make_float2(mask_ge(a[j].x, g), mask_ge(a[j].y, g));
          const float2 m1 = make_float2(mask_ge(a[j].z, g), mask_ge(a[j].w, g));
          const float2 m2 = make_float2(mask_ge(a[j + 1].x, g), mask_ge(a[j + 1].y, g));
          const float2 m3 = make_float2(mask_ge(a[j + 1].z, g), mask_ge(a[j + 1].w, g));
          s0 = __ffma2_rn(make_float2(a[j].x, a[j].y), m0, s0);
          s1 = __ffma2_rn(make_float2(a[j].z, a[j].w), m1, s1);
          c0 = __fadd2_rn(c0, m0);
          c1 = __fadd2_rn(c1, m1);
          s0 = __ffma2_rn(make_float2(a[j + 1].x, a[j + 1].y), m2, s0);
          s1 = __ffma2_rn(make_float2(a[j + 1].z, a[j + 1].w), m3, s1);
          c0 = __fadd2_rn(c0, m2);
          c1 = __fadd2_rn(c1, m3);
        }
      }
      if (g == 0.0f) {
        // x >= 0 and x <= -0 both select a zero: zeros count twice at a zero guess (rare: once
        // per row on ordinary weights, so they are counted here and not at load time)
        float2 z0 = make_float2(0.f, 0.f), z1 = z0;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          z0 = __fadd2_rn(z0, make_float2(a[j].x == 0.0f ? 1.0f : 0.0f, a[j].y == 0.0f ? 1.0f : 0.0f));
          z1 = __fadd2_rn(z1, make_float2(a[j].z == 0.0f ? 1.0f : 0.0f, a[j].w == 0.0f ? 1.0f : 0.0f));
        }
        c0 = __fadd2_rn(c0, __fadd2_rn(z0, z1));
      }
      const float2 sa = __fadd2_rn(s0, s1), ca = __fadd2_rn(c0, c1);
      float ts = sa.x + sa.y;
      const int tc_lane = static_cast<int>(ca.x + ca.y);  // exact: <= 8 * NV
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ts += __shfl_xor_sync(0xffffffffu, ts, o);
      const int cnt_i = __reduce_add_sync(0xffffffffu, tc_lane);
      const float ng = octav_update(ts, cnt_i, k);
      if (lane == 0) {
        trace[static_cast<long long>(it) * rows + row] = ng;
        if (!is_close(g, ng)) my_mask |= 1u << it;
      }
      const bool fixed = (ng == g) || (ng != ng && g != g);
      g = ng;
      if (fixed) { ++it; break; }  // warp-uniform: every lane holds the same g
    }
    if (lane == 0) {
      const bool self_close = is_close(g, g);
      for (int r = it; r < iters; ++r) {
        trace[static_cast<long long>(r) * rows + row] = g;
        if (!self_close) my_mask |= 1u << r;
      }
    }
  }
  if (lane == 0 && my_mask) atomicOr(notclose, my_mask);
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
// A warp owns 1024 consecutive floats at a time = 32 segments of 32; the tile is loaded with
// coalesced 128-bit loads, transposed through a padded shared-memory tile (row stride 33: both
// the scatter and the gather are bank-conflict free) so that lane l holds segment l in
// registers, conditioned once.  A block is BLOCK/32 adjacent lanes; its iterations run from
// registers with xor-shuffles across those lanes (none at BLOCK == 32).
template <int BLOCK>
__global__ void __launch_bounds__(256)
    octav_blocks_trace(const float* __restrict__ x, long long n, OctavConst k, int iters,
                       float* __restrict__ trace, unsigned* __restrict__ notclose) {
  constexpr int LPB = BLOCK / 32;
  __shared__ float s_t[8][32 * 33];
  __shared__ unsigned s_mask;
  if (threadIdx.x == 0) s_mask = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  float* t = s_t[warp];
  const long long nblk = n / BLOCK;
  const long long nwt = (n + 1023) / 1024;  // warp tiles
  const bool aligned = reinterpret_cast<uintptr_t>(x) % 16 == 0;
  unsigned my_mask = 0;
  for (long long wt = static_cast<long long>(blockIdx.x) * 8 + warp; wt < nwt;
       wt += static_cast<long long>(gridDim.x) * 8) {
    const long long e0 = wt * 1024;
    // ---- coalesced load + scatter into the padded tile
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = i * 32 + lane;               // float4 index inside the tile
      const long long e = e0 + 4LL * f;
      float4 v = make_float4(NAN, NAN, NAN, NAN);
      if (e + 3 < n) {
        if (aligned) v = __ldg(reinterpret_cast<const float4*>(x + e));
        else v = make_float4(x[e], x[e + 1], x[e + 2], x[e + 3]);
      }
      const int seg = f >> 3, pos = (f & 7) * 4;
      float* d = t + seg * 33 + pos;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncwarp();
    float a[32];
    // a = |x|, or -1 for a NaN (maxNum, one FMNMX): no guess >= 0 selects it
#pragma unroll
    for (int p = 0; p < 32; ++p) a[p] = fmaxf(fabsf(t[lane * 33 + p]), -1.0f);
    __syncwarp();
    const long long seg_e = e0 + 32LL * lane;    // first element of this lane's segment
    const bool valid = seg_e < n;
    const long long blk = seg_e / BLOCK;
    const bool writer = valid && (lane & (LPB - 1)) == 0;

    float g = 1.0f;
    int it = 0;
    for (; it < iters; ++it) {
      float2 sum = make_float2(0.0f, 0.0f), cnt = make_float2(0.0f, 0.0f);
      if (g == g) {
#pragma unroll
        for (int p = 0; p < 32; p += 2) acc_pair(a[p], a[p + 1], g, sum, cnt);
      }
      if (__any_sync(0xffffffffu, g == 0.0f)) {
        // x >= 0 and x <= -0 both select a zero: at a zero guess the zeros count twice.  Rare
        // (one iteration per block on ordinary weights), so they are counted here, not at load.
        float2 z = make_float2(0.f, 0.f);
#pragma unroll
        for (int p = 0; p < 32; p += 2)
          z = __fadd2_rn(z, make_float2(a[p] == 0.0f ? 1.0f : 0.0f, a[p + 1] == 0.0f ? 1.0f : 0.0f));
        if (g == 0.0f) cnt = __fadd2_rn(cnt, z);
      }
      float ts = sum.x + sum.y, tc = cnt.x + cnt.y;
#pragma unroll
      for (int o = 1; o < LPB; o <<= 1) {
        ts += __shfl_xor_sync(0xffffffffu, ts, o);
        tc += __shfl_xor_sync(0xffffffffu, tc, o);
      }
      const float ng = octav_update_f(ts, tc, k);
      if (writer) {
        trace[static_cast<long long>(it) * nblk + blk] = ng;
        if (!is_close(g, ng)) my_mask |= 1u << it;
      }
      const bool fixed = (ng == g) || (ng != ng && g != g);
      g = ng;
      // leave only when every block of the warp sits at its fixed point (shuffles need all lanes)
      if (__all_sync(0xffffffffu, fixed)) { ++it; break; }
    }
    if (writer) {
      const bool self_close = is_close(g, g);
      for (int r = it; r < iters; ++r) {
        trace[static_cast<long long>(r) * nblk + blk] = g;
        if (!self_close) my_mask |= 1u << r;
      }
    }
  }
  my_mask = __reduce_or_sync(0xffffffffu, my_mask);
  if (lane == 0 && my_mask) atomicOr(&s_mask, my_mask);
  __syncthreads();
  if (threadIdx.x == 0 && s_mask) atomicOr(notclose, s_mask);
}

// ------------------------------------------------------------------ whole tensor (TENSORWISE)
// One reduction group = the whole tensor: iteration i needs a grid-wide masked sum, so each
// iteration is two launches over the (L2-resident for <= ~100 MB) tensor: per-CTA partials, then a
// one-CTA step that folds them in fixed order, applies the update and appends to the trace.
// state: [0] current guess (float), [1] zeros count (as int bits), [2..] unused.
struct TensorPartial {
  double sum;
  long long cnt;
  long long zeros;
};

__global__ void __launch_bounds__(256)
    octav_tensor_partials(const float* __restrict__ x, long long n, const float* __restrict__ state,
                          TensorPartial* __restrict__ part) {
  __shared__ double s_sum[8];
  __shared__ long long s_cnt[8], s_zero[8];
  const float g = state[0];
  const bool live = g == g;  // a NaN guess selects nothing
  float2 sum = make_float2(0.f, 0.f), cnt = make_float2(0.f, 0.f);
  double dsum = 0.0;
  long long dcnt = 0, zeros = 0;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const uintptr_t addr = reinterpret_cast<uintptr_t>(x);
  long long head = ((16 - (addr & 15)) & 15) / 4;
  if (head > n) head = n;
  const long long nvec = (n - head) / 4;
  const float4* xv = reinterpret_cast<const float4*>(x + head);
  int run = 0;
  for (long long i = tid; i < nvec; i += nthreads) {
    const float4 v = __ldg(xv + i);
    int z = 0;
    const float a0 = condition(v.x, z), a1 = condition(v.y, z), a2 = condition(v.z, z), a3 = condition(v.w, z);
    zeros += z;
    if (live) {
      acc_pair(a0, a1, g, sum, cnt);
      acc_pair(a2, a3, g, sum, cnt);
    }
    if (++run == 64) {  // bound the fp32 partial runs, keep the long accumulation in fp64
      dsum += static_cast<double>(sum.x) + static_cast<double>(sum.y);
      dcnt += static_cast<long long>(cnt.x + cnt.y);
      sum = make_float2(0.f, 0.f); cnt = make_float2(0.f, 0.f);
      run = 0;
    }
  }
  // scalar head / tail
  for (long long i = tid; i < head + (n - head - nvec * 4); i += nthreads) {
    const long long e = i < head ? i : nvec * 4 + i;
    int z = 0;
    const float a = condition(x[e], z);
    zeros += z;
    if (live && a >= g) { dsum += static_cast<double>(a); dcnt += 1; }
  }
  dsum += static_cast<double>(sum.x) + static_cast<double>(sum.y);
  dcnt += static_cast<long long>(cnt.x + cnt.y);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    dcnt += __shfl_xor_sync(0xffffffffu, dcnt, o);
    zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_sum[warp] = dsum; s_cnt[warp] = dcnt; s_zero[warp] = zeros; }
  __syncthreads();
  if (threadIdx.x == 0) {
    TensorPartial t{0.0, 0, 0};
    for (int w = 0; w < 8; ++w) { t.sum += s_sum[w]; t.cnt += s_cnt[w]; t.zeros += s_zero[w]; }
    part[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(32)
    octav_tensor_step(const TensorPartial* __restrict__ part, int nparts, OctavConst k, int it,
                      float* __restrict__ state, float* __restrict__ trace,
                      unsigned* __restrict__ notclose) {
  if (threadIdx.x != 0) return;
  double sum = 0.0;
  long long cnt = 0, zeros = 0;
  for (int i = 0; i < nparts; ++i) { sum += part[i].sum; cnt += part[i].cnt; zeros += part[i].zeros; }
  const float g = state[0];
  if (g == 0.0f) cnt += zeros;
  const float ng = octav_update_ll(static_cast<float>(sum), cnt, k);
  trace[it] = ng;
  if (!is_close(g, ng)) atomicOr(notclose, 1u << it);
  state[0] = ng;
}

__global__ void octav_tensor_init(float* state) { state[0] = 1.0f; }

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
  // fp32 shortcut of octav_update: s_n exact in fp32, and |log2(den0 / s_n)| <= 29 for every
  // count 1..n (den0 in [1 - s, n], s_n = s * n, s >= 2^-18 for bits <= 8, n <= 2^24: the two
  // ratios are bounded by 1 / s and s * n < 2^24); den0 == 0 adds exactly.
  const float f = static_cast<float>(k.s_n);
  const bool exact = static_cast<double>(f) == k.s_n && bits >= 1 && bits <= 8 && divisor >= 1.0f &&
                     divisor <= 64.0f && n >= 1 && n <= (1LL << 24);
  k.s_n_f32 = exact ? f : NAN;
  return k;
}

template <int NV, int T>
void launch_rows_nv(const float* x, long long rows, int cols, const OctavConst& k, int iters,
                    float* trace, unsigned* notclose, int sm_count, cudaStream_t st) {
  // Measured on B200 (tools/ktime.py): insensitive to 4..16 CTAs per SM and to the prefetch; the
  // kernel is issue-bound (IPC 2.45 of 4), see profiles/.
  const int per_sm = (2048 / T) < 12 ? (2048 / T) : 12;
  long long grid = static_cast<long long>(sm_count) * per_sm;
  if (grid > rows) grid = rows;
  octav_rows_trace<NV, T, false><<<static_cast<unsigned>(grid), T, 0, st>>>(x, rows, cols, k, iters,
                                                                             trace, notclose);
}

template <int NV, bool FULL, bool SKIP>
cudaError_t launch_rows_warp_k(const float* x, long long rows, int cols, const OctavConst& k,
                               int iters, float* trace, unsigned* notclose, int sm_count,
                               cudaStream_t st) {
  constexpr int W = WarpRowsCfg<NV>::kWarps;
  const size_t smem = static_cast<size_t>(W) * static_cast<size_t>(cols) * 4;
  static PerDevice configured;  // per instantiation; the attribute is idempotent
  if (!configured.done()) {
    cudaError_t e = cudaFuncSetAttribute(octav_rows_warp<NV, FULL, SKIP>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, W * NV * 512);
    if (e != cudaSuccess) return e;
    configured.set();
  }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, octav_rows_warp<NV, FULL, SKIP>,
                                                                W * 32, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  long long grid = static_cast<long long>(sm_count) * per_sm;
  const long long need = (rows + W - 1) / W;
  if (grid > need) grid = need;
  octav_rows_warp<NV, FULL, SKIP><<<static_cast<unsigned>(grid), W * 32, smem, st>>>(
      x, rows, cols, k, iters, trace, notclose);
  return cudaSuccess;
}

template <int NV>
cudaError_t launch_rows_warp(const float* x, long long rows, int cols, const OctavConst& k,
                             int iters, float* trace, unsigned* notclose, int sm_count,
                             cudaStream_t st) {
  static const bool skip = getenv("AEQB_OCTAV_NO_SKIP") == nullptr;  // A/B runs
  if (cols == NV * 128)
    return skip ? launch_rows_warp_k<NV, true, true>(x, rows, cols, k, iters, trace, notclose, sm_count, st)
                : launch_rows_warp_k<NV, true, false>(x, rows, cols, k, iters, trace, notclose, sm_count, st);
  return skip ? launch_rows_warp_k<NV, false, true>(x, rows, cols, k, iters, trace, notclose, sm_count, st)
              : launch_rows_warp_k<NV, false, false>(x, rows, cols, k, iters, trace, notclose, sm_count, st);
}

}  // namespace

size_t octav_workspace_bytes(long long groups, int iters) {
  // mask slot + trace [+ per-CTA partials of the whole-tensor path, used when groups == 1]
  return static_cast<size_t>(groups) * static_cast<size_t>(iters) * sizeof(float) + 1024 +
         static_cast<size_t>(kTensorMaxParts) * sizeof(TensorPartial);
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
  if (rows == 1 && cols > 65536) {  // whole-tensor group: grid-wide reduction per iteration
    unsigned char* base = static_cast<unsigned char*>(ws);
    float* state = reinterpret_cast<float*>(base + 64);
    TensorPartial* part = reinterpret_cast<TensorPartial*>(base + 1024);  // past mask, state and the <= 128 B trace
    int grid = sm_count * 4;
    if (grid > kTensorMaxParts) grid = kTensorMaxParts;
    octav_tensor_init<<<1, 1, 0, st>>>(state);
    for (int it = 0; it < iters; ++it) {
      octav_tensor_partials<<<grid, 256, 0, st>>>(x, cols, state, part);
      octav_tensor_step<<<1, 32, 0, st>>>(part, grid, k, it, state, trace, notclose);
    }
    octav_select<<<1, 256, 0, st>>>(trace, notclose, 1, iters, early_stop, clip);
    return count_launch(2 + 2 * iters);
  }
  const bool vec = cols > 0 && cols % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
                   cols <= 16384;
  if (vec) {
    const int c = static_cast<int>(cols);
    const int v4 = c / 4;  // float4 per row; <= 8 per thread, thread count grows with the row
    static const bool cta_rows = getenv("AEQB_OCTAV_CTA_ROWS") != nullptr;  // A/B runs
    cudaError_t le = cudaSuccess;
    if (v4 <= 128) launch_rows_nv<1, 128>(x, rows, c, k, iters, trace, notclose, sm_count, st);
    else if (v4 <= 1024 && !cta_rows) {  // 513..4096 floats: one warp per row, row in registers
      if (v4 <= 256) le = launch_rows_warp<8>(x, rows, c, k, iters, trace, notclose, sm_count, st);
      else if (v4 <= 512) le = launch_rows_warp<16>(x, rows, c, k, iters, trace, notclose, sm_count, st);
      else if (v4 <= 768) le = launch_rows_warp<24>(x, rows, c, k, iters, trace, notclose, sm_count, st);
      else le = launch_rows_warp<32>(x, rows, c, k, iters, trace, notclose, sm_count, st);
      if (le != cudaSuccess) return le;
    }
    else if (v4 <= 256) launch_rows_nv<2, 128>(x, rows, c, k, iters, trace, notclose, sm_count, st);
    else if (v4 <= 512) launch_rows_nv<4, 128>(x, rows, c, k, iters, trace, notclose, sm_count, st);
    else if (v4 <= 1024) launch_rows_nv<16, 64>(x, rows, c, k, iters, trace, notclose, sm_count, st);
    else if (v4 <= 2048) launch_rows_nv<8, 256>(x, rows, c, k, iters, trace, notclose, sm_count, st);
    else launch_rows_nv<8, 512>(x, rows, c, k, iters, trace, notclose, sm_count, st);
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
  long long grid = ((n + 1023) / 1024 + 7) / 8;  // 8 warp tiles per CTA pass
  const long long cap = static_cast<long long>(sm_count) * 6;
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
