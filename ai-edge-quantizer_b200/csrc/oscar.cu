// OSCAR (activation-aware channel scaling + exact optimal clipping), device side of
// algorithms/uniform_quantize/oscar.py.  The reference works in float64 throughout; so do
// these kernels (B200 issues 64 DFMA / clk / SM, and every pass is a streaming read of W).
//
//   colsq_f64          out[j] = alpha * sum_i x[i, j]^2          calibrate's mu2 (:271-277) and
//                                                                 a_base (:213)
//   oscar_group_pass   per (row, column group): max_j |w_ij| s_j  _channel_scale_objective
//                      -> sum of squares per group, and the        (:172-189) and the a_eff
//                      scatter a_eff[argmax] += w^2                scatter (:223-230)
//   oscar_clip_rows / oscar_clip_blocks / tensor path             _optimal_group_clip (:62-104):
//                      sort |w s| descending with the masses m,   breakpoint scan over sorted
//                      three running sums, candidate per           magnitudes, exact argmin
//                      breakpoint, argmin
//   oscar_scale        bound -> scale                             tensor_zp_scale_from_min_max
//                                                                 (uqt:492-586, float64 inputs)
//   oscar_quantize     q = clip(rint((w s) / scale))              uniform_quantize on the scaled
//                                                                 float64 weight (:455-457)
//
// Sorting: rows (<= 16384 columns) and 32..256-wide blocks are (key, index) bitonic sorts in
// shared memory, the same segmented network as recovery.cu; one group spanning the whole
// tensor (TENSORWISE) uses cub::DeviceRadixSort + cub::DeviceScan (library code, that path
// only).  Running sums are parallel scans, NumPy's cumsum is sequential: same values to a few
// float64 ulps, which only matters for exact ties of the objective.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "aeqb_common.cuh"
#include "aeqb_kernels.h"

namespace aeqb {

namespace {

constexpr double kOscarEps = 1e-12;  // _EPS (:52)

// ------------------------------------------------------------------ column sums of squares
// part[split][j] = sum over the split's rows of x[i, j]^2 (thread per column, coalesced rows).
__global__ void __launch_bounds__(256)
    colsq_partial(const float* __restrict__ x, long long n, int d, long long rows_per_split,
                  double* __restrict__ part) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d) return;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_split;
  long long r1 = r0 + rows_per_split;
  if (r1 > n) r1 = n;
  double acc = 0.0;
  for (long long i = r0; i < r1; ++i) {
    const double v = static_cast<double>(x[i * d + j]);
    acc = fma(v, v, acc);
  }
  part[static_cast<long long>(blockIdx.y) * d + j] = acc;
}
__global__ void __launch_bounds__(256)
    colsq_reduce(const double* __restrict__ part, int splits, int d, double alpha,
                 double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d) return;
  double acc = 0.0;
  for (int s = 0; s < splits; ++s) acc += part[static_cast<long long>(s) * d + j];
  out[j] = acc * alpha;
}

// ------------------------------------------------------------------ objective / a_eff pass
// One warp per (column group gi, chunk of rows).  Lanes stride over the group's g columns.
// part_sq[chunk][gi] = sum over the chunk's rows of (max_j |w_ij| s_j)^2;
// a_eff[j*] += w[i, j*]^2 for the first arg-max column j* of every (row, group).
__global__ void __launch_bounds__(256)
    oscar_group_pass(const float* __restrict__ W, long long n, int d, int g,
                     const double* __restrict__ s, long long rows_per_chunk,
                     double* __restrict__ part_sq, double* __restrict__ a_eff) {
  const int groups = d / g;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long n_chunks = (n + rows_per_chunk - 1) / rows_per_chunk;
  if (warp >= n_chunks * groups) return;
  const int gi = static_cast<int>(warp % groups);
  const long long chunk = warp / groups;
  const long long r0 = chunk * rows_per_chunk;
  long long r1 = r0 + rows_per_chunk;
  if (r1 > n) r1 = n;
  const int c0 = gi * g;
  double acc = 0.0;
  for (long long i = r0; i < r1; ++i) {
    const float* row = W + i * d + c0;
    double best = -1.0;
    int arg = 0x7fffffff;
    for (int e = lane; e < g; e += 32) {
      const double v = fabs(static_cast<double>(row[e])) * s[c0 + e];
      if (v > best) {  // strictly greater: the first maximum wins inside a lane
        best = v;
        arg = e;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) {
        best = ob;
        arg = oa;
      }
    }
    if (lane == 0) {
      acc = fma(best, best, acc);
      if (a_eff != nullptr) {
        const double wv = static_cast<double>(row[arg]);
        atomicAdd(&a_eff[c0 + arg], wv * wv);
      }
    }
  }
  if (lane == 0) part_sq[chunk * groups + gi] = acc;
}
__global__ void __launch_bounds__(256)
    oscar_pass_reduce(const double* __restrict__ part_sq, long long n_chunks, int groups,
                      double* __restrict__ group_sq) {
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= groups) return;
  double acc = 0.0;
  for (long long c = 0; c < n_chunks; ++c) acc += part_sq[c * groups + gi];
  group_sq[gi] = acc;
}

// ------------------------------------------------------------------ optimal clip per group
struct Cand {
  double e;
  double c;
  long long k;  // candidate index: 0 = no clipping, k + 1 = "k + 1 largest magnitudes clipped"
};
__device__ __forceinline__ void cand_min(Cand& a, const Cand& b) {
  // np.argmin: first occurrence of the minimum; NaN never wins a `<`.
  if (b.e < a.e || (b.e == a.e && b.k < a.k)) a = b;
}
// Candidate after the (k+1)-th largest magnitude (inclusive running sums S_* up to it).
__device__ __forceinline__ Cand make_cand(double a_k, double a_next, double s_m, double s_am,
                                          double s_a2m, double mass, double qq, long long k) {
  // c_k = 2.0 * s_am / (mass / (6.0 * qmax * qmax) + 2.0 * s_m)       (:88)
  double c = __ddiv_rn(__dmul_rn(2.0, s_am),
                       __dadd_rn(__ddiv_rn(mass, 6.0 * qq), __dmul_rn(2.0, s_m)));
  c = fmin(fmax(c, a_next), a_k);  // np.clip(c_k, lower, a_s)          (:90)
  const double c2 = __dmul_rn(c, c);
  // e_k = c^2 * (mass / (12 q^2)) + s_a2m - 2 c * s_am + c^2 * s_m     (:91-96)
  double e = __dmul_rn(c2, __ddiv_rn(mass, 12.0 * qq));
  e = __dadd_rn(e, s_a2m);
  e = __dsub_rn(e, __dmul_rn(__dmul_rn(2.0, c), s_am));
  e = __dadd_rn(e, __dmul_rn(c2, s_m));
  Cand r;
  r.e = e;
  r.c = c;
  r.k = k + 1;
  return r;
}

// Descending bitonic sort of (key, idx) pairs in shared memory, independent segments of `seg`.
__device__ __forceinline__ void bitonic_pairs_desc(double* key, unsigned* idx, int tile, int seg) {
  const int half = tile >> 1;
  for (int k = 2; k <= seg; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int p = threadIdx.x; p < half; p += blockDim.x) {
        const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
        const int q = i | j;
        const bool down = (k == seg) || ((i & k) == 0);
        const double a = key[i], b = key[q];
        if ((a < b) == down) {
          key[i] = b;
          key[q] = a;
          const unsigned t = idx[i];
          idx[i] = idx[q];
          idx[q] = t;
        }
      }
      __syncthreads();
    }
  }
}

// One CTA per row (grid-stride), the row (d <= 16384 columns, padded to `seg`) is one group.
// Shared: key[seg] double | idx[seg] u32 | scan scratch.
__global__ void __launch_bounds__(1024)
    oscar_clip_rows(const float* __restrict__ W, long long n, int d, int seg,
                    const double* __restrict__ s, const double* __restrict__ m, double mass,
                    double qq, double* __restrict__ bound) {
  extern __shared__ __align__(16) unsigned char oscar_smem[];
  double* key = reinterpret_cast<double*>(oscar_smem);
  unsigned* idx = reinterpret_cast<unsigned*>(key + seg);
  __shared__ double w_m[32], w_am[32], w_a2m[32];
  __shared__ Cand w_c[32];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
  const int run = seg / nt;  // contiguous sorted elements per thread (seg >= nt by launch)
  for (long long row = blockIdx.x; row < n; row += gridDim.x) {
    for (int j = tid; j < seg; j += nt) {
      key[j] = j < d ? fabs(__dmul_rn(static_cast<double>(W[row * d + j]), s[j])) : -1.0;
      idx[j] = j;
    }
    __syncthreads();
    bitonic_pairs_desc(key, idx, seg, seg);
    // thread-local totals of its run
    const int k0 = tid * run;
    double t_m = 0.0, t_am = 0.0, t_a2m = 0.0;
    for (int u = 0; u < run; ++u) {
      const int k = k0 + u;
      if (k < d) {
        const double a = key[k], mm = m[idx[k]];
        t_m += mm;
        t_am += a * mm;
        t_a2m += a * a * mm;
      }
    }
    // exclusive block scan of the three totals
    double x_m = t_m, x_am = t_am, x_a2m = t_a2m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double y0 = __shfl_up_sync(0xffffffffu, x_m, o);
      const double y1 = __shfl_up_sync(0xffffffffu, x_am, o);
      const double y2 = __shfl_up_sync(0xffffffffu, x_a2m, o);
      if (lane >= o) {
        x_m += y0;
        x_am += y1;
        x_a2m += y2;
      }
    }
    if (lane == 31) {
      w_m[wid] = x_m;
      w_am[wid] = x_am;
      w_a2m[wid] = x_a2m;
    }
    __syncthreads();
    double b_m = 0.0, b_am = 0.0, b_a2m = 0.0;
    for (int w = 0; w < wid; ++w) {
      b_m += w_m[w];
      b_am += w_am[w];
      b_a2m += w_a2m[w];
    }
    double s_m = b_m + x_m - t_m, s_am = b_am + x_am - t_am, s_a2m = b_a2m + x_a2m - t_a2m;
    Cand best;
    best.e = __longlong_as_double(0x7ff0000000000000LL);
    best.c = 0.0;
    best.k = 0x7fffffffffffffffLL;
    if (tid == 0) {  // the no-clip candidate (:97-98)
      const double c0 = key[0];
      best.e = __dmul_rn(__dmul_rn(c0, c0), __ddiv_rn(mass, 12.0 * qq));
      best.c = c0;
      best.k = 0;
    }
    for (int u = 0; u < run; ++u) {
      const int k = k0 + u;
      if (k < d) {
        const double a = key[k], mm = m[idx[k]];
        s_m += mm;
        s_am += a * mm;
        s_a2m += a * a * mm;
        const double a_next = (k + 1 < d) ? key[k + 1] : 0.0;
        cand_min(best, make_cand(a, a_next, s_m, s_am, s_a2m, mass, qq, k));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Cand other;
      other.e = __shfl_xor_sync(0xffffffffu, best.e, o);
      other.c = __shfl_xor_sync(0xffffffffu, best.c, o);
      other.k = __shfl_xor_sync(0xffffffffu, best.k, o);
      cand_min(best, other);
    }
    if (lane == 0) w_c[wid] = best;
    __syncthreads();
    if (tid == 0) {
      Cand r = w_c[0];
      for (int w = 1; w < (nt >> 5); ++w) cand_min(r, w_c[w]);
      bound[row] = r.c;
    }
    __syncthreads();
  }
}

// Blocks of g = 32..256 columns: a tile of TILE floats holds TILE / g groups of one row segment;
// after the segmented sort one warp scans one group (g / 32 sorted elements per lane).
constexpr int kOscarTile = 2048;
__global__ void __launch_bounds__(256)
    oscar_clip_blocks(const float* __restrict__ W, long long total, int d, int g,
                      const double* __restrict__ s, const double* __restrict__ m,
                      const double* __restrict__ mass_g, double qq, double* __restrict__ bound) {
  __shared__ double key[kOscarTile];
  __shared__ unsigned idx[kOscarTile];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int gpt = kOscarTile / g;
  const int run = g / 32;
  const long long n_tiles = (total + kOscarTile - 1) / kOscarTile;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long base = t * kOscarTile;
    for (int e = tid; e < kOscarTile; e += blockDim.x) {
      const long long p = base + e;
      const int col = static_cast<int>(p % d);
      key[e] = p < total ? fabs(__dmul_rn(static_cast<double>(W[p]), s[col])) : -1.0;
      idx[e] = static_cast<unsigned>(col);
    }
    __syncthreads();
    bitonic_pairs_desc(key, idx, kOscarTile, g);
    for (int gi = wid; gi < gpt; gi += (blockDim.x >> 5)) {
      const long long p0 = base + static_cast<long long>(gi) * g;
      if (p0 >= total) continue;
      const double* kk = key + gi * g;
      const unsigned* ii = idx + gi * g;
      const double mass = mass_g[static_cast<int>(p0 % d) / g];
      const int k0 = lane * run;
      double t_m = 0.0, t_am = 0.0, t_a2m = 0.0;
      for (int u = 0; u < run; ++u) {
        const double a = kk[k0 + u], mm = m[ii[k0 + u]];
        t_m += mm;
        t_am += a * mm;
        t_a2m += a * a * mm;
      }
      double x_m = t_m, x_am = t_am, x_a2m = t_a2m;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double y0 = __shfl_up_sync(0xffffffffu, x_m, o);
        const double y1 = __shfl_up_sync(0xffffffffu, x_am, o);
        const double y2 = __shfl_up_sync(0xffffffffu, x_a2m, o);
        if (lane >= o) {
          x_m += y0;
          x_am += y1;
          x_a2m += y2;
        }
      }
      double s_m = x_m - t_m, s_am = x_am - t_am, s_a2m = x_a2m - t_a2m;
      Cand best;
      best.e = __longlong_as_double(0x7ff0000000000000LL);
      best.c = 0.0;
      best.k = 0x7fffffffffffffffLL;
      if (lane == 0) {
        const double c0 = kk[0];
        best.e = __dmul_rn(__dmul_rn(c0, c0), __ddiv_rn(mass, 12.0 * qq));
        best.c = c0;
        best.k = 0;
      }
      for (int u = 0; u < run; ++u) {
        const int k = k0 + u;
        const double a = kk[k], mm = m[ii[k]];
        s_m += mm;
        s_am += a * mm;
        s_a2m += a * a * mm;
        const double a_next = (k + 1 < g) ? kk[k + 1] : 0.0;
        cand_min(best, make_cand(a, a_next, s_m, s_am, s_a2m, mass, qq, k));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Cand other;
        other.e = __shfl_xor_sync(0xffffffffu, best.e, o);
        other.c = __shfl_xor_sync(0xffffffffu, best.c, o);
        other.k = __shfl_xor_sync(0xffffffffu, best.k, o);
        cand_min(best, other);
      }
      if (lane == 0) bound[p0 / g] = best.c;
    }
    __syncthreads();
  }
}

// ---- whole tensor as one group: keys / terms for the library sort + scans
__global__ void __launch_bounds__(256)
    oscar_tensor_keys(const float* __restrict__ W, long long total, int d,
                      const double* __restrict__ s, double* __restrict__ key,
                      unsigned* __restrict__ idx) {
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < total; p += step) {
    key[p] = fabs(__dmul_rn(static_cast<double>(W[p]), s[p % d]));
    idx[p] = static_cast<unsigned>(p);
  }
}
__global__ void __launch_bounds__(256)
    oscar_tensor_terms(const double* __restrict__ key, const unsigned* __restrict__ idx,
                       long long total, int d, const double* __restrict__ m,
                       double* __restrict__ t_m, double* __restrict__ t_am,
                       double* __restrict__ t_a2m) {
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < total; p += step) {
    const double a = key[p], mm = m[idx[p] % d];
    t_m[p] = mm;
    t_am[p] = a * mm;
    t_a2m[p] = a * a * mm;
  }
}
__global__ void __launch_bounds__(256)
    oscar_tensor_cands(const double* __restrict__ key, long long total,
                       const double* __restrict__ s_m, const double* __restrict__ s_am,
                       const double* __restrict__ s_a2m, double mass, double qq,
                       Cand* __restrict__ part) {
  __shared__ Cand w_c[8];
  Cand best;
  best.e = __longlong_as_double(0x7ff0000000000000LL);
  best.c = 0.0;
  best.k = 0x7fffffffffffffffLL;
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long first = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (first == 0) {
    const double c0 = key[0];
    best.e = __dmul_rn(__dmul_rn(c0, c0), __ddiv_rn(mass, 12.0 * qq));
    best.c = c0;
    best.k = 0;
  }
  for (long long p = first; p < total; p += step) {
    const double a_next = (p + 1 < total) ? key[p + 1] : 0.0;
    cand_min(best, make_cand(key[p], a_next, s_m[p], s_am[p], s_a2m[p], mass, qq, p));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Cand other;
    other.e = __shfl_xor_sync(0xffffffffu, best.e, o);
    other.c = __shfl_xor_sync(0xffffffffu, best.c, o);
    other.k = __shfl_xor_sync(0xffffffffu, best.k, o);
    cand_min(best, other);
  }
  if ((threadIdx.x & 31) == 0) w_c[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    Cand r = w_c[0];
    for (int w = 1; w < 8; ++w) cand_min(r, w_c[w]);
    part[blockIdx.x] = r;
  }
}
__global__ void oscar_tensor_pick(const Cand* __restrict__ part, int n, double* __restrict__ bound) {
  Cand r = part[0];
  for (int i = 1; i < n; ++i) cand_min(r, part[i]);
  bound[0] = r.c;
}

// ------------------------------------------------------------------ scale and quantise
// scale = max(bound, 1e-9) / qmax in float64; blockwise: -> bf16 -> fp16 -> fp32 like
// ml_dtypes (float64 goes through float32 first).  Stored as float64 either way.
__global__ void __launch_bounds__(256)
    oscar_scale(const double* __restrict__ bound, long long n, double qmax, int blockwise,
                double* __restrict__ scale) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double sc = __ddiv_rn(fmax(fabs(bound[i]), 1e-9), qmax);
  if (bound[i] != bound[i]) sc = bound[i];
  if (blockwise) {
    uint16_t bits;
    sc = static_cast<double>(round_scale_bf16_f16(static_cast<float>(sc), &bits));
  }
  scale[i] = sc;
}

// q[i, j] = clip(rint((w_ij * s_j) / scale[group(i, j)]), lo, hi); group = (i*d + j) / glen
// (glen = d: per row, glen = block, glen = n*d: whole tensor).  0 / 0 -> NaN -> 0 like the
// reference's NaN -> int8 cast.
__global__ void __launch_bounds__(256)
    oscar_quantize(const float* __restrict__ W, long long total, int d, long long glen,
                   const double* __restrict__ s, const double* __restrict__ scale, int lo, int hi,
                   int8_t* __restrict__ q) {
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < total; p += step) {
    const double v = __ddiv_rn(__dmul_rn(static_cast<double>(W[p]), s[p % d]), scale[p / glen]);
    double r = rint(v);
    int o = 0;
    if (r == r) o = static_cast<int>(fmin(fmax(r, static_cast<double>(lo)), static_cast<double>(hi)));
    q[p] = static_cast<int8_t>(o);
  }
}

int pow2_ge(long long v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}
inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }
constexpr int kTensorCandBlocks = 1024;

struct TensorLayout {
  size_t key_in, key_out, idx_in, idx_out, t0, t1, t2, part, temp, temp_bytes, total;
};
TensorLayout tensor_layout(long long total) {
  TensorLayout l;
  size_t sort_tmp = 0, scan_tmp = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_tmp, static_cast<const double*>(nullptr),
                                            static_cast<double*>(nullptr),
                                            static_cast<const unsigned*>(nullptr),
                                            static_cast<unsigned*>(nullptr), total);
  cub::DeviceScan::InclusiveSum(nullptr, scan_tmp, static_cast<const double*>(nullptr),
                                static_cast<double*>(nullptr), total);
  const size_t kd = up256(static_cast<size_t>(total) * 8), ki = up256(static_cast<size_t>(total) * 4);
  size_t off = 0;
  l.key_in = off; off += kd;
  l.key_out = off; off += kd;
  l.idx_in = off; off += ki;
  l.idx_out = off; off += ki;
  l.t0 = off; off += kd;
  l.t1 = off; off += kd;
  l.t2 = off; off += kd;
  l.part = off; off += up256(sizeof(Cand) * kTensorCandBlocks);
  l.temp = off;
  l.temp_bytes = sort_tmp > scan_tmp ? sort_tmp : scan_tmp;
  l.total = off + l.temp_bytes;
  return l;
}

}  // namespace

constexpr int kOscarMaxRow = 16384;

// ---- column second moments -------------------------------------------------------------
static int colsq_splits(long long n, int d, int sm_count) {
  const long long col_blocks = (d + 255) / 256;
  long long want = (4LL * sm_count + col_blocks - 1) / col_blocks;
  const long long by_rows = (n + 63) / 64;
  if (want > by_rows) want = by_rows;
  if (want > 256) want = 256;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}
size_t colsq_workspace_bytes(long long n, long long d, int sm_count) {
  return static_cast<size_t>(colsq_splits(n, static_cast<int>(d), sm_count)) * d * sizeof(double);
}
cudaError_t launch_colsq(const float* x, long long n, long long d, double alpha, double* out,
                         void* ws, int sm_count, cudaStream_t st) {
  if (d <= 0) return cudaSuccess;
  const int dd = static_cast<int>(d);
  const int splits = colsq_splits(n, dd, sm_count);
  const long long per = n > 0 ? (n + splits - 1) / splits : 1;
  const unsigned cb = static_cast<unsigned>((dd + 255) / 256);
  colsq_partial<<<dim3(cb, splits), 256, 0, st>>>(x, n, dd, per, static_cast<double*>(ws));
  colsq_reduce<<<cb, 256, 0, st>>>(static_cast<const double*>(ws), splits, dd, alpha, out);
  return count_launch(2);
}

// ---- objective / a_eff pass --------------------------------------------------------------
static long long pass_rows_per_chunk(long long n, long long groups, int sm_count) {
  // enough warps to fill the machine (~32 per SM), at most 64 rows per warp
  long long chunks = (32LL * sm_count + groups - 1) / groups;
  if (chunks < 1) chunks = 1;
  long long rows = (n + chunks - 1) / chunks;
  if (rows < 1) rows = 1;
  if (rows > 64) rows = 64;
  return rows;
}
size_t oscar_pass_workspace_bytes(long long n, long long d, long long g, int sm_count) {
  const long long groups = d / g;
  const long long rows = pass_rows_per_chunk(n, groups, sm_count);
  return static_cast<size_t>((n + rows - 1) / rows) * groups * sizeof(double);
}
// group_sq[d / g]; a_eff[d] (accumulated into: caller zeroes) or null.
cudaError_t launch_oscar_pass(const float* W, long long n, long long d, long long g, const double* s,
                              double* group_sq, double* a_eff, void* ws, int sm_count,
                              cudaStream_t st) {
  if (n <= 0 || d <= 0) return cudaSuccess;
  const long long groups = d / g;
  const long long rows = pass_rows_per_chunk(n, groups, sm_count);
  const long long chunks = (n + rows - 1) / rows;
  const long long warps = chunks * groups;
  const unsigned grid = static_cast<unsigned>((warps * 32 + 255) / 256);
  oscar_group_pass<<<grid, 256, 0, st>>>(W, n, static_cast<int>(d), static_cast<int>(g), s, rows,
                                         static_cast<double*>(ws), a_eff);
  oscar_pass_reduce<<<static_cast<unsigned>((groups + 255) / 256), 256, 0, st>>>(
      static_cast<const double*>(ws), chunks, static_cast<int>(groups), group_sq);
  return count_launch(2);
}

// ---- clip bounds ---------------------------------------------------------------------------
size_t oscar_clip_workspace_bytes(long long n, long long d, long long g) {
  return g == n * d && (n > 1 || d > kOscarMaxRow) ? tensor_layout(n * d).total : 0;
}
// g == d: one bound per row (mass[0]); g in 32..256: per block (mass[d / g]); g == n * d: one
// bound for the tensor (mass[0], ws required unless it is a single short row).
cudaError_t launch_oscar_clip(const float* W, long long n, long long d, long long g,
                              const double* s, const double* m, const double* mass_dev,
                              double mass0, int qmax, double* bound, void* ws, int sm_count,
                              cudaStream_t st) {
  if (n <= 0 || d <= 0) return cudaSuccess;
  const double qq = static_cast<double>(qmax) * static_cast<double>(qmax);
  if (g == n * d && (n > 1 || d > kOscarMaxRow)) {
    if (!ws) return cudaErrorInvalidValue;
    const long long total = n * d;
    const TensorLayout l = tensor_layout(total);
    unsigned char* p = static_cast<unsigned char*>(ws);
    double* kin = reinterpret_cast<double*>(p + l.key_in);
    double* kout = reinterpret_cast<double*>(p + l.key_out);
    unsigned* iin = reinterpret_cast<unsigned*>(p + l.idx_in);
    unsigned* iout = reinterpret_cast<unsigned*>(p + l.idx_out);
    double* t0 = reinterpret_cast<double*>(p + l.t0);
    double* t1 = reinterpret_cast<double*>(p + l.t1);
    double* t2 = reinterpret_cast<double*>(p + l.t2);
    Cand* part = reinterpret_cast<Cand*>(p + l.part);
    size_t tmp = l.temp_bytes;
    const unsigned grid = static_cast<unsigned>(sm_count * 8);
    oscar_tensor_keys<<<grid, 256, 0, st>>>(W, total, static_cast<int>(d), s, kin, iin);
    cudaError_t e = cub::DeviceRadixSort::SortPairsDescending(p + l.temp, tmp, kin, kout, iin, iout,
                                                              total, 0, 64, st);
    if (e != cudaSuccess) return e;
    oscar_tensor_terms<<<grid, 256, 0, st>>>(kout, iout, total, static_cast<int>(d), m, t0, t1, t2);
    // the scans run in place over the term arrays; kin is free again and not needed
    for (double* t : {t0, t1, t2}) {
      tmp = l.temp_bytes;
      e = cub::DeviceScan::InclusiveSum(p + l.temp, tmp, t, t, total, st);
      if (e != cudaSuccess) return e;
    }
    oscar_tensor_cands<<<kTensorCandBlocks, 256, 0, st>>>(kout, total, t0, t1, t2, mass0, qq, part);
    oscar_tensor_pick<<<1, 1, 0, st>>>(part, kTensorCandBlocks, bound);
    return count_launch(8);
  }
  if (g == d || g == n * d) {
    if (d > kOscarMaxRow) return cudaErrorInvalidValue;
    const int seg = pow2_ge(d) < 32 ? 32 : pow2_ge(d);
    const int threads = seg >= 1024 ? 1024 : seg;
    const size_t smem = static_cast<size_t>(seg) * 12;
    static PerDevice attr_done;
    if (!attr_done.done()) {
      cudaError_t e = cudaFuncSetAttribute(oscar_clip_rows, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kOscarMaxRow * 12);
      if (e != cudaSuccess) return e;
      attr_done.set();
    }
    long long grid = n;
    if (grid > static_cast<long long>(sm_count) * 4) grid = static_cast<long long>(sm_count) * 4;
    oscar_clip_rows<<<static_cast<unsigned>(grid), threads, smem, st>>>(
        W, n, static_cast<int>(d), seg, s, m, mass0, qq, bound);
    return count_launch();
  }
  if (g != 32 && g != 64 && g != 128 && g != 256) return cudaErrorInvalidValue;
  const long long total = n * d;
  long long grid = (total + kOscarTile - 1) / kOscarTile;
  if (grid > static_cast<long long>(sm_count) * 8) grid = static_cast<long long>(sm_count) * 8;
  oscar_clip_blocks<<<static_cast<unsigned>(grid), 256, 0, st>>>(W, total, static_cast<int>(d),
                                                                static_cast<int>(g), s, m, mass_dev,
                                                                qq, bound);
  return count_launch();
}

cudaError_t launch_oscar_scale(const double* bound, long long n, int qmax, int blockwise,
                               double* scale, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  oscar_scale<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
      bound, n, static_cast<double>(qmax), blockwise, scale);
  return count_launch();
}

cudaError_t launch_oscar_quantize(const float* W, long long n, long long d, long long glen,
                                  const double* s, const double* scale, int bits, int8_t* q,
                                  int sm_count, cudaStream_t st) {
  const long long total = n * d;
  if (total <= 0) return cudaSuccess;
  const QRange qr = qrange(bits, true);
  long long grid = (total + 255) / 256;
  if (grid > static_cast<long long>(sm_count) * 16) grid = static_cast<long long>(sm_count) * 16;
  oscar_quantize<<<static_cast<unsigned>(grid), 256, 0, st>>>(W, total, static_cast<int>(d), glen, s,
                                                            scale, qr.lo, qr.hi, q);
  return count_launch();
}

}  // namespace aeqb
