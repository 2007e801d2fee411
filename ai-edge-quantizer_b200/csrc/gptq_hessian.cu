// GPTQ, Hessian side — replaces
//   gptq.calibrate's   H = (2 / num_samples) * X^T X          (algorithms/uniform_quantize/gptq.py:100-106)
//   gptq._prepare_hessian_inverse                              (gptq.py:111-128)
//     diag 0 -> 1, + damp * mean(diag); L = cholesky(H) in float64 (np.linalg.cholesky);
//     L^-1 in float32 (scipy.linalg.lapack.strtri casts to single);
//     H^-1 = einsum("ji,jk->ik", L^-1, L^-1) in float32
//   qsv_utils._gptq_merge_hessian's weighted mean               (utils/qsv_utils.py:71-88)
//
// dtype flow mirrors NumPy 2: the GEMM accumulates in fp32 (X is
// fp32, `x.T.dot(x)` is sgemm) and the float64 scalar 2/num_samples promotes the
// product, so H is float64 holding fp32-rounded sums.
//
// Kernels (SIMT; the tensor-core version of the two dense contractions is the
// next step):
//   xtx_tile      128x128 output tile per CTA, 16-deep k-slabs of X staged in shared
//                 memory, 8x8 register tile per thread, upper-triangular tiles only,
//                 optional deterministic split over tokens (fixed-order reduce).
//   chol_panel / chol_syrk   right-looking blocked Cholesky in fp64, block 32.
//   trtri_diag / trtri_pair_gemm  lower-triangular inverse in fp32: 64 x 64 diagonal blocks,
//                 then recursive doubling with two batched SGEMMs per level.
#include "aeqb_common.cuh"
#include "aeqb_kernels.h"

#include <stdlib.h>

namespace aeqb {

namespace {

// ------------------------------------------------------------------ X^T X
constexpr int XT = 128;  // output tile edge
constexpr int XK = 16;   // tokens per slab

// Upper-triangular tile index -> (bi, bj), bi <= bj.
__device__ __forceinline__ void tri_index(int t, int nb, int& bi, int& bj) {
  int row = 0, left = t;
  while (left >= nb - row) {
    left -= nb - row;
    ++row;
  }
  bi = row;
  bj = row + left;
}

// part != nullptr: raw fp32 partial sums for split `blockIdx.y` go to part[split][K][K]
// (upper tiles only); otherwise out = alpha * sum written to both triangles.
template <typename OutT>
__global__ void __launch_bounds__(256)
    xtx_tile(const float* __restrict__ X, long long T, int K, double alpha, OutT* __restrict__ out,
             float* __restrict__ part, long long t_per_split, const int* __restrict__ gate) {
  __shared__ __align__(16) float As[2][XK][XT];
  __shared__ __align__(16) float Bs[2][XK][XT];
  // gate: set by the tensor-core path when its input held non-finite values; only then does this
  // launch (queued right behind it) do any work.
  if (gate != nullptr && *gate == 0) return;
  const int nb = (K + XT - 1) / XT;
  int bi, bj;
  tri_index(blockIdx.x, nb, bi, bj);
  const int i0 = bi * XT, j0 = bj * XT;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 8 x 8 outputs (2 x 4 + 2 x 4)
  const long long t_begin = static_cast<long long>(blockIdx.y) * t_per_split;
  long long t_end = t_begin + t_per_split;
  if (t_end > T) t_end = T;
  const bool vec = (K % 4 == 0) && (reinterpret_cast<uintptr_t>(X) % 16 == 0);

  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.0f;

  // each thread stages two float4 of A and two of B per slab: slab row = (tid*2+u)/32, col4 = (tid*2+u)%32
  auto load_slab = [&](int buf, long long t0) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int idx = tid * 2 + u;
      const int r = idx >> 5, c4 = (idx & 31) * 4;
      const long long t = t0 + r;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (t < t_end) {
        const float* row = X + t * K;
        if (vec) {
          if (i0 + c4 < K) a = __ldg(reinterpret_cast<const float4*>(row + i0 + c4));
          if (j0 + c4 < K) b = __ldg(reinterpret_cast<const float4*>(row + j0 + c4));
        } else {
          float ta[4] = {0.f, 0.f, 0.f, 0.f}, tb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (i0 + c4 + e < K) ta[e] = row[i0 + c4 + e];
            if (j0 + c4 + e < K) tb[e] = row[j0 + c4 + e];
          }
          a = make_float4(ta[0], ta[1], ta[2], ta[3]);
          b = make_float4(tb[0], tb[1], tb[2], tb[3]);
        }
      }
      *reinterpret_cast<float4*>(&As[buf][r][c4]) = a;
      *reinterpret_cast<float4*>(&Bs[buf][r][c4]) = b;
    }
  };

  int buf = 0;
  if (t_begin < t_end) load_slab(0, t_begin);
  __syncthreads();
  for (long long t0 = t_begin; t0 < t_end; t0 += XK) {
    if (t0 + XK < t_end) load_slab(buf ^ 1, t0 + XK);
#pragma unroll
    for (int k = 0; k < XK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4 + 64]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4 + 64]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
    buf ^= 1;
  }

#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int i = i0 + ty * 4 + (a & 3) + (a >> 2) * 64;
    if (i >= K) continue;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int j = j0 + tx * 4 + (b & 3) + (b >> 2) * 64;
      if (j >= K) continue;
      if (part) {
        part[(static_cast<long long>(blockIdx.y) * K + i) * K + j] = acc[a][b];
      } else {
        const OutT v = static_cast<OutT>(alpha * static_cast<double>(acc[a][b]));
        out[static_cast<long long>(i) * K + j] = v;
        if (bi != bj) out[static_cast<long long>(j) * K + i] = v;
      }
    }
  }
}

// Fixed-order sum of the token splits (deterministic), then alpha, both triangles.
template <typename OutT>
__global__ void __launch_bounds__(256)
    xtx_reduce(const float* __restrict__ part, int splits, int K, double alpha, OutT* __restrict__ out) {
  const long long n = static_cast<long long>(K) * K;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(e / K), j = static_cast<int>(e % K);
    if (j / XT < i / XT) continue;  // lower tiles are mirrored from the upper ones
    float s = 0.0f;
    for (int p = 0; p < splits; ++p) s += part[static_cast<long long>(p) * n + e];
    const OutT v = static_cast<OutT>(alpha * static_cast<double>(s));
    out[e] = v;
    if (j / XT != i / XT) out[static_cast<long long>(j) * K + i] = v;
  }
}

// ------------------------------------------------------------------ damping (gptq.py:115-117)
// diag = where(diag != 0, diag, 1); diag += damp * mean(diag).  One CTA; fp64 tree mean.
__global__ void __launch_bounds__(1024)
    damp_diagonal(double* __restrict__ A, int K, double damp) {
  __shared__ double s_part[32];
  __shared__ double s_mean;
  double acc = 0.0;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const double d = A[static_cast<long long>(i) * K + i];
    acc += (d != 0.0) ? d : 1.0;  // NaN != 0 keeps NaN, like np.where(diag, diag, 1.0)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += s_part[w];
    s_mean = t / static_cast<double>(K);
  }
  __syncthreads();
  const double add = damp * s_mean;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const long long p = static_cast<long long>(i) * K + i;
    const double d = A[p];
    A[p] = ((d != 0.0) ? d : 1.0) + add;
  }
}

__global__ void __launch_bounds__(256)
    copy_f64(const double* __restrict__ a, double* __restrict__ b, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    b[i] = a[i];
}

__global__ void __launch_bounds__(256)
    copy_diag_f64(const double* __restrict__ a, double* __restrict__ b, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K) b[static_cast<long long>(i) * K + i] = a[static_cast<long long>(i) * K + i];
}

// ------------------------------------------------------------------ Cholesky (fp64, lower, in place)
constexpr int CB = 32;    // block size
constexpr int CR = 128;   // panel rows per CTA

// Panel step at column block j: every CTA factors the CB x CB diagonal block in shared memory
// (redundantly: it is tiny) and solves its CR rows of the panel  X * L_jj^T = A[rows, j:j+CB];
// CTA 0 also stores the factored diagonal block.  info != 0 flags a non-positive pivot.
__global__ void __launch_bounds__(CR)
    chol_panel(double* __restrict__ A, int K, int j, int* __restrict__ info) {
  __shared__ double D[CB][CB + 1];
  const int tid = threadIdx.x;
  const int nb = min(CB, K - j);
  for (int e = tid; e < CB * CB; e += CR) {
    const int r = e / CB, c = e % CB;
    D[r][c] = (r < nb && c <= r) ? A[static_cast<long long>(j + r) * K + j + c] : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();
  for (int k = 0; k < nb; ++k) {  // unblocked right-looking factorisation of D
    if (tid == 0) {
      const double p = D[k][k];
      if (!(p > 0.0)) atomicExch(info, j + k + 1);
      D[k][k] = sqrt(p);
    }
    __syncthreads();
    if (tid > k && tid < nb) D[tid][k] /= D[k][k];
    __syncthreads();
    // trailing update: thread t owns row t (t > k), columns k+1..t
    if (tid > k && tid < nb) {
      const double l = D[tid][k];
      for (int c = k + 1; c <= tid; ++c) D[tid][c] -= l * D[c][k];
    }
    __syncthreads();
  }
  if (blockIdx.x == 0) {
    for (int e = tid; e < CB * CB; e += CR) {
      const int r = e / CB, c = e % CB;
      if (r < nb && c < nb) A[static_cast<long long>(j + r) * K + j + c] = (c <= r) ? D[r][c] : 0.0;
    }
  }
  const int row = j + nb + blockIdx.x * CR + tid;
  if (row < K) {
    double x[CB];
    double* a = A + static_cast<long long>(row) * K + j;
#pragma unroll
    for (int c = 0; c < CB; ++c) x[c] = c < nb ? a[c] : 0.0;
#pragma unroll
    for (int c = 0; c < CB; ++c) {
      if (c < nb) {
        double s = x[c];
#pragma unroll
        for (int m = 0; m < CB; ++m)
          if (m < c) s -= x[m] * D[c][m];
        x[c] = s / D[c][c];
      }
    }
#pragma unroll
    for (int c = 0; c < CB; ++c)
      if (c < nb) a[c] = x[c];
  }
}

// Trailing update A[r, c] -= sum_k P[r, k] P[c, k] for the lower tiles of the trailing matrix,
// P = A[j+CB:, j:j+CB].  64 x 64 tile per CTA, 4 x 4 per thread.
__global__ void __launch_bounds__(256)
    chol_syrk(double* __restrict__ A, int K, int j) {
  __shared__ double Pr[64][CB + 1];
  __shared__ double Pc[64][CB + 1];
  const int base = j + CB;
  const int nt = (K - base + 63) / 64;
  // lower-triangular tile index: blockIdx.x -> (tr >= tc)
  int tr = 0, left = blockIdx.x;
  while (left > tr) {
    left -= tr + 1;
    ++tr;
  }
  const int tc = left;
  (void)nt;
  const int r0 = base + tr * 64, c0 = base + tc * 64;
  const int tid = threadIdx.x;
  for (int e = tid; e < 64 * CB; e += 256) {
    const int r = e / CB, k = e % CB;
    Pr[r][k] = (r0 + r < K) ? A[static_cast<long long>(r0 + r) * K + j + k] : 0.0;
    Pc[r][k] = (c0 + r < K) ? A[static_cast<long long>(c0 + r) * K + j + k] : 0.0;
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  double acc[4][4] = {};
#pragma unroll 8
  for (int k = 0; k < CB; ++k) {
    double a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a[u] = Pr[ty + 16 * u][k];
      b[u] = Pc[tx + 16 * u][k];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r = r0 + ty + 16 * u;
    if (r >= K) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int c = c0 + tx + 16 * v;
      if (c <= r) A[static_cast<long long>(r) * K + c] -= acc[u][v];
    }
  }
}

// ------------------------------------------------------------------ Cholesky, two-level variant
// For large K the rank-32 trailing update is bound by the read-modify-write traffic of the
// trailing matrix (K = 11008: 344 passes over up to 970 MB).  This variant updates the trailing
// matrix once per outer block of CNB = 128 columns (a quarter of the passes) and factors the
// 32-column panels of an outer block left-looking: a panel first applies the pending rank-(j - J)
// update of the earlier panels of its outer block to its own 32 columns, then factors.  Measured
// (hessian_inverse, K = 4096 / 11008): single-level 10.4 / 91.4 ms, two-level 8.6 / 67.9 ms, so
// it is the default (kCholTwoLevelMinK = 0; AEQB_CHOL_TWO_LEVEL_MIN_K selects the other one).
// A 128 x 128 / 8 x 8-per-thread trailing update was measured slower than the 64 x 64 one
// (10.3 / 84.9 ms): one CTA per SM with single-buffered slabs exposes the global-load latency.
constexpr int CNB = 128;   // outer block width
constexpr int kCholTwoLevelMinK = 0;
constexpr int kCholDmmaMinK = 0;   // default: every size takes chol_dmma.cu
constexpr int kCholPanel2Smem = CR * (CB + 1) * 8;  // dynamic: the row staging tile

// Panel at columns [j, j + CB), outer block starting at column J <= j (w = j - J pending columns).
// Every CTA rebuilds the CB x CB diagonal block in shared memory (pending update included), warp 0
// factors it in registers (lane r = row r; pivots and multipliers travel by shuffle), inverts the
// factor, and each thread then forms its row of the panel as  (A[row, j:j+CB] - pending) * L_jj^-T
// -- a dependency-free 32 x 32 product instead of a forward substitution.
// 1 / sqrt(p) in float64 from the fp32 estimate and two Newton steps (each step squares the
// relative error: 2^-22 -> 2^-43 -> below 2^-53), a short dependent chain instead of the
// software sqrt + divide (~600 cycles per pivot on the factorisation's critical path).
__device__ __forceinline__ double rsqrt_f64(double p) {
  double r = static_cast<double>(rsqrtf(static_cast<float>(p)));
  r = r * fma(-0.5 * p, r * r, 1.5);
  r = r * fma(-0.5 * p, r * r, 1.5);
  return r;
}

__global__ void __launch_bounds__(CR)
    chol_panel2(double* __restrict__ A, int K, int j, int J, int* __restrict__ info) {
  __shared__ double D[CB][CB + 1];          // diagonal block, then its Cholesky factor
  __shared__ double Li[CB][CB + 1];         // inverse of the factor (lower triangular)
  __shared__ double Pd[CB][CNB - CB + 1];   // pending columns of the diagonal block's rows
  extern __shared__ double chol_rt_raw[];   // staging tile: 32 columns of this CTA's 128 rows
  double (*Rt)[CB + 1] = reinterpret_cast<double (*)[CB + 1]>(chol_rt_raw);
  __shared__ double rinv[CB];               // reciprocals of the factor's diagonal
  const int tid = threadIdx.x;
  const int nb = min(CB, K - j);
  const int w = j - J;
  for (int e = tid; e < CB * CB; e += CR) {
    const int r = e / CB, c = e % CB;
    D[r][c] = (r < nb && c <= r) ? A[static_cast<long long>(j + r) * K + j + c] : (r == c ? 1.0 : 0.0);
  }
  for (int e = tid; e < CB * w; e += CR) {
    const int r = e / w, k = e % w;
    Pd[r][k] = r < nb ? A[static_cast<long long>(j + r) * K + J + k] : 0.0;
  }
  __syncthreads();
  if (w > 0) {  // D -= Pd Pd^T on the lower triangle
    for (int e = tid; e < CB * CB; e += CR) {
      const int r = e / CB, c = e % CB;
      if (c <= r && r < nb) {
        double s = 0.0;
        for (int k = 0; k < w; ++k) s = fma(Pd[r][k], Pd[c][k], s);
        D[r][c] -= s;
      }
    }
    __syncthreads();
  }
  if (tid < 32) {
    const int lane = tid;
    double a[CB];
#pragma unroll
    for (int c = 0; c < CB; ++c) a[c] = D[lane][c];
#pragma unroll
    for (int k = 0; k < CB; ++k) {
      const double p = __shfl_sync(0xffffffffu, a[k], k);
      if (lane == k && k < nb && !(p > 0.0)) atomicExch(info, j + k + 1);
      const double rs = rsqrt_f64(p);
      double sq = p * rs;
      sq = fma(0.5 * rs, fma(-sq, sq, p), sq);  // one correction: sqrt(p) to the last bit or so
      if (lane == k) {
        a[k] = sq;
        rinv[k] = rs;
      }
      if (lane > k) a[k] = a[k] * rs;
#pragma unroll
      for (int c = k + 1; c < CB; ++c) {
        const double lc = __shfl_sync(0xffffffffu, a[k], c);
        if (lane >= c) a[c] = fma(-a[k], lc, a[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < CB; ++c) D[lane][c] = (c <= lane) ? a[c] : 0.0;
    __syncwarp();
    // column `lane` of the inverse by forward substitution, two partial sums per row
    double y[CB];
#pragma unroll
    for (int r = 0; r < CB; ++r) {
      double s0 = (r == lane) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
      for (int m = 0; m + 1 < r; m += 2) {
        s0 = fma(-D[r][m], y[m], s0);
        s1 = fma(-D[r][m + 1], y[m + 1], s1);
      }
      if (r & 1) s0 = fma(-D[r][r - 1], y[r - 1], s0);
      y[r] = (r < lane) ? 0.0 : (s0 + s1) * rinv[r];
    }
#pragma unroll
    for (int r = 0; r < CB; ++r) Li[r][lane] = y[r];
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    for (int e = tid; e < CB * CB; e += CR) {
      const int r = e / CB, c = e % CB;
      if (r < nb && c < nb) A[static_cast<long long>(j + r) * K + j + c] = D[r][c];
    }
  }
  // This CTA's CR rows below the diagonal block.  Global memory is only touched through the
  // staging tile, 32 consecutive columns of a row per warp request (a thread per row reading its
  // own row would scatter every request over 32 cache lines).
  const long long row0 = static_cast<long long>(j) + nb + static_cast<long long>(blockIdx.x) * CR;
  const int lane = tid & 31, wrp = tid >> 5;
  auto stage_in = [&](int col0, int ncols) {  // Rt[r][c] = A[row0 + r][col0 + c]
    __syncthreads();
    for (int r = wrp; r < CR; r += CR / 32) {
      const long long gr = row0 + r;
      Rt[r][lane] = (gr < K && lane < ncols) ? A[gr * K + col0 + lane] : 0.0;
    }
    __syncthreads();
  };
  double x[CB];
  stage_in(j, nb);
#pragma unroll
  for (int c = 0; c < CB; ++c) x[c] = Rt[tid][c];
  for (int k0 = 0; k0 < w; k0 += CB) {
    stage_in(J + k0, CB);
#pragma unroll 4
    for (int k = 0; k < CB; ++k) {
      const double v = Rt[tid][k];
#pragma unroll
      for (int c = 0; c < CB; ++c) x[c] = fma(-v, Pd[c][k0 + k], x[c]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < CB; ++c) {  // out[c] = sum_{m <= c} x[m] * Li[c][m]
    double sacc = 0.0;
#pragma unroll
    for (int m = 0; m <= c; ++m) sacc = fma(x[m], Li[c][m], sacc);
    Rt[tid][c] = sacc;
  }
  __syncthreads();
  for (int r = wrp; r < CR; r += CR / 32) {
    const long long gr = row0 + r;
    if (gr < K && lane < nb) A[gr * K + j + lane] = Rt[r][lane];
  }
}

// Trailing update A[r, c] -= sum_{k < depth} P[r, k] P[c, k] for the lower tiles of the trailing
// matrix starting at row / column `base`, P = A[base:, J:J+depth]; 32-deep slabs.
__global__ void __launch_bounds__(256)
    chol_syrk2(double* __restrict__ A, int K, int J, int depth, int base) {
  __shared__ double Pr[64][CB + 1];
  __shared__ double Pc[64][CB + 1];
  int tr = 0, left = blockIdx.x;
  while (left > tr) {
    left -= tr + 1;
    ++tr;
  }
  const int tc = left;
  const int r0 = base + tr * 64, c0 = base + tc * 64;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  double acc[4][4] = {};
  for (int k0 = 0; k0 < depth; k0 += CB) {
    if (k0 > 0) __syncthreads();
    for (int e = tid; e < 64 * CB; e += 256) {
      const int r = e / CB, k = e % CB;
      const bool in = k0 + k < depth;
      Pr[r][k] = (in && r0 + r < K) ? A[static_cast<long long>(r0 + r) * K + J + k0 + k] : 0.0;
      Pc[r][k] = (in && c0 + r < K) ? A[static_cast<long long>(c0 + r) * K + J + k0 + k] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < CB; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = Pr[ty + 16 * u][k];
        b[u] = Pc[tx + 16 * u][k];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r = r0 + ty + 16 * u;
    if (r >= K) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int c = c0 + tx + 16 * v;
      if (c <= r) A[static_cast<long long>(r) * K + c] -= acc[u][v];
    }
  }
}

// ------------------------------------------------------------------ triangular inverse (fp32, lower)
constexpr int TB = 64;

// L32 = float(L64) on the lower triangle, 0 above (strtri's float32 cast of the factor).
__global__ void __launch_bounds__(256)
    lower_to_f32(const double* __restrict__ A, float* __restrict__ L, int K) {
  const long long n = static_cast<long long>(K) * K;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(e / K), c = static_cast<int>(e % K);
    L[e] = c <= r ? static_cast<float>(A[e]) : 0.0f;
  }
}

// Inverse of each TB x TB diagonal block (forward substitution, one thread per column),
// written into the diagonal blocks of Y; the launcher zeroes Y first (the upper triangle stays 0).
__global__ void __launch_bounds__(TB)
    trtri_diag(const float* __restrict__ L, float* __restrict__ Y, int K) {
  __shared__ float D[TB][TB + 1];
  __shared__ float R[TB][TB + 1];
  const int b0 = blockIdx.x * TB;
  const int nb = min(TB, K - b0);
  const int c = threadIdx.x;
  for (int e = c; e < TB * TB; e += TB) {
    const int r = e / TB, k = e % TB;
    D[r][k] = (r < nb && k < nb) ? L[static_cast<long long>(b0 + r) * K + b0 + k] : (r == k ? 1.0f : 0.0f);
  }
  __syncthreads();
  // column c of the inverse: y[r] = (delta_rc - sum_{m<r} D[r][m] y[m]) / D[r][r]
  for (int r = 0; r < TB; ++r) {
    float s = (r == c) ? 1.0f : 0.0f;
    for (int m = c; m < r; ++m) s = fmaf(-D[r][m], R[m][c], s);
    R[r][c] = (r < c) ? 0.0f : __fdiv_rn(s, D[r][r]);
  }
  __syncthreads();
  for (int e = c; e < TB * TB; e += TB) {
    const int r = e / TB, k = e % TB;
    if (r < nb && k < nb) Y[static_cast<long long>(b0 + r) * K + b0 + k] = R[r][k];
  }
}

// Recursive doubling above the diagonal blocks: with Y holding the inverses of all b x b
// diagonal blocks, the 2b x 2b block [[A, 0], [B, C]] has inverse [[A^-1, 0], [-C^-1 B A^-1, C^-1]].
// One level is two batched SGEMMs over all block pairs p (rows / columns r0 = 2 b p):
//   mode 0:  T      = B * A^-1      B = L[r0+b.., r0..r0+b),  A^-1 = Y[r0.., r0..)   (lower tri.)
//   mode 1:  Y_off  = -C^-1 * T     C^-1 = Y[r0+b.., r0+b..)  (lower triangular)
// log2(K / 64) levels of fully parallel GEMMs replace the former per-slice recurrence, whose
// first block column was a chain of 2016 dependent 64 x 64 products (10.9 ms at K = 4096).
// 128 x 128 tile, 16-deep slabs, 8 x 8 per thread; the triangular operand trims the k range.
constexpr int GT = 128, GK = 16;
__global__ void __launch_bounds__(256)
    trtri_pair_gemm(const float* __restrict__ Lm, const float* __restrict__ Ym,
                    const float* __restrict__ Tin, float* __restrict__ Out, int K, int b, int mode) {
  __shared__ __align__(16) float As[2][GK][GT + 4];
  __shared__ __align__(16) float Bs[2][GK][GT];
  const long long r0 = static_cast<long long>(blockIdx.z) * 2 * b;
  const int Mp = static_cast<int>(min(static_cast<long long>(b), K - r0 - b));
  if (Mp <= 0) return;
  const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
  if (i0 >= Mp || j0 >= b) return;
  const float* A;
  const float* B;
  float* C = Out + (r0 + b) * K + r0;
  int Kd, k_lo, k_hi;
  float alpha;
  if (mode == 0) {
    A = Lm + (r0 + b) * K + r0;   // [Mp, b]
    B = Ym + r0 * K + r0;         // [b, b] lower triangular: rows k >= column j
    Kd = b; k_lo = j0; k_hi = b; alpha = 1.0f;
  } else {
    A = Ym + (r0 + b) * K + (r0 + b);  // [Mp, Mp] lower triangular: columns k <= row i
    B = Tin + (r0 + b) * K + r0;       // [Mp, b]
    Kd = Mp; k_lo = 0; k_hi = min(i0 + GT, Mp); alpha = -1.0f;
  }
  const int N = b;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const bool vec = (K % 4 == 0) && (reinterpret_cast<uintptr_t>(Lm) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(Ym) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(Tin) % 16 == 0);
  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[a][c] = 0.0f;

  auto load_slab = [&](int buf, int k0) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int idx = tid * 2 + u;
      {  // A tile: 128 rows x 16 k, k contiguous in memory -> transposed into As[k][row]
        const int r = idx >> 2, k4 = (idx & 3) * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (i0 + r < Mp) {
          const float* src = A + static_cast<long long>(i0 + r) * K + k0 + k4;
          if (vec && k0 + k4 + 3 < Kd) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(src));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (k0 + k4 + e < Kd) v[e] = src[e];
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) As[buf][k4 + e][r] = v[e];
      }
      {  // B tile: 16 k x 128 columns, columns contiguous
        const int kr = idx >> 5, c4 = (idx & 31) * 4;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + kr < Kd) {
          const float* src = B + static_cast<long long>(k0 + kr) * K + j0 + c4;
          if (vec && j0 + c4 + 3 < N) {
            t = __ldg(reinterpret_cast<const float4*>(src));
          } else {
            float w[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (j0 + c4 + e < N) w[e] = src[e];
            t = make_float4(w[0], w[1], w[2], w[3]);
          }
        }
        *reinterpret_cast<float4*>(&Bs[buf][kr][c4]) = t;
      }
    }
  };

  int buf = 0;
  if (k_lo < k_hi) load_slab(0, k_lo);
  __syncthreads();
  for (int k0 = k_lo; k0 < k_hi; k0 += GK) {
    if (k0 + GK < k_hi) load_slab(buf ^ 1, k0 + GK);
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4 + 64]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4 + 64]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
    }
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int i = i0 + ty * 4 + (a & 3) + (a >> 2) * 64;
    if (i >= Mp) continue;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = j0 + tx * 4 + (c & 3) + (c >> 2) * 64;
      if (j < N) C[static_cast<long long>(i) * K + j] = alpha * acc[a][c];
    }
  }
}

// out = (a * wa + b * wb) / (wa + wb) elementwise in fp64 (qsv_utils.py:84-88).
__global__ void __launch_bounds__(256)
    weighted_mean_f64(const double* __restrict__ a, double wa, const double* __restrict__ b, double wb,
                      double* __restrict__ out, long long n) {
  const double tot = wa + wb;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = __ddiv_rn(__dadd_rn(__dmul_rn(a[i], wa), __dmul_rn(b[i], wb)), tot);  // no FMA contraction
}

// AEQB_XTX_SIMT=1 keeps the contractions on the SIMT kernel (A/B timing, debugging);
// AEQB_XTX_TC_MIN_GFLOP overrides the size (2*T*K*K, default 8 GFLOP) from which the tcgen05
// path is taken: below it the split/transpose + tile-quantisation overheads lose to SIMT
// (measured: 4096 x 512 -> 52 us SIMT vs 142 us, 4100 x 1028 -> 259 us SIMT vs 152 us).
bool use_tensor_cores(long long T, long long K) {
  const char* e = getenv("AEQB_XTX_SIMT");
  if (e && e[0] == '1') return false;
  if (!xtx_tc_eligible(T, K)) return false;
  double min_gflop = 8.0;
  if (const char* m = getenv("AEQB_XTX_TC_MIN_GFLOP")) min_gflop = atof(m);
  return 2.0 * static_cast<double>(T) * static_cast<double>(K) * static_cast<double>(K) >= min_gflop * 1e9;
}

int xtx_splits(long long T, int K, int sm_count) {
  const int nb = (K + XT - 1) / XT;
  const long long tiles = static_cast<long long>(nb) * (nb + 1) / 2;
  long long want = (2LL * sm_count + tiles - 1) / tiles;
  const long long max_by_t = (T + 4 * XK - 1) / (4 * XK);
  if (want > max_by_t) want = max_by_t;
  if (want > 64) want = 64;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}

size_t xtx_split_workspace_bytes(long long T, long long K, int sm_count) {
  const int s = xtx_splits(T, static_cast<int>(K), sm_count);
  return s <= 1 ? 0 : static_cast<size_t>(s) * K * K * sizeof(float);
}

template <typename OutT>
cudaError_t xtx_impl(const float* x, long long T, long long K, double alpha, OutT* out, void* ws,
                     size_t ws_bytes, int sm_count, cudaStream_t st, int lower_tri = 0) {
  if (K <= 0) return cudaSuccess;
  const int k = static_cast<int>(K);
  const int nb = (k + XT - 1) / XT;
  const unsigned tiles = static_cast<unsigned>(static_cast<long long>(nb) * (nb + 1) / 2);
  // Large contractions: tcgen05 3xTF32 (xtx_tc.cu), then this file's SIMT kernel gated on the
  // "input was not finite" flag so that inf / NaN propagate exactly like sgemm's.
  if (ws && use_tensor_cores(T, K) && ws_bytes >= xtx_tc_workspace_bytes(T, K) &&
      reinterpret_cast<uintptr_t>(ws) % 16 == 0) {
    const int* gate = nullptr;
    cudaError_t e = launch_xtx_tc<OutT>(x, T, K, alpha, out, ws, sm_count, &gate, st, lower_tri);
    if (e == cudaSuccess) {
      xtx_tile<OutT><<<dim3(tiles, 1), 256, 0, st>>>(x, T, k, alpha, out, nullptr, T, gate);
      return count_launch();
    }
    if (e != cudaErrorNotSupported) return e;
    (void)cudaGetLastError();
  }
  const size_t split_need = xtx_split_workspace_bytes(T, K, sm_count);
  const int splits = (ws && split_need > 0 && ws_bytes >= split_need) ? xtx_splits(T, k, sm_count) : 1;
  if (splits <= 1) {
    xtx_tile<OutT><<<dim3(tiles, 1), 256, 0, st>>>(x, T, k, alpha, out, nullptr, T > 0 ? T : 1, nullptr);
    return count_launch();
  }
  long long per = (T + splits - 1) / splits;
  per = (per + XK - 1) / XK * XK;
  xtx_tile<OutT><<<dim3(tiles, splits), 256, 0, st>>>(x, T, k, alpha, out, static_cast<float*>(ws), per, nullptr);
  xtx_reduce<OutT><<<sm_count * 4, 256, 0, st>>>(static_cast<const float*>(ws), splits, k, alpha, out);
  return count_launch(2);
}

}  // namespace

size_t xtx_workspace_bytes(long long T, long long K, int sm_count) {
  if (K <= 0) return 0;
  if (use_tensor_cores(T, K)) return xtx_tc_workspace_bytes(T, K);
  return xtx_split_workspace_bytes(T, K, sm_count);
}

cudaError_t launch_xtx_f64(const float* x, long long T, long long K, double alpha, double* out,
                           void* ws, int sm_count, cudaStream_t st) {
  return xtx_impl<double>(x, T, K, alpha, out, ws, ws ? xtx_workspace_bytes(T, K, sm_count) : 0,
                          sm_count, st);
}

cudaError_t launch_xtx_f32(const float* x, long long T, long long K, double alpha, float* out,
                           void* ws, int sm_count, cudaStream_t st) {
  return xtx_impl<float>(x, T, K, alpha, out, ws, ws ? xtx_workspace_bytes(T, K, sm_count) : 0,
                         sm_count, st);
}

// ws: [ A (K*K doubles) | L32 (K*K floats) ] (dead by the time H^-1 = Y^T Y runs, so the same
// bytes, grown to what that contraction wants, are its workspace) | Y (K*K floats) | info (256 B)
static size_t hinv_scratch_bytes(long long K) {
  size_t a = static_cast<size_t>(K) * K * (sizeof(double) + sizeof(float));
  const size_t x = xtx_workspace_bytes(K, K, 148);
  if (x > a) a = x;
  return (a + 1023) / 1024 * 1024;
}
// ... | info (256 B) | pad to 256 | Linv of the current diagonal block (chol_dmma.cu)
static size_t hinv_linv_offset(long long K) {
  const size_t o = hinv_scratch_bytes(K) + static_cast<size_t>(K) * K * sizeof(float) + 256;
  return (o + 255) / 256 * 256;
}
size_t hessian_inverse_workspace_bytes(long long K) {
  return hinv_linv_offset(K) + cholesky_dmma_workspace_bytes();
}

cudaError_t launch_hessian_inverse(double* hessian, long long K, double damp, int mutate_diagonal,
                                   float* hinv, void* ws, int* info_out, int sm_count,
                                   cudaStream_t st) {
  if (K <= 0) return cudaSuccess;
  const int k = static_cast<int>(K);
  const long long n = K * K;
  unsigned char* p = static_cast<unsigned char*>(ws);
  const size_t scratch = hinv_scratch_bytes(K);
  double* A = reinterpret_cast<double*>(p);
  float* L32 = reinterpret_cast<float*>(p + n * sizeof(double));
  float* Y = reinterpret_cast<float*>(p + scratch);
  int* info = reinterpret_cast<int*>(p + scratch + n * sizeof(float));
  int launches = 0;
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  const unsigned cgrid = static_cast<unsigned>(sm_count * 8);
  copy_f64<<<cgrid, 256, 0, st>>>(hessian, A, n); ++launches;
  damp_diagonal<<<1, 1024, 0, st>>>(A, k, damp); ++launches;
  if (mutate_diagonal) {  // the reference leaves the damped diagonal in the caller's Hessian
    copy_diag_f64<<<(k + 255) / 256, 256, 0, st>>>(A, hessian, k); ++launches;
  }
  int two_level_min_k = kCholTwoLevelMinK;  // AEQB_CHOL_TWO_LEVEL_MIN_K: A/B runs and tests
  if (const char* e2 = getenv("AEQB_CHOL_TWO_LEVEL_MIN_K")) two_level_min_k = atoi(e2);
  int dmma_min_k = kCholDmmaMinK;  // AEQB_CHOL_DMMA_MIN_K: the DMMA / lookahead factorisation (chol_dmma.cu)
  if (const char* e3 = getenv("AEQB_CHOL_DMMA_MIN_K")) dmma_min_k = atoi(e3);
  if (k >= dmma_min_k) {
    e = launch_cholesky_dmma(A, k, reinterpret_cast<double*>(p + hinv_linv_offset(K)), info, st, &launches);
    if (e != cudaSuccess) return e;
  } else if (k >= two_level_min_k) {
    static PerDevice panel2_configured;
    if (!panel2_configured.done()) {
      e = cudaFuncSetAttribute(chol_panel2, cudaFuncAttributeMaxDynamicSharedMemorySize, kCholPanel2Smem);
      if (e != cudaSuccess) return e;
      panel2_configured.set();
    }
    for (int J = 0; J < k; J += CNB) {
      const int jend = J + CNB < k ? J + CNB : k;
      for (int j = J; j < jend; j += CB) {
        const int below = k - j - CB;
        const unsigned pgrid = below > 0 ? static_cast<unsigned>((below + CR - 1) / CR) : 1u;
        chol_panel2<<<pgrid, CR, kCholPanel2Smem, st>>>(A, k, j, J, info); ++launches;
      }
      const int below = k - jend;
      if (below > 0) {
        const int nt = (below + 63) / 64;
        chol_syrk2<<<static_cast<unsigned>(nt * (nt + 1) / 2), 256, 0, st>>>(A, k, J, jend - J, jend);
        ++launches;
      }
    }
  } else {
    for (int j = 0; j < k; j += CB) {
      const int below = k - j - CB;
      const unsigned pgrid = below > 0 ? static_cast<unsigned>((below + CR - 1) / CR) : 1u;
      chol_panel<<<pgrid, CR, 0, st>>>(A, k, j, info); ++launches;
      if (below > 0) {
        const int nt = (below + 63) / 64;
        chol_syrk<<<static_cast<unsigned>(nt * (nt + 1) / 2), 256, 0, st>>>(A, k, j); ++launches;
      }
    }
  }
  lower_to_f32<<<cgrid, 256, 0, st>>>(A, L32, k); ++launches;
  const int nblk = (k + TB - 1) / TB;
  e = cudaMemsetAsync(Y, 0, static_cast<size_t>(n) * sizeof(float), st);  // the upper triangle stays 0
  if (e != cudaSuccess) return e;
  trtri_diag<<<nblk, TB, 0, st>>>(L32, Y, k); ++launches;
  float* Tbuf = reinterpret_cast<float*>(A);  // the fp64 factor is dead once L32 exists
  for (long long b = TB; b < K; b *= 2) {
    const unsigned pairs = static_cast<unsigned>((K + 2 * b - 1) / (2 * b));
    const unsigned tiles = static_cast<unsigned>((b + GT - 1) / GT);
    trtri_pair_gemm<<<dim3(tiles, tiles, pairs), 256, 0, st>>>(L32, Y, Tbuf, Tbuf, k, static_cast<int>(b), 0);
    trtri_pair_gemm<<<dim3(tiles, tiles, pairs), 256, 0, st>>>(L32, Y, Tbuf, Y, k, static_cast<int>(b), 1);
    launches += 2;
  }
  e = count_launch(launches);
  if (e != cudaSuccess) return e;
  // H^-1 = Y^T Y (einsum "ji,jk->ik"): the same contraction as the Hessian itself.  The split
  // workspace may reuse A + L32, which are dead by now.
  e = xtx_impl<float>(Y, K, K, 1.0, hinv, ws, scratch, sm_count, st, /*lower_tri=*/1);  // Y = L^-1 is lower triangular
  if (e != cudaSuccess) return e;
  if (info_out) {
    e = cudaMemcpyAsync(info_out, info, sizeof(int), cudaMemcpyDeviceToDevice, st);
  }
  return e;
}

cudaError_t launch_weighted_mean_f64(const double* a, double wa, const double* b, double wb,
                                     double* out, long long n, int sm_count, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  weighted_mean_f64<<<sm_count * 8, 256, 0, st>>>(a, wa, b, wb, out, n);
  return count_launch();
}

}  // namespace aeqb
