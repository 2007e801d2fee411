// Blocked right-looking Cholesky (fp64, lower, in place) with the rank-128 updates on the
// FP64 tensor-core path (mma.sync m8n8k4.f64, "DMMA") and a one-step lookahead.
//
// Replaces np.linalg.cholesky on the damped Hessian (gptq._prepare_hessian_inverse,
// algorithms/uniform_quantize/gptq.py:111-128) for the layer sizes the benchmark runs
// (K = 4096 ... 11008).  Per 128-column block J:
//   chol_diag128   ONE CTA factors the 128 x 128 diagonal block in shared memory (four 32-column
//                  steps: a warp factors the 32 x 32 pivot block in registers, a thread per row
//                  solves the rows below it, a 6 x 6 register tile per thread updates the rest of
//                  the block) and inverts the factor in place (block recurrence), so that
//   chol_nt<32>    the panel below the block is a plain product  P = A[:, J:J+128] * Linv^T
//                  (no substitution chain across CTAs), 32 x 128 outputs per CTA;
//   chol_nt<128>   the trailing update  A[r, c] -= sum_k P[r, k] P[c, k]  as 128 x 128 tiles.
// Both products are the same "NT" tile routine: operands are K-major rows of plain row-major
// matrices, staged 32 contraction values at a time by cp.async (double-buffered, rows padded to
// 36 doubles: conflict-free fragment loads), accumulated with DMMA in registers.  The fp64 pipe
// (64 FMA / clock / SM) is the bound; everything else hides behind it.
// Lookahead: the first tile column of the trailing update (what the NEXT diagonal block and
// panel need) runs on the caller's stream, the rest on a low-priority side stream, so the serial
// diag -> panel chain overlaps the bulk of the update.  AEQB_CHOL_NO_LOOKAHEAD=1 serialises.
#include <stdio.h>
#include <stdlib.h>

#include "aeqb_common.cuh"
#include "aeqb_kernels.h"

namespace aeqb {

namespace {

constexpr int DNB = 128;         // outer block
constexpr int DLD = DNB + 4;     // row stride of the diagonal block in shared memory: 132 = 4 mod 16, so the 8-byte
                                 // words (g * DLD + t) and (t * DLD + g) of a DMMA fragment fall in 16 distinct banks
                                 // per half warp (stride 129 was 4-way conflicted: 16 wavefronts per DMMA, the bound)
constexpr int LLD = 36;          // same rule for the 32 x 32 side tiles (Li, Tm)
constexpr int DSB = 32;          // pivot sub-block
constexpr int KS = 32;           // contraction values per stage
constexpr int kCholDelay = 4;     // full passes over the trailing matrix every this many 128-column steps
constexpr int KLD = KS + 4;      // padded stage row: (g * 36 + t) mod 16 distinct over a half warp

__device__ __forceinline__ void dmma_884(double (&d)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d[0]), "+d"(d[1])
               : "d"(a), "d"(b));
}

template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(void* dst, const void* src, bool valid) {
  const int n = valid ? BYTES : 0;
  if (BYTES == 16) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 1 / sqrt(p) in float64: the hardware's double seed (rsqrt.approx.ftz.f64 = one MUFU.RSQ64H, relative
// error ~2^-22, no float64 <-> float32 conversions on the chain) and ONE cubically convergent step
//   e = 1 - p r^2,  r' = r (1 + e / 2 + 3 e^2 / 8)      (error ~ e^3: below 2^-60)
// four dependent operations after the seed; the pivot chain below is latency, not throughput.
__device__ __forceinline__ double rsqrt_f64_cubic(double p) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
  const double t = p * r;
  const double e = fma(-t, r, 1.0);
  const double q = fma(0.375, e, 0.5);
  return fma(r * e, q, r);
}

// 32 x 32 pivot block S[j0:j0+32, j0:j0+32] factored by one warp, lane = row, eight columns at a
// time in registers (pivots and multipliers travel by shuffle), the columns to the right of the
// eight updated in shared memory by a rolled loop.  The order of the statements is the schedule: a
// step first starts the reciprocal square root of ITS pivot (a chain of dependent fp64 operations,
// ~40 cycles each), then issues the column updates the PREVIOUS step left behind (independent of that
// chain, so they fill its latency), then scales its column and forms the next pivot — from the
// lane's own multiplier, so the only shuffle on the chain is the broadcast of the pivot itself.
// Eight-column groups instead of all 32 in registers keep the straight-line code near 6 KB: the
// fully unrolled version ran at instruction-fetch speed the first time a launch executed it
// (43k cycles cold against 15k warm, measured with clock64).
constexpr int PSB = 8;
__device__ __noinline__ void factor_pivot_block(double (*S)[DLD], double* rinv, int j0, int lane, int nb, int J,
                                                int* info) {
  double* row = &S[j0 + lane][j0];
#pragma unroll 1
  for (int c0 = 0; c0 < DSB; c0 += PSB) {
    double a[PSB];
#pragma unroll
    for (int u = 0; u < PSB; ++u) a[u] = row[c0 + u];
    double p = __shfl_sync(0xffffffffu, a[0], c0);
    double lprev = 0.0;  // the scaled column of the previous step (this lane's entry)
#pragma unroll
    for (int k = 0; k < PSB; ++k) {
      const int kk = c0 + k;  // pivot column = pivot lane
      const double rs = rsqrt_f64_cubic(p);
      if (k > 0) {  // deferred updates of step k - 1: columns k + 1 .. 7 of the group
#pragma unroll
        for (int c = k + 1; c < PSB; ++c) {
          const double lc = __shfl_sync(0xffffffffu, lprev, c0 + c);
          if (lane >= c0 + c) a[c] = fma(-lprev, lc, a[c]);
        }
      }
      if (lane == kk && j0 + kk < nb && !(p > 0.0)) atomicExch(info, J + j0 + kk + 1);
      double sq = p * rs;
      sq = fma(0.5 * rs, fma(-sq, sq, p), sq);  // one correction: sqrt(p) to the last bit or so
      const double lk = (lane == kk) ? sq : a[k] * rs;
      if (lane >= kk) a[k] = lk;
      if (lane == kk) rinv[kk] = rs;
      if (k + 1 < PSB) {
        const double diag_next = fma(-lk, lk, a[k + 1]);  // lane kk + 1 owns both factors of its pivot's update
        p = __shfl_sync(0xffffffffu, diag_next, kk + 1);
        const double lc = __shfl_sync(0xffffffffu, lk, kk + 1);
        if (lane > kk + 1) a[k + 1] = fma(-lk, lc, a[k + 1]);
        if (lane == kk + 1) a[k + 1] = diag_next;
      }
      lprev = lk;
    }
#pragma unroll
    for (int u = 0; u < PSB; ++u) row[c0 + u] = (c0 + u <= lane) ? a[u] : 0.0;
    // columns to the right of the group: S[r][c] -= sum_u l[r][u] l[c][u]  (each lane its own row)
#pragma unroll 1
    for (int c = c0 + PSB; c < DSB; ++c) {
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
      for (int u = 0; u < PSB; u += 2) {
        acc0 = fma(a[u], __shfl_sync(0xffffffffu, a[u], c), acc0);
        acc1 = fma(a[u + 1], __shfl_sync(0xffffffffu, a[u + 1], c), acc1);
      }
      if (lane >= c) row[c] -= acc0 + acc1;
    }
  }
  // rows above the diagonal were never written inside the groups' columns to the right: zero them
#pragma unroll 4
  for (int c = 1; c < DSB; ++c)
    if (c > lane) row[c] = 0.0;
}

// Threads 0..31: column `tid` of the inverse of the pivot block's factor (x starts as a unit vector);
// threads 32..32+below-1: one row of the block under the pivot block, X = A L^-T.  Both are the same
// right-looking substitution  x[m] *= 1 / L[m][m];  x[c] -= x[m] L[c][m]  (c > m): a chain of two
// dependent operations per step, the 31 - m updates behind it are independent.
__device__ __noinline__ void solve_rows(double (*S)[DLD], double (*Li)[LLD], const double* rinv, int j0, int below,
                                        int tid) {
  const bool inv = tid < DSB;
  const int r = j0 + DSB + (tid - DSB);
  const bool active = inv || (tid - DSB < below);
  double x[DSB];
#pragma unroll
  for (int c = 0; c < DSB; ++c) x[c] = inv ? (c == tid ? 1.0 : 0.0) : (active ? S[r][j0 + c] : 0.0);
#pragma unroll
  for (int m = 0; m < DSB; ++m) {
    x[m] *= rinv[m];
#pragma unroll
    for (int c = m + 1; c < DSB; ++c) x[c] = fma(-x[m], S[j0 + c][j0 + m], x[c]);
  }
  if (inv) {
#pragma unroll
    for (int c = 0; c < DSB; ++c) Li[c][tid] = x[c];
  } else if (active) {
#pragma unroll
    for (int c = 0; c < DSB; ++c) S[r][j0 + c] = x[c];
  }
}

// ---- DMMA phases of the diagonal block.  Everything a loop bound or a branch could depend on is a
// template parameter or clamped, so the bodies are branch-free: with a (warp-uniform) branch around
// every DMMA the compiler could not hoist the fragment loads and each DMMA paid a full shared-memory
// round trip (measured: 175 cycles per DMMA and warp instead of 16).

// Rest of the block after a 32-column step: S[r][c] -= sum_k X[r][k] X[c][k] over the lower 8 x 8
// tiles (tr >= tc) of the (8 NT) x (8 NT) square starting at b0; a warp owns tiles warp + 16 q, q < NQ,
// all in flight at once (a tile past the end is clamped to the last one and not written).
constexpr int DW = 16;  // warps of the diagonal-block kernel
template <int NQ>
__device__ __forceinline__ void block_update_tiles(double (*S)[DLD], int j0, int b0, int ntri, int warp, int g, int t) {
  int ra[NQ], rb[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    int left = min(warp + DW * q, ntri - 1), r = 0;
    while (left > r) { left -= r + 1; ++r; }
    ra[q] = b0 + r * 8 + g;     // this lane's row of the A fragment (tile row r)
    rb[q] = b0 + left * 8 + g;  // and of the B fragment (tile column)
  }
  double acc[NQ][2];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q][0] = acc[q][1] = 0.0;
#pragma unroll
  for (int k4 = 0; k4 < DSB / 4; ++k4) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) dmma_884(acc[q], S[ra[q]][j0 + k4 * 4 + t], S[rb[q]][j0 + k4 * 4 + t]);
  }
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    if (warp + DW * q < ntri) {
      const int r = ra[q], c = rb[q] - g + 2 * t;
      if (c <= r) S[r][c] -= acc[q][0];  // columns >= b0: disjoint from the X columns read above
      if (c + 1 <= r) S[r][c + 1] -= acc[q][1];
    }
  }
}

// Block row I of the in-place inverse:  T_Ij = sum_{m=j}^{I-1} L_Im X_mj  (X_jj = Li_j, X_mj in place)
// for every j < I, then X_Ij = -Li_I T_Ij.  Warp w of 16 owns tile (w / 4, w % 4) of EVERY block column
// j < I (q = j is a compile-time index), two accumulators per tile (even / odd contraction steps).
template <int I>
__device__ __forceinline__ void inverse_block_row(double (*S)[DLD], double (*Li)[DSB][LLD], double (*Tm)[DSB][LLD],
                                                  int warp, int g, int t) {
  const int tr = warp >> 2, tc = warp & 3;
  {
    double acc[I][2][2];
#pragma unroll
    for (int j = 0; j < I; ++j) acc[j][0][0] = acc[j][0][1] = acc[j][1][0] = acc[j][1][1] = 0.0;
#pragma unroll
    for (int j = 0; j < I; ++j) {
#pragma unroll
      for (int m = j; m < I; ++m) {
#pragma unroll
        for (int k4 = 0; k4 < DSB / 4; ++k4) {
          const double a = S[I * DSB + tr * 8 + g][m * DSB + k4 * 4 + t];
          const double b = (m == j) ? Li[j][k4 * 4 + t][tc * 8 + g] : S[m * DSB + k4 * 4 + t][j * DSB + tc * 8 + g];
          dmma_884(acc[j][k4 & 1], a, b);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < I; ++j) {
      Tm[j][tr * 8 + g][tc * 8 + 2 * t] = acc[j][0][0] + acc[j][1][0];
      Tm[j][tr * 8 + g][tc * 8 + 2 * t + 1] = acc[j][0][1] + acc[j][1][1];
    }
  }
  __syncthreads();
  {
    double acc[I][2][2];
#pragma unroll
    for (int j = 0; j < I; ++j) acc[j][0][0] = acc[j][0][1] = acc[j][1][0] = acc[j][1][1] = 0.0;
#pragma unroll
    for (int k4 = 0; k4 < DSB / 4; ++k4) {  // Li_I is lower triangular; its stored zeros keep this exact
#pragma unroll
      for (int j = 0; j < I; ++j)
        dmma_884(acc[j][k4 & 1], Li[I][tr * 8 + g][k4 * 4 + t], Tm[j][k4 * 4 + t][tc * 8 + g]);
    }
#pragma unroll
    for (int j = 0; j < I; ++j) {
      S[I * DSB + tr * 8 + g][j * DSB + tc * 8 + 2 * t] = -(acc[j][0][0] + acc[j][1][0]);
      S[I * DSB + tr * 8 + g][j * DSB + tc * 8 + 2 * t + 1] = -(acc[j][0][1] + acc[j][1][1]);
    }
  }
}

// ------------------------------------------------------------------ diagonal block
// 16 warps: a warp's DMMAs do not overlap each other (~64 cycles issue to issue), so the fp64 tensor
// pipe needs four warps per scheduler to stay busy.
constexpr int DIAG_THREADS = 512;
// Dynamic shared memory: S[128][132] | LiAll[4][32][36] | Tm[3][32][36] | rinv[32]
constexpr size_t kDiagSmem = (static_cast<size_t>(DNB) * DLD + 4 * DSB * (LLD) + 3 * DSB * (LLD) + DSB) * 8;

__global__ void __launch_bounds__(DIAG_THREADS, 1)
    chol_diag128(double* __restrict__ A, int K, int J, double* __restrict__ linv, int* __restrict__ info,
                 long long* __restrict__ dbg) {
  int dbg_n = 0;
#define AEQB_TICK() do { if (dbg != nullptr && threadIdx.x == 0) dbg[dbg_n++] = clock64(); } while (0)
  extern __shared__ __align__(16) double dsm[];
  double (*S)[DLD] = reinterpret_cast<double (*)[DLD]>(dsm);
  double (*Li)[DSB][LLD] = reinterpret_cast<double (*)[DSB][LLD]>(dsm + DNB * DLD);
  double (*Tm)[DSB][LLD] = reinterpret_cast<double (*)[DSB][LLD]>(dsm + DNB * DLD + 4 * DSB * (LLD));
  double* rinv = dsm + DNB * DLD + 7 * DSB * (LLD);
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int g = lane >> 2, t = lane & 3;
  const int nb = min(DNB, K - J);
  // lower triangle of the block; rows past the matrix are identity rows (harmless pivots of 1)
#pragma unroll 8
  for (int e = tid; e < DNB * DNB; e += DIAG_THREADS) {
    const int r = e >> 7, c = e & 127;
    double v = (r == c) ? 1.0 : 0.0;
    if (r < nb && c <= r) v = A[static_cast<long long>(J + r) * K + J + c];
    S[r][c] = v;
  }
  __syncthreads();
  AEQB_TICK();  // 0: block loaded

#pragma unroll 1
  for (int jj = 0; jj < DNB / DSB; ++jj) {
    const int j0 = jj * DSB;
    if (j0 >= nb) {  // nothing real left: the inverse of an identity block is the identity
      for (int e = tid; e < DSB * DSB; e += DIAG_THREADS) Li[jj][e >> 5][e & 31] = ((e >> 5) == (e & 31)) ? 1.0 : 0.0;
      continue;
    }
    if (warp == 0) factor_pivot_block(S, rinv, j0, lane, nb, J, info);
    __syncthreads();
    AEQB_TICK();  // 1 + 3 jj: pivot block factored
    const int below = DNB - j0 - DSB;  // rows of the block under the pivot block
    if (tid < DSB + below) solve_rows(S, Li[jj], rinv, j0, below, tid);
    __syncthreads();
    AEQB_TICK();  // 2 + 3 jj: rows solved, pivot block inverted
    if (below > 0) {
      // rest of the block: S[r][c] -= sum_k X[r][k] X[c][k] on DMMA, 8 x 8 tiles of the lower
      // triangle (tr >= tc) dealt to the warps two at a time (independent accumulator chains)
      const int nt = below / 8;  // 12, 8, 4
      const int b0 = j0 + DSB;
      const int ntri = nt * (nt + 1) / 2;  // 78, 36, 10 tiles
      if (jj == 0) block_update_tiles<5>(S, j0, b0, ntri, warp, g, t);
      else if (jj == 1) block_update_tiles<3>(S, j0, b0, ntri, warp, g, t);
      else block_update_tiles<1>(S, j0, b0, ntri, warp, g, t);
      __syncthreads();
    }
    AEQB_TICK();  // 3 + 3 jj: rest of the block updated
  }
  // the factor goes back to the matrix
#pragma unroll 4
  for (int e = tid; e < DNB * DNB; e += DIAG_THREADS) {
    const int r = e >> 7, c = e & 127;
    if (r < nb && c <= r) A[static_cast<long long>(J + r) * K + J + c] = S[r][c];
  }
  AEQB_TICK();  // 13: factor stored
  if (linv == nullptr) return;  // last block: nothing below it
  // ---- inverse of the factor, in place, block row by block row, the 32 x 32 block products on DMMA:
  //   X_ii = Li_i,   X_ij = -Li_i * T_ij,   T_ij = sum_{m=j}^{i-1} L_im X_mj   (X_jj = Li_j, X_mj in place)
  // A warp owns 8 x 8 output tiles (id = j * 16 + tr * 4 + tc); A fragments are rows of the left
  // factor, B fragments columns of the right one, both read straight from shared memory.
  __syncthreads();
  inverse_block_row<1>(S, Li, Tm, warp, g, t);
  __syncthreads();
  inverse_block_row<2>(S, Li, Tm, warp, g, t);
  __syncthreads();
  inverse_block_row<3>(S, Li, Tm, warp, g, t);
  __syncthreads();
#pragma unroll 4
  for (int e = tid; e < DNB * DNB; e += DIAG_THREADS) {
    const int r = e >> 7, c = e & 127;
    double v = 0.0;
    if (c <= r) v = ((r >> 5) == (c >> 5)) ? Li[r >> 5][r & 31][c & 31] : S[r][c];
    linv[e] = v;
  }
  __syncthreads();
  AEQB_TICK();  // 14: inverse done and stored
#undef AEQB_TICK
}

// Column tiles (BN wide, indices in [tc_lo, tc_hi) relative to the trailing matrix) of row tile tr (BM
// rows) that hold an element on or below the diagonal.
__host__ __device__ inline int nt_tiles_in_row(int tr, int bm, int bn, int tc_lo, int tc_hi) {
  const int last = (tr * bm + bm - 1) / bn + 1;  // column tiles [0, last) touch the triangle
  const int hi = last < tc_hi ? last : tc_hi;
  return hi > tc_lo ? hi - tc_lo : 0;
}

// ------------------------------------------------------------------ NT tile product on DMMA
// acc[BM x 128] = sum_{k < depth} X[xrow0 + i][k] * Y[yrow0 + j][k]   (X, Y row-major, K-major rows)
// MODE 0 (panel):   A[xrow0 + i][ocol0 + j]  = acc          (X = A[:, J:], Y = Linv)
// MODE 1 (update):  A[xrow0 + i][yrow0 + j] -= acc, j-th column <= row only   (X = Y = A[:, J:])
template <int MF, int NF, int WM, int WN, int MODE, int CPB>
__global__ void __launch_bounds__(WM * WN * 32, (MF * WM * NF * WN * 64 >= 128 * 128) ? 1 : 2)
    chol_nt(double* __restrict__ A, int K, const double* __restrict__ X, long long ldx, const double* __restrict__ Y,
            long long ldy, int y_rows, int depth, int base, int ocol0, int tc_lo, int tc_hi) {
  constexpr int BM = MF * WM * 8, BN = NF * WN * 8;
  constexpr int NT_THREADS = WM * WN * 32;
  extern __shared__ __align__(16) double nsm[];
  double* Xs = nsm;                          // [2][BM][KLD]
  double* Ys = nsm + 2 * BM * KLD;           // [2][BN][KLD]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp / WN, wn = warp % WN;
  // tile coordinates
  long long xrow0, yrow0;
  if (MODE == 0) {
    xrow0 = static_cast<long long>(base) + static_cast<long long>(blockIdx.x) * BM;
    yrow0 = 0;
  } else {
    // blockIdx.x enumerates, row tile by row tile (BM rows), the column tiles (BN columns, indices
    // [tc_lo, tc_hi) relative to `base`) that touch the lower triangle: tc * BN <= tr * BM + BM - 1.
    int left = blockIdx.x, tr = 0, tc = tc_lo;
    if (tc_lo == 0 && tc_hi == 1) {
      tr = left;  // the first column tile: every row tile has it
      left = 0;
    } else if (BM == 128 && BN == 64 && tc_lo == 2 && tc_hi >= 2 * ((K - base + BM - 1) / BM) - 1) {
      // row tile tr holds column tiles 2 .. 2 tr + 1: tr (tr - 1) tiles lie before it.  Closed form
      // instead of a walk over up to 85 row tiles (a quarter of this kernel's stall samples were the
      // integer instructions of that walk).
      tr = static_cast<int>((1.0f + sqrtf(1.0f + 4.0f * static_cast<float>(left))) * 0.5f);
      while (tr * (tr - 1) > left) --tr;
      while ((tr + 1) * tr <= left) ++tr;
      left -= tr * (tr - 1);
    } else if (BM == 128 && BN == 64 && tc_lo == 2 && tc_hi == 4) {
      tr = 1 + (left >> 1);  // one 128-column block: two column tiles in every row tile from the second on
      left &= 1;
    } else {
      for (;; ++tr) {
        const int cnt = nt_tiles_in_row(tr, BM, BN, tc_lo, tc_hi);
        if (left < cnt) break;
        left -= cnt;
      }
    }
    tc = tc_lo + left;
    xrow0 = static_cast<long long>(base) + static_cast<long long>(tr) * BM;
    yrow0 = static_cast<long long>(base) + static_cast<long long>(tc) * BN;
  }
  const long long x_rows = K;
  if (MODE == 1 && BM == 128) {
    // The C tile comes from HBM (the trailing matrix is far larger than L2): ask L2 for its lines
    // now, so that the epilogue's loads find them there after the products instead of waiting for
    // DRAM with nothing left to overlap (a fifth of the stall samples were those waits).
    constexpr int LPRW = BN * 8 / 128;  // 128-byte lines per tile row
    for (int e = tid; e < BM * LPRW; e += NT_THREADS) {
      const long long gr = xrow0 + e / LPRW, gc = yrow0 + (e % LPRW) * 16;
      if (gr < K && gc <= gr) asm volatile("prefetch.global.L2 [%0];" ::"l"(A + gr * K + gc));
    }
  }

  auto load_stage = [&](int s, int k0) {
    double* xs = Xs + s * BM * KLD;
    double* ys = Ys + s * BN * KLD;
    constexpr int EL = CPB / 8;          // doubles per chunk (2 or 1)
    constexpr int CPR = KS / EL;         // chunks per row
    for (int e = tid; e < BM * CPR; e += NT_THREADS) {
      const int r = e / CPR, c = (e % CPR) * EL;
      const bool ok = xrow0 + r < x_rows;
      const double* src = X + (ok ? (xrow0 + r) : xrow0) * ldx + k0 + c;
      cp_async_zfill<CPB>(xs + r * KLD + c, src, ok);
    }
    for (int e = tid; e < BN * CPR; e += NT_THREADS) {
      const int r = e / CPR, c = (e % CPR) * EL;
      const bool ok = yrow0 + r < y_rows;
      const double* src = Y + (ok ? (yrow0 + r) : yrow0) * ldy + k0 + c;
      cp_async_zfill<CPB>(ys + r * KLD + c, src, ok);
    }
    cp_async_commit();
  };

  double acc[MF][NF][2];
#pragma unroll
  for (int i = 0; i < MF; ++i)
#pragma unroll
    for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nst = depth / KS;
  load_stage(0, 0);
  for (int s = 0; s < nst; ++s) {
    if (s + 1 < nst) {
      load_stage((s + 1) & 1, (s + 1) * KS);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const double* xs = Xs + (s & 1) * BM * KLD + (wm * MF * 8 + g) * KLD + t;
    const double* ys = Ys + (s & 1) * BN * KLD + (wn * NF * 8 + g) * KLD + t;
#pragma unroll
    for (int k4 = 0; k4 < KS / 4; ++k4) {
      double a[MF], b[NF];
#pragma unroll
      for (int i = 0; i < MF; ++i) a[i] = xs[i * 8 * KLD + k4 * 4];
#pragma unroll
      for (int j = 0; j < NF; ++j) b[j] = ys[j * 8 * KLD + k4 * 4];
#pragma unroll
      for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) dmma_884(acc[i][j], a[i], b[j]);
    }
    __syncthreads();  // the stage is free for the load of s + 2
  }

  if constexpr (MODE == 1 && BM == 128 && CPB == 16) {
    // Update epilogue through shared memory: the accumulators are parked in a [128][136] tile
    // (16-byte stores, conflict-free), then every thread issues its 32 independent 16-byte loads of
    // C (a warp covers 512 contiguous bytes of a row), subtracts and stores.  Straight from the
    // fragments every DADD waited for its own 8-byte load: two thirds of the kernel's stall samples.
    constexpr int CLD = BN + 8;  // 16-byte stores of a quarter warp fall in 8 distinct 16-byte banks
    double* Ct = nsm;
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        const int r = wm * MF * 8 + i * 8 + g, c = wn * NF * 8 + j * 8 + 2 * t;
        *reinterpret_cast<double2*>(Ct + r * CLD + c) = make_double2(acc[i][j][0], acc[i][j][1]);
      }
    __syncthreads();
    constexpr int CPRW = BN / 2;                              // 16-byte chunks per tile row
    constexpr int NB16 = (BM * CPRW) / (16 * NT_THREADS);     // batches of 16 chunks per thread
    static_assert(NB16 * 16 * NT_THREADS == BM * CPRW, "chunks divide over the threads");
#pragma unroll
    for (int b = 0; b < NB16; ++b) {
      double2 v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int id = (b * 16 + u) * NT_THREADS + tid;
        const long long gr = xrow0 + id / CPRW, gc = yrow0 + (id % CPRW) * 2;
        v[u] = make_double2(0.0, 0.0);
        if (gr < K && gc <= gr) v[u] = *reinterpret_cast<const double2*>(A + gr * K + gc);
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int id = (b * 16 + u) * NT_THREADS + tid;
        const int row = id / CPRW, cc = (id % CPRW) * 2;
        const long long gr = xrow0 + row, gc = yrow0 + cc;
        if (gr < K && gc <= gr) {
          const double2 d = *reinterpret_cast<const double2*>(Ct + row * CLD + cc);
          double* o = A + gr * K + gc;
          if (gc + 1 <= gr) *reinterpret_cast<double2*>(o) = make_double2(v[u].x - d.x, v[u].y - d.y);
          else o[0] = v[u].x - d.x;
        }
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < MF; ++i) {
    const long long r = xrow0 + wm * MF * 8 + i * 8 + g;
    if (r >= K) continue;
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      const int cj = wn * NF * 8 + j * 8 + 2 * t;
      if (MODE == 0) {
        double* o = A + r * K + ocol0 + cj;
        o[0] = acc[i][j][0];
        o[1] = acc[i][j][1];
      } else {
        const long long c = yrow0 + cj;
        double* o = A + r * K + c;
        if (c <= r) o[0] -= acc[i][j][0];
        if (c + 1 <= r) o[1] -= acc[i][j][1];
      }
    }
  }
}

struct SideStream {
  int dev = -1;
  cudaStream_t caller = nullptr;
  cudaStream_t s = nullptr;
  cudaEvent_t panel = nullptr, rest = nullptr;
};

// One low-priority side stream (and its two events) per (device, caller stream): factorisations
// that the caller runs concurrently on different streams (the four Hessians of a decoder layer)
// keep independent lookahead queues.  Created on first use, kept for the life of the process; when
// the table is full the factorisation simply runs without lookahead.
SideStream* side_stream(cudaStream_t caller) {
  constexpr int kMax = 32;
  static SideStream table[kMax];
  static int used = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  for (int i = 0; i < used; ++i)
    if (table[i].dev == dev && table[i].caller == caller) return &table[i];
  if (used == kMax) return nullptr;
  SideStream ss;
  ss.dev = dev;
  ss.caller = caller;
  int lo = 0, hi = 0;
  if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) return nullptr;
  if (cudaStreamCreateWithPriority(&ss.s, cudaStreamNonBlocking, lo) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&ss.panel, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&ss.rest, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  table[used] = ss;
  return &table[used++];
}

template <int MF, int NF, int WM, int WN, int MODE, int CPB>
cudaError_t configure_nt() {
  constexpr size_t smem = static_cast<size_t>(2) * (MF * WM * 8 + NF * WN * 8) * KLD * 8;
  return cudaFuncSetAttribute(chol_nt<MF, NF, WM, WN, MODE, CPB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              static_cast<int>(smem));
}

template <int CPB>
cudaError_t cholesky_dmma_impl(double* A, int K, double* linv, int* info, cudaStream_t st, int* launches) {
  static PerDevice configured;
  cudaError_t e;
  if (!configured.done()) {
    if ((e = cudaFuncSetAttribute(chol_diag128, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(kDiagSmem))) != cudaSuccess) return e;
    configured.set();
  }
  static PerDevice nt_configured;
  if (!nt_configured.done()) {
    if ((e = configure_nt<4, 2, 1, 8, 0, CPB>()) != cudaSuccess) return e;
    if ((e = configure_nt<4, 2, 1, 8, 1, CPB>()) != cudaSuccess) return e;
    if ((e = configure_nt<4, 4, 4, 2, 1, CPB>()) != cudaSuccess) return e;
    nt_configured.set();
  }
  constexpr size_t smem32 = static_cast<size_t>(2) * (32 + 128) * KLD * 8;
  constexpr size_t smem128 = static_cast<size_t>(2) * (128 + 64) * KLD * 8;  // 128 x 64 tiles
  const bool lookahead = getenv("AEQB_CHOL_NO_LOOKAHEAD") == nullptr;
  int delay = getenv("AEQB_CHOL_DELAY") ? atoi(getenv("AEQB_CHOL_DELAY")) : kCholDelay;
  if (delay < 1) delay = 1;
  if (delay > 8) delay = 8;
  SideStream* ss = lookahead ? side_stream(st) : nullptr;
  bool rest_pending = false;
  const double* AJ;
  for (int J = 0; J < K; J += DNB) {
    const int base = J + DNB;        // first row / column after this block
    const bool last = base >= K;
    long long* dbg = nullptr;
    if (J == 0 && getenv("AEQB_CHOL_TIMING") != nullptr) {  // debug: phase clocks of the first diagonal block
      static long long* dbg_buf = nullptr;
      if (dbg_buf == nullptr && cudaMalloc(&dbg_buf, 32 * sizeof(long long)) != cudaSuccess) return cudaErrorMemoryAllocation;
      dbg = dbg_buf;
    }
    chol_diag128<<<1, DIAG_THREADS, kDiagSmem, st>>>(A, K, J, last ? nullptr : linv, info, dbg);
    ++*launches;
    if (dbg != nullptr) {
      long long h[16];
      if ((e = cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
      fprintf(stderr, "chol_diag128 phase cycles:");
      for (int i = 1; i < 15; ++i) fprintf(stderr, " %lld", h[i] - h[i - 1]);
      fprintf(stderr, "\n");
    }
    if (last) break;
    AJ = A + J;
    const int below = K - base;
    // panel: P = A[base:, J:J+128] * Linv^T
    chol_nt<4, 2, 1, 8, 0, CPB><<<(below + 31) / 32, 256, smem32, st>>>(
        A, K, AJ, K, linv, DNB, DNB, DNB, base, J, 0, 0);
    ++*launches;
    const int nct = (below + 127) / 128;  // column tiles of the trailing matrix
    if (ss != nullptr) {
      if ((e = cudaEventRecord(ss->panel, st)) != cudaSuccess) return e;
      if (rest_pending && (e = cudaStreamWaitEvent(st, ss->rest, 0)) != cudaSuccess) return e;
    }
    // first 128 columns of the trailing matrix (the next block's diagonal and panel): 32-row tiles,
    // on the caller's stream
    chol_nt<4, 2, 1, 8, 1, CPB><<<(below + 31) / 32, 256, smem32, st>>>(
        A, K, AJ, K, AJ, K, K, DNB, base, 0, 0, 1);
    ++*launches;
    if (nct > 1) {
      // The other columns: 128 x 64 tiles, two CTAs per SM (one tile's prologue and epilogue overlap the
      // other's products).  DELAYED: the read-modify-write of the trailing matrix is HBM traffic the
      // products do not hide (26.8 GB over the 86 steps at K = 11008), so a column block only has to be
      // current when its turn comes.  Every `delay`-th step updates ALL remaining columns with the last
      // `delay` panels in one pass (contraction 128 x delay); the steps in between update only the
      // block column the next step's first-column kernel needs, with the panels pending since the last
      // full pass.  Same products, same order of additions per element within a pass.
      const int i = J / DNB;
      const bool full = (i + 1) % delay == 0;
      const int s0 = full ? i + 1 - delay : (i / delay) * delay;  // first panel not yet applied
      const int depth = (i - s0 + 1) * DNB;
      const double* AS = A + static_cast<long long>(s0) * DNB;
      const int nrt = nct, ntc = (below + 63) / 64;
      const int tc_hi = full ? ntc : (ntc < 4 ? ntc : 4);
      long long ntiles = 0;
      for (int tr = 0; tr < nrt; ++tr) ntiles += nt_tiles_in_row(tr, 128, 64, 2, tc_hi);
      cudaStream_t rs = st;
      if (ss != nullptr) {
        rs = ss->s;
        if ((e = cudaStreamWaitEvent(rs, ss->panel, 0)) != cudaSuccess) return e;
      }
      if (ntiles > 0) {
        chol_nt<4, 4, 4, 2, 1, CPB><<<static_cast<unsigned>(ntiles), 256, smem128, rs>>>(
            A, K, AS, K, AS, K, K, depth, base, 0, 2, tc_hi);
        ++*launches;
      }
      if (ss != nullptr) {
        if ((e = cudaEventRecord(ss->rest, rs)) != cudaSuccess) return e;
        rest_pending = true;
      }
    }
  }
  if (ss != nullptr && rest_pending) {
    if ((e = cudaStreamWaitEvent(st, ss->rest, 0)) != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

}  // namespace

size_t cholesky_dmma_workspace_bytes() { return static_cast<size_t>(DNB) * DNB * sizeof(double); }

// A: K x K fp64, lower triangle in, Cholesky factor out (upper triangle untouched).
// linv: cholesky_dmma_workspace_bytes() of scratch, 16-byte aligned.
cudaError_t launch_cholesky_dmma(double* A, int K, double* linv, int* info, cudaStream_t st, int* launches) {
  // 16-byte cp.async needs every row start 16-byte aligned: K even and A itself aligned
  const bool a16 = (K % 2 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0);
  return a16 ? cholesky_dmma_impl<16>(A, K, linv, info, st, launches)
             : cholesky_dmma_impl<8>(A, K, linv, info, st, launches);
}

}  // namespace aeqb
