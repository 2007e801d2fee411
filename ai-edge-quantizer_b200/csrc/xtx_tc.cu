// GPTQ Hessian on the 5th-generation tensor cores:  H = alpha * X^T X
//   gptq.calibrate                (algorithms/uniform_quantize/gptq.py:100-106)
//   gptq._prepare_hessian_inverse (gptq.py:126-128, H^-1 = L^-T L^-1 is the same contraction)
//
// The reference contracts fp32 activations with sgemm (fp32 products, fp32 sums).  tcgen05 has no
// fp32 operand type; kind::tf32 keeps 10 mantissa bits.  Every operand is therefore split once
// into two TF32 planes, x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi)  (|x - hi - lo|
// <= 2^-22 |x|), and each k-step issues three MMAs into the same TMEM accumulator,
//   hi_i*hi_j + hi_i*lo_j + lo_i*hi_j      (the dropped lo*lo term is <= 2^-22 |x_i x_j|),
// which gives fp32-class products with fp32 accumulation ("3xTF32").
//
// Data flow per token chunk (<= kChunk tokens):
//   xtx_split_transpose   X[t, k] fp32 -> hi[k, t], lo[k, t]  (K-major operand planes: the
//                         contraction index t is contiguous, so one [128 rows x 32 tokens] TMA
//                         box is exactly a 128-byte-swizzled UMMA operand tile).  Also raises a
//                         device flag on non-finite / huge values (the split would turn
//                         inf * x into NaN): the flagged case is recomputed by the SIMT kernel.
//   xtx_tc_gemm           one CTA per 128 x 256 output tile that touches the upper triangle;
//                         warp 0: TMA producer (6 boxes = 96 KiB per stage, 2 stages),
//                         warp 1: TMEM allocation + single-thread tcgen05.mma issue (12 MMAs of
//                                 128x256x8 per stage), tcgen05.commit releases the stage,
//                         warps 2-9: tcgen05.ld of each finished 64-token segment from one of
//                                 two TMEM accumulators, folded into fp32 registers (the tensor
//                                 core's own accumulation truncates), finally -> fp32 partial
//                                 P[k, k] (+= for every chunk after the first).
//   xtx_tc_finish         out = alpha * P mirrored to both triangles (float64 or float32).
#include <cuda.h>

#include "aeqb_common.cuh"
#include "aeqb_kernels.h"
#include "aeqb_tc.cuh"

namespace aeqb {

using namespace tc;

namespace {

constexpr int TC_BM = 128;      // output tile rows  (UMMA M)
constexpr int TC_BN = 256;      // output tile cols  (UMMA N)
constexpr int TC_BK = 32;       // tokens per stage = one 128-byte swizzle row of fp32
constexpr int TC_UK = 8;        // tokens per tcgen05.mma kind::tf32
constexpr int TC_STAGES = 2;
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 4;           // 16 KiB per plane
constexpr uint32_t TC_B_BYTES = TC_BN * TC_BK * 4;           // 32 KiB per plane
constexpr uint32_t TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;  // 96 KiB
constexpr uint32_t TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = (2 + TC_EPI_WARPS) * 32;
constexpr int TC_SEG_KB = 2;    // stages per TMEM accumulation segment (see xtx_tc_gemm)
constexpr long long kChunk = 16384;  // tokens per split/GEMM round (bounds the plane workspace)

// Instruction descriptor: D fp32 (bits [4,6) = 1), A and B TF32 (format 2 in [7,10) and
// [10,13)), both K-major (bits 15, 16 = 0), N >> 3 in [17,23), M >> 4 in [24,29).
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) |
                                (static_cast<uint32_t>(TC_BN >> 3) << 17) |
                                (static_cast<uint32_t>(TC_BM >> 4) << 24);

// Tile t -> (bi, bj): row block bi of 128 rows, column block bj of 256 columns, only tiles with
// an element on or above the diagonal (256*bj + 255 >= 128*bi  <=>  bj >= bi / 2).
__device__ __forceinline__ void tc_tile_index(int t, int nbj, int& bi, int& bj) {
  int row = 0, left = t;
  while (left >= nbj - (row >> 1)) {
    left -= nbj - (row >> 1);
    ++row;
  }
  bi = row;
  bj = (row >> 1) + left;
}
inline long long tc_tile_count(int K) {
  const int nbi = (K + TC_BM - 1) / TC_BM, nbj = (K + TC_BN - 1) / TC_BN;
  long long n = 0;
  for (int bi = 0; bi < nbi; ++bi) n += nbj - (bi >> 1);
  return n;
}

// ------------------------------------------------------------------ split + transpose
// X[t0 + t, k] (row pitch K) -> hi[k, t], lo[k, t] (row pitch `pitch` >= tokens, zero padded).
__global__ void __launch_bounds__(256)
    xtx_split_transpose(const float* __restrict__ X, long long tokens, int K, long long pitch,
                        float* __restrict__ hi, float* __restrict__ lo, int* __restrict__ flag) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const long long t0 = static_cast<long long>(blockIdx.x) * 32;
  const int k0 = blockIdx.y * 32;
  bool bad = false;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const long long t = t0 + ty + 8 * r;
    const int k = k0 + tx;
    float v = 0.0f;
    if (t < tokens && k < K) v = __ldg(X + t * K + k);
    bad |= !(fabsf(v) < 1e37f);  // NaN, inf, or so large that hi could round up to inf
    tile[ty + 8 * r][tx] = v;
  }
  if (bad) atomicOr(flag, 1);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int k = k0 + ty + 8 * r;
    const long long t = t0 + tx;
    if (k < K && t < pitch) {
      const float v = tile[tx][ty + 8 * r];
      uint32_t h, l;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
      const float rest = v - __uint_as_float(h);  // exact
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(rest));
      hi[static_cast<long long>(k) * pitch + t] = __uint_as_float(h);
      lo[static_cast<long long>(k) * pitch + t] = __uint_as_float(l);
    }
  }
}

// ------------------------------------------------------------------ 3xTF32 GEMM
// The tensor core adds into its fp32 TMEM accumulator with truncation (measured: the error of a
// long all-positive chain grows like 2^-25.5 per MMA, i.e. 1.7e-4 of the diagonal after 16384
// tokens).  The chain is therefore cut into segments of TC_SEG_KB stages: inside a segment the
// 2^-11-sized cross terms go in first (their truncation is relative to a still-small sum), and
// after TC_SEG_KB * 4 full-sized MMAs the epilogue warps fold the segment into fp32 registers
// with round-to-nearest FADDs.  Two 256-column TMEM buffers alternate, so draining one segment
// overlaps the MMAs of the next.
__global__ void __launch_bounds__(TC_THREADS, 1)
    xtx_tc_gemm(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                float* __restrict__ P, int K, int nkb_all, int nbj, int accumulate,
                const int* __restrict__ flag, int tri) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
  uint64_t* empty = full + TC_STAGES;
  uint64_t* tmem_full = empty + TC_STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  if (*flag != 0) return;  // non-finite input: the SIMT kernel recomputes this Hessian

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bi, bj;
  tc_tile_index(blockIdx.x, nbj, bi, bj);
  const int i0 = bi * TC_BM, j0 = bj * TC_BN;
  // tri: the input is LOWER TRIANGULAR as [contraction index t, column] (H^-1 = Y^T Y with Y = L^-1):
  // Y[t, j] = 0 for t < j, so an (upper) output tile only sees contraction blocks from its first B
  // column on (j0 >= i0) -- on average a third of the range.
  const int kb0 = tri ? min(j0 / TC_BK, nkb_all - 1) : 0;
  const int nkb = nkb_all - kb0;
  const int nseg = (nkb + TC_SEG_KB - 1) / TC_SEG_KB;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_lo)) : "memory");
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], TC_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) {  // whole warp: all 512 TMEM columns = two 128-lane x 256-column accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = static_cast<uint32_t>(kb / TC_STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full[s], TC_STAGE_BYTES);
        uint8_t* base = smem + s * TC_STAGE_BYTES;
        const int t = (kb0 + kb) * TC_BK;
        tma_load_2d(base, &tm_hi, t, i0, &full[s]);                                   // A hi
        tma_load_2d(base + TC_A_BYTES, &tm_lo, t, i0, &full[s]);                      // A lo
        tma_load_2d(base + 2 * TC_A_BYTES, &tm_hi, t, j0, &full[s]);                  // B hi
        tma_load_2d(base + 3 * TC_A_BYTES, &tm_hi, t, j0 + 128, &full[s]);
        tma_load_2d(base + 2 * TC_A_BYTES + TC_B_BYTES, &tm_lo, t, j0, &full[s]);     // B lo
        tma_load_2d(base + 3 * TC_A_BYTES + TC_B_BYTES, &tm_lo, t, j0 + 128, &full[s]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer
      int kb = 0;
      for (int seg = 0; seg < nseg; ++seg) {
        const int buf = seg & 1;
        mbar_wait(&tmem_empty[buf], (static_cast<uint32_t>(seg >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + static_cast<uint32_t>(buf * TC_BN);
        const int kb_end = min(nkb, kb + TC_SEG_KB);
        bool first = true;
        for (; kb < kb_end; ++kb) {
          const int s = kb % TC_STAGES;
          const uint32_t ph = static_cast<uint32_t>(kb / TC_STAGES) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + s * TC_STAGE_BYTES);
          const uint64_t a_hi = tc_smem_desc(base);
          const uint64_t a_lo = tc_smem_desc(base + TC_A_BYTES);
          const uint64_t b_hi = tc_smem_desc(base + 2 * TC_A_BYTES);
          const uint64_t b_lo = tc_smem_desc(base + 2 * TC_A_BYTES + TC_B_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UK; ++k) {  // cross terms first
            const uint64_t off = static_cast<uint64_t>((k * TC_UK * 4) >> 4);  // 32 B per k-step
            tc_mma_tf32(tacc, a_hi + off, b_lo + off, kIdescTf32, first ? 0u : 1u);
            first = false;
            tc_mma_tf32(tacc, a_lo + off, b_hi + off, kIdescTf32, 1u);
          }
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UK; ++k) {
            const uint64_t off = static_cast<uint64_t>((k * TC_UK * 4) >> 4);
            tc_mma_tf32(tacc, a_hi + off, b_hi + off, kIdescTf32, 1u);
          }
          tc_commit(&empty[s]);  // arrives when the MMAs above have finished reading the stage
        }
        tc_commit(&tmem_full[buf]);
      }
    }
    __syncwarp();
  } else {  // ---------------- epilogue: 8 warps, lane quarter = warp % 4, column half = (warp-2)/4
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    float acc[128];
#pragma unroll
    for (int u = 0; u < 128; ++u) acc[u] = 0.0f;
    for (int seg = 0; seg < nseg; ++seg) {
      const int buf = seg & 1;
      mbar_wait(&tmem_full[buf], static_cast<uint32_t>(seg >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(buf * TC_BN + half * 128);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
#pragma unroll
        for (int u = 0; u < 32; ++u) acc[c * 32 + u] = __fadd_rn(acc[c * 32 + u], __uint_as_float(v[u]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
    const int i = i0 + q * 32 + lane;
    if (i < K) {
      float* prow = P + static_cast<long long>(i) * K;
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        const int j = j0 + half * 128 + u * 4;
        if (j < K) {  // K % 4 == 0 (launcher), so a float4 is all-in or all-out
          float4 o = make_float4(acc[4 * u], acc[4 * u + 1], acc[4 * u + 2], acc[4 * u + 3]);
          float4* dst = reinterpret_cast<float4*>(prow + j);
          if (accumulate) {
            const float4 old = *dst;
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          *dst = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u)
                 : "memory");
  }
}

// out[i, j] = out[j, i] = alpha * P[i, j] for j >= i (every such element lies in a computed tile).
template <typename OutT>
__global__ void __launch_bounds__(256)
    xtx_tc_finish(const float* __restrict__ P, int K, double alpha, OutT* __restrict__ out,
                  const int* __restrict__ flag) {
  if (*flag != 0) return;
  const long long n = static_cast<long long>(K) * K;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(e / K), j = static_cast<int>(e % K);
    if (j < i) continue;
    const OutT v = static_cast<OutT>(alpha * static_cast<double>(P[e]));
    out[e] = v;
    if (j != i) out[static_cast<long long>(j) * K + i] = v;
  }
}

// ------------------------------------------------------------------ host side
// [K rows, pitch tokens] fp32 plane, box = 32 tokens (128 B, swizzled) x 128 rows.
bool make_plane_map(CUtensorMap* m, float* plane, long long K, long long pitch) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(pitch), static_cast<cuuint64_t>(K)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch) * 4};
  const cuuint32_t box[2] = {TC_BK, 128};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, plane, dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

struct TcLayout {
  long long chunk, pitch;
  size_t p_off, hi_off, lo_off, flag_off, total;
};
TcLayout tc_layout(long long T, long long K) {
  TcLayout l;
  l.chunk = T < kChunk ? T : kChunk;
  l.pitch = round_up(l.chunk, TC_BK);
  const size_t p_bytes = static_cast<size_t>(round_up(K * K * 4, 1024));
  const size_t plane = static_cast<size_t>(round_up(K * l.pitch * 4, 1024));
  l.p_off = 0;
  l.hi_off = p_bytes;
  l.lo_off = l.hi_off + plane;
  l.flag_off = l.lo_off + plane;
  l.total = l.flag_off + 256;
  return l;
}

}  // namespace

bool xtx_tc_eligible(long long T, long long K) {
  return K >= 128 && K % 4 == 0 && T >= 64 && K <= 65536;
}

size_t xtx_tc_workspace_bytes(long long T, long long K) { return tc_layout(T, K).total; }

// Returns cudaErrorNotSupported when the tensor-map entry point is missing (caller falls back to
// the SIMT kernel).  `flag` (device int, in ws) is left non-zero when the input held non-finite
// values: then `out` was NOT written and the caller must run the SIMT path.
template <typename OutT>
cudaError_t launch_xtx_tc(const float* x, long long T, long long K, double alpha, OutT* out,
                          void* ws, int sm_count, const int** flag_out, cudaStream_t st, int lower_tri) {
  static PerDevice attr_done;
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(xtx_tc_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(TC_SMEM_BYTES));
    if (e != cudaSuccess) return e;
    attr_done.set();
  }
  const TcLayout l = tc_layout(T, K);
  unsigned char* p = static_cast<unsigned char*>(ws);
  float* P = reinterpret_cast<float*>(p + l.p_off);
  float* hi = reinterpret_cast<float*>(p + l.hi_off);
  float* lo = reinterpret_cast<float*>(p + l.lo_off);
  int* flag = reinterpret_cast<int*>(p + l.flag_off);
  CUtensorMap tm_hi, tm_lo;
  if (!make_plane_map(&tm_hi, hi, K, l.pitch) || !make_plane_map(&tm_lo, lo, K, l.pitch))
    return cudaErrorNotSupported;
  cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  const int k = static_cast<int>(K);
  const int nbj = (k + TC_BN - 1) / TC_BN;
  const unsigned tiles = static_cast<unsigned>(tc_tile_count(k));
  int launches = 0;
  int round = 0;
  const int tri = (lower_tri && T == K && T <= l.chunk) ? 1 : 0;  // one chunk: block index == contraction index
  for (long long t0 = 0; t0 < T; t0 += l.chunk, ++round) {
    const long long tokens = T - t0 < l.chunk ? T - t0 : l.chunk;
    const long long used = round_up(tokens, TC_BK);  // pad columns up to `used` are zeroed
    dim3 sgrid(static_cast<unsigned>(used / 32), static_cast<unsigned>((k + 31) / 32));
    xtx_split_transpose<<<sgrid, 256, 0, st>>>(x + t0 * K, tokens, k, l.pitch, hi, lo, flag);
    xtx_tc_gemm<<<tiles, TC_THREADS, TC_SMEM_BYTES, st>>>(tm_hi, tm_lo, P, k,
                                                          static_cast<int>(used / TC_BK), nbj,
                                                          round > 0 ? 1 : 0, flag, tri);
    launches += 2;
  }
  xtx_tc_finish<OutT><<<sm_count * 8, 256, 0, st>>>(P, k, alpha, out, flag);
  ++launches;
  *flag_out = flag;
  return count_launch(launches);
}

template cudaError_t launch_xtx_tc<double>(const float*, long long, long long, double, double*,
                                           void*, int, const int**, cudaStream_t, int);
template cudaError_t launch_xtx_tc<float>(const float*, long long, long long, double, float*, void*,
                                          int, const int**, cudaStream_t, int);

}  // namespace aeqb
