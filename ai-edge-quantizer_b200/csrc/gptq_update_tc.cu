// GPTQ's inter-block update on the 5th-generation tensor cores, LEFT-looking.
//
// gptq._apply_gptq (algorithms/uniform_quantize/gptq.py:208-214) is right-looking: after every
// 64-column block it subtracts  Err_block @ Hinv[block, end:]  from all the columns still to come,
// i.e. a contraction of length 64 per launch and one read-modify-write of the trailing weight
// per block (R*K^2/64 bytes of traffic, 15 TFLOP/s on SIMT FMAs).  The same numbers, regrouped:
// just before block b is quantised its 64 columns need
//      W[:, b] -= Err[:, :64 b] @ Hinv[:64 b, b]
// ONE product per block with a contraction as long as everything quantised so far, whose output
// is only [R, 64] and is consumed once.  Err lives as two TF32 planes [R, K] (written by the
// column kernel as it produces them), Hinv is symmetric, so BOTH operands are K-major tiles of
// plain row-major planes:  C[r, c] = sum_l Err[r, l] * Hinv[64 b + c, l].
//
// One CTA per (128-row tile, split of the contraction): warp 0 lane 0 feeds a 4-stage ring with
// TMA boxes (A hi/lo [128 x 32], B hi/lo [64 x 32], SWIZZLE_128B), warp 1 lane 0 issues 3xTF32
// tcgen05.mma 128x64x8 into one of two 64-column TMEM accumulators, warps 2-5 drain each
// finished segment into fp32 registers with round-to-nearest adds (the tensor core truncates its
// own accumulation, see xtx_tc.cu) and finally write the CTA's partial product to
// part[split][R][64]; the column kernel subtracts the partials in split order, so the result
// does not depend on scheduling.  Tolerance: the reference rounds after every 64-long sgemm
// block, this sums the whole contraction first (tools/gptq_left_looking_study.py).
#include "aeqb_common.cuh"
#include "aeqb_kernels.h"
#include "aeqb_tc.cuh"

namespace aeqb {

using namespace tc;

namespace {

constexpr int GU_BM = 128;     // rows per tile (UMMA M)
constexpr int GU_BN = 64;      // the block's columns (UMMA N)
constexpr int GU_BK = 32;      // contraction values per stage = one 128-byte swizzle row
constexpr int GU_UK = 8;       // per tcgen05.mma kind::tf32
constexpr int GU_STAGES = 4;
constexpr uint32_t GU_A_BYTES = GU_BM * GU_BK * 4;   // 16 KiB per plane
constexpr uint32_t GU_B_BYTES = GU_BN * GU_BK * 4;   // 8 KiB per plane
constexpr uint32_t GU_STAGE_BYTES = 2 * GU_A_BYTES + 2 * GU_B_BYTES;  // 48 KiB
constexpr uint32_t gu_smem_bytes(int stages) { return stages * GU_STAGE_BYTES + 1024 + 256; }
constexpr int GU_EPI_WARPS = 4;
constexpr int GU_THREADS = (2 + GU_EPI_WARPS) * 32;
constexpr int GU_SEG_KB = 4;   // stages per accumulation segment (128 contraction values)
constexpr uint32_t kIdescGu = (1u << 4) | (2u << 7) | (2u << 10) |
                              (static_cast<uint32_t>(GU_BN >> 3) << 17) |
                              (static_cast<uint32_t>(GU_BM >> 4) << 24);

// x -> TF32 planes (elementwise): Hinv once per call.
__global__ void __launch_bounds__(256)
    split_planes_kernel(const float* __restrict__ x, long long n, float* __restrict__ hi,
                        float* __restrict__ lo) {
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += step) {
    const float v = x[i];
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    const float rest = v - __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(rest));
    hi[i] = __uint_as_float(h);
    lo[i] = __uint_as_float(l);
  }
}

// part[blockIdx.y][r, c] = sum_{l in this split} Err[r, l] * Hinv[c0 + c, l]
template <int STAGES>
__global__ void __launch_bounds__(GU_THREADS, 1)
    gptq_update_tc_kernel(const __grid_constant__ CUtensorMap tm_ehi, const __grid_constant__ CUtensorMap tm_elo,
                          const __grid_constant__ CUtensorMap tm_hhi, const __grid_constant__ CUtensorMap tm_hlo,
                          float* __restrict__ part, int R, int c0, int kb_total) {
  extern __shared__ uint8_t gu_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(gu_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * GU_STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * GU_BM;
  // this split's stage range: an even share of the kb_total stages, earlier splits take the rest
  const int ns = gridDim.y, sp = blockIdx.y;
  const int base_kb = kb_total / ns, extra = kb_total % ns;
  const int kb0 = sp * base_kb + min(sp, extra);
  const int nkb = base_kb + (sp < extra ? 1 : 0);
  const int nseg = (nkb + GU_SEG_KB - 1) / GU_SEG_KB;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_ehi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_elo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_hhi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_hlo)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], GU_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) {  // two 128-lane x 64-column fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(128u)
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
        const int s = kb % STAGES;
        const uint32_t ph = static_cast<uint32_t>(kb / STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full[s], GU_STAGE_BYTES);
        uint8_t* base = smem + s * GU_STAGE_BYTES;
        const int l = (kb0 + kb) * GU_BK;
        tma_load_2d(base, &tm_ehi, l, r0, &full[s]);                                 // A hi
        tma_load_2d(base + GU_A_BYTES, &tm_elo, l, r0, &full[s]);                    // A lo
        tma_load_2d(base + 2 * GU_A_BYTES, &tm_hhi, l, c0, &full[s]);                // B hi
        tma_load_2d(base + 2 * GU_A_BYTES + GU_B_BYTES, &tm_hlo, l, c0, &full[s]);   // B lo
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
        const uint32_t tacc = tmem_base + static_cast<uint32_t>(buf * GU_BN);
        const int kb_end = min(nkb, kb + GU_SEG_KB);
        bool first = true;
        for (; kb < kb_end; ++kb) {
          const int s = kb % STAGES;
          const uint32_t ph = static_cast<uint32_t>(kb / STAGES) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + s * GU_STAGE_BYTES);
          const uint64_t a_hi = tc_smem_desc(base);
          const uint64_t a_lo = tc_smem_desc(base + GU_A_BYTES);
          const uint64_t b_hi = tc_smem_desc(base + 2 * GU_A_BYTES);
          const uint64_t b_lo = tc_smem_desc(base + 2 * GU_A_BYTES + GU_B_BYTES);
#pragma unroll
          for (int k = 0; k < GU_BK / GU_UK; ++k) {  // cross terms first
            const uint64_t off = static_cast<uint64_t>((k * GU_UK * 4) >> 4);
            tc_mma_tf32(tacc, a_hi + off, b_lo + off, kIdescGu, first ? 0u : 1u);
            first = false;
            tc_mma_tf32(tacc, a_lo + off, b_hi + off, kIdescGu, 1u);
          }
#pragma unroll
          for (int k = 0; k < GU_BK / GU_UK; ++k) {
            const uint64_t off = static_cast<uint64_t>((k * GU_UK * 4) >> 4);
            tc_mma_tf32(tacc, a_hi + off, b_hi + off, kIdescGu, 1u);
          }
          tc_commit(&empty[s]);
        }
        tc_commit(&tmem_full[buf]);
      }
    }
    __syncwarp();
  } else {  // ---------------- epilogue: warp w reads TMEM lane quarter w % 4
    const int q = warp & 3;
    float acc[GU_BN];
#pragma unroll
    for (int u = 0; u < GU_BN; ++u) acc[u] = 0.0f;
    for (int seg = 0; seg < nseg; ++seg) {
      const int buf = seg & 1;
      mbar_wait(&tmem_full[buf], static_cast<uint32_t>(seg >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * GU_BN);
#pragma unroll
      for (int c = 0; c < GU_BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
#pragma unroll
        for (int u = 0; u < 32; ++u) acc[c * 32 + u] = __fadd_rn(acc[c * 32 + u], __uint_as_float(v[u]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
    const int r = r0 + q * 32 + lane;
    if (r < R) {
      float4* dst = reinterpret_cast<float4*>(part + (static_cast<long long>(sp) * R + r) * GU_BN);
#pragma unroll
      for (int u = 0; u < GU_BN / 4; ++u) dst[u] = make_float4(acc[4 * u], acc[4 * u + 1], acc[4 * u + 2], acc[4 * u + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u)
                 : "memory");
  }
}

// [rows, K] fp32 plane (row pitch K), box = 32 contraction values (128 B, swizzled) x box_rows.
bool make_map(CUtensorMap* m, const float* plane, long long rows, long long K, int box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * 4};
  const cuuint32_t box[2] = {GU_BK, static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(plane), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool gptq_update_tc_eligible(long long R, long long K) {
  static const bool off = getenv("AEQB_GPTQ_SIMT") && atoi(getenv("AEQB_GPTQ_SIMT"));
  static const long long min_k = getenv("AEQB_GPTQ_TC_MIN_K") ? atoll(getenv("AEQB_GPTQ_TC_MIN_K")) : 1024;
  return !off && K % 64 == 0 && K >= min_k && R >= 32 && encode_tiled_fn() != nullptr;
}

int gptq_update_tc_max_splits() { return 8; }

// Splits of the contraction so that (row tiles x splits) fills the SMs once; every split keeps
// at least two stages.
int gptq_update_tc_splits(long long R, int kb_total, int sm_count) {
  const long long tiles = (R + GU_BM - 1) / GU_BM;
  long long s = sm_count / tiles;
  if (s < 1) s = 1;
  if (s > gptq_update_tc_max_splits()) s = gptq_update_tc_max_splits();
  if (s > kb_total / 2) s = kb_total / 2;
  if (s < 1) s = 1;
  return static_cast<int>(s);
}

cudaError_t launch_split_planes(const float* x, long long n, float* hi, float* lo, int sm_count, cudaStream_t st) {
  split_planes_kernel<<<sm_count * 8, 256, 0, st>>>(x, n, hi, lo);
  return count_launch();
}

struct GptqTcMaps {
  CUtensorMap ehi, elo, hhi, hlo;
};

cudaError_t gptq_update_tc_prepare(GptqTcMapsOpaque* out, const float* err_hi, const float* err_lo,
                                   const float* h_hi, const float* h_lo, long long R, long long K) {
  static_assert(sizeof(GptqTcMapsOpaque) >= sizeof(GptqTcMaps), "opaque storage too small");
  GptqTcMaps* m = reinterpret_cast<GptqTcMaps*>(out);
  static PerDevice attr_done;
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(gptq_update_tc_kernel<GU_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(gu_smem_bytes(GU_STAGES)));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gptq_update_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(gu_smem_bytes(2)));
    if (e != cudaSuccess) return e;
    attr_done.set();
  }
  if (!make_map(&m->ehi, err_hi, R, K, GU_BM) || !make_map(&m->elo, err_lo, R, K, GU_BM) ||
      !make_map(&m->hhi, h_hi, K, K, GU_BN) || !make_map(&m->hlo, h_lo, K, K, GU_BN))
    return cudaErrorNotSupported;
  return cudaSuccess;
}

// part[s][R][64] = split s of  Err[:, :L] @ Hinv[:L, c0 : c0 + 64]   (L a multiple of 64).
// light: two pipeline stages (97 KiB of shared memory instead of 193) so that the column kernel's CTAs
// fit beside it on the same SMs — the lookahead runs this product WHILE the previous block's columns
// are being quantised.
cudaError_t launch_gptq_update_tc(const GptqTcMapsOpaque* maps, float* part, long long R, int c0, int L,
                                  int n_splits, cudaStream_t st, int light) {
  const GptqTcMaps* m = reinterpret_cast<const GptqTcMaps*>(maps);
  const dim3 grid(static_cast<unsigned>((R + GU_BM - 1) / GU_BM), static_cast<unsigned>(n_splits));
  if (light)
    gptq_update_tc_kernel<2><<<grid, GU_THREADS, gu_smem_bytes(2), st>>>(m->ehi, m->elo, m->hhi, m->hlo, part,
                                                                        static_cast<int>(R), c0, L / GU_BK);
  else
    gptq_update_tc_kernel<GU_STAGES><<<grid, GU_THREADS, gu_smem_bytes(GU_STAGES), st>>>(
        m->ehi, m->elo, m->hhi, m->hlo, part, static_cast<int>(R), c0, L / GU_BK);
  return count_launch();
}

}  // namespace aeqb
