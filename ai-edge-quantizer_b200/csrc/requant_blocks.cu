// Fused blockwise symmetric requantisation (block = 32/64/128/256 along the
// input-feature axis): |x| max per block -> scale = bound/qmax rounded
// fp32->bf16->fp16 -> q = clip(rint(x/scale)) -> two nibbles per byte.
//
// Replaces, for FC / EMBEDDING weights with a BLOCKWISE_* granularity:
//   common_quantize.init_tensor_min_max, blockwise branch (common_quantize.py:1345-1358)
//   uniform_quantize_tensor.tensor_zp_scale_from_min_max    (uqt:492-586, :577-581)
//   uniform_quantize_tensor.uniform_quantize + _broadcast_scale_zp_for_blockwise
//                                                           (uqt:222-362)
//   transformation_utils.pack_data                           (:293-353)
//   quantize_tensor._perform_blockwise_quantization fp16 scale (:129-133)
// HBM traffic: 4 B read + 0.5 B packed + 2/block B fp16 scale per weight.
//
// Because cols % block == 0, the [rows, cols] matrix is a flat stream of
// blocks and the scale tensor [rows, cols/block] is that stream's block index.
// Tile-stream as in requant_rows.cu; here every lane owns 8 consecutive floats
// (two float4, loaded in a swizzled order so that LDS.128 is bank-conflict
// free), block/8 lanes share a block, and the block never leaves registers:
// shared memory is read once.
#include "aeqb_common.cuh"
#include "aeqb_kernels.h"

namespace aeqb {

namespace {

constexpr int kStages = 3;
constexpr int kStageBytes = 32768;
constexpr int kStageFloats = kStageBytes / 4;
constexpr int kNW = 8;
constexpr int kSlice = 256;  // floats handled by one warp iteration

struct BlockQ {
  DivBy div;
  float scale;
  uint16_t f16;
};

// uqt:552-563 + :577-581 for one block.
__device__ __forceinline__ BlockQ finalize_block(const BlocksArgs& a, long long blk, float amax) {
  const QRange qr = qrange(a.bits, true);
  float bound = max_nan(amax, 1e-9f);
  if (a.clip) {
    // uqt:529-550: with clipping values the bound is also kept inside what an
    // fp16 scale can represent.
    const float c = a.clip[blk];
    const float f16_hi = 65280.0f * static_cast<float>((1 << a.bits) - 1);
    const float f16_lo = -65280.0f * static_cast<float>(1 << a.bits);
    const float hi = min_nan(c, f16_hi);
    const float lo = max_nan(-c, f16_lo);
    bound = min_nan(max_nan(bound, lo), hi);
  }
  BlockQ r;
  r.scale = round_scale_bf16_f16(__fdiv_rn(bound, qr.qmax), &r.f16);
  r.div = make_div(r.scale, amax);
  return r;
}

template <int LPB>  // lanes per block = block / 8
__device__ __forceinline__ float group_max_nan(float v) {
#pragma unroll
  for (int o = 1; o < LPB; o <<= 1) v = max_nan(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <int BLOCK>
__global__ void __launch_bounds__((kNW + 1) * 32)
    requant_blocks_stream(const __grid_constant__ BlocksArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  constexpr int LPB = BLOCK / 8;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const long long n_tiles = a.n_tiles;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kNW);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == kNW) {  // ---------------- producer
    if (lane == 0) {
      long long it = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int s = static_cast<int>(it % kStages);
        const long long round = it / kStages;
        if (round > 0) mbar_wait(&empty_bar[s], static_cast<uint32_t>((round - 1) & 1));
        const long long e0 = tile * kStageFloats;
        const long long ne = min(static_cast<long long>(kStageFloats), a.n - e0);
        const uint32_t bytes = static_cast<uint32_t>(ne * 4);
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        bulk_g2s(smem_raw + static_cast<size_t>(s) * kStageBytes, a.x + e0, bytes, &full_bar[s]);
      }
    }
    return;
  }

  // ---------------- consumers
  const QRange qr = qrange(a.bits, true);
  const int sw = (lane >> 2) & 1;  // load-order swizzle
  long long it = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int s = static_cast<int>(it % kStages);
    const uint32_t ph = static_cast<uint32_t>((it / kStages) & 1);
    const long long e0 = tile * kStageFloats;
    const int ne = static_cast<int>(min(static_cast<long long>(kStageFloats), a.n - e0));
    const int nslices = (ne + kSlice - 1) / kSlice;
    const float4* t4 = reinterpret_cast<const float4*>(smem_raw + static_cast<size_t>(s) * kStageBytes);

    mbar_wait(&full_bar[s], ph);

    for (int sl = warp; sl < nslices; sl += kNW) {
      const int f0 = sl * kSlice + lane * 8;  // first float of this lane inside the tile
      const bool valid = f0 < ne;             // whole blocks are valid or not (ne % BLOCK == 0)
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (valid) {
        va = t4[(f0 >> 2) + sw];        // sw == 0: elements 0..3, else 4..7
        vb = t4[(f0 >> 2) + (sw ^ 1)];
      }
      float amax = absmax4(absmax4(0.0f, va), vb);
      amax = group_max_nan<LPB>(amax);
      const long long blk = (e0 + f0) / BLOCK;
      BlockQ bq;
      if (valid) bq = finalize_block(a, blk, amax);
      if (!valid) continue;

      int qa[4], qb[4];
      if (bq.div.fast) {
        qa[0] = clampi(rni(div_fast(va.x, bq.div)), qr.lo, qr.hi);
        qa[1] = clampi(rni(div_fast(va.y, bq.div)), qr.lo, qr.hi);
        qa[2] = clampi(rni(div_fast(va.z, bq.div)), qr.lo, qr.hi);
        qa[3] = clampi(rni(div_fast(va.w, bq.div)), qr.lo, qr.hi);
        qb[0] = clampi(rni(div_fast(vb.x, bq.div)), qr.lo, qr.hi);
        qb[1] = clampi(rni(div_fast(vb.y, bq.div)), qr.lo, qr.hi);
        qb[2] = clampi(rni(div_fast(vb.z, bq.div)), qr.lo, qr.hi);
        qb[3] = clampi(rni(div_fast(vb.w, bq.div)), qr.lo, qr.hi);
      } else {
        qa[0] = clampi(rni(__fdiv_rn(va.x, bq.scale)), qr.lo, qr.hi);
        qa[1] = clampi(rni(__fdiv_rn(va.y, bq.scale)), qr.lo, qr.hi);
        qa[2] = clampi(rni(__fdiv_rn(va.z, bq.scale)), qr.lo, qr.hi);
        qa[3] = clampi(rni(__fdiv_rn(va.w, bq.scale)), qr.lo, qr.hi);
        qb[0] = clampi(rni(__fdiv_rn(vb.x, bq.scale)), qr.lo, qr.hi);
        qb[1] = clampi(rni(__fdiv_rn(vb.y, bq.scale)), qr.lo, qr.hi);
        qb[2] = clampi(rni(__fdiv_rn(vb.z, bq.scale)), qr.lo, qr.hi);
        qb[3] = clampi(rni(__fdiv_rn(vb.w, bq.scale)), qr.lo, qr.hi);
      }
      const long long e = e0 + f0;
      if (a.q) {
        const uint32_t wa = pack_i8x4(qa[0], qa[1], qa[2], qa[3]);
        const uint32_t wb = pack_i8x4(qb[0], qb[1], qb[2], qb[3]);
        *reinterpret_cast<uint2*>(a.q + e) = sw ? make_uint2(wb, wa) : make_uint2(wa, wb);
      }
      if (a.packed) {  // bits == 4
        const uint32_t ha = (qa[0] & 0xF) | ((qa[1] & 0xF) << 4) | ((qa[2] & 0xF) << 8) | ((qa[3] & 0xF) << 12);
        const uint32_t hb = (qb[0] & 0xF) | ((qb[1] & 0xF) << 4) | ((qb[2] & 0xF) << 8) | ((qb[3] & 0xF) << 12);
        *reinterpret_cast<uint32_t*>(a.packed + (e >> 1)) = sw ? (hb | (ha << 16)) : (ha | (hb << 16));
      }
      if ((lane & (LPB - 1)) == 0) {
        if (a.scale) a.scale[blk] = bq.scale;
        if (a.scale_f16) a.scale_f16[blk] = bq.f16;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }
}

// Generic fallback: one warp per block, scalar global loads, any alignment.
__global__ void __launch_bounds__(256) requant_blocks_generic(const __grid_constant__ BlocksArgs a) {
  const int lane = threadIdx.x & 31;
  const long long blk = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long n_blocks = a.n / a.block;
  if (blk >= n_blocks) return;
  const QRange qr = qrange(a.bits, true);
  const float* x = a.x + blk * a.block;
  float amax = 0.0f;
  for (int i = lane; i < a.block; i += 32) amax = max_nan(amax, fabsf(x[i]));
  amax = warp_max_nan(amax);
  const BlockQ bq = finalize_block(a, blk, amax);
  if (lane == 0) {
    if (a.scale) a.scale[blk] = bq.scale;
    if (a.scale_f16) a.scale_f16[blk] = bq.f16;
  }
  for (int i = lane * 2; i < a.block; i += 64) {
    const int q0 = clampi(rni(div_any(x[i], bq.div)), qr.lo, qr.hi);
    const int q1 = clampi(rni(div_any(x[i + 1], bq.div)), qr.lo, qr.hi);
    const long long e = blk * a.block + i;
    if (a.q) {
      a.q[e] = static_cast<int8_t>(q0);
      a.q[e + 1] = static_cast<int8_t>(q1);
    }
    if (a.packed) a.packed[e >> 1] = static_cast<uint8_t>((q0 & 0xF) | ((q1 & 0xF) << 4));
  }
}

template <int BLOCK>
cudaError_t launch_stream(BlocksArgs a, int sm_count, cudaStream_t st) {
  auto kern = requant_blocks_stream<BLOCK>;
  const int smem = kStages * kStageBytes;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  a.n_tiles = (a.n + kStageFloats - 1) / kStageFloats;
  long long grid = static_cast<long long>(sm_count) * 2;
  if (grid > a.n_tiles) grid = a.n_tiles;
  kern<<<static_cast<unsigned>(grid), (kNW + 1) * 32, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_requant_blocks(BlocksArgs a, int sm_count, cudaStream_t st) {
  if (a.n <= 0) return cudaSuccess;
  const bool aligned = (reinterpret_cast<uintptr_t>(a.x) % 16 == 0) &&
                       (!a.q || reinterpret_cast<uintptr_t>(a.q) % 8 == 0) &&
                       (!a.packed || reinterpret_cast<uintptr_t>(a.packed) % 4 == 0);
  if (aligned) {
    switch (a.block) {
      case 32: return launch_stream<32>(a, sm_count, st);
      case 64: return launch_stream<64>(a, sm_count, st);
      case 128: return launch_stream<128>(a, sm_count, st);
      case 256: return launch_stream<256>(a, sm_count, st);
      default: break;
    }
  }
  const int warps = 8;
  const long long n_blocks = a.n / a.block;
  const long long grid = (n_blocks + warps - 1) / warps;
  requant_blocks_generic<<<static_cast<unsigned>(grid), warps * 32, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace aeqb
