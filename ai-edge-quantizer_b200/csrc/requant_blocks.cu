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
// Because cols % block == 0, a [rows, cols] matrix is a flat stream of blocks and
// its scale tensor [rows, cols/block] is that stream's block index; a batch of
// tensors is the concatenation of their streams (one persistent launch for a
// whole model's weight buffers).  Tile-stream: one producer lane keeps a 3-stage
// shared-memory ring full with 1-D TMA bulk copies; every consumer lane owns 8
// consecutive floats (two float4, loaded in a swizzled order so LDS.128 is
// bank-conflict free), block/8 lanes share a block, and the data never leaves
// registers between the |x| max and the quantise step: shared memory is read once.
//
// Fast path (no clipping constants, fp16-normal scale): x/scale via the hoisted
// exact divide (3 FFMA), rint via one magic-number FADD, nibbles assembled by
// IMAD Horner steps + one PRMT; the clip to [qmin, qmax] provably never binds
// (|x|/scale <= qmax / (1 - 2^-9) < qmax + 0.5).  Anything else (clip given,
// scale flushed to 0 / subnormal / inf / NaN) takes the IEEE-divide path.
#include "aeqb_common.cuh"
#include "aeqb_kernels.h"

namespace aeqb {

namespace {

constexpr int kStages = 3;
// 64 KiB stages, 16 consumer warps, ONE CTA per SM (round 2, measured back to back on one box with
// tools/sustain_ab.py, SUSTAIN_KIND=blocks, 477 x [4096,4096] INT4 block-32): 0.98 of the measured copy peak
// over 0.5 s (1.00-1.01 at the full clock, 0.97 power-capped) against 0.947 flat with 32 KiB stages, 8 warps
// and two CTAs per SM -- fewer, larger bulk copies per byte, as with the 96 KiB tiles of the rows kernel.
constexpr int kStageBytes = 65536;
constexpr int kStageFloats = kStageBytes / 4;
constexpr int kNW = 16;
constexpr int kSlice = 256;  // floats handled by one warp iteration

struct StageDesc {
  BlocksJob job;
  long long e0;  // first element of the tile inside job.x
  int ne;        // elements in the tile
};

// uqt:552-563 + :577-581 for one block (clip / degenerate path).
__device__ __forceinline__ float block_scale_slow(float amax, const float* clip, long long blk,
                                                  int bits, float qmax, uint16_t* f16) {
  float bound = max_nan(amax, 1e-9f);
  if (clip) {
    // uqt:529-550: with clipping values the bound also stays inside what an fp16
    // scale can represent.
    const float c = clip[blk];
    const float hi = min_nan(c, 65280.0f * static_cast<float>((1 << bits) - 1));
    const float lo = max_nan(-c, -65280.0f * static_cast<float>(1 << bits));
    bound = min_nan(max_nan(bound, lo), hi);
  }
  return round_scale_bf16_f16(__fdiv_rn(bound, qmax), f16);
}

template <int LPB>  // lanes per block = block / 8
__device__ __forceinline__ float group_max_nan(float v) {
#pragma unroll
  for (int o = 1; o < LPB; o <<= 1) v = max_nan(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// MIRROR: the fp16 scales of a sharded job set also go to the other ranks' copies of the gathered
// scale buffer (NVLink peer memory, `b.peers`), which is the all-gather of blockwise scales
// (2 B per 32 weights, the one exchange of a blockwise model: quantize_tensor.py:107-147 needs
// every tensor's scales to serialise it) without a collective launch.  To make those remote
// stores worth their NVLink packets a warp then takes CONTIGUOUS slices of the tile, parks its
// scales in a private shared-memory strip and sends the strip to every peer as one run of
// coalesced 4-byte stores (64 B per peer and tile for block 32) instead of 2 bytes at a time.
template <int BLOCK, bool OUT_Q, bool OUT_P, bool MIRROR>
__global__ void __launch_bounds__((kNW + 1) * 32, 1)
    requant_blocks_stream(const __grid_constant__ BlocksBatch b) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ StageDesc desc[kStages];
  __shared__ __align__(16) uint16_t strip[MIRROR ? kNW : 1][MIRROR ? (kStageFloats / BLOCK / kNW) : 1];
  constexpr int LPB = BLOCK / 8;
  constexpr int kSliceBlocks = kSlice / BLOCK;               // scales one slice produces
  constexpr int kSlicesPerWarp = kStageFloats / kSlice / kNW;  // contiguous split of a full tile

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler
  const int lane = tid & 31;
  const long long n_tiles = b.n_tiles;

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
      int j = 0;
      const BlocksJob* jobs = b.table != nullptr ? b.table : b.jobs;  // device table or kernel parameters
      const long long* ends = b.table_ends;
      BlocksJob job = jobs[0];
      int s = 0;
      uint32_t round = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
        if (ends != nullptr) {
          if (tile >= job.tile_end) {
            while (tile >= __ldg(ends + j)) ++j;
            job = jobs[j];
          }
        } else {
          while (tile >= job.tile_end) job = jobs[++j];
        }
        const long long e0 = (tile - job.tile0) * kStageFloats;
        const long long ne = min(static_cast<long long>(kStageFloats), job.n - e0);
        desc[s].job = job;
        desc[s].e0 = e0;
        desc[s].ne = static_cast<int>(ne);
        const uint32_t bytes = static_cast<uint32_t>(ne * 4);
        mbar_arrive_expect_tx(&full_bar[s], bytes);  // release: desc is visible to waiters
        bulk_g2s(smem_raw + static_cast<size_t>(s) * kStageBytes, job.x + e0, bytes, &full_bar[s]);
        if (++s == kStages) { s = 0; ++round; }
      }
    }
    return;
  }

  // ---------------- consumers
  const QRange qr = qrange(b.bits, true);
  const DivBy dq = make_recip(qr.qmax);  // bound / qmax with a hoisted reciprocal
  const int sw = (lane >> 2) & 1;            // load-order swizzle
  const uint32_t sel = sw ? 0x1054u : 0x5410u;  // PRMT: which float4 holds the low elements
  const bool leader = (lane & (LPB - 1)) == 0;
  int s = -1;
  uint32_t ph = 1;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (++s == kStages) s = 0;
    if (s == 0) ph ^= 1;
    const float4* t4 = reinterpret_cast<const float4*>(smem_raw + static_cast<size_t>(s) * kStageBytes);

    mbar_wait(&full_bar[s], ph);

    const long long e0 = desc[s].e0;
    const int ne = desc[s].ne;
    const float* clip = desc[s].job.clip;
    int8_t* qp = OUT_Q ? desc[s].job.q + e0 : nullptr;
    uint8_t* pp = OUT_P ? desc[s].job.packed + (e0 >> 1) : nullptr;
    float* sp = desc[s].job.scale;
    uint16_t* hp = desc[s].job.scale_f16;
    const long long blk0 = e0 / BLOCK;
    const bool w_scale = leader && sp != nullptr;
    const bool w_f16 = leader && hp != nullptr;
    const int nslices = (ne + kSlice - 1) / kSlice;

    const int sl_first = MIRROR ? warp * kSlicesPerWarp : warp;
    const int sl_last = MIRROR ? min(nslices, sl_first + kSlicesPerWarp) : nslices;
    for (int sl = sl_first; sl < sl_last; sl += (MIRROR ? 1 : kNW)) {
      const int f0 = sl * kSlice + lane * 8;  // first float of this lane inside the tile
      const bool valid = f0 < ne;             // whole blocks are valid or not (ne % BLOCK == 0)
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (valid) {
        va = t4[(f0 >> 2) + sw];        // sw == 0: elements 0..3, else 4..7
        vb = t4[(f0 >> 2) + (sw ^ 1)];
      }
      const float amax = group_max_nan<LPB>(absmax4(absmax4(0.0f, va), vb));
      if (!valid) continue;
      const int blk = f0 / BLOCK;

      // ---- scale (fast: bound/qmax through the hoisted divide, fp16-normal result)
      const float bound = max_nan(amax, 1e-9f);
      uint16_t h16;
      float scale = round_scale_bf16_f16(div_fast(bound, dq), &h16);
      const bool fast = (clip == nullptr) && (bound <= 1.0e30f) && (scale >= 6.103515625e-05f) &&
                        (scale <= 65504.0f);
      uint32_t word_q0 = 0, word_q1 = 0, word_p = 0;
      if (fast) {
        const DivBy d = make_recip(scale);
        const float2 t01 = div_fast2(make_float2(va.x, va.y), d.b, d.y);
        const float2 t23 = div_fast2(make_float2(va.z, va.w), d.b, d.y);
        const float2 t45 = div_fast2(make_float2(vb.x, vb.y), d.b, d.y);
        const float2 t67 = div_fast2(make_float2(vb.z, vb.w), d.b, d.y);
        if (OUT_P) {
          const uint2 r01 = rmagic2(t01, kMagicPlus8), r23 = rmagic2(t23, kMagicPlus8);
          const uint2 r45 = rmagic2(t45, kMagicPlus8), r67 = rmagic2(t67, kMagicPlus8);
          const uint32_t ha = nibbles4_biased(r01.x, r01.y, r23.x, r23.y);
          const uint32_t hb = nibbles4_biased(r45.x, r45.y, r67.x, r67.y);
          word_p = __byte_perm(ha, hb, sel) ^ 0x88888888u;
        }
        if (OUT_Q) {
          const uint2 r01 = rmagic2(t01, kMagic), r23 = rmagic2(t23, kMagic);
          const uint2 r45 = rmagic2(t45, kMagic), r67 = rmagic2(t67, kMagic);
          word_q0 = bytes4(r01.x, r01.y, r23.x, r23.y);
          word_q1 = bytes4(r45.x, r45.y, r67.x, r67.y);
        }
      } else {
        scale = block_scale_slow(amax, clip, blk0 + blk, b.bits, qr.qmax, &h16);
        int qa[4], qb[4];
        qa[0] = clampi(rni(__fdiv_rn(va.x, scale)), qr.lo, qr.hi);
        qa[1] = clampi(rni(__fdiv_rn(va.y, scale)), qr.lo, qr.hi);
        qa[2] = clampi(rni(__fdiv_rn(va.z, scale)), qr.lo, qr.hi);
        qa[3] = clampi(rni(__fdiv_rn(va.w, scale)), qr.lo, qr.hi);
        qb[0] = clampi(rni(__fdiv_rn(vb.x, scale)), qr.lo, qr.hi);
        qb[1] = clampi(rni(__fdiv_rn(vb.y, scale)), qr.lo, qr.hi);
        qb[2] = clampi(rni(__fdiv_rn(vb.z, scale)), qr.lo, qr.hi);
        qb[3] = clampi(rni(__fdiv_rn(vb.w, scale)), qr.lo, qr.hi);
        if (OUT_Q) {
          word_q0 = pack_i8x4(qa[0], qa[1], qa[2], qa[3]);
          word_q1 = pack_i8x4(qb[0], qb[1], qb[2], qb[3]);
        }
        if (OUT_P) {
          const uint32_t ha = (qa[0] & 0xF) | ((qa[1] & 0xF) << 4) | ((qa[2] & 0xF) << 8) | ((qa[3] & 0xF) << 12);
          const uint32_t hb = (qb[0] & 0xF) | ((qb[1] & 0xF) << 4) | ((qb[2] & 0xF) << 8) | ((qb[3] & 0xF) << 12);
          word_p = __byte_perm(ha, hb, sel);
        }
      }
      if (OUT_Q)
        *reinterpret_cast<uint2*>(qp + f0) = sw ? make_uint2(word_q1, word_q0) : make_uint2(word_q0, word_q1);
      if (OUT_P) *reinterpret_cast<uint32_t*>(pp + (f0 >> 1)) = word_p;
      if (w_scale) sp[blk0 + blk] = scale;
      if (w_f16) hp[blk0 + blk] = h16;
      if (MIRROR && leader) strip[warp][blk - sl_first * kSliceBlocks] = h16;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);  // the tile is consumed; the strip is this warp's own
    if (MIRROR && hp != nullptr) {
      const int b_first = sl_first * kSliceBlocks;
      const int nb = min(ne / BLOCK, b_first + kSlicesPerWarp * kSliceBlocks) - b_first;  // scales in the strip
      if (nb > 0) {
        uint16_t* dst0 = hp + blk0 + b_first;
        const int np = b.peers.n;
        if ((nb & 1) == 0 && (reinterpret_cast<uintptr_t>(dst0) & 3) == 0) {
          const int words = nb >> 1;
          const uint32_t* src = reinterpret_cast<const uint32_t*>(strip[warp]);
          for (int i = lane; i < np * words; i += 32) {
            const int p = i / words, wd = i - p * words;
            reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(dst0) + b.peers.delta[p])[wd] = src[wd];
          }
        } else {
          for (int i = lane; i < np * nb; i += 32) {
            const int p = i / nb, k = i - p * nb;
            reinterpret_cast<uint16_t*>(reinterpret_cast<char*>(dst0) + b.peers.delta[p])[k] = strip[warp][k];
          }
        }
      }
      __syncwarp();
    }
  }
}

// Generic fallback: one warp per block, scalar global loads, any alignment.
__global__ void __launch_bounds__(256)
    requant_blocks_generic(const BlocksJob a, int block, int bits) {
  const int lane = threadIdx.x & 31;
  const long long blk = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long n_blocks = a.n / block;
  if (blk >= n_blocks) return;
  const QRange qr = qrange(bits, true);
  const float* x = a.x + blk * block;
  float amax = 0.0f;
  for (int i = lane; i < block; i += 32) amax = max_nan(amax, fabsf(x[i]));
  amax = warp_max_nan(amax);
  uint16_t h16;
  const float scale = block_scale_slow(amax, a.clip, blk, bits, qr.qmax, &h16);
  if (lane == 0) {
    if (a.scale) a.scale[blk] = scale;
    if (a.scale_f16) a.scale_f16[blk] = h16;
  }
  for (int i = lane * 2; i < block; i += 64) {
    const int q0 = clampi(rni(__fdiv_rn(x[i], scale)), qr.lo, qr.hi);
    const int q1 = clampi(rni(__fdiv_rn(x[i + 1], scale)), qr.lo, qr.hi);
    const long long e = blk * block + i;
    if (a.q) {
      a.q[e] = static_cast<int8_t>(q0);
      a.q[e + 1] = static_cast<int8_t>(q1);
    }
    if (a.packed) a.packed[e >> 1] = static_cast<uint8_t>((q0 & 0xF) | ((q1 & 0xF) << 4));
  }
}

template <int BLOCK, bool OUT_Q, bool OUT_P, bool MIRROR = false>
cudaError_t launch_stream(const BlocksBatch& b, int sm_count, cudaStream_t st) {
  auto kern = requant_blocks_stream<BLOCK, OUT_Q, OUT_P, MIRROR>;
  const int smem = kStages * kStageBytes;
  static PerDevice configured;
  if (!configured.done()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured.set();
  }
  long long grid = static_cast<long long>(sm_count);  // one CTA (3 x 64 KiB of stages) per SM
  if (grid > b.n_tiles) grid = b.n_tiles;
  kern<<<static_cast<unsigned>(grid), (kNW + 1) * 32, smem, st>>>(b);
  return count_launch();
}

template <int BLOCK>
cudaError_t launch_stream_out(const BlocksBatch& b, bool out_q, bool out_p, int sm_count,
                              cudaStream_t st) {
  if (b.peers.n > 0) {  // sharded model: packed payload stays local, fp16 scales reach every rank
    if (out_q || !out_p) return cudaErrorInvalidValue;
    return launch_stream<BLOCK, false, true, true>(b, sm_count, st);
  }
  if (out_q && out_p) return launch_stream<BLOCK, true, true>(b, sm_count, st);
  if (out_p) return launch_stream<BLOCK, false, true>(b, sm_count, st);
  if (out_q) return launch_stream<BLOCK, true, false>(b, sm_count, st);
  return launch_stream<BLOCK, false, false>(b, sm_count, st);  // scales only
}

}  // namespace

long long blocks_job_tiles(long long n) { return (n + kStageFloats - 1) / kStageFloats; }

bool blocks_job_streamable(const BlocksJob& j) {
  return (reinterpret_cast<uintptr_t>(j.x) % 16 == 0) &&
         (!j.q || reinterpret_cast<uintptr_t>(j.q) % 8 == 0) &&
         (!j.packed || reinterpret_cast<uintptr_t>(j.packed) % 4 == 0);
}

cudaError_t launch_requant_blocks_stream(const BlocksBatch& b, bool out_q, bool out_p,
                                         int sm_count, cudaStream_t st) {
  if (b.n_tiles <= 0) return cudaSuccess;
  switch (b.block) {
    case 32: return launch_stream_out<32>(b, out_q, out_p, sm_count, st);
    case 64: return launch_stream_out<64>(b, out_q, out_p, sm_count, st);
    case 128: return launch_stream_out<128>(b, out_q, out_p, sm_count, st);
    case 256: return launch_stream_out<256>(b, out_q, out_p, sm_count, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_requant_blocks_generic(const BlocksJob& j, int block, int bits,
                                          cudaStream_t st) {
  if (j.n <= 0) return cudaSuccess;
  const int warps = 8;
  const long long n_blocks = j.n / block;
  const long long grid = (n_blocks + warps - 1) / warps;
  requant_blocks_generic<<<static_cast<unsigned>(grid), warps * 32, 0, st>>>(j, block, bits);
  return count_launch();
}

}  // namespace aeqb
