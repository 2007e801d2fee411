// Fused per-channel (and, with broadcast min/max, per-tensor) requantisation:
//   min/max over each row -> scale / zero-point -> q = clip(rint(x/scale + zp))
//   -> int8 (one value per byte) and/or INT4 / INT2 packed.
//
// Replaces, for a 2-D fp32 weight [rows, cols] quantised along dim 0:
//   common_quantize.init_tensor_min_max         (common_quantize.py:1311-1359)
//   uniform_quantize_tensor.tensor_zp_scale_from_min_max   (uqt:492-586)
//   uniform_quantize_tensor.uniform_quantize               (uqt:273-362)
//   transformation_utils.pack_data                          (:293-353)
// i.e. naive_min_max_quantize.get_tensor_quant_params (:34-110) in one pass
// over HBM: 4 B read + 1 B (or 0.5 B) written per weight.
//
// Tile-stream design (sm_100a): a persistent CTA owns every gridDim-th tile of
// whole rows.  One producer lane keeps a 3-stage shared-memory ring full with
// 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx); consumer warps
// read the tile twice from shared memory (pass 1: NaN-propagating |x| max or
// min&max per row; pass 2: exact divide, rint, clip, pack) so HBM is touched
// once.  Rows are cut into 128-float warp chunks: a warp instruction reads 512
// contiguous bytes (conflict-free LDS.128) and writes 128 contiguous bytes.
#include "aeqb_common.cuh"
#include "aeqb_kernels.h"

namespace aeqb {

namespace {

constexpr int kStages = 3;
constexpr int kMaxRowsPerTile = 64;
constexpr int kChunk = 128;  // floats per warp chunk

struct RowAcc {  // merged across warps with shared-memory atomics
  unsigned amax_bits;
  int mn_ord, mx_ord;
  int nan;
};

struct RowQ {  // per-row quantisation constants for pass 2
  DivBy div;
  float zp;
};

__device__ __forceinline__ void acc_reset(RowAcc& a) {
  a.amax_bits = 0u;
  a.mn_ord = 0x7f800000;             // f2ord(+inf)
  a.mx_ord = (int)0x807fffff;        // f2ord(-inf)
  a.nan = 0;
}

// uqt:492-586 for one row.  `xmax` bounds |x| over the row (for the divide
// window); pass +inf when unknown.
__device__ __forceinline__ RowQ finalize_row(const RowsArgs& a, long long row, float mn, float mx,
                                            float xmax) {
  const QRange qr = qrange(a.bits, a.symmetric != 0);
  float scale, zpf = 0.0f;
  if (a.symmetric) {
    float bound = max_nan(max_nan(fabsf(mn), fabsf(mx)), 1e-9f);
    if (a.clip) {
      const float c = a.clip[row * a.clip_stride];
      bound = min_nan(max_nan(bound, -c), c);
    }
    scale = __fdiv_rn(bound, qr.qmax);
  } else {
    const float bmax = max_nan(mx, 0.0f);
    const float bmin = min_nan(mn, 0.0f);
    float bound = max_nan(__fsub_rn(bmax, bmin), 1e-9f);
    if (a.clip) {
      const float c = a.clip[row * a.clip_stride];
      bound = min_nan(max_nan(bound, -c), c);
    }
    scale = __fdiv_rn(bound, __fsub_rn(qr.qmax, qr.qmin));
    zpf = rintf(__fsub_rn(qr.qmin, __fdiv_rn(bmin, scale)));
  }
  if (a.scale) a.scale[row * a.out_stride] = scale;
  if (a.zp) a.zp[row * a.out_stride] = rni(zpf);
  RowQ r;
  r.div = make_div(scale, xmax);
  // int8 cast of the zero point (uqt:585) is the identity for finite inputs.
  r.zp = static_cast<float>(rni(zpf));
  return r;
}

__device__ __forceinline__ int quant1(float x, const RowQ& rq, bool sym, int lo, int hi) {
  float t = div_any(x, rq.div);
  if (!sym) t = __fadd_rn(t, rq.zp);
  return clampi(rni(t), lo, hi);
}

// ------------------------------------------------------------------ TMA tile stream
template <int STAGE_BYTES, int NW>
__global__ void __launch_bounds__((NW + 1) * 32)
    requant_rows_stream(const __grid_constant__ RowsArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ RowAcc s_acc[2][kMaxRowsPerTile];
  __shared__ RowQ s_rq[kMaxRowsPerTile];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int cols = a.cols;
  const int cpr = cols / kChunk;
  const int rpt = a.rows_per_tile;
  const long long n_tiles = a.n_tiles;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], NW);
    }
    mbar_fence_init();
  }
  if (tid < 2 * kMaxRowsPerTile) acc_reset(s_acc[tid / kMaxRowsPerTile][tid % kMaxRowsPerTile]);
  __syncthreads();

  if (warp == NW) {  // ---------------- producer
    if (lane == 0) {
      long long it = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int s = static_cast<int>(it % kStages);
        const long long round = it / kStages;
        if (round > 0) mbar_wait(&empty_bar[s], static_cast<uint32_t>((round - 1) & 1));
        const long long row0 = tile * rpt;
        const long long nrows = min(static_cast<long long>(rpt), a.rows - row0);
        const uint32_t bytes = static_cast<uint32_t>(nrows * cols * 4);
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        bulk_g2s(smem_raw + static_cast<size_t>(s) * STAGE_BYTES, a.x + row0 * cols, bytes,
                 &full_bar[s]);
      }
    }
    return;
  }

  // ---------------- consumers
  const bool sym = a.symmetric != 0;
  const bool given = a.given_min != nullptr;
  const QRange qr = qrange(a.bits, sym);
  long long it = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int s = static_cast<int>(it % kStages);
    const uint32_t ph = static_cast<uint32_t>((it / kStages) & 1);
    const int buf = static_cast<int>(it & 1);
    const long long row0 = tile * rpt;
    const int nrows = static_cast<int>(min(static_cast<long long>(rpt), a.rows - row0));
    const int nchunks = nrows * cpr;
    const float4* t4 = reinterpret_cast<const float4*>(smem_raw + static_cast<size_t>(s) * STAGE_BYTES);

    mbar_wait(&full_bar[s], ph);

    // ---- pass 1: per-row statistics
    if (!given) {
      int row = 0, rem = warp;
      while (rem >= cpr) { rem -= cpr; ++row; }
      int cur = row;
      bool any = false;
      float amax = 0.0f, mn = INFINITY, mx = -INFINITY;
      auto flush = [&](int r) {
        if (sym) {
          const float m = warp_max_nan(amax);
          if (lane == 0) atomicMax(&s_acc[buf][r].amax_bits, __float_as_uint(m));
        } else {
          const float lo = warp_min_nan(mn), hi = warp_max_nan(mx);
          if (lane == 0) {
            if (lo != lo || hi != hi) {
              atomicOr(&s_acc[buf][r].nan, 1);
            } else {
              atomicMin(&s_acc[buf][r].mn_ord, f2ord(lo));
              atomicMax(&s_acc[buf][r].mx_ord, f2ord(hi));
            }
          }
        }
      };
      for (int c = warp; c < nchunks; c += NW) {
        if (row != cur) {
          flush(cur);
          cur = row;
          amax = 0.0f; mn = INFINITY; mx = -INFINITY;
        }
        const float4 v = t4[c * 32 + lane];
        if (sym) {
          amax = absmax4(amax, v);
        } else {
          mn = min_nan(min_nan(mn, v.x), min_nan(v.y, min_nan(v.z, v.w)));
          mx = max_nan(max_nan(mx, v.x), max_nan(v.y, max_nan(v.z, v.w)));
        }
        any = true;
        rem += NW;
        while (rem >= cpr) { rem -= cpr; ++row; }
      }
      if (any) flush(cur);
    }
    named_bar_sync(1, NW * 32);

    // ---- per-row scale / zero point
    if (tid < nrows) {
      const long long grow = row0 + tid;
      float mn, mx, xmax;
      if (given) {
        mn = a.given_min[grow * a.mm_stride];
        mx = a.given_max[grow * a.mm_stride];
        xmax = INFINITY;  // row not scanned: always take the IEEE divide
      } else if (sym) {
        mx = __uint_as_float(s_acc[buf][tid].amax_bits);
        mn = -mx;
        xmax = mx;
      } else {
        const RowAcc acc = s_acc[buf][tid];
        mn = acc.nan ? NAN : ord2f(acc.mn_ord);
        mx = acc.nan ? NAN : ord2f(acc.mx_ord);
        xmax = max_nan(fabsf(mn), fabsf(mx));
      }
      s_rq[tid] = finalize_row(a, grow, mn, mx, xmax);
    }
    if (tid < kMaxRowsPerTile) acc_reset(s_acc[buf ^ 1][tid]);
    named_bar_sync(1, NW * 32);

    // ---- pass 2: quantise + store
    {
      int row = 0, rem = warp;
      while (rem >= cpr) { rem -= cpr; ++row; }
      const long long tile_elem0 = row0 * cols;
      for (int c = warp; c < nchunks; c += NW) {
        const RowQ rq = s_rq[row];
        const float4 v = t4[c * 32 + lane];
        int q0, q1, q2, q3;
        if (rq.div.fast && sym) {
          q0 = clampi(rni(div_fast(v.x, rq.div)), qr.lo, qr.hi);
          q1 = clampi(rni(div_fast(v.y, rq.div)), qr.lo, qr.hi);
          q2 = clampi(rni(div_fast(v.z, rq.div)), qr.lo, qr.hi);
          q3 = clampi(rni(div_fast(v.w, rq.div)), qr.lo, qr.hi);
        } else {
          q0 = quant1(v.x, rq, sym, qr.lo, qr.hi);
          q1 = quant1(v.y, rq, sym, qr.lo, qr.hi);
          q2 = quant1(v.z, rq, sym, qr.lo, qr.hi);
          q3 = quant1(v.w, rq, sym, qr.lo, qr.hi);
        }
        const long long e = tile_elem0 + static_cast<long long>(c) * kChunk + lane * 4;
        if (a.q) *reinterpret_cast<uint32_t*>(a.q + e) = pack_i8x4(q0, q1, q2, q3);
        if (a.packed) {
          if (a.bits == 4) {
            uint32_t h = (q0 & 0xF) | ((q1 & 0xF) << 4) | ((q2 & 0xF) << 8) | ((q3 & 0xF) << 12);
            const uint32_t o = __shfl_down_sync(0xffffffffu, h, 1);
            if ((lane & 1) == 0) *reinterpret_cast<uint32_t*>(a.packed + e / 2) = h | (o << 16);
          } else {  // bits == 2
            uint32_t b = (q0 & 3) | ((q1 & 3) << 2) | ((q2 & 3) << 4) | ((q3 & 3) << 6);
            const uint32_t o1 = __shfl_down_sync(0xffffffffu, b, 1);
            const uint32_t o2 = __shfl_down_sync(0xffffffffu, b, 2);
            const uint32_t o3 = __shfl_down_sync(0xffffffffu, b, 3);
            if ((lane & 3) == 0)
              *reinterpret_cast<uint32_t*>(a.packed + e / 4) = b | (o1 << 8) | (o2 << 16) | (o3 << 24);
          }
        }
        rem += NW;
        while (rem >= cpr) { rem -= cpr; ++row; }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }
}

// ------------------------------------------------------------------ generic fallback
// One warp per row, scalar loads straight from global, two passes (the second
// one normally hits L2).  Any cols, any alignment.  Packed output is written
// bytewise: a byte never straddles two lanes because each lane owns aligned
// groups of 8/bits consecutive elements; packing is over the FLAT tensor
// (pack_data ravel()s first), so a byte may straddle two rows when cols is
// odd — those tensors go through aeqb_pack_bits instead (host side decides).
__global__ void __launch_bounds__(256) requant_rows_generic(const __grid_constant__ RowsArgs a) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  const bool sym = a.symmetric != 0;
  const QRange qr = qrange(a.bits, sym);
  const float* x = a.x + row * a.cols;
  float mn, mx, xmax;
  if (a.given_min) {
    mn = a.given_min[row * a.mm_stride];
    mx = a.given_max[row * a.mm_stride];
    xmax = INFINITY;
  } else {
    mn = INFINITY;
    mx = -INFINITY;
    for (int c = lane; c < a.cols; c += 32) {
      const float v = x[c];
      mn = min_nan(mn, v);
      mx = max_nan(mx, v);
    }
    mn = warp_min_nan(mn);
    mx = warp_max_nan(mx);
    xmax = max_nan(fabsf(mn), fabsf(mx));
  }
  RowsArgs b = a;  // only lane 0 publishes scale / zp
  if (lane != 0) { b.scale = nullptr; b.zp = nullptr; }
  const RowQ rq = finalize_row(b, row, mn, mx, xmax);
  const int per = a.packed ? 8 / a.bits : 1;  // elements per packed byte
  for (int c0 = lane * per; c0 < a.cols; c0 += 32 * per) {
    unsigned byte = 0;
    for (int j = 0; j < per && c0 + j < a.cols; ++j) {
      const int q = quant1(x[c0 + j], rq, sym, qr.lo, qr.hi);
      if (a.q) a.q[row * a.cols + c0 + j] = static_cast<int8_t>(q);
      byte |= (static_cast<unsigned>(q) & ((1u << a.bits) - 1u)) << (a.bits * j);
    }
    if (a.packed) a.packed[(row * a.cols + c0) / per] = static_cast<uint8_t>(byte);
  }
}

template <int STAGE_BYTES, int NW>
cudaError_t launch_stream(RowsArgs a, int sm_count, int ctas_per_sm, cudaStream_t st) {
  auto kern = requant_rows_stream<STAGE_BYTES, NW>;
  const int smem = kStages * STAGE_BYTES;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const long long row_bytes = static_cast<long long>(a.cols) * 4;
  int rpt = static_cast<int>(STAGE_BYTES / row_bytes);
  if (rpt > kMaxRowsPerTile) rpt = kMaxRowsPerTile;
  a.rows_per_tile = rpt;
  a.n_tiles = (a.rows + rpt - 1) / rpt;
  long long grid = static_cast<long long>(sm_count) * ctas_per_sm;
  if (grid > a.n_tiles) grid = a.n_tiles;
  kern<<<static_cast<unsigned>(grid), (NW + 1) * 32, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_requant_rows(RowsArgs a, int sm_count, cudaStream_t st) {
  if (a.rows <= 0 || a.cols <= 0) return cudaSuccess;
  const long long row_bytes = static_cast<long long>(a.cols) * 4;
  const bool aligned = (reinterpret_cast<uintptr_t>(a.x) % 16 == 0) &&
                       (!a.q || reinterpret_cast<uintptr_t>(a.q) % 4 == 0) &&
                       (!a.packed || reinterpret_cast<uintptr_t>(a.packed) % 4 == 0);
  const bool packed_rows_ok = !a.packed || (a.cols % (8 / a.bits) == 0);
  if (aligned && a.cols % kChunk == 0 && row_bytes <= 32768)
    return launch_stream<32768, 8>(a, sm_count, 2, st);
  if (aligned && a.cols % kChunk == 0 && row_bytes <= 65536)
    return launch_stream<65536, 16>(a, sm_count, 1, st);
  if (!packed_rows_ok) return cudaErrorInvalidValue;  // caller packs separately
  const int warps = 8;
  const long long grid = (a.rows + warps - 1) / warps;
  requant_rows_generic<<<static_cast<unsigned>(grid), warps * 32, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace aeqb
