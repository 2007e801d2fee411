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
// whole rows of a batch of tensors (job table passed by value in the kernel
// parameters).  One producer lane keeps a 3-stage shared-memory ring full with
// 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx).  Rows are cut into
// 128-float warp chunks; each consumer warp pulls its <= 8 chunks of the tile
// into registers once (conflict-free LDS.128), reduces |x| max (or min & max)
// per row with NaN propagation, merges across warps with shared-memory atomics,
// and after one named barrier quantises from registers: hoisted exact divide
// (3 FFMA), magic-number FADD for rint, PRMT byte packing, 128 contiguous bytes
// written per warp instruction.  HBM is touched once, shared memory once.
#include <stdlib.h>

#include "aeqb_common.cuh"
#include "aeqb_kernels.h"

namespace aeqb {

namespace {

constexpr int kMaxRowsPerTile = 64;
constexpr int kRowsOnePollerDefault = 0;
constexpr int kChunk = 128;   // floats per warp chunk

struct RowAcc {  // merged across warps with shared-memory atomics
  unsigned amax_bits;
  int mn_ord, mx_ord;
  int nan;
};

enum RowMode : int {
  kSlow = 0,       // IEEE divide, integer clamp (degenerate scales, unknown range)
  kFastClamp = 1,  // hoisted divide + zero point + integer clamp
  kFastSym = 2,    // hoisted divide + magic rounding, clamp provably idle
};

struct RowQ {  // per-row quantisation constants for pass 2 (16 bytes: one LDS.128)
  float b, y, zp;
  int mode;
};

struct StageDesc {
  RowsJob job;
  long long row0;
  int nrows;
};

__device__ __forceinline__ void acc_reset(RowAcc& a) {
  a.amax_bits = 0u;
  a.mn_ord = 0x7f800000;       // f2ord(+inf)
  a.mx_ord = (int)0x807fffff;  // f2ord(-inf)
  a.nan = 0;
}

// uqt:492-586 for one row.  `xmax` bounds |x| over the row (for the divide
// window); pass +inf when the row was not scanned.
__device__ __forceinline__ void publish_scale(float* p, float v, const PeerMirror& pm) {
  *p = v;
  for (int i = 0; i < pm.n; ++i)  // peer copies of the gathered scale buffer (NVLink stores)
    *reinterpret_cast<float*>(reinterpret_cast<char*>(p) + pm.delta[i]) = v;
}

__device__ __forceinline__ RowQ finalize_row(const RowsJob& a, int bits, bool sym, long long row,
                                            float mn, float mx, float xmax, bool publish,
                                            const PeerMirror& pm) {
  const QRange qr = qrange(bits, sym);
  float scale, zpf = 0.0f;
  if (a.given_scale) {  // uniform_quantize (uqt:273-362) with the caller's parameters, no statistics
    RowQ g;
    g.b = a.given_scale[row * a.mm_stride];
    g.zp = a.given_zp ? static_cast<float>(a.given_zp[row * a.mm_stride]) : 0.0f;
    const DivBy d = make_div(g.b, xmax);
    g.y = d.y;
    g.mode = d.fast ? kFastClamp : kSlow;
    return g;
  }
  if (sym) {
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
  // uqt:585: the zero point is cast to the quantised dtype (int8 for <= 8 bits)
  // without clipping, i.e. it wraps; identity unless a clip shrank the range.
  const int zpi = static_cast<int>(static_cast<int8_t>(rni(zpf)));
  if (publish) {
    if (a.scale) publish_scale(&a.scale[row * a.out_stride], scale, pm);
    if (a.zp) a.zp[row * a.out_stride] = zpi;
  }
  RowQ r;
  const DivBy d = make_div(scale, xmax);
  r.b = scale;
  r.y = d.y;
  r.zp = static_cast<float>(zpi);
  const bool clamp_idle = sym && a.clip == nullptr && a.given_min == nullptr;
  r.mode = !d.fast ? kSlow : (clamp_idle ? kFastSym : kFastClamp);
  return r;
}

__device__ __forceinline__ int quant_slow(float x, const RowQ& rq, bool sym, int lo, int hi) {
  (void)sym;
  // x / scale + zp; adding a zero zero-point only turns -0 into +0, which rounds the same.
  const float t = __fadd_rn(__fdiv_rn(x, rq.b), rq.zp);
  return clampi(rni(t), lo, hi);
}

__device__ __forceinline__ float div_row(float x, const RowQ& rq) {
  const float q0 = x * rq.y;
  return fmaf(rq.y, fmaf(-rq.b, q0, x), q0);
}

// Pass 2 of a tile whose rows are all symmetric / unclipped / inside the divide window.
// Branch-free; partial tiles (row length not a multiple of NW chunks, tensor tails) compute their
// unused chunk slots on zeros and predicate only the store.  OUT = 1: int8 only, 2: packed INT4
// only, 3: both -- a compile-time choice: with run-time pointer tests every slot issued the other
// output's seven instructions predicated off (7 of 20 issue slots per chunk slot of the INT8 launch),
// and the partial-tile variant wrapped the packed half in a divergence-protected branch per slot
// (packed INT4 on 2560 / 3584 / 5120 / 11008-wide rows: 0.76-0.81 of peak against 0.97 on full tiles).
__device__ __forceinline__ void stg_u32(void* p, uint32_t v) {
  asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void stg_u16(void* p, uint32_t v) {
  asm volatile("st.global.u16 [%0], %1;" ::"l"(p), "h"(static_cast<unsigned short>(v)) : "memory");
}

template <bool FULL, bool CLAMP, int NW, int SLOTS, int OUT>
__device__ __forceinline__ void tight_pass2(const float4 (&v)[SLOTS], const float2* by_slots,
                                            int8_t* ql, uint8_t* pl, int lane, int warp, int nchunks,
                                            float lo, float hi) {
  (void)lane;
#pragma unroll
  for (int j = 0; j < SLOTS; ++j) {
    const bool valid = FULL || warp + j * NW < nchunks;  // partial tiles: only the store is predicated
    const float2 by = by_slots[j];
    // The hoisted exact divide on the packed-fp32 pipe: two elements per FMUL2 / FFMA2
    // (same three roundings per element as the scalar sequence, so still bit-identical).
    float2 qa = div_fast2(make_float2(v[j].x, v[j].y), by.x, by.y);
    float2 qb = div_fast2(make_float2(v[j].z, v[j].w), by.x, by.y);
    if (CLAMP) {
      // Clipped rows (OCTAV, caller-supplied ranges): clamp(rint(t)) == rint(clamp(t)) for integer
      // bounds, so the clip happens in the float domain and the magic-number rounding stays.
      qa.x = fminf(fmaxf(qa.x, lo), hi); qa.y = fminf(fmaxf(qa.y, lo), hi);
      qb.x = fminf(fmaxf(qb.x, lo), hi); qb.y = fminf(fmaxf(qb.y, lo), hi);
    }
    if (OUT & 1) {
      const uint2 ra = rmagic2(qa, kMagic), rb = rmagic2(qb, kMagic);
      if (valid) stg_u32(ql + j * NW * kChunk, bytes4(ra.x, ra.y, rb.x, rb.y));
    }
    if (OUT & 2) {
      const uint2 ra = rmagic2(qa, kMagicPlus8), rb = rmagic2(qb, kMagicPlus8);
      // Two bytes per lane, 64 contiguous bytes per warp store: no cross-lane merge on the
      // pass-2 critical path (the nibble order inside the 16 bits is already final).
      const uint32_t h = nibbles4_biased(ra.x, ra.y, rb.x, rb.y) ^ 0x8888u;
      if (valid) stg_u16(pl + j * NW * (kChunk / 2), h);
    }
  }
}

template <bool CLAMP, int NW, int SLOTS, int OUT>
__device__ __forceinline__ void tight_pass2_tile(bool full_tile, const float4 (&v)[SLOTS], const float2* by_slots,
                                                 int8_t* ql, uint8_t* pl, int lane, int warp, int nchunks,
                                                 float lo, float hi) {
  if (full_tile) tight_pass2<true, CLAMP, NW, SLOTS, OUT>(v, by_slots, ql, pl, lane, warp, nchunks, lo, hi);
  else tight_pass2<false, CLAMP, NW, SLOTS, OUT>(v, by_slots, ql, pl, lane, warp, nchunks, lo, hi);
}

// Shared-memory max reduction by the calling lane only (inline PTX: the compiler does not wrap it in
// its warp-aggregation sequence, ~10 instructions per call site when only lane 0 is active anyway).
__device__ __forceinline__ void smem_red_max(unsigned* p, unsigned v) {
  asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// Pass 1 of a full tile whose rows split evenly over the consumer warps: SPR consecutive chunk slots of
// every warp belong to one row.  Two independent maxima per row keep the FMNMX chains short.
template <int SPR, int NW, int SLOTS>
__device__ __forceinline__ void fold_rows(const float4 (&v)[SLOTS], unsigned (*part)[NW], int warp, int lane) {
#pragma unroll
  for (int r = 0; r < SLOTS / SPR; ++r) {
    float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
    for (int k = 0; k < SPR; ++k) {
      if (k & 1) a1 = absmax4(a1, v[r * SPR + k]);
      else a0 = absmax4(a0, v[r * SPR + k]);
    }
    const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(max_nan(a0, a1)));
    if (lane == 0) part[r][warp] = m;
  }
}

template <int NW, int SLOTS>
__device__ __forceinline__ bool fold_rows_any(int spr, const float4 (&v)[SLOTS], unsigned (*part)[NW], int warp,
                                              int lane) {
#define AEQB_FOLD_CASE(S)                                             \
  if constexpr (SLOTS % S == 0) {                                     \
    if (spr == S) {                                                   \
      fold_rows<S, NW, SLOTS>(v, part, warp, lane);                   \
      return true;                                                    \
    }                                                                 \
  }
  AEQB_FOLD_CASE(1) AEQB_FOLD_CASE(2) AEQB_FOLD_CASE(3) AEQB_FOLD_CASE(4) AEQB_FOLD_CASE(6)
  AEQB_FOLD_CASE(8) AEQB_FOLD_CASE(12) AEQB_FOLD_CASE(16)
#undef AEQB_FOLD_CASE
  return false;
}

// ------------------------------------------------------------------ TMA tile stream
// SLOTS = chunk slots (128 floats each) per consumer warp and tile, STAGES = depth of the ring.
// Classes 1-3 use 8 slots and 3 stages; the wide class (4) uses 12 slots and 2 stages of 96 KiB so
// that one round trip of a CTA moves up to 96 KiB (two 44 KiB rows, four 24 KiB rows).
// RICH = false is the min/max hot path exactly; RICH = true adds the modes whose extra code and
// shared memory measurably slow that path down when compiled into it (0.96 -> 0.92 of peak):
// the RMS row statistic of MSE and the clamped tight pass 2 of clipped / given-scale rows.  The
// launcher picks the instantiation per batch.
template <int STAGE_BYTES, int NW, int SLOTS, int STAGES, bool RICH>
__global__ void __launch_bounds__((NW + 1) * 32, (STAGE_BYTES <= 16384 ? 4 : (STAGE_BYTES <= 32768 ? 2 : 1)))
    requant_rows_stream(const __grid_constant__ RowsBatch b) {
  static_assert(STAGE_BYTES / (kChunk * 4) == NW * SLOTS, "chunks per warp");
  static_assert(SLOTS <= 32, "one finalising lane per chunk slot");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ StageDesc desc[STAGES];
  __shared__ RowAcc s_acc[3][kMaxRowsPerTile];
  __shared__ __align__(16) unsigned s_part[3][SLOTS][NW];  // FOLD: per-warp |x| maxima of each row of the tile
  __shared__ float2 s_by[NW][SLOTS];  // per-warp (scale, reciprocal) of each chunk slot
  __shared__ double s_sq[RICH ? 3 : 1][RICH ? NW * SLOTS : 1];  // MSE: per-chunk fp64 sums of fp32 squares

  const int tid = threadIdx.x;
  // Broadcast from lane 0: tells the compiler the warp index is warp-uniform, so the per-chunk
  // guards below (`c < nchunks`, c = warp + j * NW) compile to uniform branches instead of
  // divergence-protected regions around every REDUX / SHFL (measured on the partial-tile path:
  // 23 % of the executed instructions were ISETP / BRA / BSSY / BSYNC).
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int lane = tid & 31;
  const long long n_tiles = b.n_tiles;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], NW);
    }
    mbar_fence_init();
  }
  for (int e = tid; e < 3 * kMaxRowsPerTile; e += (NW + 1) * 32)
    acc_reset(s_acc[e / kMaxRowsPerTile][e % kMaxRowsPerTile]);
  __syncthreads();

  if (warp == NW) {  // ---------------- producer
    if (lane == 0) {
      int j = 0;
      const RowsJob* jobs = b.table != nullptr ? b.table : b.jobs;  // device table or kernel parameters
      const long long* ends = b.table_ends;
      RowsJob job = jobs[0];
      int s = 0;
      uint32_t round = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
        if (ends != nullptr) {  // walk the packed tile ends (L1-resident lines), fetch a job only when it changes
          if (tile >= job.tile_end) {
            while (tile >= __ldg(ends + j)) ++j;
            job = jobs[j];
          }
        } else {
          while (tile >= job.tile_end) job = jobs[++j];
        }
        const long long row0 = (tile - job.tile0) * job.rows_per_tile;
        const long long nrows = min(static_cast<long long>(job.rows_per_tile), job.rows - row0);
        desc[s].job = job;
        desc[s].row0 = row0;
        desc[s].nrows = static_cast<int>(nrows);
        const uint32_t bytes = static_cast<uint32_t>(nrows * job.cols * 4);
        mbar_arrive_expect_tx(&full_bar[s], bytes);  // release: desc visible to waiters
        bulk_g2s(smem_raw + static_cast<size_t>(s) * STAGE_BYTES, job.x + row0 * job.cols, bytes,
                 &full_bar[s]);
        if (++s == STAGES) { s = 0; ++round; }
      }
    }
    return;
  }

  // ---------------- consumers
  const bool sym = b.symmetric != 0;
  const int bits = b.bits;
  const QRange qr = qrange(bits, sym);
  int s = -1, buf = -1;
  uint32_t ph = 1;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (++s == STAGES) s = 0;
    if (s == 0) ph ^= 1;
    if (++buf == 3) buf = 0;
    const float4* t4 = reinterpret_cast<const float4*>(smem_raw + static_cast<size_t>(s) * STAGE_BYTES);

    if (b.one_poller) {
      // One warp polls the mbarrier; the others wait on a hardware named barrier, which issues
      // nothing while it blocks (the polling loops were a quarter of the executed instructions of
      // the 477-tensor launch).  Their own wait afterwards succeeds at once: it is the acquire that
      // makes the bulk copy's bytes visible to them.
      if (warp == 0) mbar_wait(&full_bar[s], ph);
      named_bar_sync(2, NW * 32);
      if (warp != 0) mbar_wait(&full_bar[s], ph);
    } else {
      mbar_wait(&full_bar[s], ph);
    }

    const RowsJob& job = desc[s].job;
    const long long row0 = desc[s].row0;
    const int nrows = desc[s].nrows;
    const int cols = job.cols;
    const int cpr = cols / kChunk;
    const int nchunks = nrows * cpr;
    // Caller-supplied min / max: no scan.  Caller-supplied scale: |x| max is still scanned (free
    // at HBM speed) because it decides whether the hoisted divide may be used.
    const bool gscale = job.given_scale != nullptr;
    const bool given = job.given_min != nullptr && !gscale;
    const bool mse = RICH && job.mse_k != 0.0f;  // scale from the row's RMS; |x| max still scanned (divide window)
    const bool abs_scan = gscale || sym || mse;

    // ---- pass 1: pull this warp's chunks into registers; one REDUX + one shared
    // atomic per chunk merges the row statistics (no row-change bookkeeping).
    const unsigned magic = job.cpr_magic;  // row of chunk c = (c * magic) >> 20
    const bool full_tile = nchunks == NW * SLOTS;
    const bool plain = abs_scan && !given && !mse;
    float4 v[SLOTS];
    // FOLD: when a row's chunks are a multiple of the consumer warps (cpr = spr * NW: 2048 / 4096 / 8192-float
    // rows ...), every warp holds the same slots of every row (slot j -> row j / spr), so it folds its slots
    // of a row in registers and publishes ONE partial maximum per row with a plain store; the finalising
    // lane maxes the NW partials.  No shared-memory atomics (a sixth of the executed instructions of the
    // 477-tensor launch went into the per-chunk atomicMax and its divergence region).  AEQB_ROWS_NO_FOLD=1
    // for A/B runs.  Every other plain tile (row length not a multiple of NW chunks, tensor tails) folds the
    // RUN of consecutive slots that share a row and merges once per run with a shared-memory reduction
    // (red.shared: no warp-aggregation code around it); unused chunk slots load zeros and form runs of
    // their own, which are dropped.  The row index is made provably warp-uniform (lane 0's view of the stage
    // descriptor), so the run boundaries are uniform branches.
    const int cpr_u = __shfl_sync(0xffffffffu, cpr, 0);
    const int spr = cpr_u / NW;
    bool fold = false;
    if (plain) {
#pragma unroll
      for (int j = 0; j < SLOTS; ++j) {
        const int c = warp + j * NW;
        v[j] = (full_tile || c < nchunks) ? t4[c * 32 + lane] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      }
      if (b.fold != 0 && full_tile && spr * NW == cpr_u)
        fold = fold_rows_any<NW, SLOTS>(spr, v, s_part[buf], warp, lane);
      if (!fold && cpr_u < NW) {
        // rows shorter than NW chunks: every slot of a warp is a row of its own, nothing to fold --
        // one REDUX + one shared atomic per slot, branch-free (invalid slots hold zeros and are predicated)
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) {
          const int c = warp + j * NW;
          const int r = static_cast<int>((static_cast<unsigned>(c) * magic) >> 20);
          const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(absmax4(0.0f, v[j])));
          if (lane == 0 && (full_tile || c < nchunks)) atomicMax(&s_acc[buf][r].amax_bits, m);
        }
      } else if (!fold) {
        const unsigned magic_u = __shfl_sync(0xffffffffu, magic, 0);
        const int nrows_u = __shfl_sync(0xffffffffu, nrows, 0);
        float acc = 0.0f;
        int run_r = static_cast<int>((static_cast<unsigned>(warp) * magic_u) >> 20);
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) {
          const int r = static_cast<int>((static_cast<unsigned>(warp + j * NW) * magic_u) >> 20);
          if (j > 0 && r != run_r) {  // warp-uniform: close the run
            const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(acc));
            if (lane == 0 && run_r < nrows_u) smem_red_max(&s_acc[buf][run_r].amax_bits, m);
            acc = 0.0f;
            run_r = r;
          }
          acc = absmax4(acc, v[j]);
        }
        const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(acc));
        if (lane == 0 && run_r < nrows_u) smem_red_max(&s_acc[buf][run_r].amax_bits, m);
      }
    } else if (RICH && mse) {
      // mse.get_tensor_quant_params (mse.py:100-108): fp32 squares summed in fp64.  A lane adds up
      // the chunks of one row it sees in a run of consecutive slots and the warp reduces once per
      // run (eight times fewer fp64 shuffles on 4096-wide rows); the run's total lands in the
      // slot of its first chunk, zeros in the others, and the row is later summed in chunk
      // order, so the result does not depend on timing.
      double run = 0.0;
      int run_c = warp, run_r = 0;
#pragma unroll
      for (int j = 0; j < SLOTS; ++j) {
        const int c = warp + j * NW;
        const bool valid = c < nchunks;
        v[j] = valid ? t4[c * 32 + lane] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        const int r = static_cast<int>((static_cast<unsigned>(c) * magic) >> 20);
        const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(absmax4(0.0f, v[j])));
        if (lane == 0 && valid) atomicMax(&s_acc[buf][r].amax_bits, m);
        if (j > 0 && (r != run_r || !valid)) {  // warp-uniform: close the run
          double t = run;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (lane == 0 && run_c < nchunks) s_sq[buf][run_c] = t;
          run = 0.0;
          run_c = c;
        } else if (j > 0 && lane == 0 && valid) {
          s_sq[buf][c] = 0.0;
        }
        run_r = r;
        double sq = static_cast<double>(__fmul_rn(v[j].x, v[j].x)) + static_cast<double>(__fmul_rn(v[j].y, v[j].y));
        sq += static_cast<double>(__fmul_rn(v[j].z, v[j].z)) + static_cast<double>(__fmul_rn(v[j].w, v[j].w));
        run += sq;
      }
      {
        double t = run;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0 && run_c < nchunks) s_sq[buf][run_c] = t;
      }
    } else
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
      const int c = warp + j * NW;
      if (c < nchunks) {
        v[j] = t4[c * 32 + lane];
        if (!given) {
          const int r = static_cast<int>((static_cast<unsigned>(c) * magic) >> 20);
          if (abs_scan) {
            const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(absmax4(0.0f, v[j])));
            if (lane == 0) atomicMax(&s_acc[buf][r].amax_bits, m);
          } else {
            const float lo = min_nan(min_nan(v[j].x, v[j].y), min_nan(v[j].z, v[j].w));
            const float hi = max_nan(max_nan(v[j].x, v[j].y), max_nan(v[j].z, v[j].w));
            const bool bad = __any_sync(0xffffffffu, lo != lo);
            const int mn = __reduce_min_sync(0xffffffffu, f2ord(lo));
            const int mx = __reduce_max_sync(0xffffffffu, f2ord(hi));
            if (lane == 0) {
              if (bad) {
                atomicOr(&s_acc[buf][r].nan, 1);
              } else {
                atomicMin(&s_acc[buf][r].mn_ord, mn);
                atomicMax(&s_acc[buf][r].mx_ord, mx);
              }
            }
          }
        }
      }
    }
    // Everything pass 2 needs from the stage descriptor is copied to registers,
    // then the stage goes back to the producer early: the tile lives in registers.
    int8_t* const qp = job.q ? job.q + row0 * cols : nullptr;
    uint8_t* const pp = job.packed ? job.packed + ((row0 * cols * bits) >> 3) : nullptr;
    // Lane l < 8 will finalise the row of this warp's l-th chunk: it needs the job.
    const int my_c = warp + (lane < SLOTS ? lane : 0) * NW;
    const bool my_valid = lane < SLOTS && my_c < nchunks;
    RowsJob jcopy;
    if (my_valid) jcopy = job;
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);  // asm volatile + memory clobber: no smem read sinks below
    named_bar_sync(1, NW * 32);  // every warp's partials are merged in s_acc[buf]

    // ---- per-row scale / zero point, one lane per chunk slot (no second barrier):
    // the row's first chunk owner publishes scale / zp to global memory.
    RowQ mine;
    mine.b = 1.0f; mine.y = 1.0f; mine.zp = 0.0f; mine.mode = kFastSym;  // unused slots must not veto the tight path
    auto row_amax = [&](int r) -> float {  // |x| maximum of row r of this tile
      if (!fold) return __uint_as_float(s_acc[buf][r].amax_bits);
      unsigned m = 0u;
#pragma unroll
      for (int w = 0; w < NW; ++w) m = max(m, s_part[buf][r][w]);
      return __uint_as_float(m);
    };
    if (my_valid) {
      const int r = static_cast<int>((static_cast<unsigned>(my_c) * magic) >> 20);
      const long long grow = row0 + r;
      float mn, mx, xmax;
      if (RICH && mse) {
        double t = 0.0;
        const int cfirst = r * cpr;
        for (int k2 = 0; k2 < cpr; ++k2) t += s_sq[buf][cfirst + k2];
        const float mean = __fdiv_rn(static_cast<float>(t), static_cast<float>(cols));
        const float sc = __fmul_rn(jcopy.mse_k, __fsqrt_rn(mean));
        xmax = row_amax(r);
        const DivBy dv = make_div(sc, xmax);
        mine.b = sc; mine.y = dv.y; mine.zp = 0.0f;
        mine.mode = dv.fast ? kFastClamp : kSlow;
        if (my_c == r * cpr) {
          if (jcopy.scale) publish_scale(&jcopy.scale[grow * jcopy.out_stride], sc, b.peers);
          if (jcopy.zp) jcopy.zp[grow * jcopy.out_stride] = 0;
        }
      } else {
      if (gscale) {
        mx = row_amax(r);  // finalize_row only uses xmax here
        mn = -mx;
        xmax = mx;
      } else if (given) {
        mn = jcopy.given_min[grow * jcopy.mm_stride];
        mx = jcopy.given_max[grow * jcopy.mm_stride];
        xmax = INFINITY;  // row not scanned: always take the IEEE divide
      } else if (sym) {
        mx = row_amax(r);
        mn = -mx;
        xmax = mx;
      } else {
        const RowAcc acc = s_acc[buf][r];
        mn = acc.nan ? NAN : ord2f(acc.mn_ord);
        mx = acc.nan ? NAN : ord2f(acc.mx_ord);
        xmax = max_nan(fabsf(mn), fabsf(mx));
      }
      mine = finalize_row(jcopy, bits, sym, grow, mn, mx, xmax, my_c == r * cpr, b.peers);
      }
    }
    // Buffer (it+2)%3 was last read during the previous tile, which every warp has
    // left (they all passed the barrier above); it is next written two tiles from now.
    if (tid < kMaxRowsPerTile) acc_reset(s_acc[buf == 0 ? 2 : buf - 1][tid]);

    // ---- pass 2: quantise from registers + store
    const bool all_fast = __all_sync(0xffffffffu, lane >= SLOTS || mine.mode == kFastSym);
    // symmetric rows with a live clamp (zero point 0, divisor inside the hoisted-divide window)
    const bool all_clampable =
        RICH && sym && __all_sync(0xffffffffu, lane >= SLOTS || (mine.mode != kSlow && mine.zp == 0.0f));
    if ((all_fast || all_clampable) && !(pp && bits != 4)) {
      // Tight path: every row of the tile is symmetric / unclipped / in the divide
      // window.  (scale, reciprocal) per chunk slot via one broadcast LDS.64.
      if (lane < SLOTS) s_by[warp][lane] = make_float2(mine.b, mine.y);
      __syncwarp();
      int8_t* const ql = qp ? qp + warp * kChunk + lane * 4 : nullptr;
      uint8_t* const pl = pp ? pp + ((warp * kChunk + lane * 4) >> 1) : nullptr;
      const float lo_f = static_cast<float>(qr.lo), hi_f = static_cast<float>(qr.hi);
      // which outputs the job has, as a warp-uniform value (lane 0's view of the stage descriptor)
      const int outs = __shfl_sync(0xffffffffu, (qp != nullptr ? 1 : 0) | (pp != nullptr ? 2 : 0), 0);
      if (!RICH || all_fast) {
        if (outs == 1) tight_pass2_tile<false, NW, SLOTS, 1>(full_tile, v, s_by[warp], ql, pl, lane, warp, nchunks, lo_f, hi_f);
        else if (outs == 2) tight_pass2_tile<false, NW, SLOTS, 2>(full_tile, v, s_by[warp], ql, pl, lane, warp, nchunks, lo_f, hi_f);
        else if (outs == 3) tight_pass2_tile<false, NW, SLOTS, 3>(full_tile, v, s_by[warp], ql, pl, lane, warp, nchunks, lo_f, hi_f);
      } else if constexpr (RICH) {
        if (outs == 1) tight_pass2_tile<true, NW, SLOTS, 1>(full_tile, v, s_by[warp], ql, pl, lane, warp, nchunks, lo_f, hi_f);
        else if (outs == 2) tight_pass2_tile<true, NW, SLOTS, 2>(full_tile, v, s_by[warp], ql, pl, lane, warp, nchunks, lo_f, hi_f);
        else if (outs == 3) tight_pass2_tile<true, NW, SLOTS, 3>(full_tile, v, s_by[warp], ql, pl, lane, warp, nchunks, lo_f, hi_f);
      }
      __syncwarp();  // s_by[warp] is rewritten next tile
    } else {
#pragma unroll
      for (int j = 0; j < SLOTS; ++j) {
        const int c = warp + j * NW;
        RowQ rq;
        rq.b = __shfl_sync(0xffffffffu, mine.b, j);
        rq.y = __shfl_sync(0xffffffffu, mine.y, j);
        rq.zp = __shfl_sync(0xffffffffu, mine.zp, j);
        rq.mode = __shfl_sync(0xffffffffu, mine.mode, j);
        if (c < nchunks) {
          const int e = c * kChunk + lane * 4;  // inside the tile
          if (rq.mode == kFastSym && !(pp && bits == 2)) {
            const float t0 = div_row(v[j].x, rq), t1 = div_row(v[j].y, rq),
                        t2 = div_row(v[j].z, rq), t3 = div_row(v[j].w, rq);
            if (qp) *reinterpret_cast<uint32_t*>(qp + e) = bytes4(rmagic(t0), rmagic(t1), rmagic(t2), rmagic(t3));
            if (pp) {  // bits == 4
              const uint32_t h =
                  (nibbles4_biased(rmagic8(t0), rmagic8(t1), rmagic8(t2), rmagic8(t3)) ^ 0x8888u) & 0xFFFFu;
              const uint32_t o = __shfl_down_sync(0xffffffffu, h, 1);
              if ((lane & 1) == 0) *reinterpret_cast<uint32_t*>(pp + (e >> 1)) = h | (o << 16);
            }
          } else {
            int q0, q1, q2, q3;
            if (rq.mode != kSlow) {
              float t0 = div_row(v[j].x, rq), t1 = div_row(v[j].y, rq), t2 = div_row(v[j].z, rq),
                    t3 = div_row(v[j].w, rq);
              // + zero point (0 when symmetric: only turns -0 into +0, which rounds the same)
              t0 = __fadd_rn(t0, rq.zp); t1 = __fadd_rn(t1, rq.zp);
              t2 = __fadd_rn(t2, rq.zp); t3 = __fadd_rn(t3, rq.zp);
              q0 = clampi(rni(t0), qr.lo, qr.hi); q1 = clampi(rni(t1), qr.lo, qr.hi);
              q2 = clampi(rni(t2), qr.lo, qr.hi); q3 = clampi(rni(t3), qr.lo, qr.hi);
            } else {
              q0 = quant_slow(v[j].x, rq, sym, qr.lo, qr.hi);
              q1 = quant_slow(v[j].y, rq, sym, qr.lo, qr.hi);
              q2 = quant_slow(v[j].z, rq, sym, qr.lo, qr.hi);
              q3 = quant_slow(v[j].w, rq, sym, qr.lo, qr.hi);
            }
            if (qp) *reinterpret_cast<uint32_t*>(qp + e) = pack_i8x4(q0, q1, q2, q3);
            if (pp) {
              if (bits == 4) {
                const uint32_t h = (q0 & 0xF) | ((q1 & 0xF) << 4) | ((q2 & 0xF) << 8) | ((q3 & 0xF) << 12);
                const uint32_t o = __shfl_down_sync(0xffffffffu, h, 1);
                if ((lane & 1) == 0) *reinterpret_cast<uint32_t*>(pp + (e >> 1)) = h | (o << 16);
              } else {  // bits == 2
                const uint32_t by = (q0 & 3) | ((q1 & 3) << 2) | ((q2 & 3) << 4) | ((q3 & 3) << 6);
                const uint32_t o1 = __shfl_down_sync(0xffffffffu, by, 1);
                const uint32_t o2 = __shfl_down_sync(0xffffffffu, by, 2);
                const uint32_t o3 = __shfl_down_sync(0xffffffffu, by, 3);
                if ((lane & 3) == 0)
                  *reinterpret_cast<uint32_t*>(pp + (e >> 2)) = by | (o1 << 8) | (o2 << 16) | (o3 << 24);
              }
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ generic fallback
// One warp per row, scalar loads straight from global, two passes (the second
// one normally hits L2).  Any cols, any alignment.  Packed output is written
// bytewise; packing is over the FLAT tensor (pack_data ravel()s first), so the
// launcher only allows it when a byte cannot straddle two rows.
__global__ void __launch_bounds__(256)
    requant_rows_generic(const RowsJob a, int bits, int symmetric) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  const bool sym = symmetric != 0;
  const QRange qr = qrange(bits, sym);
  const float* x = a.x + row * a.cols;
  float mn, mx, xmax;
  if (a.given_scale) {
    mn = mx = 0.0f;
    xmax = INFINITY;
  } else if (a.given_min) {
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
  PeerMirror none;
  none.n = 0;
  RowQ rq = finalize_row(a, bits, sym, row, mn, mx, xmax, lane == 0, none);
  rq.mode = kSlow;
  const int per = a.packed ? 8 / bits : 1;  // elements per packed byte
  for (int c0 = lane * per; c0 < a.cols; c0 += 32 * per) {
    unsigned byte = 0;
    for (int j = 0; j < per && c0 + j < a.cols; ++j) {
      const int q = quant_slow(x[c0 + j], rq, sym, qr.lo, qr.hi);
      if (a.q) a.q[row * a.cols + c0 + j] = static_cast<int8_t>(q);
      byte |= (static_cast<unsigned>(q) & ((1u << bits) - 1u)) << (bits * j);
    }
    if (a.packed) a.packed[(row * a.cols + c0) / per] = static_cast<uint8_t>(byte);
  }
}

template <int STAGE_BYTES, int NW, int SLOTS, int STAGES, bool RICH>
cudaError_t launch_stream_as(const RowsBatch& b, int sm_count, int ctas_per_sm, cudaStream_t st) {
  auto kern = requant_rows_stream<STAGE_BYTES, NW, SLOTS, STAGES, RICH>;
  const int smem = STAGES * STAGE_BYTES;
  static PerDevice configured;
  if (!configured.done()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured.set();
  }
  long long grid = static_cast<long long>(sm_count) * ctas_per_sm;
  if (grid > b.n_tiles) grid = b.n_tiles;
  static const int one_poller = getenv("AEQB_ROWS_ONE_POLLER") ? atoi(getenv("AEQB_ROWS_ONE_POLLER")) : kRowsOnePollerDefault;
  static const int fold = (getenv("AEQB_ROWS_NO_FOLD") && atoi(getenv("AEQB_ROWS_NO_FOLD"))) ? 0 : 1;
  if (b.one_poller != one_poller || b.fold != fold) {
    RowsBatch bb = b;
    bb.one_poller = one_poller;
    bb.fold = fold;
    kern<<<static_cast<unsigned>(grid), (NW + 1) * 32, smem, st>>>(bb);
  } else {
    kern<<<static_cast<unsigned>(grid), (NW + 1) * 32, smem, st>>>(b);
  }
  return count_launch();
}

// Plain min/max batches take the lean instantiation; anything with an RMS statistic, clipping
// constants or caller-supplied scales takes the rich one.
template <int STAGE_BYTES, int NW, int SLOTS, int STAGES>
cudaError_t launch_stream(const RowsBatch& b, int sm_count, int ctas_per_sm, cudaStream_t st) {
  bool rich = b.table != nullptr && b.rich != 0;
  for (int i = 0; b.table == nullptr && i < b.n_jobs; ++i)
    rich |= b.jobs[i].mse_k != 0.0f || b.jobs[i].clip != nullptr || b.jobs[i].given_scale != nullptr;
  return rich ? launch_stream_as<STAGE_BYTES, NW, SLOTS, STAGES, true>(b, sm_count, ctas_per_sm, st)
              : launch_stream_as<STAGE_BYTES, NW, SLOTS, STAGES, false>(b, sm_count, ctas_per_sm, st);
}

}  // namespace

// Stage size and consumer warps per class.  Smaller stages mean more CTAs per SM, i.e. more
// independent pass-1 / barrier / pass-2 pipelines to hide each other's bubbles.
//   1: rows <= 16 KiB, 4 consumer warps, 3 x 16 KiB stages, 4 CTAs / SM
//   2: rows <= 32 KiB, 8 consumer warps, 3 x 32 KiB stages, 2 CTAs / SM
//   3: rows <= 64 KiB, 16 consumer warps, 3 x 64 KiB stages, 1 CTA / SM
//   4: rows <= 96 KiB, 16 consumer warps x 12 chunk slots, 2 x 96 KiB stages, 1 CTA / SM
static int min_stream_class() {
  static const int v = getenv("AEQB_ROWS_MIN_CLASS") ? atoi(getenv("AEQB_ROWS_MIN_CLASS")) : 1;
  return v;
}

// Whole rows per tile of a class (stage bytes / row bytes, capped).
// AEQB_ROWS_SMALL_TILES=1: the batch-size rule of rows_job_class is off (A/B runs).
static bool small_tiles_only() {
  static const bool v = getenv("AEQB_ROWS_SMALL_TILES") && atoi(getenv("AEQB_ROWS_SMALL_TILES"));
  return v;
}

static int class_rows_per_tile(int cols, int klass) {
  const long long stage = klass == 1 ? 16384 : (klass == 2 ? 32768 : (klass == 3 ? 65536 : 98304));
  long long rpt = stage / (static_cast<long long>(cols) * 4);
  if (rpt > kMaxRowsPerTile) rpt = kMaxRowsPerTile;
  return static_cast<int>(rpt);
}

// Every tile costs its CTA one pass-1 -> barrier -> pass-2 round trip whatever it holds, so the
// class is the one that moves the most bytes per round trip and SM: (CTAs per SM) x (rows per
// tile) x (row bytes), counted up to 64 KiB (beyond that HBM is the limit); ties go to the
// smaller class (more independent pipelines).  Measured: a 20 KiB row reaches 0.62 of the HBM
// peak in class 2 (2 CTAs x 20 KiB) and 0.96 in class 4 (4 rows, 80 KiB); a 44 KiB row 0.68 in
// class 3 (one row per tile) and 1.02 in class 4 (two rows).
int rows_job_class(const RowsJob& j, int bits, long long batch_bytes) {
  const long long row_bytes = static_cast<long long>(j.cols) * 4;
  const bool aligned = (reinterpret_cast<uintptr_t>(j.x) % 16 == 0) &&
                       (!j.q || reinterpret_cast<uintptr_t>(j.q) % 4 == 0) &&
                       (!j.packed || reinterpret_cast<uintptr_t>(j.packed) % 4 == 0);
  (void)bits;
  if (!aligned || j.cols % kChunk != 0 || row_bytes > 98304) return 0;
  int best = 0;
  long long best_bytes = -1;
  for (int k = min_stream_class() < 1 ? 1 : min_stream_class(); k <= 4; ++k) {
    const int ctas = k == 1 ? 4 : (k == 2 ? 2 : 1);
    long long bytes = static_cast<long long>(ctas) * class_rows_per_tile(j.cols, k) * row_bytes;
    // 64 KiB per round trip saturates HBM with int8 output; the packed INT4 / INT2 pass 2 is
    // longer, so its round trip wants 96 KiB (4096-wide rows: 0.84 of peak in class 1, 0.95 in 4).
    // A batch of 96 MiB or more takes the 96 KiB tiles whatever it writes: measured back to back on one
    // box (tools/sustain_ab.py, 477 x [4096,4096] INT8), class 4 ran at 1.00-1.02 of the measured copy
    // peak at the full clock and 0.94-0.95 power-capped, class 2 at 0.92 / 0.96-0.98 and class 1 at 0.90
    // flat.  Where it starts to pay (tools/class_threshold.py, n x [4096,4096], L2 flushed between
    // launches, class 2 -> class 4): 64 MiB 23.6 -> 23.6 us, 128 MiB 39.9 -> 35.9, 256 MiB 68.6 -> 64.5,
    // 512 MiB 123.9 -> 117.8, 1 GiB 236 -> 222 (a single 64 MiB tensor is 4.6 class-4 tiles per CTA: its
    // last partial wave costs what the larger tiles gain).
    const bool big_batch = batch_bytes >= (96ll << 20) && !small_tiles_only();
    const long long cap = ((j.packed && !j.q) || big_batch) ? 98304 : 65536;
    if (bytes > cap) bytes = cap;
    if (bytes > best_bytes) {
      best_bytes = bytes;
      best = k;
    }
  }
  // A row that fills a class-1 stage exactly (4096 floats) ties between class 1 (four CTAs of one row)
  // and class 2 (two CTAs of two rows).  Measured at the end of round 2 (tools/shape_bench.py,
  // tools/sustain_probe.py, same box, back to back): class 2 is 0.5 % faster at the full SM clock,
  // 2.7 % over a 2 GB stack and 3.5 % once the power cap has pulled the clock down (half as many tile
  // round trips per byte leave more issue slots for the arithmetic).  AEQB_ROWS_TIE_SMALL=1 restores
  // the old rule.
  static const bool tie_small = getenv("AEQB_ROWS_TIE_SMALL") && atoi(getenv("AEQB_ROWS_TIE_SMALL"));
  if (best == 1 && !tie_small && row_bytes == 16384 && min_stream_class() <= 2) best = 2;
  return best;
}

int rows_job_rows_per_tile(const RowsJob& j, int klass) { return class_rows_per_tile(j.cols, klass); }

cudaError_t launch_requant_rows_stream(const RowsBatch& b, int klass, int sm_count,
                                       cudaStream_t st) {
  if (b.n_tiles <= 0) return cudaSuccess;
  if (klass == 1) return launch_stream<16384, 4, 8, 3>(b, sm_count, 4, st);
  if (klass == 2) {
    // AEQB_ROWS_CLASS2_W4=1 (experiment): four consumer warps with 16 chunk slots each instead of eight
    // with 8 -- half as many (warp, tile) round trips per byte.
    static const bool w4 = getenv("AEQB_ROWS_CLASS2_W4") && atoi(getenv("AEQB_ROWS_CLASS2_W4"));
    if (w4) return launch_stream<32768, 4, 16, 3>(b, sm_count, 2, st);
    return launch_stream<32768, 8, 8, 3>(b, sm_count, 2, st);
  }
  if (klass == 3) return launch_stream<65536, 16, 8, 3>(b, sm_count, 1, st);
  return launch_stream<98304, 16, 12, 2>(b, sm_count, 1, st);
}

cudaError_t launch_requant_rows_generic(const RowsJob& j, int bits, int symmetric,
                                        cudaStream_t st) {
  if (j.rows <= 0 || j.cols <= 0) return cudaSuccess;
  if (j.packed && (j.cols % (8 / bits) != 0)) return cudaErrorInvalidValue;  // caller packs separately
  const int warps = 8;
  const long long grid = (j.rows + warps - 1) / warps;
  requant_rows_generic<<<static_cast<unsigned>(grid), warps * 32, 0, st>>>(j, bits, symmetric);
  return count_launch();
}

}  // namespace aeqb
