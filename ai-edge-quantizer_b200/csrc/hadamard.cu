// Block-diagonal normalised Hadamard rotation of the last axis — replaces
// hadamard_rotation._rotate_with_diagonal_hadamard
// (algorithms/uniform_quantize/hadamard_rotation.py:93-134):
//     W.reshape(-1, n) @ (H_n / sqrt(n)),   H_n = Sylvester (kron of [[1,1],[1,-1]]), n = 2^m | cols
// The reference multiplies by the dense fp32 matrix (2*n flop per element); here
// each n-long segment gets a fast Walsh-Hadamard transform in natural (Sylvester)
// order — log2(n) add/sub per element — followed by one multiply with the same
// fp32 constant fl(1 / fl(sqrt(n))) the reference's matrix entries hold
// (hadamard_rotation.py:79-89).  The kernel is HBM-bound: 4 B read + 4 B written.
//
// Three kernels, picked by shape in launch_hadamard_rows: `hadamard_tiles` (n = 256..4096 on
// tensors that are whole 16 KiB tiles: one warp per tile, bulk copies in and out, packed-fp32
// butterflies in registers — the hot shapes), `hadamard_rows_reg` (rows of 256-float groups, n up
// to 8192) and the shared-memory butterfly below for everything else.
//
// Shared-memory butterfly: a CTA owns one row at a time (cols <= 16384 floats =
// 64 KiB), loads it with 128-bit loads, runs radix-4 passes (two butterfly stages
// per pass, each thread owns a 4-point group in registers, so shared memory is
// touched once per two stages), a final radix-2 pass when m is odd, then scales
// and streams the row out.  Rounding differs from the reference's sgemm only in
// summation order (and in "sum then scale" vs "scale then sum" when n is not a
// power of 4); DESIGN.md states the tolerance.
#include "aeqb_common.cuh"
#include "aeqb_kernels.h"

#include <cstdlib>

namespace aeqb {

namespace {

__global__ void __launch_bounds__(512)
    hadamard_rows_smem(const float* __restrict__ x, long long rows, int cols, int n, float norm,
                       float* __restrict__ out) {
  extern __shared__ __align__(16) float s_row[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int nvec = cols >> 2;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const float4* src = reinterpret_cast<const float4*>(x + row * cols);
    float4* s4 = reinterpret_cast<float4*>(s_row);
    for (int i = tid; i < nvec; i += nt) s4[i] = __ldg(src + i);
    __syncthreads();
    int h = 1;
    // radix-4 passes: stages h and 2h
    for (; h * 4 <= n; h *= 4) {
      if (h == 1) {
        for (int q = tid; q < nvec; q += nt) {
          float4 v = s4[q];
          const float a0 = v.x + v.y, a1 = v.x - v.y, a2 = v.z + v.w, a3 = v.z - v.w;
          v.x = a0 + a2; v.y = a1 + a3; v.z = a0 - a2; v.w = a1 - a3;
          s4[q] = v;
        }
      } else {
        for (int q = tid; q < nvec; q += nt) {
          const int base = (q / h) * (4 * h) + (q % h);
          const float v0 = s_row[base], v1 = s_row[base + h], v2 = s_row[base + 2 * h],
                      v3 = s_row[base + 3 * h];
          const float a0 = v0 + v1, a1 = v0 - v1, a2 = v2 + v3, a3 = v2 - v3;
          s_row[base] = a0 + a2;
          s_row[base + h] = a1 + a3;
          s_row[base + 2 * h] = a0 - a2;
          s_row[base + 3 * h] = a1 - a3;
        }
      }
      __syncthreads();
    }
    if (h * 2 <= n) {  // odd log2(n): one radix-2 stage
      for (int p = tid; p < (cols >> 1); p += nt) {
        const int i = (p / h) * (2 * h) + (p % h);
        const float a = s_row[i], b = s_row[i + h];
        s_row[i] = a + b;
        s_row[i + h] = a - b;
      }
      __syncthreads();
    }
    float4* dst = reinterpret_cast<float4*>(out + row * cols);
    for (int i = tid; i < nvec; i += nt) {
      float4 v = s4[i];
      v.x *= norm; v.y *= norm; v.z *= norm; v.w *= norm;
      dst[i] = v;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ register FWHT
// The fast path for rows that are a multiple of 256 floats (every FC / embedding width in
// practice): shared memory is crossed ONCE instead of once per two stages.
//   phase 1  each warp transforms 256-float groups entirely in registers: lane l holds elements
//            l + 32 j (j < 8, coalesced 128-byte loads); strides 1..16 are xor-shuffles, strides
//            32..128 are register butterflies.  Stages at strides >= n are skipped.
//   phase 2  (n > 256) the remaining strides 256 .. n/2 couple the SAME position of different
//            groups: a thread takes 4 consecutive positions x G = n / 256 groups from shared
//            memory (conflict-free LDS.128), finishes the butterflies in registers, scales and
//            stores 16 bytes per lane straight to global memory.
__device__ __forceinline__ void warp_fwht256(float (&v)[8], int lane, int n) {
#pragma unroll
  for (int h = 1; h < 32; h <<= 1) {
    if (h < n) {
      const bool upper = (lane & h) != 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float o = __shfl_xor_sync(0xffffffffu, v[j], h);
        v[j] = upper ? o - v[j] : v[j] + o;
      }
    }
  }
#pragma unroll
  for (int hj = 1; hj < 8; hj <<= 1) {
    if (hj * 32 < n) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if ((j & hj) == 0) {
          const float a = v[j], b = v[j | hj];
          v[j] = a + b;
          v[j | hj] = a - b;
        }
      }
    }
  }
}

template <int G>  // groups of 256 per Hadamard segment (n = 256 * G); G == 1: phase 1 only
__global__ void __launch_bounds__(256)
    hadamard_rows_reg(const float* __restrict__ x, long long rows, int cols, int n, float norm,
                      float* __restrict__ out) {
  extern __shared__ __align__(16) float s_row[];
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int nthreads = blockDim.x, nwarps = blockDim.x >> 5;
  const int ngroups = cols >> 8;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const float* src = x + row * cols;
    float* dst = out + row * cols;
    for (int g = warp; g < ngroups; g += nwarps) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __ldg(src + g * 256 + j * 32 + lane);
      warp_fwht256(v, lane, n);
      if (G == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[g * 256 + j * 32 + lane] = v[j] * norm;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) s_row[g * 256 + j * 32 + lane] = v[j];
      }
    }
    if (G > 1) {
      __syncthreads();
      // items: (segment, 4-position slot); 64 slots per segment
      const int nitems = (cols / (256 * G)) * 64;
      for (int it = tid; it < nitems; it += nthreads) {
        const int seg = it >> 6, p4 = (it & 63) * 4;
        const int base = seg * 256 * G + p4;
        float4 v[G];
#pragma unroll
        for (int g = 0; g < G; ++g) v[g] = *reinterpret_cast<const float4*>(&s_row[base + g * 256]);
#pragma unroll
        for (int hg = 1; hg < G; hg <<= 1) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if ((g & hg) == 0) {
              const float4 a = v[g], b = v[g | hg];
              v[g] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
              v[g | hg] = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
            }
          }
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
          float4 o = v[g];
          o.x *= norm; o.y *= norm; o.z *= norm; o.w *= norm;
          *reinterpret_cast<float4*>(&dst[base + g * 256]) = o;
        }
      }
      __syncthreads();  // s_row is rewritten by the next row
    }
  }
}

template <int G>
cudaError_t launch_reg(const float* x, long long rows, int cols, int n, float norm, float* out,
                       int sm_count, cudaStream_t st) {
  auto kern = hadamard_rows_reg<G>;
  const int smem = G > 1 ? cols * 4 : 0;
  static PerDevice configured;
  if (!configured.done() && G > 1) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    if (e != cudaSuccess) return e;
    configured.set();
  }
  // Wide segments keep 4 * G floats per thread in phase 2: smaller CTAs keep more rows in flight.
  const int threads = G >= 8 ? 128 : 256;
  long long per_sm = G > 1 ? 200000 / (smem + 1024) : 8;
  if (per_sm > 12) per_sm = 12;
  if (per_sm < 1) per_sm = 1;
  long long grid = static_cast<long long>(sm_count) * per_sm;
  if (grid > rows) grid = rows;
  kern<<<static_cast<unsigned>(grid), threads, smem, st>>>(x, rows, cols, n, norm, out);
  return count_launch();
}

// ------------------------------------------------------------------ 4096-float tiles, one warp each
// n = 128..4096 with the tensor a whole number of 16 KiB tiles: a warp owns a tile (1..32 whole
// segments) and keeps it in 128 registers per lane.  The tile arrives by ONE 16 KiB bulk copy
// (cp.async.bulk + mbarrier; the next tile is queued as soon as this one sits in registers) and
// leaves by 32 bulk stores of 512 B, so the warp issues no global loads or stores itself and HBM
// latency never reaches the scoreboard.  Element index e = 12 bits:
//   map 1 (after the load):   float4 slot j*32 + lane  -> registers hold e[1:0] and e[11:7]
//   map 2 (after the exchange): float4 slot lane*32 + k -> registers hold e[6:0]
// so stages 1, 2 and 128..n/2 run in map 1, stages 4..64 in map 2, every butterfly on packed
// fp32 pairs (FADD2 / FFMA2, half an issue slot per add) except stage 1.  The one exchange goes
// through a shared-memory tile whose 512-byte rows are padded to 528 bytes: both access patterns
// are conflict-free 128-bit accesses, and a row is exactly one lane's output run, which that
// lane hands to the bulk-store engine.  8.3 issue slots per element instead of 33.
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {  // a - b: b * -1 is exact, one rounding
  return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a);
}

constexpr int kTileFloats = 4096;
constexpr int kTileWarps = 2;                       // warps per CTA, each with its own buffers
constexpr int kTileLandBytes = kTileFloats * 4;     // bulk-copy landing buffer
constexpr int kTileRowF4 = 33;                      // 32 float4 + 1 pad per exchange row
constexpr int kTileXchgBytes = 32 * kTileRowF4 * 16;
constexpr int kTileSmemPerWarp = kTileLandBytes + kTileXchgBytes;

template <int LOG2N>
__global__ void __launch_bounds__(kTileWarps * 32, 3)
    hadamard_tiles(const float* __restrict__ x, long long ntiles, float norm, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ __align__(8) uint64_t s_bar[kTileWarps];
  const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  unsigned char* base = s_raw + static_cast<size_t>(warp) * kTileSmemPerWarp;
  const float4* land = reinterpret_cast<const float4*>(base);
  float4* xchg = reinterpret_cast<float4*>(base + kTileLandBytes);
  uint64_t* bar = &s_bar[warp];
  const long long stride = static_cast<long long>(gridDim.x) * kTileWarps;
  long long tile = static_cast<long long>(blockIdx.x) * kTileWarps + warp;
  if (lane == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    if (tile < ntiles) {
      mbar_arrive_expect_tx(bar, kTileLandBytes);
      bulk_g2s(base, x + tile * kTileFloats, kTileLandBytes, bar);
    }
  }
  __syncwarp();
  unsigned parity = 0;
  for (; tile < ntiles; tile += stride) {
    mbar_wait(bar, parity);
    parity ^= 1u;
    float2 lo[32], hi[32];  // float4 slot = (lo, hi)
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float4 v = land[j * 32 + lane];
      lo[j] = make_float2(v.x, v.y);
      hi[j] = make_float2(v.z, v.w);
    }
    __syncwarp();  // the landing buffer is drained: queue the next tile into it
    if (lane == 0 && tile + stride < ntiles) {
      mbar_arrive_expect_tx(bar, kTileLandBytes);
      bulk_g2s(base, x + (tile + stride) * kTileFloats, kTileLandBytes, bar);
    }
    // ---- map 1: stages 1 and 2 inside a float4, stages 128..n/2 across slots
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float2 a = make_float2(lo[j].x + lo[j].y, lo[j].x - lo[j].y);
      const float2 b = make_float2(hi[j].x + hi[j].y, hi[j].x - hi[j].y);
      lo[j] = add2(a, b);
      hi[j] = sub2(a, b);
    }
#pragma unroll
    for (int hb = 1; hb < (1 << (LOG2N - 7)); hb <<= 1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if ((j & hb) == 0) {
          const float2 a0 = lo[j], a1 = hi[j], b0 = lo[j | hb], b1 = hi[j | hb];
          lo[j] = add2(a0, b0); hi[j] = add2(a1, b1);
          lo[j | hb] = sub2(a0, b0); hi[j | hb] = sub2(a1, b1);
        }
      }
    }
    // the previous tile's bulk stores must have read the exchange rows before they are rewritten
    bulk_wait_read0();
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; ++j)
      xchg[j * kTileRowF4 + lane] = make_float4(lo[j].x, lo[j].y, hi[j].x, hi[j].y);
    __syncwarp();
    // ---- map 2: lane owns 128 consecutive floats; stages 4..64 across its 32 slots
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float4 v = xchg[lane * kTileRowF4 + k];
      lo[k] = make_float2(v.x, v.y);
      hi[k] = make_float2(v.z, v.w);
    }
#pragma unroll
    for (int hb = 1; hb < 32; hb <<= 1) {
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        if ((k & hb) == 0) {
          const float2 a0 = lo[k], a1 = hi[k], b0 = lo[k | hb], b1 = hi[k | hb];
          lo[k] = add2(a0, b0); hi[k] = add2(a1, b1);
          lo[k | hb] = sub2(a0, b0); hi[k | hb] = sub2(a1, b1);
        }
      }
    }
    const float2 nn = make_float2(norm, norm);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float2 p = __fmul2_rn(lo[k], nn), q = __fmul2_rn(hi[k], nn);
      xchg[lane * kTileRowF4 + k] = make_float4(p.x, p.y, q.x, q.y);  // the lane's own row
    }
    fence_async_smem();  // generic-proxy writes -> visible to the bulk-copy engine
    bulk_s2g(out + tile * kTileFloats + lane * 128, xchg + lane * kTileRowF4, 512);
    bulk_commit();
  }
  bulk_wait_read0();  // shared memory must outlive the last stores' reads
}

template <int LOG2N>
cudaError_t launch_tiles_k(const float* x, long long ntiles, float norm, float* out, int sm_count,
                           cudaStream_t st) {
  constexpr int smem = kTileWarps * kTileSmemPerWarp;
  static PerDevice configured;
  if (!configured.done()) {
    cudaError_t e = cudaFuncSetAttribute(hadamard_tiles<LOG2N>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured.set();
  }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hadamard_tiles<LOG2N>,
                                                                kTileWarps * 32, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  long long grid = static_cast<long long>(sm_count) * per_sm;
  const long long need = (ntiles + kTileWarps - 1) / kTileWarps;
  if (grid > need) grid = need;
  hadamard_tiles<LOG2N><<<static_cast<unsigned>(grid), kTileWarps * 32, smem, st>>>(x, ntiles, norm,
                                                                                    out);
  return count_launch();
}

// Fallback for rows longer than 64 KiB or unaligned / odd shapes: one butterfly
// stage per launch, in place in global memory (`buf` already holds a copy of x).
__global__ void __launch_bounds__(256)
    hadamard_stage_global(float* __restrict__ buf, long long pairs, long long h) {
  for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < pairs;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = (p / h) * (2 * h) + (p % h);
    const float a = buf[i], b = buf[i + h];
    buf[i] = a + b;
    buf[i + h] = a - b;
  }
}

__global__ void __launch_bounds__(256)
    scale_copy_global(const float* in, float* out, long long n, float s) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = in[i] * s;
}

}  // namespace

cudaError_t launch_hadamard_rows(const float* x, long long rows, long long cols, long long n,
                                 float* out, int sm_count, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  if (n < 1 || (n & (n - 1)) || cols % n) return cudaErrorInvalidValue;
  // fl(1 / fl(sqrt(n))) in fp32, like `h / np.sqrt(n, dtype=np.float32)` entry by entry.
  const float norm = 1.0f / sqrtf(static_cast<float>(n));
  const bool smem_ok = cols <= 16384 && cols % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
                       reinterpret_cast<uintptr_t>(out) % 16 == 0;
  static const bool no_tiles = getenv("AEQB_HADAMARD_NO_TILES") != nullptr;  // A/B runs
  if (!no_tiles && n >= 256 && n <= kTileFloats && (rows * cols) % kTileFloats == 0 &&
      reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0) {
    // segments are consecutive n-float runs of the flattened tensor: 16 KiB tiles hold whole ones
    const long long ntiles = rows * cols / kTileFloats;
    switch (n) {
      case 256: return launch_tiles_k<8>(x, ntiles, norm, out, sm_count, st);
      case 512: return launch_tiles_k<9>(x, ntiles, norm, out, sm_count, st);
      case 1024: return launch_tiles_k<10>(x, ntiles, norm, out, sm_count, st);
      case 2048: return launch_tiles_k<11>(x, ntiles, norm, out, sm_count, st);
      default: return launch_tiles_k<12>(x, ntiles, norm, out, sm_count, st);
    }
  }
  if (smem_ok && cols % 256 == 0 && n <= 8192) {  // n = 16384 would need 256 data registers
    const int c = static_cast<int>(cols), nn = static_cast<int>(n);
    switch (n > 256 ? nn / 256 : 1) {
      case 1: return launch_reg<1>(x, rows, c, nn, norm, out, sm_count, st);
      case 2: return launch_reg<2>(x, rows, c, nn, norm, out, sm_count, st);
      case 4: return launch_reg<4>(x, rows, c, nn, norm, out, sm_count, st);
      case 8: return launch_reg<8>(x, rows, c, nn, norm, out, sm_count, st);
      case 16: return launch_reg<16>(x, rows, c, nn, norm, out, sm_count, st);
      default: return launch_reg<32>(x, rows, c, nn, norm, out, sm_count, st);
    }
  }
  if (smem_ok) {
    const int smem = static_cast<int>(cols) * 4;
    static PerDevice configured;
    if (!configured.done()) {
      cudaError_t e = cudaFuncSetAttribute(hadamard_rows_smem,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
      if (e != cudaSuccess) return e;
      configured.set();
    }
    const int threads = cols >= 8192 ? 512 : (cols >= 1024 ? 256 : 128);
    long long per_sm = 200000 / (smem + 1024);
    if (per_sm > 2048 / threads) per_sm = 2048 / threads;
    if (per_sm < 1) per_sm = 1;
    long long grid = static_cast<long long>(sm_count) * per_sm;
    if (grid > rows) grid = rows;
    hadamard_rows_smem<<<static_cast<unsigned>(grid), threads, smem, st>>>(
        x, rows, static_cast<int>(cols), static_cast<int>(n), norm, out);
    return count_launch();
  }
  const long long total = rows * cols;
  long long grid = (total + 255) / 256;
  if (grid > sm_count * 16LL) grid = sm_count * 16LL;
  scale_copy_global<<<static_cast<unsigned>(grid), 256, 0, st>>>(x, out, total, 1.0f);
  int launches = 1;
  for (long long h = 1; h < n; h *= 2) {
    hadamard_stage_global<<<static_cast<unsigned>(grid), 256, 0, st>>>(out, total / 2, h);
    ++launches;
  }
  scale_copy_global<<<static_cast<unsigned>(grid), 256, 0, st>>>(out, out, total, norm);
  return count_launch(launches + 1);
}

}  // namespace aeqb
