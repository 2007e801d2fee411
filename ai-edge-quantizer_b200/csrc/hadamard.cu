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
// Shared-memory butterfly: a CTA owns one row at a time (cols <= 16384 floats =
// 64 KiB), loads it with 128-bit loads, runs radix-4 passes (two butterfly stages
// per pass, each thread owns a 4-point group in registers, so shared memory is
// touched once per two stages), a final radix-2 pass when m is odd, then scales
// and streams the row out.  Rounding differs from the reference's sgemm only in
// summation order (and in "sum then scale" vs "scale then sum" when n is not a
// power of 4); DESIGN.md states the tolerance.
#include "aeqb_common.cuh"
#include "aeqb_kernels.h"

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
  static bool configured = false;
  if (!configured && G > 1) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    if (e != cudaSuccess) return e;
    configured = true;
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
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(hadamard_rows_smem,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
      if (e != cudaSuccess) return e;
      configured = true;
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
