// Neighbour-table build, tensor-core variant (entrypoints/generate_codebook.py:53-60) for K << N.
//
//   1. dist_gemm_kernel — tcgen05: D~[i][j] = |e_i|^2 + |e_j|^2 - 2 e_i.e_j with the cross term on the 5th-gen
//      tensor cores (kind::tf32, operands rounded to TF32 while they are staged into the canonical K-major
//      shared-memory layout, fp32 accumulators in TMEM, one elected thread issues the MMAs, tcgen05.commit ->
//      mbarrier, tcgen05.ld epilogue adds the norms).  D~ is an fp32 [N, N] scratch in HBM (1 GiB at N = 16384).
//   2. nbr_select_kernel — per codebook row: exact K-th smallest approximate distance t (select.cuh), candidate set
//      {j : D~ <= t + 2*eps} where eps bounds |D~ - d^2| (TF32 rounding + fp32 accumulation), then the oracle's
//      arithmetic on the candidates only: squared distance by direct differences in fp64, bitonic sort by
//      (distance, id), first K ids out.  Every true K-nearest neighbour is a candidate (d~ <= d^2 + eps <=
//      D_K + eps <= t + 2 eps), so the result is bit-identical to the exact kernel in neighbors.cu.
//      Candidates are also checked against eps; a violation (or a candidate overflow) raises a flag and the
//      caller falls back to the exact kernel.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "select.cuh"

namespace lantern {

constexpr int kGemmThreads = 128;
constexpr int kTileM = 128, kTileN = 256, kChunkK = 32;   // K is consumed in chunks of 32 (4 MMAs of K = 8)
constexpr int kSelThreads = 512;
constexpr int kSelNE = 32;         // register-resident row: N <= 16384
constexpr int kCandMax = 4096;

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// K-major, no-swizzle ("interleave") canonical layout: core matrix = 8 rows x 16 bytes, rows 16 bytes apart;
// 16-byte K chunks are the outer dimension.  SBO = 128 B (next 8-row group), LBO = rows * 16 B (next K chunk).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                 // base offset 0, layout type 0 = SWIZZLE_NONE
}

// kind::tf32, fp32 accumulate, A and B K-major, M x N tile.
__host__ __device__ constexpr uint32_t make_instr_desc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void row_norms_kernel(const float* __restrict__ E, int N, int d, float* __restrict__ norms) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float s = 0.f;
  for (int k = 0; k < d; ++k) s = fmaf(E[(size_t)i * d + k], E[(size_t)i * d + k], s);
  norms[i] = s;
}

__global__ void __launch_bounds__(kGemmThreads) dist_gemm_kernel(const float* __restrict__ E,
                                                                 const float* __restrict__ norms, int N, int d,
                                                                 float* __restrict__ Dout, int ld) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t* sA = reinterpret_cast<uint32_t*>(smem);                                  // [kChunkK/4][kTileM][4]
  uint32_t* sB = reinterpret_cast<uint32_t*>(smem + kChunkK * kTileM * 4);           // [kChunkK/4][kTileN][4]
  float* sNb = reinterpret_cast<float*>(smem + kChunkK * (kTileM + kTileN) * 4);     // [kTileN]
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kTileM;
  const int dpad = (d + kChunkK - 1) / kChunkK * kChunkK;
  const int n_col_tiles = (N + kTileN - 1) / kTileN;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "n"(kTileN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) mbar_init(&mbar, 1);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t idesc = make_instr_desc(kTileM, kTileN);
  uint32_t parity = 0;
  const int my_row = row0 + tid;   // accumulator lane == thread
  const float na = my_row < N ? norms[my_row] : 0.f;

  for (int ct = blockIdx.y; ct < n_col_tiles; ct += gridDim.y) {
    const int col0 = ct * kTileN;
    for (int j = tid; j < kTileN; j += kGemmThreads) sNb[j] = (col0 + j < N) ? norms[col0 + j] : 0.f;
    for (int kc0 = 0; kc0 < dpad; kc0 += kChunkK) {
      // ---- stage the operand chunks: [K chunk of 4][row][4 tf32] ----
      for (int idx = tid; idx < kTileM * (kChunkK / 4); idx += kGemmThreads) {
        const int r = idx % kTileM, c4 = idx / kTileM;
        uint4 v = make_uint4(0, 0, 0, 0);
        const int gr = row0 + r, k0 = kc0 + c4 * 4;
        if (gr < N) {
          const float* src = E + (size_t)gr * d + k0;
          v.x = k0 + 0 < d ? to_tf32(src[0]) : 0u;
          v.y = k0 + 1 < d ? to_tf32(src[1]) : 0u;
          v.z = k0 + 2 < d ? to_tf32(src[2]) : 0u;
          v.w = k0 + 3 < d ? to_tf32(src[3]) : 0u;
        }
        *reinterpret_cast<uint4*>(sA + ((size_t)c4 * kTileM + r) * 4) = v;
      }
      for (int idx = tid; idx < kTileN * (kChunkK / 4); idx += kGemmThreads) {
        const int r = idx % kTileN, c4 = idx / kTileN;
        uint4 v = make_uint4(0, 0, 0, 0);
        const int gr = col0 + r, k0 = kc0 + c4 * 4;
        if (gr < N) {
          const float* src = E + (size_t)gr * d + k0;
          v.x = k0 + 0 < d ? to_tf32(src[0]) : 0u;
          v.y = k0 + 1 < d ? to_tf32(src[1]) : 0u;
          v.z = k0 + 2 < d ? to_tf32(src[2]) : 0u;
          v.w = k0 + 3 < d ? to_tf32(src[3]) : 0u;
        }
        *reinterpret_cast<uint4*>(sB + ((size_t)c4 * kTileN + r) * 4) = v;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async-proxy (MMA) reads
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
        for (int s = 0; s < kChunkK / 8; ++s) {
          const uint64_t adesc = make_smem_desc(smem_u32(sA) + s * 2 * (kTileM * 16), kTileM * 16, 128);
          const uint64_t bdesc = make_smem_desc(smem_u32(sB) + s * 2 * (kTileN * 16), kTileN * 16, 128);
          const uint32_t accumulate = (kc0 > 0 || s > 0) ? 1u : 0u;
          asm volatile(
              "{\n"
              ".reg .pred p;\n"
              "setp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
              "}\n" ::"r"(tmem_base),
              "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
      }
      mbar_wait(&mbar, parity);   // the MMAs of this chunk are done: operands may be overwritten
      parity ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;");
    }
    // ---- epilogue: TMEM -> registers, add the norms, store the tile row by row ----
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int cb = 0; cb < kTileN; cb += 32) {
      uint32_t r[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(lane_addr + (uint32_t)cb));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (my_row < N) {
        float* dst = Dout + (size_t)my_row * ld + col0 + cb;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float o[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int col = col0 + cb + j + u;
            float v = na + sNb[cb + j + u] - 2.0f * __uint_as_float(r[j + u]);
            if (col == my_row) v = INFINITY;   // self is never a neighbour
            o[u] = v;
          }
          if (col0 + cb + j + 3 < ld) *reinterpret_cast<float4*>(dst + j) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();   // all TMEM reads done before the next tile's MMAs overwrite the accumulator
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTileN));
}

// One CTA per codebook row: exact top-K from the approximate distance row.
__global__ void __launch_bounds__(kSelThreads) nbr_select_kernel(const float* __restrict__ E,
                                                                 const float* __restrict__ norms,
                                                                 const float* __restrict__ Dapprox, int ld, int N,
                                                                 int d, int K, float max_norm,
                                                                 int32_t* __restrict__ out, int* __restrict__ flags) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* key = reinterpret_cast<double*>(smem_raw);                       // [kCandMax]
  int* cidx = reinterpret_cast<int*>(smem_raw + (size_t)kCandMax * 8);      // [kCandMax]
  __shared__ SelectSmem sm;
  __shared__ int n_cand, bad;
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* drow = Dapprox + (size_t)r * ld;
  float v[kSelNE];
#pragma unroll
  for (int e = 0; e < kSelNE; ++e) {
    const int j = e * kSelThreads + tid;
    v[e] = j < N ? -drow[j] : -INFINITY;   // negated: K-th smallest distance = K-th largest value (self is -inf)
  }
  float vmin = INFINITY, vmax = -INFINITY;
#pragma unroll
  for (int e = 0; e < kSelNE; ++e) {
    if (v[e] > -INFINITY) vmin = fminf(vmin, v[e]);
    vmax = fmaxf(vmax, v[e]);
  }
  vmin = -block_reduce(-vmin, OpMaxF(), -INFINITY, sm.f4[2]);
  vmax = block_reduce(vmax, OpMaxF(), -INFINITY, sm.f4[3]);
  // tier 2/3 selectors expect finite extremes; -inf entries (self, padding) simply rank last
  float kth;
  {
    float tmp[kSelNE];
#pragma unroll
    for (int e = 0; e < kSelNE; ++e) tmp[e] = v[e] > -INFINITY ? v[e] : vmin - 1.0f;
    kth = select_slow<kSelNE>(tmp, K, vmin - 1.0f, vmax, sm);
  }
  // error bound of the TF32 cross term + fp32 norms (see header): |D~ - d^2| <= eps
  const float na = sqrtf(norms[r]);
  const float eps = 1.5f * (0.0025f * na * max_norm + 6e-5f * (na * na + max_norm * max_norm)) + 1e-30f;
  const float cut = -kth + 2.0f * eps;
  if (tid == 0) { n_cand = 0; bad = 0; }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < kSelNE; ++e) {
    const int j = e * kSelThreads + tid;
    if (j < N && j != r && -v[e] <= cut) {
      const int p = atomicAdd(&n_cand, 1);
      if (p < kCandMax) cidx[p] = j;
    }
  }
  __syncthreads();
  const int nc = n_cand;
  if (nc > kCandMax || nc < K) {   // overflow (massive ties) or an inconsistent GEMM: let the exact kernel handle it
    if (tid == 0) atomicAdd(&flags[0], 1);
    return;
  }
  int np = 1;
  while (np < nc) np <<= 1;
  // exact squared distances of the candidates, oracle arithmetic (fp64, dimension order, no contraction)
  const float* er = E + (size_t)r * d;
  for (int c = tid; c < np; c += kSelThreads) {
    if (c < nc) {
      const int j = cidx[c];
      const float* ej = E + (size_t)j * d;
      double acc = 0.0;
      for (int k = 0; k < d; ++k) {
        const double diff = __dsub_rn((double)er[k], (double)ej[k]);
        acc = __dadd_rn(acc, __dmul_rn(diff, diff));
      }
      key[c] = acc;
      if (fabs(acc - (double)drow[j]) > (double)eps) bad = 1;   // the bound must hold, otherwise D~ is not trustworthy
    } else {
      key[c] = INFINITY;
      cidx[c] = N + c;
    }
  }
  __syncthreads();
  if (bad) {
    if (tid == 0) atomicAdd(&flags[1], 1);
    return;
  }
  for (int size = 2; size <= np; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (np >> 1); t += kSelThreads) {
        const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
        const bool up = (lo & size) == 0;
        const double ka = key[lo], kb = key[hi];
        const int ia = cidx[lo], ib = cidx[hi];
        const bool a_gt_b = (ka > kb) || (ka == kb && ia > ib);
        if (a_gt_b == up) { key[lo] = kb; key[hi] = ka; cidx[lo] = ib; cidx[hi] = ia; }
      }
      __syncthreads();
    }
  }
  for (int c = tid; c < K; c += kSelThreads) out[(size_t)r * K + c] = cidx[c];
}

__global__ void max_norm_kernel(const float* __restrict__ norms, int N, float* __restrict__ out) {
  __shared__ float scr[33];
  float m = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) m = fmaxf(m, norms[i]);
  m = block_reduce(m, OpMaxF(), 0.f, scr);
  if (threadIdx.x == 0) out[0] = sqrtf(m);
}

}  // namespace lantern

using namespace lantern;

// Debug / test hook: the approximate distance matrix alone (fp32 [N, ld], ld = N rounded up to 4).
extern "C" LANTERN_API int lantern_debug_dist_gemm(const float* E_dev, int32_t N, int32_t d, float* D_dev, int32_t ld,
                                                   void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!E_dev || !D_dev || N < 2 || d < 1 || ld < N || ld % 4) {
    set_error("lantern_debug_dist_gemm: bad argument");
    return LANTERN_E_INVALID;
  }
  float* norms = nullptr;
  LANTERN_CUDA(cudaMallocAsync(&norms, (size_t)N * 4, s));
  row_norms_kernel<<<(N + 255) / 256, 256, 0, s>>>(E_dev, N, d, norms);
  const size_t smem = (size_t)kChunkK * (kTileM + kTileN) * 4 + kTileN * 4;
  LANTERN_CUDA(cudaFuncSetAttribute(dist_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int row_blocks = (N + kTileM - 1) / kTileM, col_tiles = (N + kTileN - 1) / kTileN;
  const int gy = std::max(1, std::min(col_tiles, (2 * kNumSMs + row_blocks - 1) / row_blocks));
  dist_gemm_kernel<<<dim3(row_blocks, gy), kGemmThreads, smem, s>>>(E_dev, norms, N, d, D_dev, ld);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(norms, s);
  LANTERN_CUDA(e);
  return LANTERN_OK;
}

// Returns LANTERN_OK and *fell_back = 0 when the tensor-core path produced the table, *fell_back = 1 when the caller
// must run the exact kernel (unsupported shape, candidate overflow or a violated error bound).
int build_neighbors_tensor_core(const float* E_dev, int N, int d, int K, int32_t* out_dev, cudaStream_t s,
                                int* fell_back) {
  *fell_back = 1;
  if (N > kSelNE * kSelThreads || K > kCandMax / 2 || K * 2 > N) return LANTERN_OK;
  const int ld = (N + 3) & ~3;
  float *norms = nullptr, *D = nullptr, *mx = nullptr;
  int* flags = nullptr;
  {   // keep the (up to 1 GiB) scratch in the stream-ordered pool between calls instead of returning it to the OS
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  const bool timing = getenv("LANTERN_NBR_TIMING") != nullptr;
  cudaEvent_t ev[3];
  if (timing) for (auto& e : ev) cudaEventCreate(&e);
  LANTERN_CUDA(cudaMallocAsync(&norms, (size_t)N * 4, s));
  LANTERN_CUDA(cudaMallocAsync(&mx, 4, s));
  LANTERN_CUDA(cudaMallocAsync(&flags, 8, s));
  LANTERN_CUDA(cudaMallocAsync(&D, (size_t)N * ld * 4, s));
  LANTERN_CUDA(cudaMemsetAsync(flags, 0, 8, s));
  if (timing) cudaEventRecord(ev[0], s);
  row_norms_kernel<<<(N + 255) / 256, 256, 0, s>>>(E_dev, N, d, norms);
  max_norm_kernel<<<1, 512, 0, s>>>(norms, N, mx);
  const size_t smem = (size_t)kChunkK * (kTileM + kTileN) * 4 + kTileN * 4;
  LANTERN_CUDA(cudaFuncSetAttribute(dist_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int row_blocks = (N + kTileM - 1) / kTileM, col_tiles = (N + kTileN - 1) / kTileN;
  const int gy = std::max(1, std::min(col_tiles, (2 * kNumSMs + row_blocks - 1) / row_blocks));
  dist_gemm_kernel<<<dim3(row_blocks, gy), kGemmThreads, smem, s>>>(E_dev, norms, N, d, D, ld);
  if (timing) cudaEventRecord(ev[1], s);
  float h_mx = 0.f;
  LANTERN_CUDA(cudaMemcpyAsync(&h_mx, mx, 4, cudaMemcpyDeviceToHost, s));
  LANTERN_CUDA(cudaStreamSynchronize(s));
  const size_t sel_smem = (size_t)kCandMax * 12;
  LANTERN_CUDA(cudaFuncSetAttribute(nbr_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
  nbr_select_kernel<<<N, kSelThreads, sel_smem, s>>>(E_dev, norms, D, ld, N, d, K, h_mx, out_dev, flags);
  if (timing) cudaEventRecord(ev[2], s);
  int h_flags[2] = {0, 0};
  LANTERN_CUDA(cudaMemcpyAsync(h_flags, flags, 8, cudaMemcpyDeviceToHost, s));
  LANTERN_CUDA(cudaStreamSynchronize(s));
  if (timing) {
    float g = 0.f, q = 0.f;
    cudaEventElapsedTime(&g, ev[0], ev[1]);
    cudaEventElapsedTime(&q, ev[1], ev[2]);
    fprintf(stderr, "[lantern] neighbours N=%d d=%d K=%d: distance GEMM %.3f ms, select+rerank %.3f ms, flags overflow=%d bound=%d\n", N, d, K, g, q, h_flags[0], h_flags[1]);
    for (auto& e : ev) cudaEventDestroy(e);
  }
  cudaFreeAsync(norms, s); cudaFreeAsync(mx, s); cudaFreeAsync(flags, s); cudaFreeAsync(D, s);
  LANTERN_CUDA(cudaGetLastError());
  *fell_back = (h_flags[0] || h_flags[1]) ? 1 : 0;
  return LANTERN_OK;
}
