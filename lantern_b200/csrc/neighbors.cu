// Neighbour-table build (entrypoints/generate_codebook.py:53-60), exact variant.
//
// Definition (oracle/lantern_oracle.py neighbor_table): squared L2 distance by direct differences accumulated
// in fp64 in dimension order, self excluded, order by (distance, id).  One CTA per codebook row: distances to
// all N rows are computed straight into shared memory from a transposed copy of the codebook (coalesced
// reads), then the (distance, id) pairs are bitonic-sorted in shared memory and the first K ids written out.
// The N x N distance matrix never touches HBM.
#include <cstdlib>

#include "common.cuh"

namespace lantern {

constexpr int kNbrThreads = 1024;

__global__ void transpose_kernel(const float* __restrict__ E, float* __restrict__ Et, int N, int d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (int64_t)N * d) {
    const int r = (int)(i / d), c = (int)(i % d);
    Et[(int64_t)c * N + r] = E[i];
  }
}

__global__ void __launch_bounds__(kNbrThreads) nbr_rows_kernel(const float* __restrict__ E,
                                                               const float* __restrict__ Et, int N, int d, int K,
                                                               int NP /*pow2 >= N*/, int32_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* key = reinterpret_cast<double*>(smem_raw);
  int* idx = reinterpret_cast<int*>(smem_raw + (size_t)NP * 8);
  float* er = reinterpret_cast<float*>(smem_raw + (size_t)NP * 12);
  const int r = blockIdx.x, tid = threadIdx.x;
  for (int c = tid; c < d; c += kNbrThreads) er[c] = E[(int64_t)r * d + c];
  __syncthreads();
  for (int c = tid; c < NP; c += kNbrThreads) {
    double acc = 0.0;
    if (c < N && c != r) {
      for (int k = 0; k < d; ++k) {
        const double diff = __dsub_rn((double)er[k], (double)Et[(int64_t)k * N + c]);
        acc = __dadd_rn(acc, __dmul_rn(diff, diff));
      }
      idx[c] = c;
    } else {
      acc = INFINITY;
      idx[c] = c == r ? N : N + 1 + c;   // self and padding sort after every real neighbour
    }
    key[c] = acc;
  }
  __syncthreads();
  // bitonic sort ascending by (key, idx)
  for (int size = 2; size <= NP; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (NP >> 1); t += kNbrThreads) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const double ka = key[lo], kb = key[hi];
        const int ia = idx[lo], ib = idx[hi];
        const bool a_gt_b = (ka > kb) || (ka == kb && ia > ib);
        if (a_gt_b == up) {
          key[lo] = kb; key[hi] = ka;
          idx[lo] = ib; idx[hi] = ia;
        }
      }
      __syncthreads();
    }
  }
  for (int c = tid; c < K; c += kNbrThreads) out[(int64_t)r * K + c] = idx[c];
}

}  // namespace lantern

static thread_local int g_last_path = 0;   // 1 = tensor-core candidates + exact re-rank, 2 = all-fp64 kernel

extern "C" LANTERN_API int lantern_debug_neighbors_path(void) { return g_last_path; }

int build_neighbors_tensor_core(const float* E_dev, int N, int d, int K, int32_t* out_dev, cudaStream_t s, int* fell_back);

extern "C" int lantern_build_neighbors(const float* E_dev, int32_t N, int32_t d, int32_t K, int32_t* out_dev,
                                       void* stream) {
  using namespace lantern;
  if (!E_dev || !out_dev || N < 2 || d < 1 || K < 1 || K > N - 1) {
    set_error("lantern_build_neighbors: bad argument (need N >= 2, d >= 1, 1 <= K <= N-1)");
    return LANTERN_E_INVALID;
  }
  int NP = 1;
  while (NP < N) NP <<= 1;
  const size_t smem = (size_t)NP * 12 + (size_t)((d + 3) & ~3) * 4;
  if (smem > 227 * 1024) {
    set_error("lantern_build_neighbors: N=%d needs %zu bytes of shared memory (> 227 KB)", N, smem);
    return LANTERN_E_UNSUPPORTED;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (getenv("LANTERN_NBR_EXACT_ONLY") == nullptr) {   // tensor-core path for K << N (bit-identical results)
    int fell_back = 1;
    const int rc = build_neighbors_tensor_core(E_dev, N, d, K, out_dev, s, &fell_back);
    if (rc != LANTERN_OK) return rc;
    if (!fell_back) { g_last_path = 1; return LANTERN_OK; }
  }
  g_last_path = 2;
  float* Et = nullptr;
  LANTERN_CUDA(cudaMallocAsync(&Et, (size_t)N * d * sizeof(float), s));
  const int64_t n = (int64_t)N * d;
  transpose_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(E_dev, Et, N, d);
  if (smem > 48 * 1024)
    LANTERN_CUDA(cudaFuncSetAttribute(nbr_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nbr_rows_kernel<<<N, kNbrThreads, smem, s>>>(E_dev, Et, N, d, K, NP, out_dev);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(Et, s);
  LANTERN_CUDA(e);
  return LANTERN_OK;
}
