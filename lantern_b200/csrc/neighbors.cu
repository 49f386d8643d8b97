// Neighbour-table build (entrypoints/generate_codebook.py:53-60), exact variant.
//
// Definition (oracle/lantern_oracle.py neighbor_table): squared L2 distance by direct differences accumulated
// in fp64 in dimension order, self excluded, order by (distance, id).  One CTA per codebook row: distances to
// all N rows are computed straight into shared memory from a transposed copy of the codebook (coalesced
// reads), then the (distance, id) pairs are bitonic-sorted in shared memory and the first K ids written out.
// The N x N distance matrix never touches HBM.
#include <algorithm>
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

// Rows the tensor-core route marked, gathered into a list (order is irrelevant).
__global__ void nbr_redo_list_kernel(const unsigned char* __restrict__ row_redo, int N, int* __restrict__ list,
                                     int* __restrict__ count) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
    if (row_redo[i]) list[atomicAdd(count, 1)] = i;
}

// Persistent CTAs: every row (redo_list == NULL) or only the listed rows (none in the common case: the CTAs exit at once).
__global__ void __launch_bounds__(kNbrThreads) nbr_rows_kernel(const float* __restrict__ E,
                                                               const float* __restrict__ Et, int N, int d, int K,
                                                               int NP /*pow2 >= N*/, int32_t* __restrict__ out,
                                                               const int* __restrict__ redo_list,
                                                               const int* __restrict__ redo_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* key = reinterpret_cast<double*>(smem_raw);
  int* idx = reinterpret_cast<int*>(smem_raw + (size_t)NP * 8);
  float* er = reinterpret_cast<float*>(smem_raw + (size_t)NP * 12);
  const int tid = threadIdx.x;
  const int n_work = redo_list ? *redo_count : N;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
  const int r = redo_list ? redo_list[w] : w;
  __syncthreads();   // the previous row's readers of the shared arrays are done
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
}

}  // namespace lantern

bool neighbors_tc_eligible(int N, int d, int K);
size_t neighbors_tc_workspace_bytes(int N, int d);
int build_neighbors_tensor_core(const float* E_dev, int N, int d, int K, int32_t* out_dev, void* workspace,
                                unsigned char** row_redo_out, int** flags_out, int** redo_list_out, cudaStream_t s);

// route_dev[0] = 1: every row came from the tensor-core route; 2: the exact kernel computed some (or all) rows;
// route_dev[1] = rows the exact kernel computed.
__global__ void nbr_route_kernel(const int* __restrict__ flags, int N, int tc, int32_t* __restrict__ route) {
  const int redo = tc ? flags[0] + flags[1] : N;
  route[0] = redo == 0 ? 1 : 2;
  route[1] = redo;
}

static size_t exact_bytes(int N, int d) { return ((size_t)N * d * sizeof(float) + 255) & ~size_t(255); }
static bool use_tc(int N, int d, int K) {
  return getenv("LANTERN_NBR_EXACT_ONLY") == nullptr && neighbors_tc_eligible(N, d, K);
}

extern "C" size_t lantern_build_neighbors_workspace_bytes(int32_t N, int32_t d, int32_t K) {
  if (N < 2 || d < 1 || K < 1 || K > N - 1) return 0;
  return exact_bytes(N, d) + (use_tc(N, d, K) ? neighbors_tc_workspace_bytes(N, d) : 0);
}

extern "C" int lantern_build_neighbors(const float* E_dev, int32_t N, int32_t d, int32_t K, int32_t* out_dev,
                                       void* workspace_dev, size_t workspace_bytes, int32_t* route_dev, void* stream) {
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
  const size_t need = lantern_build_neighbors_workspace_bytes(N, d, K);
  if (!workspace_dev || workspace_bytes < need) {
    set_error("lantern_build_neighbors: workspace too small (%zu < %zu)", workspace_bytes, need);
    return LANTERN_E_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* Et = static_cast<float*>(workspace_dev);
  unsigned char* redo = nullptr;
  int *flags = nullptr, *redo_list = nullptr;
  const bool tc = use_tc(N, d, K);
  if (tc) {   // tensor-core candidates + exact re-rank (bit-identical results); rows it cannot finish are marked
    const int rc = build_neighbors_tensor_core(E_dev, N, d, K, out_dev,
                                               static_cast<unsigned char*>(workspace_dev) + exact_bytes(N, d), &redo,
                                               &flags, &redo_list, s);
    if (rc != LANTERN_OK) return rc;
  }
  // exact kernel: every row (no tensor-core route for this shape) or the marked rows only (the other CTAs exit at once)
  const int64_t n = (int64_t)N * d;
  transpose_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(E_dev, Et, N, d);
  if (smem > 48 * 1024)
    LANTERN_CUDA(cudaFuncSetAttribute(nbr_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (tc) nbr_redo_list_kernel<<<kNumSMs, 256, 0, s>>>(redo, N, redo_list, flags + 2);
  nbr_rows_kernel<<<std::min(N, kNumSMs), kNbrThreads, smem, s>>>(E_dev, Et, N, d, K, NP, out_dev, tc ? redo_list : nullptr,
                                                                 tc ? flags + 2 : nullptr);
  if (route_dev) nbr_route_kernel<<<1, 1, 0, s>>>(flags, N, tc ? 1 : 0, route_dev);
  LANTERN_CUDA(cudaGetLastError());
  return LANTERN_OK;
}
