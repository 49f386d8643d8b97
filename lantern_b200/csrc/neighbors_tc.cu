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

constexpr int kGemmThreads = 320;   // 8 epilogue warps + TMA producer warp + MMA issuer warp
constexpr int kMaxStages = 6;
constexpr int kGemmMaxDim = 256;   // the resident A block needs dpad * 512 bytes of shared memory
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

// Packs the codebook into the tensor-core operand layout once: TF32-rounded, zero-padded to [dpad/4][Npad][4], i.e.
// for every 16-byte K chunk the rows are contiguous.  A 128- or 256-row operand tile of one K chunk is then one
// contiguous 2 / 4 KiB run that a single cp.async.bulk moves straight into the canonical K-major SWIZZLE_NONE
// shared-memory layout (core matrix = 8 rows x 16 bytes, SBO = 128 B, LBO = rows * 16 B).
__global__ void pack_tf32_kernel(const float* __restrict__ E, int N, int d, int Npad, int dpad,
                                 uint32_t* __restrict__ Epk) {
  const int n4 = dpad / 4;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (size_t)Npad * n4;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % n4), row = (int)(idx / n4);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row < N) {
      const float* src = E + (size_t)row * d;
      const int k0 = c4 * 4;
      if (k0 + 0 < d) v.x = to_tf32(src[k0 + 0]);
      if (k0 + 1 < d) v.y = to_tf32(src[k0 + 1]);
      if (k0 + 2 < d) v.z = to_tf32(src[k0 + 2]);
      if (k0 + 3 < d) v.w = to_tf32(src[k0 + 3]);
    }
    *reinterpret_cast<uint4*>(Epk + ((size_t)c4 * Npad + row) * 4) = v;
  }
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One thread's 32 accumulator columns [cb, cb+32) of its row: D~ = |a|^2 + |b|^2 - 2 a.b, stored as full 32-byte sectors.
__device__ __forceinline__ void epilogue_block(const uint32_t (&r)[32], float na, const float* __restrict__ nbh,
                                               float* __restrict__ drow, int cb, int colbase, int ld, bool row_ok,
                                               bool wide) {
  if (!row_ok) return;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const float4 n0 = *reinterpret_cast<const float4*>(nbh + cb + j);
    const float4 n1 = *reinterpret_cast<const float4*>(nbh + cb + j + 4);
    float o[8];
    o[0] = fmaf(-2.0f, __uint_as_float(r[j + 0]), na + n0.x);
    o[1] = fmaf(-2.0f, __uint_as_float(r[j + 1]), na + n0.y);
    o[2] = fmaf(-2.0f, __uint_as_float(r[j + 2]), na + n0.z);
    o[3] = fmaf(-2.0f, __uint_as_float(r[j + 3]), na + n0.w);
    o[4] = fmaf(-2.0f, __uint_as_float(r[j + 4]), na + n1.x);
    o[5] = fmaf(-2.0f, __uint_as_float(r[j + 5]), na + n1.y);
    o[6] = fmaf(-2.0f, __uint_as_float(r[j + 6]), na + n1.z);
    o[7] = fmaf(-2.0f, __uint_as_float(r[j + 7]), na + n1.w);
    const int c = colbase + cb + j;
    if (wide) {   // one full 32-byte sector per thread (256-bit store, sm_100)
      if (c + 7 < ld)
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(drow + cb + j), "f"(o[0]),
                     "f"(o[1]), "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
                     : "memory");
    } else {
      if (c + 3 < ld) *reinterpret_cast<float4*>(drow + cb + j) = make_float4(o[0], o[1], o[2], o[3]);
      if (c + 7 < ld) *reinterpret_cast<float4*>(drow + cb + j + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// Warp-specialised persistent distance GEMM.  Grid (row blocks of 128, column-tile groups); per CTA:
//   warp 8 (one lane)  TMA producer: the resident A block once, then B chunks through a kStages-deep ring
//   warp 9 (one lane)  tcgen05.mma issuer: 4 MMAs (K = 8 each) per chunk into one of two 256-column TMEM accumulators
//   warps 0-7          epilogue: tcgen05.ld of the other accumulator, + |a|^2 + |b|^2, fp32 stores (the diagonal
//                      is left as computed, ~0; consumers exclude self by index)
// so loads, tensor-core math and the store of the previous tile overlap.
__global__ void __launch_bounds__(kGemmThreads) dist_gemm_kernel(const uint32_t* __restrict__ Epk,
                                                                 const float* __restrict__ norms, int N, int Npad,
                                                                 int dpad, int n_stages, float* __restrict__ Dout,
                                                                 int ld) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int n_chunks = dpad / kChunkK;
  constexpr uint32_t kStageBytes = kChunkK * kTileN * 4;
  uint32_t* sA = reinterpret_cast<uint32_t*>(smem);                                 // [dpad/4][kTileM][4]
  unsigned char* sB = smem + (size_t)dpad * kTileM * 4;                             // [n_stages][kChunkK/4][kTileN][4]
  float* sNb = reinterpret_cast<float*>(sB + (size_t)n_stages * kStageBytes);       // [2][kTileN]
  __shared__ __align__(8) uint64_t bar_a, bar_full[kMaxStages], bar_empty[kMaxStages], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kTileM;
  const int n_col_tiles = (N + kTileN - 1) / kTileN;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "n"(2 * kTileN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar_a, 1);
    for (int i = 0; i < n_stages; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_tfull[i], 1); mbar_init(&bar_tempty[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(&bar_a, (uint32_t)dpad * kTileM * 4);
      for (int c4 = 0; c4 < dpad / 4; ++c4)
        bulk_g2s(sA + (size_t)c4 * kTileM * 4, Epk + ((size_t)c4 * Npad + row0) * 4, kTileM * 16, &bar_a);
      int it = 0;
      for (int ct = blockIdx.y; ct < n_col_tiles; ct += gridDim.y) {
        const int col0 = ct * kTileN;
        for (int c = 0; c < n_chunks; ++c, ++it) {
          const int st = it % n_stages;
          const uint32_t ph = (uint32_t)(it / n_stages) & 1u;
          mbar_wait(&bar_empty[st], ph ^ 1u);   // first pass over the ring falls through
          mbar_expect_tx(&bar_full[st], kStageBytes);
          unsigned char* dst = sB + (size_t)st * kStageBytes;
#pragma unroll
          for (int q = 0; q < kChunkK / 4; ++q)
            bulk_g2s(dst + (size_t)q * kTileN * 16, Epk + ((size_t)(c * (kChunkK / 4) + q) * Npad + col0) * 4,
                     kTileN * 16, &bar_full[st]);
        }
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      const uint32_t idesc = make_instr_desc(kTileM, kTileN);
      mbar_wait(&bar_a, 0);
      int it = 0, t = 0;
      for (int ct = blockIdx.y; ct < n_col_tiles; ct += gridDim.y, ++t) {
        const int b = t & 1;
        mbar_wait(&bar_tempty[b], ((uint32_t)(t >> 1) & 1u) ^ 1u);   // the epilogue drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t d_tmem = tmem_base + (uint32_t)(b * kTileN);
        for (int c = 0; c < n_chunks; ++c, ++it) {
          const int st = it % n_stages;
          mbar_wait(&bar_full[st], (uint32_t)(it / n_stages) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;");
          const uint32_t b_addr = smem_u32(sB + (size_t)st * kStageBytes);
#pragma unroll
          for (int q = 0; q < kChunkK / 8; ++q) {
            const uint32_t a_addr = smem_u32(sA) + (uint32_t)((c * (kChunkK / 4) + q * 2) * (kTileM * 16));
            const uint64_t adesc = make_smem_desc(a_addr, kTileM * 16, 128);
            const uint64_t bdesc = make_smem_desc(b_addr + q * 2 * (kTileN * 16), kTileN * 16, 128);
            const uint32_t accumulate = (c > 0 || q > 0) ? 1u : 0u;
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "setp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
                "}\n" ::"r"(d_tmem),
                "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
          }
          // frees the ring slot once the MMAs that read it have completed
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&bar_empty[st]))
                       : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&bar_tfull[b]))
                     : "memory");
      }
    }
  } else {
    // ---- epilogue warps 0-7: warp w reads TMEM lanes [32(w%4), +32) = tile rows, column half w/4 ----
    const int quarter = warp & 3, half = warp >> 2;
    const int my_row = row0 + quarter * 32 + lane;
    const float na = my_row < N ? norms[my_row] : 0.f;
    const bool wide = (ld % 8 == 0) && (reinterpret_cast<uintptr_t>(Dout) % 32 == 0);
    int t = 0;
    for (int ct = blockIdx.y; ct < n_col_tiles; ct += gridDim.y, ++t) {
      const int b = t & 1, col0 = ct * kTileN;
      float* nb = sNb + b * kTileN;
      // nb was last read two tiles ago by these same warps; the named barrier orders those reads before the refill
      asm volatile("bar.sync 1, 256;" ::: "memory");
      nb[tid] = (col0 + tid < N) ? norms[col0 + tid] : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&bar_tfull[b], (uint32_t)(t >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * kTileN + half * 128);
      float* drow = Dout + (size_t)my_row * ld + col0 + half * 128;
      const float* nbh = nb + half * 128;
      // software pipeline over the four 32-column blocks: the TMEM load of block i+1 is in flight while block i is
      // finished and stored (tcgen05.wait::ld covers every outstanding load, so it sits right before the issue)
      uint32_t ra[32], rb[32];
      tmem_ld32(lane_addr, ra);
      tmem_ld_wait();
      tmem_ld32(lane_addr + 32u, rb);
      epilogue_block(ra, na, nbh, drow, 0, col0 + half * 128, ld, my_row < N, wide);
      tmem_ld_wait();
      tmem_ld32(lane_addr + 64u, ra);
      epilogue_block(rb, na, nbh, drow, 32, col0 + half * 128, ld, my_row < N, wide);
      tmem_ld_wait();
      tmem_ld32(lane_addr + 96u, rb);
      epilogue_block(ra, na, nbh, drow, 64, col0 + half * 128, ld, my_row < N, wide);
      tmem_ld_wait();
      epilogue_block(rb, na, nbh, drow, 96, col0 + half * 128, ld, my_row < N, wide);
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[b]);   // 8 arrivals (one per epilogue warp) release the accumulator
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * kTileN));
}

// Bitonic sort of EPT * kSelThreads (key, id) pairs, ascending by (key, id).  Element i = tid * EPT + e lives in a
// register of thread tid: compare-exchange partners at distance j < EPT are in the same thread, at EPT <= j < 32 * EPT in
// another lane of the warp (shuffles), beyond that in another warp (shared memory, two barriers).  For 2048 pairs that is
// 10 shared-memory phases instead of the 66 of the all-shared-memory network.
template <int EPT>
__device__ __forceinline__ void bitonic_hybrid(double* __restrict__ key, int* __restrict__ idx) {
  const int tid = threadIdx.x;
  constexpr int NP = EPT * kSelThreads;
  double k[EPT];
  int ix[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) { k[e] = key[tid * EPT + e]; ix[e] = idx[tid * EPT + e]; }
  __syncthreads();
#pragma unroll 1
  for (int size = 2; size <= NP; size <<= 1) {
#pragma unroll 1
    for (int j = size >> 1; j > 0; j >>= 1) {
      if (j >= 32 * EPT) {                       // partner in another warp
#pragma unroll
        for (int e = 0; e < EPT; ++e) { key[tid * EPT + e] = k[e]; idx[tid * EPT + e] = ix[e]; }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          const int i = tid * EPT + e;
          const double pk = key[i ^ j];
          const int pix = idx[i ^ j];
          const bool mine_gt = (k[e] > pk) || (k[e] == pk && ix[e] > pix);
          const bool keep_min = ((i & j) == 0) == ((i & size) == 0);
          if (mine_gt == keep_min) { k[e] = pk; ix[e] = pix; }
        }
        __syncthreads();
      } else if (j >= EPT) {                     // partner in another lane
        const int lx = j / EPT;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          const int i = tid * EPT + e;
          const double pk = __shfl_xor_sync(0xffffffffu, k[e], lx);
          const int pix = __shfl_xor_sync(0xffffffffu, ix[e], lx);
          const bool mine_gt = (k[e] > pk) || (k[e] == pk && ix[e] > pix);
          const bool keep_min = ((i & j) == 0) == ((i & size) == 0);
          if (mine_gt == keep_min) { k[e] = pk; ix[e] = pix; }
        }
      } else {                                   // partner in the same thread
#pragma unroll
        for (int JJ = EPT >> 1; JJ > 0; JJ >>= 1) {
          if (j == JJ) {
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
              if ((e & JJ) == 0) {
                const int i = tid * EPT + e;
                const bool up = (i & size) == 0;
                const bool a_gt_b = (k[e] > k[e | JJ]) || (k[e] == k[e | JJ] && ix[e] > ix[e | JJ]);
                if (a_gt_b == up) {
                  const double tk = k[e]; k[e] = k[e | JJ]; k[e | JJ] = tk;
                  const int ti = ix[e]; ix[e] = ix[e | JJ]; ix[e | JJ] = ti;
                }
              }
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < EPT; ++e) { key[tid * EPT + e] = k[e]; idx[tid * EPT + e] = ix[e]; }
  __syncthreads();
}

// One CTA per codebook row: exact top-K from the approximate distance row.
template <int NE>
__global__ void __launch_bounds__(kSelThreads, 2) nbr_select_kernel(const float* __restrict__ E,
                                                                 const float* __restrict__ norms,
                                                                 const float* __restrict__ Dapprox, int ld, int N,
                                                                 int d, int K, const float* __restrict__ max_norm_dev,
                                                                 float z_guess, int32_t* __restrict__ out,
                                                                 int* __restrict__ flags,
                                                                 unsigned char* __restrict__ row_redo) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* key = reinterpret_cast<double*>(smem_raw);                       // [kCandMax]
  int* cidx = reinterpret_cast<int*>(smem_raw + (size_t)kCandMax * 8);      // [kCandMax]
  double* erd = reinterpret_cast<double*>(smem_raw + (size_t)kCandMax * 12);   // [kGemmMaxDim]
  __shared__ SelectSmem sm;
  __shared__ int n_cand, bad;
  const int r = blockIdx.x, tid = threadIdx.x;
  const float max_norm = __ldg(max_norm_dev);
  const float* drow = Dapprox + (size_t)r * ld;
  float v[NE];
  // element e of thread tid is column col_of(e): 128-bit loads (rows are 16-byte aligned, ld % 4 == 0)
#pragma unroll
  for (int e4 = 0; e4 < NE / 4; ++e4) {
    const int j = (e4 * kSelThreads + tid) * 4;
    float4 f = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
    if (j < ld) f = *reinterpret_cast<const float4*>(drow + j);
    // negated: K-th smallest distance = K-th largest value (self is -inf)
    v[4 * e4 + 0] = (j + 0 < N && j + 0 != r) ? -f.x : -INFINITY;
    v[4 * e4 + 1] = (j + 1 < N && j + 1 != r) ? -f.y : -INFINITY;
    v[4 * e4 + 2] = (j + 2 < N && j + 2 != r) ? -f.z : -INFINITY;
    v[4 * e4 + 3] = (j + 3 < N && j + 3 != r) ? -f.w : -INFINITY;
  }
  float vmin = INFINITY, vmax = -INFINITY, vsum = 0.f, vsq = 0.f;
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    if (v[e] > -INFINITY) {
      vmin = fminf(vmin, v[e]);
      vsum += v[e];
      vsq = fmaf(v[e], v[e], vsq);
    }
    vmax = fmaxf(vmax, v[e]);
  }
  vmin = -block_reduce(-vmin, OpMaxF(), -INFINITY, sm.f4[2]);
  vmax = block_reduce(vmax, OpMaxF(), -INFINITY, sm.f4[3]);
  vsum = block_reduce(vsum, OpSum(), 0.f, sm.f4[0]);
  vsq = block_reduce(vsq, OpSum(), 0.f, sm.f4[1]);
  // tier 2/3 selectors expect finite extremes; -inf entries (self, padding) simply rank last
  float kth;
  {
    float tmp[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) tmp[e] = v[e] > -INFINITY ? v[e] : vmin - 1.0f;
    // First classification range: one standard deviation around the Gaussian guess of the K-th largest value instead
    // of [min, max] - values outside fall into the two edge fields, and if the target is there the selector re-centres
    // on that field's exact range, so any guess keeps the result exact; a good one saves one or two passes.
    const float inv_n = 1.0f / (float)(N - 1);
    const float mean = vsum * inv_n;
    const float sd = sqrtf(fmaxf(vsq * inv_n - mean * mean, 0.f));
    float lo = fmaxf(vmin - 1.0f, mean + (z_guess - 0.5f) * sd), hi = fminf(vmax, mean + (z_guess + 0.5f) * sd);
    if (!(sd > 0.f) || !(lo < hi) || !isfinite(lo) || !isfinite(hi)) { lo = vmin - 1.0f; hi = vmax; }
    kth = select_slow<NE>(tmp, K, lo, hi, sm);
  }
  // error bound of the TF32 cross term + fp32 norms (see header): |D~ - d^2| <= eps
  const float na = sqrtf(norms[r]);
  const float eps = 1.5f * (0.0025f * na * max_norm + 6e-5f * (na * na + max_norm * max_norm)) + 1e-30f;
  const float cut = -kth + 2.0f * eps;
  if (tid == 0) { n_cand = 0; bad = 0; }
  __syncthreads();
  {   // one shared atomic per warp: the thread counts its candidates, the warp reserves a range, lanes fill their slices
    int mine = 0;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int j = ((e >> 2) * kSelThreads + tid) * 4 + (e & 3);
      mine += (j < N && j != r && -v[e] <= cut) ? 1 : 0;
    }
    const int lane = tid & 31;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    int base = 0;
    if (lane == 31) base = atomicAdd(&n_cand, incl);
    int p = __shfl_sync(0xffffffffu, base, 31) + incl - mine;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int j = ((e >> 2) * kSelThreads + tid) * 4 + (e & 3);
      if (j < N && j != r && -v[e] <= cut) {
        if (p < kCandMax) cidx[p] = j;
        ++p;
      }
    }
  }
  __syncthreads();
  const int nc = n_cand;
  if (nc > kCandMax || nc < K) {   // overflow (massive ties) or an inconsistent GEMM: let the exact kernel handle it
    if (tid == 0) { atomicAdd(&flags[0], 1); row_redo[r] = 1; }
    return;
  }
  int np = 1;
  while (np < nc) np <<= 1;
  // exact squared distances of the candidates, oracle arithmetic (fp64, dimension order, no contraction)
  // The query row is staged once as fp64; each thread then streams its candidate's row with 128-bit loads (the
  // accumulation order over the dimensions is the oracle's, so the sum cannot be split across lanes).
  const float* er = E + (size_t)r * d;
  for (int k = tid; k < d; k += kSelThreads) erd[k] = (double)er[k];
  __syncthreads();
  const bool vec4 = (d % 4 == 0) && (reinterpret_cast<uintptr_t>(E) % 16 == 0);
  for (int c = tid; c < np; c += kSelThreads) {
    if (c < nc) {
      const int j = cidx[c];
      const float* ej = E + (size_t)j * d;
      double acc = 0.0;
      if (vec4) {
        const float4* ej4 = reinterpret_cast<const float4*>(ej);
#pragma unroll 4
        for (int k4 = 0; k4 < d / 4; ++k4) {
          const float4 f = __ldg(ej4 + k4);
          double diff = __dsub_rn(erd[4 * k4 + 0], (double)f.x);
          acc = __dadd_rn(acc, __dmul_rn(diff, diff));
          diff = __dsub_rn(erd[4 * k4 + 1], (double)f.y);
          acc = __dadd_rn(acc, __dmul_rn(diff, diff));
          diff = __dsub_rn(erd[4 * k4 + 2], (double)f.z);
          acc = __dadd_rn(acc, __dmul_rn(diff, diff));
          diff = __dsub_rn(erd[4 * k4 + 3], (double)f.w);
          acc = __dadd_rn(acc, __dmul_rn(diff, diff));
        }
      } else {
        for (int k = 0; k < d; ++k) {
          const double diff = __dsub_rn(erd[k], (double)ej[k]);
          acc = __dadd_rn(acc, __dmul_rn(diff, diff));
        }
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
    if (tid == 0) { atomicAdd(&flags[1], 1); row_redo[r] = 1; }
    return;
  }
  if (np == 8 * kSelThreads) bitonic_hybrid<8>(key, cidx);
  else if (np == 4 * kSelThreads) bitonic_hybrid<4>(key, cidx);
  else if (np == 2 * kSelThreads) bitonic_hybrid<2>(key, cidx);
  else {
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

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// Scratch layout of the tensor-core route (all caller-owned, carved from one workspace).
struct TcScratch {
  size_t norms, mx, flags, redo, redo_list, epk, D, total;
  int ld, dpad, Npad;
};
static TcScratch tc_scratch(int N, int d) {
  TcScratch t;
  t.ld = (N + 7) & ~7;   // rows start on 32-byte sectors (256-bit epilogue stores)
  t.dpad = (d + kChunkK - 1) / kChunkK * kChunkK;
  t.Npad = (N + kTileN - 1) / kTileN * kTileN;
  size_t o = 0;
  t.norms = o; o += align256((size_t)N * 4);
  t.mx = o;    o += 256;
  t.flags = o; o += 256;
  t.redo = o;  o += align256((size_t)N);
  t.redo_list = o; o += align256((size_t)N * 4);
  t.epk = o;   o += align256((size_t)t.Npad * t.dpad * 4);
  t.D = o;     o += align256((size_t)N * t.ld * 4);
  t.total = o;
  return t;
}

bool neighbors_tc_eligible(int N, int d, int K) {
  return !(N > kSelNE * kSelThreads || K > kCandMax / 2 || K * 2 > N || d > kGemmMaxDim);
}
size_t neighbors_tc_workspace_bytes(int N, int d) { return tc_scratch(N, d).total; }

// norms must already hold |e_i|^2.  Packs E into Epk, then runs the persistent GEMM; D is fp32 [N, ld].
static int launch_dist_gemm(const float* E_dev, const float* norms, int N, int d, uint32_t* Epk, float* D, int ld,
                            cudaStream_t s) {
  const int dpad = (d + kChunkK - 1) / kChunkK * kChunkK;
  const int Npad = (N + kTileN - 1) / kTileN * kTileN;
  pack_tf32_kernel<<<2 * kNumSMs, 256, 0, s>>>(E_dev, N, d, Npad, dpad, Epk);
  const size_t a_bytes = (size_t)dpad * kTileM * 4, stage = (size_t)kChunkK * kTileN * 4, tail = 2 * kTileN * 4;
  const size_t budget = 227 * 1024 - 512;   // static shared memory (barriers) comes out of the same 227 KiB
  const int n_stages = (int)std::min<size_t>(kMaxStages, (budget - a_bytes - tail) / stage);
  const size_t smem = a_bytes + n_stages * stage + tail;   // > 113 KiB always: one CTA (and one 512-column TMEM allocation) per SM
  LANTERN_CUDA(cudaFuncSetAttribute(dist_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int row_blocks = (N + kTileM - 1) / kTileM, col_tiles = (N + kTileN - 1) / kTileN;
  const int gy = std::max(1, std::min(col_tiles, kNumSMs / row_blocks));
  dist_gemm_kernel<<<dim3(row_blocks, gy), kGemmThreads, smem, s>>>(Epk, norms, N, Npad, dpad, n_stages, D, ld);
  LANTERN_CUDA(cudaGetLastError());
  return LANTERN_OK;
}

// Debug / test hook: the approximate distance matrix alone (fp32 [N, ld], ld = N rounded up to 4).  The only entry
// that allocates (stream-ordered scratch for the packed codebook and the norms): it is not on any product path.
extern "C" LANTERN_API int lantern_debug_dist_gemm(const float* E_dev, int32_t N, int32_t d, float* D_dev, int32_t ld,
                                                   void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!E_dev || !D_dev || N < 2 || d < 1 || d > kGemmMaxDim || ld < N || ld % 4) {
    set_error("lantern_debug_dist_gemm: bad argument");
    return LANTERN_E_INVALID;
  }
  const TcScratch t = tc_scratch(N, d);
  float* norms = nullptr;
  uint32_t* Epk = nullptr;
  LANTERN_CUDA(cudaMallocAsync(&norms, (size_t)N * 4, s));
  LANTERN_CUDA(cudaMallocAsync(&Epk, (size_t)t.Npad * t.dpad * 4, s));
  row_norms_kernel<<<(N + 255) / 256, 256, 0, s>>>(E_dev, N, d, norms);
  int rc = launch_dist_gemm(E_dev, norms, N, d, Epk, D_dev, ld, s);
  cudaFreeAsync(norms, s);
  cudaFreeAsync(Epk, s);
  return rc;
}

// Tensor-core route: candidates from the tcgen05 distance GEMM, exact fp64 re-rank per row.  Rows the route cannot
// finish (candidate overflow, violated error bound) are marked in row_redo for the exact kernel; nothing here
// synchronises or allocates.  flags_dev[0..1] count those rows.
int build_neighbors_tensor_core(const float* E_dev, int N, int d, int K, int32_t* out_dev, void* workspace,
                                unsigned char** row_redo_out, int** flags_out, int** redo_list_out, cudaStream_t s) {
  const TcScratch t = tc_scratch(N, d);
  unsigned char* w = static_cast<unsigned char*>(workspace);
  float* norms = reinterpret_cast<float*>(w + t.norms);
  float* mx = reinterpret_cast<float*>(w + t.mx);
  int* flags = reinterpret_cast<int*>(w + t.flags);
  unsigned char* redo = w + t.redo;
  uint32_t* Epk = reinterpret_cast<uint32_t*>(w + t.epk);
  float* D = reinterpret_cast<float*>(w + t.D);
  LANTERN_CUDA(cudaMemsetAsync(w + t.flags, 0, 256 + align256((size_t)N), s));   // flags + row_redo (adjacent)
  row_norms_kernel<<<(N + 255) / 256, 256, 0, s>>>(E_dev, N, d, norms);
  max_norm_kernel<<<1, 512, 0, s>>>(norms, N, mx);
  { int rc = launch_dist_gemm(E_dev, norms, N, d, Epk, D, t.ld, s); if (rc != LANTERN_OK) return rc; }
  const size_t sel_smem = (size_t)kCandMax * 12 + (size_t)kGemmMaxDim * 8;
  // Gaussian guess (in standard deviations of the negated distance row) of the K-th largest of N - 1 values
  float z_guess = 0.f;
  {
    const double p = (double)K / (double)(N - 1);
    double zl = -8.0, zh = 8.0;
    for (int it = 0; it < 60; ++it) {
      const double zm = 0.5 * (zl + zh);
      if (0.5 * erfc(zm / 1.4142135623730951) > p) zl = zm; else zh = zm;
    }
    z_guess = (float)(0.5 * (zl + zh));
  }
  auto run_select = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem);
    if (e != cudaSuccess) return e;
    kern<<<N, kSelThreads, sel_smem, s>>>(E_dev, norms, D, t.ld, N, d, K, mx, z_guess, out_dev, flags, redo);
    return cudaGetLastError();
  };
  if (N <= 8 * kSelThreads) LANTERN_CUDA(run_select(nbr_select_kernel<8>));
  else if (N <= 16 * kSelThreads) LANTERN_CUDA(run_select(nbr_select_kernel<16>));
  else LANTERN_CUDA(run_select(nbr_select_kernel<kSelNE>));
  *row_redo_out = redo;
  *flags_out = flags;
  *redo_list_out = reinterpret_cast<int*>(w + t.redo_list);
  return LANTERN_OK;
}
